"""The oracle (oracle/lm_oracle.py + the host sequence surgery of ssr_speech_b200/seq.py, which the CUDA path shares) against
the UNMODIFIED reference executed live, over a seeded sweep of span geometries the committed fixtures cannot enumerate:
spans starting at frame 0, ending at the last frame, adjacent spans, empty (insertion) spans, 1-3 spans, TTS, aug_context
with and without CFG rows, kvcache 0/1, greedy and sampled.  Small shapes (1 layer would change the weights: the 2-layer tiny
model is kept, short texts bound the roll-out through the reference's own length guard ssr.py:739) keep a case at a few tens
of milliseconds.

Container only: /root/reference does not exist on the GPU box, so the whole module is skipped there (and it is not a gpu test).
Bar: integer outputs (tokens, marks, both interval lists) identical."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_loader  # noqa: E402
from lm_oracle import LMOracle  # noqa: E402
from ssr_speech_b200 import seq  # noqa: E402
from ssr_speech_b200.config import cfg_tiny  # noqa: E402
from ssr_speech_b200.synth import make_lm_state_dict  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="needs the reference tree (build container only)")

SILENCE = [3, 17, 40]


@pytest.fixture(scope="module")
def pair():
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    cfg = cfg_tiny()
    sd = make_lm_state_dict(cfg, seed=7)
    ssr = ref_loader.load_reference_ssr()
    ref = ssr.SSR_Speech(cfg.to_namespace()).eval()
    ref.load_state_dict(sd, strict=True)
    return cfg, ref, LMOracle(cfg, sd)


def random_spans(rng, T, n):
    """n sorted, non-overlapping [a, b) intervals inside [0, T]; empty intervals and touching neighbours allowed."""
    cuts = sorted(rng.integers(0, T + 1, size=2 * n).tolist())
    return [[cuts[2 * i], cuts[2 * i + 1]] for i in range(n)]


def edge_spans(T):
    return [[[0, 0]], [[0, 3]], [[T - 2, T]], [[T, T]], [[0, T]], [[2, 5], [5, 9]], [[0, 1], [1, 2], [T - 1, T]],
            [[4, 4], [4, 4]], [[0, 0], [T, T]], [[3, 7], [7, 7], [7, 12]]]


def run_case(pair, seed, T, Lx, spans, kw, ctx=None):
    cfg, ref, oracle = pair
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, cfg.text_vocab_size, (1, Lx), generator=g)
    y = torch.randint(0, cfg.audio_vocab_size, (1, T, cfg.n_codebooks), generator=g)
    if ctx:
        px = torch.randint(0, cfg.text_vocab_size, (1, ctx[0]), generator=g)
        pr = torch.randint(0, cfg.audio_vocab_size, (1, ctx[1], cfg.n_codebooks), generator=g)
    else:
        px, pr = x, y
    mi = torch.tensor([spans], dtype=torch.long)
    torch.manual_seed(seed)
    with torch.no_grad():
        res, marks, masks, nmi = ref.inference(x, torch.tensor([Lx]), px, torch.tensor([px.shape[1]]), y, pr, mask_interval=mi,
                                               silence_tokens=SILENCE, **kw)
    use_ctx = bool(kw.get("aug_context")) and sum(b - a for a, b in spans) < 2 * 50
    xo = torch.cat([px[0], x[0]]) if use_ctx else x[0]
    yo = np.concatenate([pr[0].T.numpy(), y[0].T.numpy()], 1) if use_ctx else y[0].T.numpy().copy()
    prep = seq.prepare(cfg, yo, spans, out_len=pr.shape[1] if use_ctx else 0)
    okw = {k: v for k, v in kw.items() if k not in ("kvcache", "aug_context")}
    torch.manual_seed(seed)
    got = oracle.inference(xo, torch.from_numpy(prep.prompt_tokens), prep.num_spans, silence_tokens=SILENCE,
                           incremental=bool(kw["kvcache"]), **okw)
    # the iteration bound lm.SSR_Speech uses to size its last chunk of decode iterations covers the reference's roll-out
    assert sum(len(sp) for sp in got) <= prep.num_spans * seq.expected_steps(cfg, xo.shape[0], prep.prompt_tokens.shape[1])
    ores, omarks, omasks, onmi = seq.finalize(cfg, prep, got)
    tag = (seed, T, Lx, spans, kw, ctx)
    assert ores.shape == tuple(res[0].shape), tag
    assert np.array_equal(ores, res[0].numpy()), tag
    assert np.array_equal(omarks, marks[0].numpy()), tag
    assert omasks == [tuple(int(v) for v in m) for m in masks], tag
    assert onmi == [tuple(int(v) for v in m) for m in nmi], tag
    return int(omarks.sum())


GREEDY = dict(top_k=1, top_p=1.0, temperature=1.0, stop_repetition=-1, kvcache=1, cfg_coef=1.5, cfg_stride=1, aug_text=False)
SAMPLED_CFG = dict(top_k=0, top_p=0.9, temperature=1.0, stop_repetition=2, kvcache=1, cfg_coef=1.5, cfg_stride=2, aug_text=True)


def test_edge_span_geometries(pair):
    T, n_gen = 14, 0
    for i, spans in enumerate(edge_spans(T)):
        n_gen += run_case(pair, 100 + i, T, 3, spans, GREEDY)
        run_case(pair, 200 + i, T, 2, spans, SAMPLED_CFG)
    assert n_gen > 0                                     # the sweep does generate frames (not only forced EOGs)


def test_random_span_geometries(pair):
    rng = np.random.default_rng(5)
    for i in range(24):
        T = int(rng.integers(6, 26))
        n = int(rng.integers(1, 4))
        kw = dict(SAMPLED_CFG if i % 2 else GREEDY)
        kw["kvcache"] = int(i % 3 != 0)
        kw["cfg_stride"] = 1 + i % 3 if kw["aug_text"] else 1
        run_case(pair, 300 + i, T, int(rng.integers(2, 5)), random_spans(rng, T, n), kw)


def test_aug_context_geometries(pair):
    rng = np.random.default_rng(6)
    for i in range(8):
        T = int(rng.integers(8, 20))
        kw = dict(SAMPLED_CFG if i % 2 else GREEDY, aug_context=True)
        spans = random_spans(rng, T, 1 + i % 2)
        run_case(pair, 400 + i, T, 3, spans, kw, ctx=(int(rng.integers(1, 4)), int(rng.integers(2, 9))))
