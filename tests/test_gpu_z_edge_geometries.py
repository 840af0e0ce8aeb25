"""CUDA decode path (fp32 parity mode, through the C ABI) against the oracle on the span geometries of
tests/test_oracle_vs_reference_sweep.py — spans at frame 0, at the last frame, adjacent, empty, whole-utterance, 1-3 spans —
where the oracle itself is pinned against the unmodified reference in the build container.  Explicit Exp(1) noise makes the
sampled cases exact."""
import os

import numpy as np
import pytest
import torch

from lm_oracle import LMOracle
from ssr_speech_b200 import seq
from ssr_speech_b200.config import cfg_tiny
from ssr_speech_b200.synth import make_lm_state_dict
from test_gpu_lm import make_model
from test_oracle_vs_reference_sweep import GREEDY, SAMPLED_CFG, SILENCE, edge_spans, random_spans

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model_fp32():
    return make_model("fp32")


@pytest.fixture(scope="module")
def oracle():
    cfg = cfg_tiny()
    return LMOracle(cfg, make_lm_state_dict(cfg, seed=7))


def check(model, oracle, seed, T, Lx, spans, kw):
    cfg = cfg_tiny()
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, cfg.text_vocab_size, (1, Lx), generator=g)
    y = torch.randint(0, cfg.audio_vocab_size, (1, T, cfg.n_codebooks), generator=g)
    un = torch.randint(0, cfg.n_text_tokens, (Lx,), generator=g) if kw["aug_text"] else None
    noise = torch.empty(len(spans) * (10 * Lx + 8), cfg.n_codebooks, cfg.n_audio_tokens).exponential_(1, generator=g)
    prep = seq.prepare(cfg, y[0].T.numpy().copy(), spans)
    okw = {k: v for k, v in kw.items() if k != "kvcache"}
    want = oracle.inference(x[0], torch.from_numpy(prep.prompt_tokens), prep.num_spans, silence_tokens=SILENCE, uncond_x=un,
                            noise=noise, **okw)
    wres, wmarks, wmasks, wnmi = seq.finalize(cfg, prep, want)
    res, marks, masks, nmi = model.inference(x.cuda(), torch.tensor([Lx]), x.cuda(), torch.tensor([Lx]), y.cuda(), y.cuda(),
                                             mask_interval=torch.tensor([spans]), silence_tokens=SILENCE, _uncond_x=un,
                                             _noise=noise, **kw)
    tag = (seed, T, Lx, spans)
    assert np.array_equal(res[0].cpu().numpy(), wres), tag
    assert np.array_equal(marks[0].numpy(), wmarks) and list(masks) == wmasks and list(nmi) == wnmi, tag


def test_edge_span_geometries(model_fp32, oracle):
    T = 14
    for i, spans in enumerate(edge_spans(T)):
        check(model_fp32, oracle, 100 + i, T, 3, spans, GREEDY)
        check(model_fp32, oracle, 200 + i, T, 2, spans, SAMPLED_CFG)


def test_random_span_geometries(model_fp32, oracle):
    rng = np.random.default_rng(5)
    for i in range(16):
        T = int(rng.integers(6, 26))
        kw = dict(SAMPLED_CFG if i % 2 else GREEDY)
        kw["cfg_stride"] = 1 + i % 3 if kw["aug_text"] else 1
        check(model_fp32, oracle, 300 + i, T, int(rng.integers(2, 5)), random_spans(rng, T, int(rng.integers(1, 4))), kw)
