"""The EXPERIMENTAL persistent whole-iteration decode kernel for small batches (csrc/lm_mega.cu, SSRB_MEGA=1; R <= 16 transformer
rows = batch 1 and 8 of BASELINE's metric with their CFG rows) against the oracle.

Roll-outs run on the production path (CUDA graph: embedding, ONE persistent kernel for 16 layers + heads, sampler); at
checkpoints along the roll-out the raw head outputs are read back and compared with the bf16-storage oracle teacher-forced on
the GPU's OWN tokens — exactly the logits models/ssr.py:673-689 computes at that iteration.  Tolerance 2e-2 (SURVEY §8c, bf16
production mode).  Covered: R = 1 (no CFG), 2, 10 (ragged, utterances finishing at different iterations), 16; K-chunks of 512 /
1024 / 2 x 1024 / 8 x 1024; the 830M model at batch 1 and batch 8."""
import ctypes as C

import numpy as np
import pytest
import torch

from lm_oracle import LMOracle
from ssr_speech_b200 import _lib, seq
from ssr_speech_b200.config import cfg_830m, cfg_tiny
from ssr_speech_b200.lm import SSR_Speech
from ssr_speech_b200.synth import make_lm_state_dict

pytestmark = pytest.mark.gpu
TOL = 2e-2


@pytest.fixture(autouse=True)
def _select_persistent_kernel(monkeypatch):
    """The kernel is experimental and off by default (it does not beat the per-GEMM chain yet: profiles/r02b_mega_kernel.md);
    the engine reads the switch whenever it is created."""
    monkeypatch.setenv("SSRB_MEGA", "1")


def last_logits(oracle, x, fed):
    xi = oracle.embed_text(x)
    yi = oracle.embed_audio_tokens(fed) + oracle.alpha_a * oracle._pe(fed.shape[1])
    h, _ = oracle.stack(torch.cat([xi, yi], 0), None)
    return oracle.heads(h[-1]).numpy()


def rollout_check(cfg, sd, lx, tt, aug_text, checkpoints, watch, seed=0):
    K = cfg.n_codebooks
    U = len(lx)
    g = torch.Generator().manual_seed(100 + seed)
    xs = [torch.randint(0, 100, (n,), generator=g) for n in lx]
    ys = [torch.randint(0, cfg.audio_vocab_size, (t, K), generator=g) for t in tt]
    un = [torch.randint(0, 101, (n,), generator=g) for n in lx]
    m = SSR_Speech(cfg.to_namespace(), precision="bf16")
    m.load_state_dict(sd)
    m.to("cuda").eval()
    oracle = LMOracle(cfg, sd, round_weights_to_bf16=True, round_acts_to_bf16=True)
    lib = _lib.load()
    ob = m.open_batch(xs, ys, [[[t, t]] for t in tt], top_k=0, top_p=0.8, temperature=1.0, stop_repetition=2, cfg_coef=1.5,
                      cfg_stride=3, aug_text=aug_text, uncond_xs=un if aug_text else None, seed=7)
    assert m.decode_path() == 3, "the batch must run through the persistent small-batch kernel"
    rpu = 2 if aug_text else 1
    preps = ob["preps"]
    worst, n_cmp = 0.0, 0
    with torch.cuda.device(m._device):
        st = _lib.stream_ptr()
        it, nd = C.c_int(0), C.c_int(0)
        flags = np.zeros(U, dtype=np.int32)
        _lib.check(lib.ssrb_lm_poll(m._h, st, C.byref(nd), C.byref(it)), "poll")
        for cp in checkpoints:
            if cp > it.value:
                _lib.check(lib.ssrb_lm_decode(m._h, cp - it.value, st), "decode")
            _lib.check(lib.ssrb_lm_poll_flags(m._h, st, C.c_void_p(flags.ctypes.data), C.byref(it)), "poll_flags")
            assert it.value == cp
            raw = m.last_raw_logits().numpy()
            for u in watch:
                buf = np.zeros((m._cap[3], K), dtype=np.int32)
                sl, nt = (C.c_int32 * _lib.MAX_SPANS)(), C.c_int(0)
                _lib.check(lib.ssrb_lm_read_tokens(m._h, st, u, C.c_void_p(buf.ctypes.data), buf.shape[0], C.byref(nt), sl), "read")
                if flags[u] and nt.value < cp:
                    continue                                 # finished earlier: its rows are no longer computed
                assert nt.value == cp
                fed = np.concatenate([preps[u].prompt_tokens, np.full((K, 1), cfg.mts), buf[:cp - 1].T.astype(np.int64)], 1)
                for j in range(rpu):
                    want = last_logits(oracle, xs[u] if j == 0 else un[u], torch.from_numpy(fed))
                    err = float(np.abs(raw[rpu * u + j] - want).max())
                    worst = max(worst, err)
                    assert np.isfinite(raw[rpu * u + j]).all()
                    assert err <= TOL, (cp, u, j, err)
                    n_cmp += 1
        # drain: every utterance ends where the reference's length guard puts it (ssr.py:739)
        n_max = max(seq.expected_steps(cfg, lx[u], preps[u].prompt_tokens.shape[1]) for u in range(U))
        if n_max > it.value:
            _lib.check(lib.ssrb_lm_decode(m._h, n_max - it.value, st), "decode")
        _lib.check(lib.ssrb_lm_poll(m._h, st, C.byref(nd), C.byref(it)), "poll")
        assert nd.value == U
    return worst, n_cmp


@pytest.fixture(scope="module")
def tiny512():
    cfg = cfg_tiny(d_model=512, nhead=4, num_layers=3, audio_vocab_size=2048)      # K chunks: 512 (d_model), 2 x 1024 (FFN2), 1024 (heads)
    return cfg, make_lm_state_dict(cfg, seed=31, pin_eog_bias=True)


@pytest.mark.parametrize("lx,tt,aug", [
    ([9], [40], False),                                              # R = 1
    ([9], [40], True),                                               # R = 2
    ([9, 12, 7, 10, 11], [40, 70, 33, 90, 64], True),                # R = 10, ragged: utterances finish between iteration 33 and 58
    ([8, 9, 10, 11, 12, 13, 14, 15], [30, 40, 50, 60, 70, 80, 90, 100], True),    # R = 16
], ids=["R1", "R2", "R10-ragged", "R16"])
def test_small_batch_kernel_matches_oracle_tiny(tiny512, lx, tt, aug):
    cfg, sd = tiny512
    U = len(lx)
    ends = [seq.expected_steps(cfg, lx[u], tt[u] + 10 - 1) for u in range(U)]
    cps = sorted(set([1, 2, 3, 17] + [e for e in ends if e > 3] + [min(ends) // 2]))
    worst, n = rollout_check(cfg, sd, lx, tt, aug, cps, list(range(U)))
    assert n >= 4 * U


def test_small_batch_kernel_matches_oracle_tiny_d256():
    """cfg_tiny itself (d_model 256: one 256-k chunk per feature group, 72 audio classes: most CTAs own no head-2 feature)."""
    cfg = cfg_tiny()
    sd = make_lm_state_dict(cfg, seed=7, pin_eog_bias=True)
    worst, n = rollout_check(cfg, sd, [6, 7, 5], [20, 31, 12], True, [1, 2, 9, 25], [0, 1, 2])
    assert n >= 12


@pytest.mark.parametrize("U", [1, 8], ids=["batch1", "batch8"])
def test_small_batch_kernel_matches_oracle_830m(U):
    """BASELINE's batch 1 and batch 8 points (830M, CFG rows): 3 s prompt, checkpoints up to iteration 120 (S up to ~330)."""
    cfg = cfg_830m()
    sd = make_lm_state_dict(cfg, seed=0, pin_eog_bias=True)
    lx = [41 - (i % 3) for i in range(U)]
    tt = [150 - 10 * (i % 2) for i in range(U)]
    watch = [0] if U == 1 else [0, 7]
    worst, n = rollout_check(cfg, sd, lx, tt, True, [1, 2, 60, 120], watch)
    assert n >= 8 * len(watch)
