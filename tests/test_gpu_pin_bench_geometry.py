"""The bf16 PRODUCTION path pinned to the oracle at the geometry bench.py measures (BASELINE configs[2]: 830M, 32 utterances
x cond/uncond rows, 10 s prompt -> 10 s generation, S 611 -> 1115; codec 500-frame encode, 1001-frame wmdecode).

  (a) teacher-forced logits of one 1115-position sequence (Lx = 101, Ty = 1014) through the prefill kernels
      (gemm_flat / attn_prefill_mma) against LMOracle with bf16-rounded weights / operands / KV: max-abs <= 2e-2;
  (b) a ragged R = 64 CFG roll-out on the decode chain (gemm_dec / attn_decode_tma under the CUDA graph): raw head outputs
      read back at checkpoints along the roll-out (first, every ~100th, last iteration) and compared with the oracle
      teacher-forced on the GPU's OWN tokens (models/ssr.py:673-689 computes exactly these logits each iteration):
      max-abs <= 2e-2, while other utterances of the batch have already finished (their rows carry no tiles);
  (c) codec: 32 x 10 s fp32 encode (latents <= 1e-4 * max, RVQ indices identical except near-ties) and 32 x 1001-frame bf16
      decode / wmdecode (<= 2e-2 * max, correlation > 0.9995) against CodecOracle on sampled utterances of the batch
      (wmencodec.py:324-375; the LSTM recurrence runs 500 / 1001 sequential steps here, not the <= 100 of the fixtures).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from codec_oracle import CodecOracle
from lm_oracle import LMOracle
from ssr_speech_b200 import _lib, seq
from ssr_speech_b200.codec import WMEncodecModel
from ssr_speech_b200.config import CodecConfig, cfg_830m
from ssr_speech_b200.lm import SSR_Speech
from ssr_speech_b200.synth import make_codec_state_dict, make_lm_state_dict
from test_gpu_codec import near_tie_only

pytestmark = pytest.mark.gpu
TOL = 2e-2


@pytest.fixture(scope="module")
def sd830():
    return make_lm_state_dict(cfg_830m(), seed=0, pin_eog_bias=True)


@pytest.fixture(scope="module")
def oracle830(sd830):
    torch.set_num_threads(max(1, torch.get_num_threads()))
    return LMOracle(cfg_830m(), sd830, round_weights_to_bf16=True, round_acts_to_bf16=True)


def new_model(sd):
    m = SSR_Speech(cfg_830m().to_namespace(), precision="bf16")
    m.load_state_dict(sd)
    return m.to("cuda").eval()


def test_teacher_forced_logits_1115_positions(sd830, oracle830):
    g = torch.Generator().manual_seed(31)
    Lx, Ty = 101, 1014
    x = torch.randint(0, 100, (Lx,), generator=g)
    toks = torch.randint(0, 2048, (4, Ty), generator=g)
    m = new_model(sd830)
    got = m.teacher_forced_logits(x, toks).numpy()
    del m
    want = oracle830.teacher_forced_logits(x, toks).numpy()
    err = np.abs(got - want).max(axis=(1, 2))            # per audio position
    assert err.max() <= TOL, (float(err.max()), int(err.argmax()))
    # no drift with position: the last 64-key tiles are as good as the first
    assert err[-128:].max() <= TOL and err[:128].max() <= TOL


def oracle_last_logits(oracle, x, fed):
    """Raw head outputs at the LAST audio position of [x ; fed] — what one loop iteration computes (ssr.py:673-689)."""
    xi = oracle.embed_text(x)
    yi = oracle.embed_audio_tokens(fed) + oracle.alpha_a * oracle._pe(fed.shape[1])
    h, _ = oracle.stack(torch.cat([xi, yi], 0), None)
    return oracle.heads(h[-1]).numpy()


def test_ragged_cfg_rollout_logits_along_the_decode_chain(sd830, oracle830):
    cfg = cfg_830m()
    K = cfg.n_codebooks
    g = torch.Generator().manual_seed(32)
    U = 32
    # ragged: text 101/94/87/80 phonemes, prompts 500/460/420 frames -> utterances finish between iteration ~300 and 505
    lx = [101 - 7 * (i % 4) for i in range(U)]
    tt = [500 - 40 * (i % 3) for i in range(U)]
    xs = [torch.randint(0, 100, (n,), generator=g) for n in lx]
    ys = [torch.randint(0, 2048, (t, K), generator=g) for t in tt]
    un = [torch.randint(0, 101, (n,), generator=g) for n in lx]
    mis = [[[t, t]] for t in tt]
    m = new_model(sd830)
    lib = _lib.load()
    ob = m.open_batch(xs, ys, mis, top_k=0, top_p=0.8, temperature=1.0, stop_repetition=2, cfg_coef=1.5, cfg_stride=5,
                      aug_text=True, uncond_xs=un, seed=9)
    preps = ob["preps"]
    watch = [0, 5, 31]                                   # (Lx, T) = (101, 500), (94, 420), (80, 460)
    last = {u: seq.expected_steps(cfg, lx[u], tt[u] + 10 - 1) for u in watch}      # 505, 515, 335 (length guard, ssr.py:739)
    assert sorted(last.values()) == [335, 505, 515]
    checkpoints = sorted(set([1, 2, 300, 400] + list(last.values())))
    n_cmp, n_cmp_with_finished_peers = 0, 0
    worst = 0.0
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
            raw = m.last_raw_logits().numpy()            # [R, K, V] of iteration `cp`
            for u in watch:
                buf = np.zeros((m._cap[3], K), dtype=np.int32)
                sl = (C.c_int32 * _lib.MAX_SPANS)()
                nt = C.c_int(0)
                _lib.check(lib.ssrb_lm_read_tokens(m._h, st, u, C.c_void_p(buf.ctypes.data), buf.shape[0], C.byref(nt), sl), "read")
                if flags[u] and nt.value < cp:
                    continue                             # finished before this checkpoint: its logits rows are stale by design
                assert nt.value == cp
                fed = np.concatenate([preps[u].prompt_tokens, np.full((K, 1), cfg.mts), buf[:cp - 1].T.astype(np.int64)], 1)
                for j, xr in enumerate((xs[u], un[u])):
                    want = oracle_last_logits(oracle830, xr, torch.from_numpy(fed))
                    err = float(np.abs(raw[2 * u + j] - want).max())
                    worst = max(worst, err)
                    assert err <= TOL, (cp, u, j, err)
                    n_cmp += 1
                    n_cmp_with_finished_peers += int(flags.sum() > 0)
    assert n_cmp >= 30 and n_cmp_with_finished_peers >= 6, (n_cmp, n_cmp_with_finished_peers)
    # the whole batch ends where the reference's length guard puts it (ssr.py:739): 10 * Lx - Y0 + 1 new frames + EOG drain
    n_max = max(seq.expected_steps(cfg, lx[u], tt[u] + 10 - 1) for u in range(U))     # 585: Lx 101 with the 420-frame prompt
    _lib.check(lib.ssrb_lm_decode(m._h, n_max - it.value, st), "decode")
    _lib.check(lib.ssrb_lm_poll(m._h, st, C.byref(nd), C.byref(it)), "poll")
    assert nd.value == U and it.value == n_max
    for u in range(U):
        res = m._collect(u, preps[u])[0]
        assert res.shape[-1] == tt[u] + 10 * lx[u] - (tt[u] + 10) + 1


# ---- codec at the bench geometry ------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def codec_pair(gold_dir):
    import os
    g = np.load(os.path.join(gold_dir, "codec_small.npz"))
    cfg = CodecConfig()
    sd = make_codec_state_dict(cfg, seed=int(g["weights_seed"]), codebook_mu=g["codebook_mu"], codebook_sigma=g["codebook_sigma"])
    return cfg, sd, CodecOracle(cfg, sd)


def test_encode_32x10s_fp32_vs_oracle(codec_pair):
    cfg, sd, o = codec_pair
    m = WMEncodecModel(cfg, max_batch_chunk=32, precision="bf16")     # the bench's model; encode() stays fp32 by design
    m.load_state_dict(sd)
    m.to("cuda")
    B, T = 32, 160000
    wav = torch.stack([0.1 * torch.randn(1, T, generator=torch.Generator().manual_seed(1234 + i)) for i in range(B)])
    codes, scale, emb = m.encode(wav.cuda())
    assert tuple(codes.shape) == (B, 4, 500) and scale is None
    for i in (0, 17, 31):                                # first / middle / last chunk of the batch
        oc, _, oe = o.encode(wav[i:i + 1])
        assert float((emb[i:i + 1].cpu() - oe).abs().max()) <= 1e-4 * float(oe.abs().max())
        got = codes[i:i + 1].cpu().numpy()
        assert (got == oc.numpy()).mean() >= 0.98
        assert near_tie_only(o, oe.numpy(), got, oc.numpy())
        assert torch.equal(m.quantize(oe.cuda()).cpu(), oc)


def test_decode_and_wmdecode_32x1001_frames_bf16_vs_oracle(codec_pair):
    cfg, sd, o = codec_pair
    m = WMEncodecModel(cfg, max_batch_chunk=32, precision="bf16")
    m.load_state_dict(sd)
    m.to("cuda")
    B, Tf = 32, 1001
    gen = torch.Generator().manual_seed(4)
    codes = torch.randint(0, cfg.bins, (B, 4, Tf), generator=gen)
    marks = torch.zeros(B, Tf, dtype=torch.long)
    marks[:, 500:] = 1                                    # TTS: the generated half carries the watermark bit
    wav = 0.1 * torch.randn(B, 1, Tf * 320, generator=gen)
    wav[:, :, 500 * 320:] = 0                             # inference_scale.py:66-78: zeros where frames were generated
    dec = m.decode(codes.cuda()).cpu()
    wm, _ = m.wmdecode(codes.cuda(), marks.cuda(), wav.cuda(), return_marks=False)
    wm = wm.cpu()
    for i in (0, 31):
        od = o.decode(codes[i:i + 1])
        ow, _ = o.wmdecode(codes[i:i + 1], marks[i:i + 1], wav[i:i + 1])
        for name, got, want in (("decode", dec[i:i + 1], od), ("wmdecode", wm[i:i + 1], ow)):
            err = float((got - want).abs().max())
            assert err <= TOL * float(want.abs().max()), (name, i, err, float(want.abs().max()))
            # no drift along the 1001 recurrent steps: the last second is as good as the first
            tail = float((got[..., -16000:] - want[..., -16000:]).abs().max())
            assert tail <= TOL * float(want.abs().max()), (name, i, "tail", tail)
            assert np.corrcoef(got.numpy().ravel(), want.numpy().ravel())[0, 1] > 0.9995, (name, i)
