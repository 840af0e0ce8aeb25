"""Host-side (integer) logic: sequence surgery, config parsing, sharding, padding/length formulas."""
import math
from argparse import Namespace

import numpy as np
import pytest
import torch

from ssr_speech_b200 import seq
from ssr_speech_b200.config import CodecConfig, SSRConfig, cfg_830m, cfg_tiny
from ssr_speech_b200.dist import shard_range


def test_config_from_reference_namespace():
    cfg = cfg_830m()
    ns = cfg.to_namespace()
    assert isinstance(ns.audio_vocab_size, str)           # e830M.sh passes a string that ssr.py evals
    back = SSRConfig.from_args(ns)
    assert back == cfg and back.n_audio_tokens == 2056 and back.n_text_tokens == 101 and back.ffn_dim == 8192


def test_config_asserts_like_reference():
    ns = cfg_830m().to_namespace()
    ns.eog = 7
    with pytest.raises(AssertionError):
        SSRConfig.from_args(ns)


@pytest.mark.parametrize("n", [0, 1, 5, 17])
def test_delay_pattern_roundtrip(n):
    cfg = cfg_tiny()
    rng = np.random.default_rng(n)
    t = rng.integers(0, cfg.audio_vocab_size, (4, n))
    d = seq.delay_pattern(t, cfg.empty_token)
    assert d.shape == (4, n + 3)
    for q in range(4):
        assert (d[q, :q] == cfg.empty_token).all() and (d[q, q + n:] == cfg.empty_token).all()
    assert np.array_equal(seq.revert_delay_pattern(d, cfg.empty_token), t)


def test_prepare_tts_layout():
    cfg = cfg_tiny()
    T = 12
    y = np.arange(4 * T).reshape(4, T) % cfg.audio_vocab_size
    p = seq.prepare(cfg, y, [[T, T]])
    # sos + T frames delayed (T+1+3), <mts0>, [eos] delayed (4)  -> Y0-1 = T + 9 (Y0 = T + 10)
    assert p.prompt_tokens.shape == (4, T + 9)
    assert (p.prompt_tokens[:, T + 4] == cfg.mts).all()
    assert p.prompt_tokens[0, 0] == cfg.sos and p.prompt_tokens[0, T + 5] == cfg.eos
    assert p.non_mask_intervals == [(0, T), (T, T)] and p.num_spans == 1


def test_prepare_edit_two_spans_and_finalize_shapes():
    cfg = cfg_tiny()
    T = 30
    rng = np.random.default_rng(0)
    y = rng.integers(0, cfg.audio_vocab_size, (4, T))
    p = seq.prepare(cfg, y, [[5, 9], [20, 28]])
    assert p.num_spans == 2 and p.non_mask_intervals == [(0, 5), (9, 20), (28, 30)]
    # generated spans: n tokens + eog column, delayed -> n + 4 iterations
    spans = []
    for n in (3, 6):
        g = rng.integers(0, cfg.audio_vocab_size, (4, n))
        g = np.concatenate([g, np.full((4, 1), cfg.eog)], 1)
        spans.append(seq.delay_pattern(g, cfg.empty_token).T)
    res, marks, masks, nmi = seq.finalize(cfg, p, spans)
    assert res.shape == (4, 5 + 3 + 11 + 6 + 2)
    assert marks.sum() == 9 and masks == [(0, 5), (8, 19), (25, 27)]
    assert np.array_equal(res[:, :5], y[:, :5]) and np.array_equal(res[:, -2:], y[:, 28:])


def test_expected_steps_matches_survey():
    cfg = cfg_830m()
    # T=150, Lx=60 -> 441 generated frames, 445 iterations
    assert seq.expected_steps(cfg, 60, 150 + 9) == 445
    assert seq.expected_steps(cfg, 40, 150 + 9) == 245


def test_shard_range_partition():
    for n in (0, 1, 7, 256):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


# length formulas the reference's own tests pin (audiocraft/tests/modules/test_conv.py:151-203)
@pytest.mark.parametrize("k,s", [(4, 1), (4, 2), (10, 5), (7, 1), (16, 8), (8, 4)])
@pytest.mark.parametrize("T", [10, 23, 320, 1001])
def test_conv_and_convtr_length_rules(k, s, T):
    from codec_oracle import CodecOracle
    sd = {"c.weight": torch.randn(3, 2, k), "c.bias": torch.zeros(3), "t.weight": torch.randn(2, 3, k), "t.bias": torch.zeros(3)}
    o = CodecOracle(CodecConfig(), sd)
    x = torch.randn(1, 2, T)
    y = o.conv(x, "c.", s)
    assert y.shape[-1] == math.ceil(T / s)
    if k == 2 * s:
        z = o.convtr(x, "t.", s)
        assert z.shape[-1] == T * s


def test_code_file_format(tmp_path):
    """data/encode.py:53-57: K lines of space-separated integers, no trailing newline."""
    from ssr_speech_b200.encode_dataset import write_code_rows
    fn = tmp_path / "seg.txt"
    write_code_rows([[1, 22, 333], [4, 5, 6]], str(fn))
    assert fn.read_text() == "1 22 333\n4 5 6"
    write_code_rows(np.asarray([[7, 8]]).tolist(), str(fn))
    assert fn.read_text() == "7 8"


def test_audio_plumbing_matches_reference_rules(tmp_path):
    """data/tokenizer.py:87-97,141-159: zero-pad to a multiple of 320, stereo -> mono by averaging, mono stays, 16 kHz is not resampled;
    the wav reader returns float32 [C, T] in [-1, 1) like torchaudio.load."""
    from scipy.io import wavfile
    from ssr_speech_b200.codec import convert_audio, load_wav, pad_to_multiple
    rng = np.random.default_rng(0)
    pcm = (rng.standard_normal((1000, 2)) * 8000).astype(np.int16)
    fn = str(tmp_path / "a.wav")
    wavfile.write(fn, 16000, pcm)
    wav, sr = load_wav(fn)
    assert sr == 16000 and wav.dtype == torch.float32 and tuple(wav.shape) == (2, 1000)
    assert torch.allclose(wav, torch.from_numpy(pcm.T.astype(np.float32) / 32768.0))
    part, _ = load_wav(fn, offset=100, num_frames=300)
    assert torch.equal(part, wav[:, 100:400])
    padded = pad_to_multiple(wav, 320)
    assert padded.shape[-1] == 1280 and torch.equal(padded[:, :1000], wav) and not padded[:, 1000:].any()
    assert pad_to_multiple(padded, 320) is padded                       # already a multiple: untouched
    mono = convert_audio(padded, 16000, 16000, 1)
    assert tuple(mono.shape) == (1, 1280) and torch.allclose(mono[0], padded.mean(0))
    assert torch.equal(convert_audio(mono, 16000, 16000, 1), mono)
    assert tuple(convert_audio(mono, 16000, 16000, 2).shape) == (2, 1280)
    with pytest.raises(AssertionError):
        convert_audio(torch.zeros(3, 640), 16000, 16000, 1)


def test_decode_chunk_schedule_stops_at_the_iteration_bound():
    """lm.SSR_Speech.inference_batch polls `done` every poll_every iterations; the last chunk is cut at the batch's iteration bound
    (bench batch: 505 iterations, the prefill's sample being the first) instead of overshooting to 513."""
    from ssr_speech_b200.lm import SSR_Speech
    it, chunks = 1, []
    while it < 505:
        n = SSR_Speech._next_chunk(16, 505, it)
        chunks.append(n)
        it += n
    assert it == 505 and chunks[:-1] == [16] * 31 and chunks[-1] == 8
    assert SSR_Speech._next_chunk(16, 505, 505) == 1 and SSR_Speech._next_chunk(16, 10, 400) == 1     # past the bound: keep stepping
    assert SSR_Speech._next_chunk(1, 505, 3) == 1


def _xp_cfg(**over):
    """A resolved cfg shaped like the one audiocraft stores in the checkpoint (config/model/encodec/default.yaml +
    encodec_large_nq4_s320.yaml), as plain dicts (OmegaConf is absent in this image; DictConfig answers `in` / [] alike)."""
    cfg = {"sample_rate": 16000, "channels": 1, "compression_model": "wmencodec",
           "encodec": {"autoencoder": "seanet", "quantizer": "rvq", "sample_rate": 16000, "channels": 1, "causal": False,
                       "renormalize": False},
           "seanet": {"dimension": 128, "channels": 1, "causal": False, "n_filters": 64, "n_residual_layers": 1,
                      "ratios": [8, 5, 4, 2], "activation": "ELU", "activation_params": {"alpha": 1.0}, "norm": "weight_norm",
                      "norm_params": {}, "kernel_size": 7, "residual_kernel_size": 3, "last_kernel_size": 7, "dilation_base": 2,
                      "pad_mode": "constant", "true_skip": True, "compress": 2, "lstm": 2, "disable_norm_outer_blocks": 0},
           "rvq": {"n_q": 4, "bins": 2048, "q_dropout": False}}
    for k, v in over.items():
        sec, _, key = k.partition("__")
        if key:
            if v is None:
                cfg[sec].pop(key)
            else:
                cfg[sec][key] = v
        elif v is None:
            cfg.pop(sec)
        else:
            cfg[sec] = v
    return cfg


def test_codec_checkpoint_cfg_is_read_strictly():
    """wmcompression.py:281-315: the checkpoint's 'xp.cfg' decides the codec geometry.  A good cfg is mirrored; a malformed or
    unsupported one raises (it used to fall back to the default geometry silently)."""
    from ssr_speech_b200.codec import _cfg_from_xp
    assert _cfg_from_xp(_xp_cfg()) == CodecConfig()
    c = _cfg_from_xp(_xp_cfg(seanet__ratios=[8, 5, 4, 4], rvq__bins=1024, seanet__n_filters=32))
    assert c.ratios == (8, 5, 4, 4) and c.bins == 1024 and c.n_filters == 32 and c.hop_length == 640
    c = _cfg_from_xp(Namespace(**{k: (Namespace(**v) if isinstance(v, dict) and k in ("seanet", "rvq", "encodec") else v)
                                  for k, v in _xp_cfg().items()}))          # attribute-style access works too
    assert c == CodecConfig()
    for bad in (dict(seanet=None), dict(rvq__bins=None), dict(seanet__ratios=None), dict(seanet__causal=True),
                dict(seanet__norm="none"), dict(seanet__n_residual_layers=3), dict(encodec__renormalize=True),
                dict(seanet__ratios="8-5-4-2"), dict(sample_rate=None)):
        with pytest.raises(ValueError):
            _cfg_from_xp(_xp_cfg(**bad))
    with pytest.raises(ValueError):
        _cfg_from_xp(None)


def test_audio_tokenizer_loads_a_reference_style_checkpoint(tmp_path):
    """AudioTokenizer(signature=path) (data/tokenizer.py:99-113 -> wmcompression.py:281-315): {'xp.cfg', 'best_state': {'model'}}
    is loaded, the cfg read, the state dict kept for the engine (no CUDA call until the first encode / decode)."""
    from ssr_speech_b200.codec import AudioTokenizer
    from ssr_speech_b200.synth import make_codec_state_dict
    ccfg = CodecConfig(n_filters=8, ratios=(4, 2), bins=32)
    sd = make_codec_state_dict(ccfg, seed=1)
    path = tmp_path / "wmencodec.th"
    torch.save({"xp.cfg": _xp_cfg(seanet__n_filters=8, seanet__ratios=[4, 2], rvq__bins=32), "best_state": {"model": sd},
                "version": "test", "exported": True}, path)
    tok = AudioTokenizer(signature=str(path), device="cuda:0")
    assert tok.sample_rate == 16000 and tok.channels == 1 and tok.device == torch.device("cuda:0")
    assert tok.codec.cfg == ccfg and tok.codec.cfg.hop_length == 8
    assert set(tok.codec.state_dict()) == set(sd)
    torch.save({"xp.cfg": _xp_cfg(seanet=None), "best_state": {"model": sd}}, path)
    with pytest.raises(ValueError):
        AudioTokenizer(signature=str(path), device="cuda:0")
    torch.save({"best_state": {"model": sd}}, path)
    with pytest.raises(AssertionError):
        AudioTokenizer(signature=str(path), device="cuda:0")


def test_silence_token_list_longer_than_the_engine_table_raises():
    """ssr.py:727 accepts any list; the engine's table holds 8 — a longer list must fail loudly, not be truncated."""
    from ssr_speech_b200.lm import SSR_Speech
    from ssr_speech_b200.synth import make_lm_state_dict
    cfg = cfg_tiny()
    m = SSR_Speech(cfg.to_namespace(), precision="fp32")
    m.load_state_dict(make_lm_state_dict(cfg, seed=7))
    m._device = torch.device("cuda", 0)
    m._ensure_engine = lambda *a, **k: None          # host-side argument handling only: no engine, no CUDA
    x = torch.zeros(5, dtype=torch.long)
    y = torch.zeros(10, 4, dtype=torch.long)
    with pytest.raises(ValueError, match="silence"):
        m.open_batch([x], [y], [[[10, 10]]], silence_tokens=list(range(9)))


def test_max_n_spans_beyond_the_engine_state_is_rejected():
    ns = cfg_tiny().to_namespace()
    ns.max_n_spans = 4
    with pytest.raises(AssertionError):
        SSRConfig.from_args(ns)


@pytest.mark.parametrize("spans", [[[30, 30]], [[0, 4]], [[5, 12], [20, 30]], [[0, 3], [10, 10], [28, 30]]])
def test_device_splice_of_the_watermark_input_equals_the_host_splice(spans):
    """pipeline.splice_original_device (slice copies into a zeroed buffer, used on tensors that are already in HBM) against
    pipeline.splice_original (inference_scale.py:66-78) on the masks / ori_masks seq.finalize produces for TTS, an edit at frame 0,
    two spans and three spans with an empty one.  Runs on CPU tensors: the function is device-agnostic."""
    from ssr_speech_b200 import pipeline
    cfg = cfg_tiny()
    rng = np.random.default_rng(len(spans))
    T = 30
    y = rng.integers(0, cfg.audio_vocab_size, size=(cfg.n_codebooks, T))
    prep = seq.prepare(cfg, y, spans)
    gen = [rng.integers(0, cfg.audio_vocab_size, size=(int(rng.integers(3, 9)) + cfg.n_codebooks, cfg.n_codebooks)) for _ in range(prep.num_spans)]
    res, marks, masks, nmi = seq.finalize(cfg, prep, gen)
    n_frames = res.shape[-1]
    wav = torch.from_numpy(rng.standard_normal((1, T * 320)).astype(np.float32))
    want = pipeline.splice_original(wav, n_frames, masks, nmi)
    got = torch.zeros(1, n_frames * 320)
    pipeline.splice_original_device(wav, got, masks, nmi)
    assert torch.equal(got, want)
    kept = int((np.asarray(marks).reshape(-1) == 0).sum())
    assert int((want != 0).sum()) <= kept * 320


def test_fast_elu_for_bf16_outputs_stays_inside_half_a_bf16_ulp():
    """csrc/common.cuh elu1_bf16: x > 0 ? x : (x > -1e-3 ? x + x^2/2 : exp(x) - 1 through MUFU.EX2).  Restated in float32 with the
    hardware approximation's worst-case relative error (2^-22 on exp2) injected, against expm1 in float64: the result must differ
    by far less than half a bf16 ulp (2^-9 relative) everywhere, which is what makes it safe in front of a bf16 store."""
    x = -np.logspace(-8, np.log10(30.0), 20000).astype(np.float32)
    want = np.expm1(x.astype(np.float64))
    series = (x + np.float32(0.5) * x * x).astype(np.float32)
    worst = 0.0
    for sign in (-1.0, 1.0):
        ex = (np.exp(x.astype(np.float64)) * (1.0 + sign * 2.0 ** -22)).astype(np.float32)
        got = np.where(x > np.float32(-1e-3), series, (ex - np.float32(1.0)).astype(np.float32)).astype(np.float64)
        worst = max(worst, float(np.max(np.abs(got - want) / np.abs(want))))
    assert worst < 2.0 ** -9 / 4, worst
    assert worst < 5e-4, worst


@pytest.mark.parametrize("predict_mask_token,predict_all", [(True, False), (False, False), (True, True), (False, True)])
def test_training_loss_flags_match_the_oracle_masks(predict_mask_token, predict_all):
    """seq.loss_flags (the host-side integer logic of SSR_Speech.forward, ssr.py:330-345, vectorised) against the oracle's
    loop restatement — which is pinned to the unmodified reference's loss by tests/test_oracle_golden.py — on random token
    grids with zero, one and several <mts> occurrences per codebook row, empty / pad / eog tokens mixed in."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from lm_oracle import LMOracle
    cfg = cfg_tiny()
    rng = np.random.default_rng(5)
    specials = [cfg.empty_token, cfg.eog, cfg.audio_pad_token, cfg.mts, cfg.mts + 1, cfg.mts + 2]
    for trial in range(40):
        T = int(rng.integers(2, 40))
        y = rng.integers(0, cfg.audio_vocab_size, size=(cfg.n_codebooks, T))
        n_special = int(rng.integers(0, max(1, T // 2)))
        for _ in range(n_special):
            y[int(rng.integers(0, cfg.n_codebooks)), int(rng.integers(0, T))] = specials[int(rng.integers(0, len(specials)))]
        flags = seq.loss_flags(cfg, y, predict_mask_token, predict_all)
        tg, mask, tmp = LMOracle.loss_masks(cfg, torch.from_numpy(y), predict_mask_token, predict_all)
        assert flags.dtype == np.uint8 and flags.shape == (cfg.n_codebooks, T - 1)
        assert np.array_equal(flags & 1, tmp.numpy().astype(np.uint8)), trial
        assert np.array_equal((flags >> 1) & 1, mask.numpy().astype(np.uint8)), trial


@pytest.mark.parametrize("tag", ["default", "all"])
def test_training_forward_wrapper_combines_like_the_reference(monkeypatch, tag):
    """SSR_Speech.forward end to end on the CPU with the device call replaced by the oracle: the wrapper's batch handling, loss
    flags, per-utterance accumulation and the final combination (loss = sum_k mean-CE_k * ntokens_k * weight_k, top-10 accuracy
    x ntokens; ssr.py:352-372) must reproduce the numbers recorded from the unmodified reference.  (The CUDA kernels behind
    ssrb_lm_forward_loss are checked against the same golden on the GPU: tests/test_gpu_zzz_train_forward.py.)"""
    import contextlib
    import ctypes
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from lm_oracle import LMOracle
    from ssr_speech_b200 import _lib, lm as lm_mod
    from ssr_speech_b200.synth import make_lm_state_dict
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lm_train_forward.npz"))
    cfg = cfg_tiny()
    oracle = LMOracle(cfg, make_lm_state_dict(cfg, seed=int(g["weights_seed"])))
    K, V = cfg.n_codebooks, cfg.n_audio_tokens

    class FakeLib:
        @staticmethod
        def ssrb_lm_forward_loss(h, text_p, Lx, audio_p, Ty, flags_p, out_p, stream):
            text = np.ctypeslib.as_array(ctypes.cast(text_p, ctypes.POINTER(ctypes.c_int32)), (Lx,))
            audio = np.ctypeslib.as_array(ctypes.cast(audio_p, ctypes.POINTER(ctypes.c_int32)), (K, Ty))
            flags = np.ctypeslib.as_array(ctypes.cast(flags_p, ctypes.POINTER(ctypes.c_uint8)), (K, Ty - 1))
            out = np.ctypeslib.as_array(ctypes.cast(out_p, ctypes.POINTER(ctypes.c_double)), (K, 4))
            logits = oracle.teacher_forced_logits(torch.from_numpy(text.astype(np.int64)), torch.from_numpy(audio.astype(np.int64)))[:-1].double()
            for k in range(K):
                sel = torch.from_numpy((flags[k] & 1).astype(bool))
                tgt = torch.from_numpy(audio[k, 1:].astype(np.int64))
                lk = logits[:, k][sel]
                out[k, 0] = float(torch.nn.functional.cross_entropy(lk, tgt[sel], reduction="sum")) if sel.any() else 0.0
                out[k, 1] = int(sel.sum())
                out[k, 2] = int((lk.topk(10, dim=-1).indices == tgt[sel][:, None]).any(-1).sum()) if sel.any() else 0
                out[k, 3] = int(((flags[k] >> 1) & 1).sum())
            return 0

    ns = cfg.to_namespace()
    ns.predict_mask_token, ns.predict_all = int(g[f"{tag}_predict_mask_token"]), int(g[f"{tag}_predict_all"])
    ns.codebook_weight = None if tag == "all" else str(g[f"{tag}_codebook_weight"].tolist())
    m = lm_mod.SSR_Speech(ns, precision="fp32")
    m._device = torch.device("cpu")
    m._h = ctypes.c_void_p(1)
    monkeypatch.setattr(m, "_ensure_engine", lambda *a, **k: None)
    monkeypatch.setattr(_lib, "load", lambda: FakeLib)
    monkeypatch.setattr(_lib, "stream_ptr", lambda: None)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    out = m.forward({"x": torch.from_numpy(g["x"]), "x_lens": torch.from_numpy(g["x_lens"]), "y": torch.from_numpy(g["y"]),
                     "y_lens": torch.from_numpy(g["y_lens"])})
    m._h = None
    assert int(out["effective_ntoken"]) == int(g[f"{tag}_ntoken"])
    assert abs(float(out["loss"]) - float(g[f"{tag}_loss"])) <= 1e-5 * abs(float(g[f"{tag}_loss"]))
    assert abs(float(out["top10acc"]) - float(g[f"{tag}_top10acc"])) <= 1e-3
    assert np.allclose([float(v) for v in out["top10acc_by_codebook"]], g[f"{tag}_top10acc_by_codebook"], atol=1e-3)
