"""End-to-end hot path through the reference-facing API (pipeline.inference_one_sample / inference_batch):
encode -> SSR_Speech.inference -> wmdecode|decode -> TTS trim, against the oracle fed with the same codes and noise."""
import numpy as np
import pytest
import torch

from codec_oracle import CodecOracle
from lm_oracle import LMOracle
from ssr_speech_b200 import pipeline, seq
from ssr_speech_b200.codec import AudioTokenizer, WMEncodecModel
from ssr_speech_b200.config import CodecConfig, cfg_tiny
from ssr_speech_b200.lm import SSR_Speech
from ssr_speech_b200.synth import calibrate_codebooks, make_codec_state_dict, make_lm_state_dict

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def stack():
    cfg = cfg_tiny(audio_vocab_size=2048)
    sd = make_lm_state_dict(cfg, seed=21)
    model = SSR_Speech(cfg.to_namespace(), precision="fp32")
    model.load_state_dict(sd)
    model.to("cuda").eval()
    ccfg = CodecConfig()
    codec = WMEncodecModel(ccfg)
    codec.load_state_dict(make_codec_state_dict(ccfg, seed=5))
    codec.to("cuda")
    cal = 0.1 * torch.randn(2, 1, 16000, generator=torch.Generator().manual_seed(9))
    _, _, emb = codec.encode(cal.cuda())
    mu, sg = calibrate_codebooks(ccfg, 5, emb.permute(0, 2, 1).reshape(-1, ccfg.dimension).cpu())
    csd = make_codec_state_dict(ccfg, seed=5, codebook_mu=mu, codebook_sigma=sg)
    codec.load_state_dict(csd)
    codec.to("cuda")
    return cfg, sd, model, ccfg, csd, AudioTokenizer(model=codec, device="cuda")


DC = {"top_k": 0, "top_p": 0.9, "temperature": 1.0, "stop_repetition": 2, "kvcache": 1, "codec_sr": 50}


@pytest.mark.parametrize("use_watermark,tts,span", [(True, True, None), (False, True, None), (True, False, (10, 22))])
def test_inference_one_sample_matches_oracle(stack, use_watermark, tts, span):
    cfg, sd, model, ccfg, csd, tok = stack
    g = torch.Generator().manual_seed(3)
    wav = 0.1 * torch.randn(1, 16000 - 37, generator=g)              # not a multiple of 320: exercises the padding
    text = torch.randint(0, cfg.text_vocab_size, (8,), generator=g)
    Tf = 50
    mi = torch.tensor([[Tf, Tf]] if span is None else [list(span)])
    noise = torch.empty(120, 4, cfg.n_audio_tokens).exponential_(1, generator=g)
    uncond = torch.randint(0, cfg.n_text_tokens, (8,), generator=g)
    # the mirror API draws uncond text from the global RNG and noise from Philox; inject both for an exact comparison
    orig = model.inference

    def patched(*a, **k):
        return orig(*a, _uncond_x=uncond, _noise=noise, **k)
    model.inference = patched
    try:
        out = pipeline.inference_one_sample(model, cfg.to_namespace(), None, None, tok, wav, text, text, mi, 1.5, 2, True, False,
                                            use_watermark, tts, "cuda", DC)
    finally:
        model.inference = orig
    # ---- oracle on the codes OUR encoder produced -------------------------------------------------------------
    wav_p = torch.nn.functional.pad(wav, (0, 37))
    codes, _, _ = tok.encode(wav_p[None])
    y = codes[0].cpu().numpy()
    prep = seq.prepare(cfg, y, mi.tolist())
    spans = LMOracle(cfg, sd).inference(text, torch.from_numpy(prep.prompt_tokens), prep.num_spans, top_k=0, top_p=0.9,
                                        stop_repetition=2, cfg_coef=1.5, cfg_stride=2, aug_text=True, uncond_x=uncond, noise=noise)
    res, marks, masks, nmi = seq.finalize(cfg, prep, spans)
    co = CodecOracle(ccfg, csd)
    rt = torch.from_numpy(res)[None]
    if use_watermark:
        new_wav = pipeline.splice_original(wav_p, res.shape[1], masks, nmi)
        want, _ = co.wmdecode(rt, torch.from_numpy(marks)[None], new_wav[None])
    else:
        want = co.decode(rt)
    if tts:
        want = want[:, :, masks[0][1] * 320:]
    assert tuple(out.shape) == tuple(want.shape)
    assert (out.cpu() - want).abs().max() <= 1e-4 * want.abs().max()


def test_inference_batch_host_buffers(stack):
    cfg, sd, model, ccfg, csd, tok = stack
    g = torch.Generator().manual_seed(4)
    wavs = [0.1 * torch.randn(1, 9600, generator=g) for _ in range(3)]
    texts = [torch.randint(0, cfg.text_vocab_size, (n,), generator=g) for n in (6, 9, 7)]
    spans = [[[30, 30]], [[30, 30]], [[5, 12]]]
    tm = {}
    torch.manual_seed(0)          # the uncond (CFG) phonemes come from the global CPU RNG, like the reference (ssr.py:574)
    outs, results = pipeline.inference_batch(model, tok, wavs, texts, spans, DC, cfg_coef=1.5, cfg_stride=2, aug_text=True,
                                             use_watermark=True, tts=False, seed=77, timings=tm)
    torch.manual_seed(0)
    again, _ = pipeline.inference_batch(model, tok, wavs, texts, spans, DC, cfg_coef=1.5, cfg_stride=2, aug_text=True,
                                        use_watermark=True, tts=False, seed=77)
    assert len(outs) == 3 and all(not o.is_cuda for o in outs)
    for o, a, (res, marks, masks, nmi) in zip(outs, again, results):
        assert o.shape == (1, res.shape[-1] * 320) and torch.equal(o, a)          # seed-deterministic end to end
        assert marks.shape[-1] == res.shape[-1]
    assert tm["gen_frames"] > 0 and tm["lm_ms"] > 0


def test_inference_batch_pinned_result_buffer_and_device_splice(stack):
    """host_out (a pinned buffer the caller reuses) must give the same waveforms as the default path, and the device-side
    splice of the watermark decoder's new_wav must equal the host one (inference_scale.py:66-78) on every span geometry."""
    cfg, sd, model, ccfg, csd, tok = stack
    g = torch.Generator().manual_seed(5)
    wavs = [0.1 * torch.randn(1, 9600, generator=g) for _ in range(3)]
    texts = [torch.randint(0, cfg.text_vocab_size, (n,), generator=g) for n in (6, 9, 7)]
    spans = [[[30, 30]], [[0, 4]], [[5, 12], [20, 30]]]
    torch.manual_seed(0)
    want, results = pipeline.inference_batch(model, tok, wavs, texts, spans, DC, cfg_coef=1.5, cfg_stride=2, aug_text=True,
                                             use_watermark=True, tts=False, seed=78)
    longest = max(r[0].shape[-1] for r in results) * 320
    host_out = torch.empty(3, 1, longest + 640).pin_memory()
    torch.manual_seed(0)
    got, _ = pipeline.inference_batch(model, tok, [w.pin_memory() for w in wavs], texts, spans, DC, cfg_coef=1.5, cfg_stride=2,
                                      aug_text=True, use_watermark=True, tts=False, seed=78, host_out=host_out)
    for a, b in zip(want, got):
        assert a.shape == b.shape and torch.equal(a, b)
        assert b.data_ptr() >= host_out.data_ptr() and b.data_ptr() < host_out.data_ptr() + host_out.numel() * 4
    for w, (res, marks, masks, nmi) in zip(wavs, results):
        n_frames = res.shape[-1]
        ref = pipeline.splice_original(w, n_frames, masks, nmi)
        dev = torch.zeros(1, n_frames * 320, device="cuda")
        pipeline.splice_original_device(w.cuda(), dev, masks, nmi)
        assert torch.equal(dev.cpu(), ref)
