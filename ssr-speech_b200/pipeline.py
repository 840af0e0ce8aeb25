"""`inference_one_sample` (reference inference_scale.py:17-88) and its batched extension.

Hot path top to bottom: phoneme ids -> tokenize_audio (WM-Encodec encode) -> SSR_Speech.inference ->
AudioTokenizer.wmdecode | decode -> TTS trim.  Phonemisation (espeak) is untouched reference territory:
pass `text_tokenizer=None` together with pre-computed phoneme id tensors to skip it.
"""
from __future__ import annotations

import logging
import time
from typing import Dict, List, Optional, Sequence

import torch

from .codec import AudioTokenizer, load_wav, pad_to_multiple, tokenize_audio


def _phonemes(text_tokenizer, phn2num, text):
    if torch.is_tensor(text):                      # already phoneme ids
        return text.to(torch.int64).reshape(1, -1)
    phones = text_tokenizer([text.strip()])[0]     # data/tokenizer.py:82-85 tokenize_text
    return torch.LongTensor([phn2num[p] for p in phones if p in phn2num]).unsqueeze(0)


@torch.no_grad()
def inference_one_sample(model, model_args, phn2num, text_tokenizer, audio_tokenizer, audio_fn, prompt_text, target_text,
                         mask_interval, cfg_coef, cfg_stride, aug_text, aug_context, use_watermark, tts, device,
                         decode_config):
    """Same argument list and return value as the reference.  `audio_fn` may be a path or a float tensor
    [C, T]; `prompt_text` / `target_text` may be strings (phonemised with `text_tokenizer`) or id tensors."""
    text_tokens = _phonemes(text_tokenizer, phn2num, target_text)
    text_tokens_lens = torch.LongTensor([text_tokens.shape[-1]])
    prompt_text_tokens = _phonemes(text_tokenizer, phn2num, prompt_text)
    prompt_text_tokens_lens = torch.LongTensor([prompt_text_tokens.shape[-1]])

    encoded_frames, scale, emb = tokenize_audio(audio_tokenizer, audio_fn)
    original_audio = encoded_frames.transpose(2, 1)                                  # [1,T,K]
    K = model_args.n_codebooks
    assert original_audio.ndim == 3 and original_audio.shape[0] == 1 and original_audio.shape[2] == K, original_audio.shape
    logging.info(f"original audio length: {original_audio.shape[1]} codec frames, "
                 f"which is {original_audio.shape[1] / decode_config['codec_sr']:.2f} sec.")
    stime = time.time()
    encoded_frames, marks, masks, ori_masks = model.inference(
        text_tokens.to(device), text_tokens_lens.to(device), prompt_text_tokens.to(device), prompt_text_tokens_lens.to(device),
        original_audio[..., :K].to(device), original_audio[..., :K].to(device),
        mask_interval=mask_interval.unsqueeze(0).to(device), top_k=decode_config["top_k"], top_p=decode_config["top_p"],
        temperature=decode_config["temperature"], stop_repetition=decode_config["stop_repetition"],
        kvcache=decode_config["kvcache"], cfg_coef=cfg_coef, cfg_stride=cfg_stride, aug_text=aug_text,
        **({"aug_context": aug_context} if aug_context else {}))
    logging.info(f"inference on one sample take: {time.time() - stime:.4f} sec.")
    if use_watermark:
        if torch.is_tensor(audio_fn):
            wav = audio_fn
        else:
            wav, _ = load_wav(audio_fn)
        wav = pad_to_multiple(wav, 320)
        new_wav = splice_original(wav, encoded_frames.shape[-1], masks, ori_masks)
        generated = audio_tokenizer.wmdecode(encoded_frames, marks.to(encoded_frames.device),
                                             new_wav.unsqueeze(0).to(encoded_frames.device), scale)
    else:
        generated = audio_tokenizer.decode(encoded_frames, scale)
    if tts:
        generated = generated[:, :, masks[0][1] * 320:]
    return generated


def splice_original(wav: torch.Tensor, n_frames: int, masks, ori_masks, hop: int = 320) -> torch.Tensor:
    """inference_scale.py:66-78: original samples copied into the kept intervals, zeros in generated ones."""
    new_wav = torch.zeros(1, n_frames * hop)
    ori = [(max(a, 0), b) for a, b in ori_masks]
    new = [(max(a, 0), b) for a, b in masks]
    for (na, nb), (oa, ob) in zip(new, ori):
        new_wav[:, na * hop:nb * hop] = wav[:1, oa * hop:ob * hop]
    return new_wav


def splice_original_device(wav_dev: torch.Tensor, out: torch.Tensor, masks, ori_masks, hop: int = 320) -> None:
    """splice_original on tensors that are already in HBM: wav_dev [1, T] is the utterance's original waveform (the encode input),
    out [1, n_frames * hop] is zero-initialised; the kept intervals are device-to-device slice copies."""
    for (na, nb), (oa, ob) in zip(masks, ori_masks):
        na, oa = max(na, 0), max(oa, 0)
        if nb > na:
            out[:, na * hop:nb * hop] = wav_dev[:1, oa * hop:ob * hop]


@torch.no_grad()
def inference_batch(model, audio_tokenizer: AudioTokenizer, wavs: List[torch.Tensor], text_ids: List[torch.Tensor],
                    mask_intervals: List, decode_config: Dict, cfg_coef: float = 1.5, cfg_stride: int = 5,
                    aug_text: bool = True, use_watermark: bool = True, tts: bool = True, seed: Optional[int] = None,
                    timings: Optional[Dict[str, float]] = None, to_host: bool = True,
                    host_out: Optional[torch.Tensor] = None):
    """Batched hot path with HOST inputs/outputs (this is what bench.py's `e2e` measures).

    wavs[i]: float32 [1, T_i] host tensors at 16 kHz (T_i multiple of 320, equal within the batch; pinned tensors are copied
    asynchronously); text_ids[i]: int64 [Lx_i]; mask_intervals[i]: [M_i, 2] frames.  Returns a list of host waveforms [1, T_out_i].
    The original waveforms cross the bus ONCE: the watermark decoder's `new_wav` (inference_scale.py:66-78) is spliced from the
    device copy the encoder already used.  `host_out` (optional, pinned, [U, 1, >= longest output]): the results are copied into it
    asynchronously and the returned tensors are views of it — a serving loop reuses one such buffer instead of paying a pageable
    allocation + staged copy per batch.  `to_host=False` leaves the results in HBM (multi-GPU jobs gather on the device first
    and copy to the host once, dist.gather_waveforms)."""
    dev = audio_tokenizer.device
    U = len(wavs)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    T = wavs[0].shape[-1]
    assert all(w.shape[-1] == T for w in wavs), "equal-length prompts per batch (pad on the caller side)"
    ev[0].record()
    wav_dev = torch.empty(U, 1, T, dtype=torch.float32, device=dev)
    for i, w in enumerate(wavs):
        wav_dev[i].copy_(w.reshape(1, T).to(torch.float32), non_blocking=True)
    codes, scale, _ = audio_tokenizer.encode(wav_dev)                                 # [U,K,Tf]
    ev[1].record()
    codes_h = codes.to("cpu")                                                         # ONE read-back: the prologue is host-side integer work
    ys = [codes_h[i].transpose(0, 1) for i in range(U)]                               # [Tf,K]
    results = model.inference_batch(text_ids, ys, mask_intervals, top_k=decode_config["top_k"], top_p=decode_config["top_p"],
                                    temperature=decode_config["temperature"], stop_repetition=decode_config["stop_repetition"],
                                    silence_tokens=decode_config.get("silence_tokens", (1388, 1898, 131)), cfg_coef=cfg_coef,
                                    cfg_stride=cfg_stride, aug_text=aug_text, seed=seed, device=dev)
    ev[2].record()
    outs = []
    # group utterances by output length so the codec runs batched
    by_len: Dict[int, List[int]] = {}
    for i, (res, marks, masks, nmi) in enumerate(results):
        by_len.setdefault(res.shape[-1], []).append(i)
    out_host: List[Optional[torch.Tensor]] = [None] * U
    for n_frames, idxs in by_len.items():
        fr = torch.cat([results[i][0] for i in idxs], 0).to(dev)
        if use_watermark:
            mk = torch.cat([results[i][1] for i in idxs], 0).to(dev, non_blocking=True)
            nw = torch.zeros(len(idxs), 1, n_frames * 320, dtype=torch.float32, device=dev)
            for j, i in enumerate(idxs):
                splice_original_device(wav_dev[i], nw[j], results[i][2], results[i][3])
            gen = audio_tokenizer.wmdecode(fr, mk, nw, scale)
        else:
            gen = audio_tokenizer.decode(fr, scale)
        if not to_host:
            gen_h = gen
        elif host_out is not None:
            assert host_out.is_pinned() and host_out.shape[0] >= U and host_out.shape[-1] >= gen.shape[-1], "host_out: pinned [U, 1, >= T_out]"
            for j, i in enumerate(idxs):
                host_out[i, :, :gen.shape[-1]].copy_(gen[j], non_blocking=True)       # completes before the synchronize below
            gen_h = None
        else:
            gen_h = gen.to("cpu")
        for j, i in enumerate(idxs):
            g = gen_h[j] if gen_h is not None else host_out[i, :, :gen.shape[-1]]
            if tts:
                g = g[:, results[i][2][0][1] * 320:]
            out_host[i] = g
    ev[3].record()
    torch.cuda.synchronize()
    if timings is not None:
        timings["encode_ms"] = ev[0].elapsed_time(ev[1])
        timings["lm_ms"] = ev[1].elapsed_time(ev[2])
        timings["decode_ms"] = ev[2].elapsed_time(ev[3])
        timings["gen_frames"] = float(sum(int(r[1].sum()) for r in results))
        timings.update({f"lm_{k}": v for k, v in getattr(model, "last_stats", {}).items()})
    return out_host, results
