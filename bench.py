#!/usr/bin/env python
"""bench.py — SSR-Speech hot path on B200: codec-tokens/s (whole job) and RTF.

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's CPU implementation

Workload (BASELINE.json configs[2], the configuration the metric is quoted on; configs[4] = the same per GPU at N=8):
English TTS, random-init 830M SSR-Speech, cfg_coef 1.5 / cfg_stride 5 / aug_text, top_p 0.8, 10 s synthetic
prompt -> 10 s generation, batch 32 per GPU (weak scaling).  One *step* = one pass of the whole hot path over one batch:
WM-Encodec encode -> prefill + 505 decode iterations -> WM-Encodec wmdecode.  Generation length is fixed by the
reference's own length guard (ssr.py:739) with the EOG bias pinned in the synthetic checkpoint.

value  = codec tokens of all ranks / device time of the step with inputs resident in HBM
e2e    = the same through pipeline.inference_batch with HOST buffers (H2D of waveforms/text, D2H of waveforms inside
         the timed region; at N>1 it also contains the single NCCL all_gather of finished waveforms)
roofline = decode iteration (the dominant cost): algorithmic bytes (weights once + KV of every active row)
         / CUDA-event time of the decode loop inside the timed region, against MEASURED_PEAKS.json hbm_gbs
cpu_baseline / --impl reference = the reference's algorithm on the host cores (the unmodified reference when
         /root/reference exists, else the oracle port), on a bounded sample extrapolated to the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

K_CODEBOOKS = 4
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="utterances per GPU")
    ap.add_argument("--prompt-sec", type=float, default=10.0)
    ap.add_argument("--lx", type=int, default=101, help="phonemes per utterance (generation length = 10*lx - prompt frames - 9)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--codec-precision", default="bf16", choices=["bf16", "fp32"], help="decoder/wmdecode convolutions: bf16 tcgen05 or fp32 CUDA cores (encode is always fp32)")
    ap.add_argument("--codec-chunk", type=int, default=32, help="utterances per codec pass (bounds the activation arena)")
    ap.add_argument("--no-watermark", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batch-sweep", action="store_true", help="skip the batch 1 / 8 points of BASELINE's metric (reported under other_batches)")
    ap.add_argument("--cpu-iters", type=int, default=12)
    ap.add_argument("--profile-iters", type=int, default=6)
    return ap.parse_args()


def decode_config():
    return {"top_k": 0, "top_p": 0.8, "temperature": 1.0, "stop_repetition": 2, "kvcache": 1, "codec_sr": 50,
            "silence_tokens": (1388, 1898, 131)}


def synth_inputs(batch: int, prompt_sec: float, lx: int, rank: int):
    T = int(round(prompt_sec * 50)) * 320
    wavs, texts, spans = [], [], []
    for i in range(batch):
        g = torch.Generator().manual_seed(1234 + rank * 100000 + i)
        wavs.append(0.1 * torch.randn(1, T, generator=g))
        texts.append(torch.randint(0, 100, (lx,), generator=g))
        spans.append([[T // 320, T // 320]])
    return wavs, texts, spans


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {getattr(nv, n): n.replace("nvmlClocksThrottleReason", "").replace("nvmlClocksEventReason", "")
                     for n in dir(nv) if n.startswith("nvmlClocksThrottleReason") and isinstance(getattr(nv, n), int)}
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, nm in names.items():
                        if bit and (r & bit) and nm not in ("None", "GpuIdle", "All"):
                            self.reasons.add(nm)
                except Exception:
                    pass
                time.sleep(0.2)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# =====================================================================================================================
# reference / oracle CPU arm
# =====================================================================================================================
def cpu_reference_sample(args, n_iters: int):
    """Times the reference's algorithm for ONE utterance of the workload on the host cores: the full prefill
    (2 CFG rows x (lx + prompt frames + 10) positions) + `n_iters` decode iterations, then extrapolates to the full
    generation length.  The reference has no batched inference (inference_v2.py:331-333 loops over utterances), so
    the batch figure equals the single-utterance figure by construction."""
    from ssr_speech_b200.config import cfg_830m
    from ssr_speech_b200.synth import make_lm_state_dict
    from ssr_speech_b200 import seq
    import ref_loader
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = cfg_830m()
    sd = make_lm_state_dict(cfg, seed=0, pin_eog_bias=True)
    T = int(round(args.prompt_sec * 50))
    g = torch.Generator().manual_seed(1234)
    x = torch.randint(0, 100, (args.lx,), generator=g)
    y = torch.randint(0, 2048, (T, K_CODEBOOKS), generator=g)      # codes (the CPU codec is not part of this sample)
    prep = seq.prepare(cfg, y.T.contiguous().numpy(), [[T, T]])
    gen_frames = 10 * args.lx - (prep.prompt_tokens.shape[1] + 1) + 2 - 1
    n_total = gen_frames + K_CODEBOOKS                              # iterations of the full loop
    kw = dict(top_k=0, top_p=0.8, temperature=1.0, stop_repetition=2, cfg_coef=1.5, cfg_stride=5, aug_text=True)
    if ref_loader.reference_available():
        kind = "reference"
        ssr = ref_loader.load_reference_ssr()
        model = ssr.SSR_Speech(cfg.to_namespace()).eval()
        model.load_state_dict(sd)
        del sd
        # bounded run of the UNMODIFIED reference: count forward calls, stop after n_iters decode iterations
        calls = {"n": 0, "t": []}
        orig = model.dec_forward

        class _Stop(Exception):
            pass

        def timed(*a, **k):
            t0 = time.perf_counter()
            out = orig(*a, **k)
            calls["t"].append(time.perf_counter() - t0)
            calls["n"] += 1
            if calls["n"] > n_iters:
                raise _Stop()
            return out
        model.dec_forward = timed
        t0 = time.perf_counter()
        try:
            with torch.no_grad():
                model.inference(x[None], torch.tensor([args.lx]), x[None], torch.tensor([args.lx]), y[None], y[None],
                                mask_interval=torch.tensor([[[T, T]]]), kvcache=1, **kw)
        except _Stop:
            pass
        wall = time.perf_counter() - t0
        t_prefill = calls["t"][0]
        t_iter = (wall - t_prefill) / max(calls["n"] - 1, 1)       # includes the reference's per-step mask/cat overheads
    else:
        kind = "port"
        from lm_oracle import LMOracle
        oracle = LMOracle(cfg, sd)
        del sd
        # "all the host threads it can use": torch CPU GEMV/LSTM work stops scaling (and then collapses) far below the
        # core count of a large host, so pick the thread count that makes the reference's arithmetic fastest.
        best = (None, 1e30)
        h = torch.randn(2, cfg.d_model)
        w = oracle.sd["decoder.layers.0.linear1.weight"]
        for nt in sorted({1, 4, 8, 16, 32, 64, os.cpu_count() or 1}):
            if nt > (os.cpu_count() or 1):
                continue
            torch.set_num_threads(nt)
            torch.nn.functional.linear(h, w)
            t0 = time.perf_counter()
            for _ in range(5):
                torch.nn.functional.linear(h, w)
            dt = time.perf_counter() - t0
            if dt < best[1]:
                best = (nt, dt)
        torch.set_num_threads(best[0])
        # one bounded run; the prediction heads are evaluated once per row per iteration, so their call times mark the
        # iteration boundaries (row 0 and row 1 of the CFG pair)
        stamps = []
        orig_heads = oracle.heads

        def heads_timed(hvec):
            out = orig_heads(hvec)
            stamps.append(time.perf_counter())
            return out
        oracle.heads = heads_timed
        t0 = time.perf_counter()
        oracle.inference(x, torch.from_numpy(prep.prompt_tokens), 1, max_steps=1 + n_iters, **kw)
        oracle.heads = orig_heads
        R = 2
        t_prefill = stamps[R - 1] - t0
        t_iter = (stamps[-1] - stamps[R - 1]) / max(len(stamps) // R - 1, 1)
    t_codec = cpu_codec_seconds(args, T, gen_frames)
    t_utt = t_codec + t_prefill + (n_total - 1) * t_iter
    tokens = K_CODEBOOKS * gen_frames
    return {"value": tokens / t_utt, "unit": "codec-tokens/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"1 utterance of the workload, fp32: WM-Encodec encode + wmdecode ({t_codec:.2f} s) + full prefill "
                      f"({t_prefill:.2f} s) + {n_iters} decode iterations ({t_iter * 1e3:.1f} ms each) extrapolated to "
                      f"{n_total} iterations; the batch-B figure is identical by construction (the reference loops over "
                      f"utterances sequentially, inference_v2.py:331-333)",
            "t_prefill_s": t_prefill, "t_iter_s": t_iter, "t_codec_s": t_codec, "gen_frames": gen_frames}


_CODEC_S = {}


def cpu_codec_seconds(args, T: int, gen_frames: int) -> float:
    """encode(prompt) + wmdecode(prompt + generation) of one utterance on the host cores (reference code when present)."""
    key = (T, gen_frames)
    if key in _CODEC_S:
        return _CODEC_S[key]
    from ssr_speech_b200.config import CodecConfig
    from ssr_speech_b200.synth import make_codec_state_dict
    import ref_loader
    ccfg = CodecConfig()
    sd = make_codec_state_dict(ccfg, seed=0)
    wav = 0.1 * torch.randn(1, 1, T * 320, generator=torch.Generator().manual_seed(1))
    n_out = T + gen_frames
    codes = torch.randint(0, ccfg.bins, (1, ccfg.n_q, n_out), generator=torch.Generator().manual_seed(2))
    marks = torch.zeros(1, n_out, dtype=torch.long)
    marks[:, T:] = 1
    new_wav = torch.zeros(1, 1, n_out * 320)
    new_wav[..., :T * 320] = wav
    with torch.no_grad():
        if ref_loader.reference_available():
            m = ref_loader.build_reference_codec()
            m.load_state_dict(sd)
            t0 = time.perf_counter()
            m.encode(wav)
            if args.no_watermark:
                m.decode(codes, None)
            else:
                m.wmdecode(codes, marks, new_wav, None)
        else:
            from codec_oracle import CodecOracle
            o = CodecOracle(ccfg, sd)
            t0 = time.perf_counter()
            o.encode(wav)
            if args.no_watermark:
                o.decode(codes)
            else:
                o.wmdecode(codes, marks, new_wav)
    _CODEC_S[key] = time.perf_counter() - t0
    return _CODEC_S[key]


def cpu_reference_full(args):
    """ONE full utterance of the workload on the host cores, start to end, nothing extrapolated: WM-Encodec encode of the prompt,
    the whole `SSR_Speech.inference` roll-out (prefill of both CFG rows + every decode iteration up to the reference's own length
    guard, ssr.py:739) and wmdecode of prompt + generation.  The unmodified reference when its tree is present
    (oracle/ref_loader.py: SSRB_REFERENCE_ROOT, /root/reference, baseline/_ref), else the oracle port (same arithmetic through torch
    CPU kernels).  The batch-B figure equals this one by construction (the reference loops over utterances, inference_v2.py:331-333)."""
    from ssr_speech_b200.config import cfg_830m
    from ssr_speech_b200.synth import make_lm_state_dict
    from ssr_speech_b200 import seq
    import ref_loader
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = cfg_830m()
    sd = make_lm_state_dict(cfg, seed=0, pin_eog_bias=True)
    T = int(round(args.prompt_sec * 50))
    g = torch.Generator().manual_seed(1234)
    x = torch.randint(0, 100, (args.lx,), generator=g)
    y = torch.randint(0, 2048, (T, K_CODEBOOKS), generator=g)
    prep = seq.prepare(cfg, y.T.contiguous().numpy(), [[T, T]])
    kw = dict(top_k=0, top_p=0.8, temperature=1.0, stop_repetition=2, cfg_coef=1.5, cfg_stride=5, aug_text=True)
    torch.manual_seed(0)
    if ref_loader.reference_available():
        kind = "reference"
        ssr = ref_loader.load_reference_ssr()
        model = ssr.SSR_Speech(cfg.to_namespace()).eval()
        model.load_state_dict(sd)
        del sd
        t0 = time.perf_counter()
        with torch.no_grad():
            res = model.inference(x[None], torch.tensor([args.lx]), x[None], torch.tensor([args.lx]), y[None], y[None],
                                  mask_interval=torch.tensor([[[T, T]]]), kvcache=1, **kw)[0]
        t_lm = time.perf_counter() - t0
        gen_frames = int(res.shape[-1]) - T
    else:
        kind = "port"
        from lm_oracle import LMOracle
        oracle = LMOracle(cfg, sd)
        del sd
        t0 = time.perf_counter()
        spans = oracle.inference(x, torch.from_numpy(prep.prompt_tokens), 1, **kw)
        t_lm = time.perf_counter() - t0
        gen_frames = int(spans[0].shape[0]) - K_CODEBOOKS
    t_codec = cpu_codec_seconds(args, T, gen_frames)
    t_utt = t_codec + t_lm
    tokens = K_CODEBOOKS * gen_frames
    return {"value": tokens / t_utt, "unit": "codec-tokens/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"ONE full utterance of the workload, fp32, nothing extrapolated: WM-Encodec encode + wmdecode ({t_codec:.2f} s) + "
                      f"the whole inference roll-out ({gen_frames + K_CODEBOOKS} iterations, {t_lm:.1f} s) for {gen_frames} generated frames; "
                      f"the batch-B figure is identical by construction (the reference loops over utterances sequentially, "
                      f"inference_v2.py:331-333)",
            "t_utt_s": t_utt, "t_lm_s": t_lm, "t_codec_s": t_codec, "gen_frames": gen_frames}


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the path on the host cores.  The measurement is ONE FULL utterance
    (cpu_reference_full); warm-up steps and any further timed steps are bounded samples (full prefill + a few decode iterations,
    extrapolated) that keep the K / W contract without running for an hour — their spread is reported next to the value."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(args.warmup):
        cpu_reference_sample(args, 2)
    full = cpu_reference_full(args)
    extra = [cpu_reference_sample(args, max(2, args.cpu_iters // 2))["value"] for _ in range(max(0, args.steps - 1))]
    v = float(full["value"])
    cb = {k: full[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if extra:
        cb["bounded_samples_codec_tokens_per_s"] = [round(e, 2) for e in extra]
    line = {"impl": "reference", "metric": "codec_tokens_per_sec", "value": v, "unit": "codec-tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * full["t_utt_s"] * args.batch, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args), "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "codec-tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args):
    T = int(round(args.prompt_sec * 50))
    exp = {k: os.environ[k] for k in ("SSRB_LAYER_KERNEL", "SSRB_FLAT_2CTA") if os.environ.get(k, "0") not in ("", "0")}
    return {**({"experimental_switches": exp} if exp else {}),      # A/B lines of kernels that are off by default name themselves
            "workload": f"BASELINE configs[2]: English TTS 830M random-init, cfg_coef=1.5 cfg_stride=5 aug_text, top_p=0.8, "
                        f"{args.prompt_sec:g} s prompt -> {(10 * args.lx - T - 9) / 50:g} s generation, batch {args.batch}/GPU",
            "batch_per_gpu": args.batch, "prompt_frames": T, "text_len": args.lx, "rows_per_gpu": 2 * args.batch,
            "precision": args.precision, "codec_decoder_precision": args.codec_precision, "watermark_decode": not args.no_watermark,
            "l2_policy": "inputs larger than L2: each decode iteration streams 1.65 GB of weights + the KV cache",
            "parallelism": f"dp{args.gpus} (utterances sharded, weights replicated)"}


# =====================================================================================================================
# our arm
# =====================================================================================================================
def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    import torch.distributed as dist
    from ssr_speech_b200 import _lib, pipeline
    from ssr_speech_b200.codec import AudioTokenizer, WMEncodecModel
    from ssr_speech_b200.config import CodecConfig, cfg_830m
    from ssr_speech_b200.dist import gather_waveforms
    from ssr_speech_b200.lm import SSR_Speech
    from ssr_speech_b200.synth import calibrate_codebooks, make_codec_state_dict, make_lm_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    # ---- models (random-init, seeded; identical on every rank) -------------------------------------------------------
    cfg = cfg_830m()
    model = SSR_Speech(cfg.to_namespace(), precision=args.precision)
    model.load_state_dict(make_lm_state_dict(cfg, seed=0, pin_eog_bias=True))
    model.to(dev).eval()
    ccfg = CodecConfig()
    codec = WMEncodecModel(ccfg, max_batch_chunk=args.codec_chunk, precision=args.codec_precision)
    codec.load_state_dict(make_codec_state_dict(ccfg, seed=0))
    codec.to(dev)
    cal = 0.1 * torch.randn(4, 1, 32000, generator=torch.Generator().manual_seed(7))
    _, _, emb = codec.encode(cal.to(dev))                              # calibrate the synthetic codebooks on real latents
    mu, sigma = calibrate_codebooks(ccfg, 0, emb.permute(0, 2, 1).reshape(-1, ccfg.dimension).cpu())
    codec.load_state_dict(make_codec_state_dict(ccfg, seed=0, codebook_mu=mu, codebook_sigma=sigma))
    codec.to(dev)
    tok = AudioTokenizer(model=codec, device=dev)

    wavs, texts, spans = synth_inputs(args.batch, args.prompt_sec, args.lx, rank)
    wavs = [w.pin_memory() for w in wavs]
    dc = decode_config()
    T = wavs[0].shape[-1]
    # one pinned result buffer for the whole run, as a serving loop would keep (results are copied into it asynchronously)
    host_out = torch.empty(args.batch, 1, (T // 320 + 10 * args.lx) * 320, dtype=torch.float32).pin_memory()
    gen_frames = 10 * args.lx - (T // 320 + 10) + 1
    tokens_per_step = args.batch * K_CODEBOOKS * gen_frames            # per rank

    def step_host(timings):
        # N > 1: the finished waveforms stay in HBM, ONE NCCL all_gather over NVLink, then ONE device-to-host copy
        out, results = pipeline.inference_batch(model, tok, wavs, texts, spans, dc, cfg_coef=1.5, cfg_stride=5, aug_text=True,
                                                use_watermark=not args.no_watermark, tts=True, seed=1000, timings=timings,
                                                to_host=(world == 1), host_out=host_out if world == 1 else None)
        if world > 1:
            out = gather_waveforms(out, device=dev, to_host=True)
        return out, results

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also builds engines, captures the decode graph) -------------------------------------------------------
    for _ in range(args.warmup):
        tm = {}
        out, results = step_host(tm)
    got = int(results[0][1].sum())
    assert got == gen_frames, f"generated {got} frames, expected {gen_frames}"

    # ---- timed region A: host buffers (e2e) --------------------------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.ssrb_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    acc = {"encode_ms": 0.0, "lm_ms": 0.0, "decode_ms": 0.0, "lm_prefill_ms": 0.0, "lm_decode_ms": 0.0}
    for _ in range(args.steps):
        tm = {}
        out, results = step_host(tm)
        for k in acc:
            acc[k] += tm.get(k, 0.0)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    launches = lib.ssrb_launch_count() - launches0
    # the waveforms cross the bus once (the watermark decoder's new_wav is spliced on the device from the encoder's copy);
    # text ids as int32 for the cond and the uncond row
    h2d = sum(w.numel() * 4 for w in wavs) + sum(t.numel() * 4 for t in texts) * 2
    # per rank: its own waveforms at N = 1; at N > 1 every rank reads the whole gathered job back in one copy
    d2h = sum(o.numel() * 4 for o in out) + args.batch * K_CODEBOOKS * (gen_frames + 4) * 4

    # ---- timed region B: inputs resident in HBM (value) ----------------------------------------------------------------------
    wav_dev = torch.stack(wavs, 0).to(dev)
    barrier()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dec_ms, dec_bytes = 0.0, 0.0
    d0.record()
    for _ in range(args.steps):
        codes, _, _ = tok.encode(wav_dev)
        ys = [codes[i].transpose(0, 1) for i in range(args.batch)]
        res = model.inference_batch(texts, ys, spans, top_k=dc["top_k"], top_p=dc["top_p"], temperature=dc["temperature"],
                                    stop_repetition=dc["stop_repetition"], cfg_coef=1.5, cfg_stride=5, aug_text=True, seed=1000,
                                    device=dev)
        dec_ms += model.last_stats["decode_ms"]
        fr = torch.cat([r[0] for r in res], 0)
        if args.no_watermark:
            _ = tok.decode(fr, None)
        else:
            mk = torch.cat([r[1] for r in res], 0).to(dev)
            new_wav = torch.zeros(args.batch, 1, fr.shape[-1] * 320, device=dev)
            new_wav[:, :, :T] = wav_dev
            _ = tok.wmdecode(fr, mk, new_wav, None)
    d1.record()
    barrier()
    dev_ms = d0.elapsed_time(d1)
    clocks = sampler.stop()

    # ---- roofline of the decode iteration (algorithmic bytes: SURVEY §8d) ------------------------------------------------------
    n_iter = gen_frames + K_CODEBOOKS                       # loop iterations incl. the prefill sample
    wb, _ = model.step_bytes()
    R = 2 * args.batch
    S0 = args.lx + T // 320 + 10
    kv_per_pos = cfg.num_decoder_layers * 2 * cfg.d_model * (2 if args.precision == "bf16" else 4)
    # iterations 2..n_iter each read S0+j-1 cached positions + write 1, for all R rows (rows finish together here)
    kv_total = sum(R * kv_per_pos * ((S0 + j) + 1) for j in range(1, n_iter))
    dec_bytes = (n_iter - 1) * wb + kv_total
    dec_s = dec_ms / args.steps / 1e3
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", FALLBACK_HBM_GBS))
    achieved = dec_bytes / dec_s / 1e9
    # per-kernel-class breakdown of one iteration at full context (un-graphed, CUDA events around every class)
    breakdown = None
    try:
        # re-open the batch, advance to mid-generation with the graph, then profile a few un-graphed iterations
        breakdown = profile_breakdown(model, lib, texts, ys, spans, dc, args, gen_frames)
    except Exception as e:  # pragma: no cover
        breakdown = {"error": repr(e)}

    # ---- the other batch sizes BASELINE.json's metric names (1 and 8 per GPU), same workload, e2e from host buffers -------
    other = {}
    if not args.no_batch_sweep and args.batch == 32 and world == 1:
        for B in (1, 8):
            try:
                other[str(B)] = batch_point(B, model, tok, wavs, texts, spans, dc, args, gen_frames, wb, kv_per_pos, S0, n_iter, peak, host_out=host_out)
            except Exception as e:  # pragma: no cover
                other[str(B)] = {"error": repr(e)}

    # ---- the other single-GPU configurations BASELINE.json lists: configs[1] (TTS 3 s -> 5 s, batch 1, greedy, no CFG) and
    # configs[3] (mid-span edit [200, 300) of a 10 s context, batch 8, CFG), each with its decode-loop roofline fraction ----------
    other_cfg = {}
    if not args.no_batch_sweep and args.batch == 32 and world == 1:
        for name, kw in (("configs[1] tts 3s->5s batch1 greedy no-cfg", dict(B=1, T=150, lx=41, span=None, aug_text=False, top_k=1, top_p=1.0, stop_rep=-1)),
                         ("configs[3] edit [200,300) of 10s batch8 cfg", dict(B=8, T=500, lx=51, span=(200, 300), aug_text=True, top_k=0, top_p=0.8, stop_rep=2))):
            try:
                other_cfg[name] = config_point(model, tok, cfg, args, wb, peak, **kw)
            except Exception as e:  # pragma: no cover
                other_cfg[name] = {"error": repr(e)}

    # ---- aggregate over ranks -------------------------------------------------------------------------------------------------
    t = torch.tensor([e2e_ms, dev_ms, dec_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms, dev_ms, dec_ms_max = [float(v) for v in t]
    total_tokens = tokens_per_step * args.steps * world
    value = total_tokens / (dev_ms / 1e3)
    e2e_val = total_tokens / (e2e_ms / 1e3)
    gen_audio_s = gen_frames / 50.0
    # DRAM traffic of one decode iteration from the committed `ncu --set full` captures (profiles/ncu_traffic.json): the
    # capture sits at iteration 251 of 505, whose algorithmic bytes equal the loop average within 1 %; only quoted for
    # the geometry it was captured on
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
        if args.batch == 32 and args.lx == 101 and args.prompt_sec == 10.0 and args.precision == "bf16":
            traffic = float(tj["iteration"]["dram_bytes"])
            traffic_src = ("profiles/ncu_traffic.json: dram read+write of the 16 attention + 64 layer GEMM launches of iteration 251 "
                           f"(algorithmic {tj['iteration']['algorithmic_bytes']} B at that iteration)")
    except Exception:  # pragma: no cover
        pass
    line = {
        "metric": "codec_tokens_per_sec", "value": value, "unit": "codec-tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": workload_config(args), "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "codec-tokens/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_ms / args.steps,
                "phase_ms_per_step": {k: v / args.steps for k, v in acc.items()}},
        "gpu_launches": int(launches),
        "per_gpu": {"codec_tokens_per_sec": value / world, "rtf_batch": (dev_ms / args.steps / 1e3) / gen_audio_s,
                    "rtf_per_utterance_amortised": (dev_ms / args.steps / 1e3) / (gen_audio_s * args.batch),
                    "e2e_rtf_batch": (e2e_ms / args.steps / 1e3) / gen_audio_s,
                    "lm_only_codec_tokens_per_sec": tokens_per_step * args.steps / ((acc["lm_ms"]) / 1e3) if acc["lm_ms"] else None,
                    "decode_iterations_per_sec": (n_iter - 1) / dec_s},
        "roofline": {"bound": "hbm", "kernel": "decode iteration (CUDA graph: 16 layers + heads + sampler)", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_iteration_avg": dec_bytes / (n_iter - 1),
                     "weight_bytes_per_iteration": wb, "iteration_ms_avg": 1e3 * dec_s / (n_iter - 1), "breakdown": breakdown},
    }
    if other:
        line["other_batches"] = other
    if other_cfg:
        line["other_configs"] = other_cfg
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cb = cpu_reference_sample(args, args.cpu_iters)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # pragma: no cover
            line["cpu_baseline"] = {"value": None, "unit": "codec-tokens/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {e!r}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def batch_point(B, model, tok, wavs, texts, spans, dc, args, gen_frames, wb, kv_per_pos, S0, n_iter, peak, host_out=None):
    """One point of BASELINE's batch sweep on this rank: 1 warm-up + 2 timed passes of the whole hot path (host buffers in,
    waveforms out) at batch B, with the decode loop's HBM roofline fraction at that batch."""
    from ssr_speech_b200 import pipeline
    run = lambda tm: pipeline.inference_batch(model, tok, wavs[:B], texts[:B], spans[:B], dc, cfg_coef=1.5, cfg_stride=5,
                                              aug_text=True, use_watermark=not args.no_watermark, tts=True, seed=1000, timings=tm,
                                              host_out=None if host_out is None else host_out[:B])
    run({})
    torch.cuda.synchronize()
    n, acc = 2, {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        tm = {}
        run(tm)
        for k, v in tm.items():
            acc[k] = acc.get(k, 0.0) + v
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    dec_s = acc.get("lm_decode_ms", 0.0) / n / 1e3
    R = 2 * B
    bytes_dec = (n_iter - 1) * wb + sum(R * kv_per_pos * ((S0 + j) + 1) for j in range(1, n_iter))
    return {"batch_per_gpu": B, "e2e_codec_tokens_per_sec": B * K_CODEBOOKS * gen_frames / (ms / 1e3), "ms_per_step": ms,
            "e2e_rtf_batch": (ms / 1e3) / (gen_frames / 50.0), "phase_ms_per_step": {k: v / n for k, v in acc.items()},
            "decode_iteration_ms_avg": 1e3 * dec_s / (n_iter - 1) if dec_s else None,
            "roofline_frac": bytes_dec / dec_s / 1e9 / peak if dec_s else None}


def config_point(model, tok, cfg, args, wb, peak, B, T, lx, span, aug_text, top_k, top_p, stop_rep):
    """One of BASELINE.json's other single-GPU configurations through the whole hot path (host waveforms in, host waveforms out):
    1 warm-up + 2 timed passes.  Generation length = 10 * lx - Y0 + 1 frames (the reference's length guard, ssr.py:739), with
    Y0 = T + 10 for TTS and T + 10 - (b - a) for a one-span edit."""
    from ssr_speech_b200 import pipeline
    wavs, texts, _ = synth_inputs(B, T / 50.0, lx, rank=7)
    wavs = [w.pin_memory() for w in wavs]
    host_out = torch.empty(B, 1, (T + 10 * lx) * 320, dtype=torch.float32).pin_memory()
    tts = span is None
    spans = [[[T, T]] if tts else [list(span)]] * B
    dc = {"top_k": top_k, "top_p": top_p, "temperature": 1.0, "stop_repetition": stop_rep, "kvcache": 1, "codec_sr": 50,
          "silence_tokens": (1388, 1898, 131)}
    y0 = T + 10 - (0 if tts else span[1] - span[0])
    gen_frames = 10 * lx - y0 + 1
    run = lambda tm: pipeline.inference_batch(model, tok, wavs, texts, spans, dc, cfg_coef=1.5, cfg_stride=5, aug_text=aug_text,
                                              use_watermark=not args.no_watermark, tts=tts, seed=1000, timings=tm, host_out=host_out)
    tm0 = {}
    _, res = run(tm0)
    got = int(res[0][1].sum())
    assert got == gen_frames, f"generated {got} frames, expected {gen_frames}"
    torch.cuda.synchronize()
    n, acc = 2, {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        tm = {}
        run(tm)
        for k, v in tm.items():
            acc[k] = acc.get(k, 0.0) + v
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    dec_s = acc.get("lm_decode_ms", 0.0) / n / 1e3
    R = (2 if aug_text else 1) * B
    S0 = lx + y0
    n_iter = gen_frames + K_CODEBOOKS
    kv_per_pos = cfg.num_decoder_layers * 2 * cfg.d_model * (2 if args.precision == "bf16" else 4)
    bytes_dec = (n_iter - 1) * wb + sum(R * kv_per_pos * ((S0 + j) + 1) for j in range(1, n_iter))
    return {"batch": B, "rows": R, "prompt_frames": T, "text_len": lx, "generated_frames": gen_frames, "decode_iterations": n_iter,
            "e2e_codec_tokens_per_sec": B * K_CODEBOOKS * gen_frames / (ms / 1e3), "ms_per_step": ms,
            "e2e_rtf_batch": (ms / 1e3) / (gen_frames / 50.0), "phase_ms_per_step": {k: v / n for k, v in acc.items()},
            "decode_iteration_ms_avg": 1e3 * dec_s / (n_iter - 1) if dec_s else None,
            "roofline_frac": bytes_dec / dec_s / 1e9 / peak if dec_s else None}


def profile_breakdown(model, lib, texts, ys, spans, dc, args, gen_frames):
    """Critical-path decomposition of one decode iteration near mid-generation, measured INSIDE the CUDA-graph / PDL chain with
    in-kernel %globaltimer stamps (ssr_speech_b200.timeline): per kernel class, the sum over its launches of
    max(CTA exit) - min(dependency resolved).  HBM rates: attention over the KV bytes, GEMMs over the weight bytes."""
    from ssr_speech_b200 import _lib, timeline
    half = gen_frames // 2
    model.open_batch(texts, ys, spans, top_k=dc["top_k"], top_p=dc["top_p"], temperature=dc["temperature"],
                     stop_repetition=dc["stop_repetition"], cfg_coef=1.5, cfg_stride=5, aug_text=True, seed=1000)
    st = _lib.stream_ptr()
    _lib.check(lib.ssrb_lm_decode(model._h, half, st), "decode")
    wb, kb = model.step_bytes()
    n = max(2, args.profile_iters)
    cp = timeline.critical_path(timeline.capture(model, n), n)
    out = {"at_iteration": half, "method": "in-kernel globaltimer stamps under the CUDA graph (critical path per class)",
           "iteration_us": cp["wall_us"], "attention_us": cp["attention_us"], "gemm_us": cp["gemm_us"],
           "layernorm_us": cp["layernorm_us"], "embed_us": cp["embed_us"], "sample_us": cp["sample_us"],
           "gemm_launches": cp["gemm_launches"], "attention_launches": cp["attention_launches"],
           "attention_GBps": kb / (cp["attention_us"] * 1e-6) / 1e9 if cp["attention_us"] else None,
           "gemm_GBps": wb / (cp["gemm_us"] * 1e-6) / 1e9 if cp["gemm_us"] else None, "kv_bytes": kb, "weight_bytes": wb}
    _lib.check(lib.ssrb_lm_decode(model._h, gen_frames, st), "decode")   # drain
    torch.cuda.synchronize()
    return out


if __name__ == "__main__":
    main()
