#!/usr/bin/env python
"""Per-phase timeline of ONE iteration of the persistent small-batch decode kernel (csrc/lm_mega.cu): every CTA's consumer thread 0
stamps %globaltimer at {phase start, activations staged, chunks consumed, phase end}.  Prints, per phase kind, medians over CTAs
and layers of: barrier wait (previous end -> start), staging, chunk consumption, epilogue.
    python tools/mega_trace.py --batch 1 [--skip 250]"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--skip", type=int, default=250)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    from ssr_speech_b200 import _lib
    from ssr_speech_b200.config import cfg_830m
    from ssr_speech_b200.lm import SSR_Speech
    from ssr_speech_b200.synth import make_lm_state_dict
    cfg = cfg_830m()
    m = SSR_Speech(cfg.to_namespace(), precision="bf16")
    m.load_state_dict(make_lm_state_dict(cfg, seed=0, pin_eog_bias=True))
    m.to("cuda:0")
    g = torch.Generator().manual_seed(0)
    B = args.batch
    xs = [torch.randint(0, 100, (101,), generator=g) for _ in range(B)]
    ys = [torch.randint(0, 2048, (500, 4), generator=g) for _ in range(B)]
    m.open_batch(xs, ys, [[[500, 500]]] * B, top_k=0, top_p=0.8, stop_repetition=2, cfg_coef=1.5, cfg_stride=5, aug_text=True, seed=1)
    assert m.decode_path() == 3
    lib, st = _lib.load(), _lib.stream_ptr()
    _lib.check(lib.ssrb_lm_decode(m._h, args.skip, st), "decode")
    torch.cuda.synchronize()
    cap, G = 512, C.c_int(0)
    buf = torch.zeros(256 * cap, dtype=torch.int64, device="cuda")
    _lib.check(lib.ssrb_debug_mega_trace(C.c_void_p(buf.data_ptr()), cap, C.byref(G)), "trace")
    _lib.check(lib.ssrb_lm_decode(m._h, 1, st), "decode")
    torch.cuda.synchronize()
    _lib.check(lib.ssrb_debug_mega_trace(None, 0, C.byref(G)), "trace")
    t = buf.view(256, cap)[:G.value].cpu().numpy().astype(np.float64)
    if args.out:
        np.save(args.out, t)
    L = cfg.num_decoder_layers
    # stamp layout per CTA: entry, dep | layer l: [QKV: (start if l > 0), staged, consumed, end] [ATTN: start, end] [OUT|FFN1|FFN2: start, staged, consumed, end] | H1, H2 | exit
    t0 = t[:, 1].min()
    print(f"grid {G.value} CTAs; kernel entry -> dependency resolved: median {np.median(t[:, 1] - t[:, 0]) / 1e3:.2f} us")
    idx = 2
    rows = {k: [] for k in ("QKV", "ATTN", "OUT", "FFN1", "FFN2", "H1", "H2")}
    prev_end = t[:, 1].copy()
    def gemv(name, has_start):
        nonlocal idx, prev_end
        start = t[:, idx] if has_start else prev_end
        if has_start:
            idx += 1
        staged, consumed, end = t[:, idx], t[:, idx + 1], t[:, idx + 2]
        idx += 3
        rows[name].append((np.median(start - prev_end), np.median(staged - start), np.median(consumed - staged), np.median(end - consumed),
                           (end.max() - start.min())))
        prev_end = end
    for l in range(L):
        gemv("QKV", l > 0)
        start, end = t[:, idx], t[:, idx + 1]
        idx += 2
        rows["ATTN"].append((np.median(start - prev_end), 0.0, np.median(end - start), 0.0, end.max() - start.min()))
        prev_end = end
        gemv("OUT", True); gemv("FFN1", True); gemv("FFN2", True)
    gemv("H1", True); gemv("H2", True)
    total = (t[:, idx].max() - t0) / 1e3
    print(f"{'phase':6s} {'barrier wait':>13s} {'stage acts':>11s} {'chunks':>9s} {'epilogue':>9s} {'phase span (all CTAs)':>22s}   (us, medians over CTAs, mean over layers)")
    for k, v in rows.items():
        a = np.asarray(v) / 1e3
        print(f"{k:6s} {a[:, 0].mean():13.2f} {a[:, 1].mean():11.2f} {a[:, 2].mean():9.2f} {a[:, 3].mean():9.2f} {a[:, 4].mean():22.2f}   x{len(v)}")
    print(f"iteration (dependency resolved -> last exit): {total:.1f} us")


if __name__ == "__main__":
    main()
