#!/usr/bin/env python
"""In-kernel timeline of a few decode iterations (no nsys in this image): arms ssrb_debug_timeline, replays the CUDA graph,
and prints per-kernel {first CTA start, dependency resolved, last CTA end} relative to the iteration start."""
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
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--skip", type=int, default=250)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--out", default="gpurun_out/timeline.npy")
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
    xs = [torch.randint(0, 100, (101,), generator=g) for _ in range(args.batch)]
    ys = [torch.randint(0, 2048, (500, 4), generator=g) for _ in range(args.batch)]
    m.open_batch(xs, ys, [[[500, 500]]] * args.batch, top_k=0, top_p=0.8, stop_repetition=2, cfg_coef=1.5, cfg_stride=5, aug_text=True, seed=1)
    lib, st = _lib.load(), _lib.stream_ptr()
    _lib.check(lib.ssrb_lm_decode(m._h, args.skip, st), "decode")
    torch.cuda.synchronize()
    cap = 400000
    buf = torch.zeros(cap * 10, dtype=torch.int64, device="cuda")
    idx = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(lib.ssrb_debug_timeline(C.c_void_p(buf.data_ptr()), C.c_void_p(idx.data_ptr()), cap), "timeline")
    _lib.check(lib.ssrb_lm_decode(m._h, args.iters, st), "decode")
    torch.cuda.synchronize()
    _lib.check(lib.ssrb_debug_timeline(None, None, 0), "timeline")
    n = int(idx.item())
    rec = buf[:n * 10].view(n, 10).cpu().numpy()
    np.save(args.out, rec)
    names = {1: "embed", 2: "ln", 3: "gemm", 4: "attn", 5: "sample"}
    t0 = rec[:, 2].min()
    # group consecutive records of the same kernel launch: sort by start, split when kernel id changes or a big gap
    order = np.argsort(rec[:, 2], kind="stable")
    rec = rec[order]
    groups, cur = [], [0]
    for i in range(1, n):
        if rec[i, 0] != rec[cur[-1], 0] or (rec[i, 0] in (1, 2, 5) and rec[i, 1] <= rec[cur[-1], 1] and rec[i, 1] == 0):
            groups.append(cur); cur = [i]
        else:
            cur.append(i)
    groups.append(cur)
    print(f"{n} CTA records, {len(groups)} kernel groups; times in us relative to the first record")
    prev_end = None
    for gi, gidx in enumerate(groups[:140]):
        r = rec[gidx]
        s0, dep, e1 = (r[:, 2].min() - t0) / 1e3, (r[:, 3][r[:, 3] > 0].min() - t0) / 1e3 if (r[:, 3] > 0).any() else float("nan"), (r[:, 4].max() - t0) / 1e3
        gap = "" if prev_end is None else f" gap_after_prev_end={s0 - prev_end:7.2f}"
        print(f"{gi:3d} {names.get(int(r[0, 0]), '?'):6s} ctas={len(gidx):5d} start={s0:9.2f} dep={dep:9.2f} end={e1:9.2f} dur={e1 - s0:7.2f}{gap}")
        prev_end = e1
    # GEMM phase breakdown (medians over CTAs, us): dep->loads issued->accum ready->parked->cluster barrier->reduced->exit
    g = rec[rec[:, 0] == 3]
    g = g[np.argsort(g[:, 3])]
    cuts = np.where(np.diff(g[:, 3]) > 3000)[0] + 1
    for li, idxs in enumerate(np.split(np.arange(len(g)), cuts)[8:24]):
        r = g[idxs].astype(np.float64)
        med = lambda a: float(np.median(a)) / 1e3
        d = r[:, 3]
        print(f"gemm#{li} ctas={len(r):4d} start-dep={med(r[:,2]-d):7.2f} loads_issued={med(r[:,5]-d):6.2f} accum={med(r[:,6]-d):6.2f} "
              f"parked={med(r[:,7]-d):6.2f} cbar={med(r[:,8]-d):6.2f} reduced={med(r[:,9]-d):6.2f} exit={med(r[:,4]-d):6.2f} (max exit {float((r[:,4]-d).max())/1e3:6.2f})")


if __name__ == "__main__":
    main()
