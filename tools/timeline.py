#!/usr/bin/env python
"""Prints the in-kernel timeline (ssr_speech_b200.timeline) of a few decode iterations of the benchmark batch."""
import argparse
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
    from ssr_speech_b200 import _lib, timeline
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
    rec = timeline.capture(m, args.iters)
    np.save(args.out, rec)
    t0 = rec[:, 2].min()
    print(f"{len(rec)} CTA records; us relative to the first record; dep = dependency resolved (griddepcontrol.wait returned)")
    for i, (kid, r) in enumerate(timeline.launches(rec)):
        dep = np.where(r[:, 3] > 0, r[:, 3], r[:, 2])
        line = (f"{i:3d} {timeline.KERNELS[kid]:9s} ctas={len(r):5d} first_start={(r[:, 2].min() - t0) / 1e3:9.2f} "
                f"dep={(dep.min() - t0) / 1e3:9.2f} last_exit={(r[:, 4].max() - t0) / 1e3:9.2f} dep->exit={(r[:, 4].max() - dep.min()) / 1e3:7.2f}")
        if kid == 3 and not r[:, 8].any() and not r[:, 9].any() and r[:, 5].any() and len(r) >= 64:
            # persistent per-layer kernel (gemm_layer.cu, SSRB_LAYER_KERNEL=1): aux0..2 = producer thread past grid barrier 1..3
            d = dep.astype(np.float64)
            med = lambda a: float(np.median(a[a > 0])) / 1e3 if (a > 0).any() else float("nan")
            line += (f" | LAYER kernel, medians after dep: past_barrier1={med(r[:, 5] - d):6.2f} past_barrier2={med(r[:, 6] - d):6.2f} "
                     f"past_barrier3={med(r[:, 7] - d):6.2f} exit={med(r[:, 4] - d):6.2f} start={float(np.median(r[:, 2] - d)) / 1e3:6.2f}")
        elif kid == 3:
            d = dep.astype(np.float64)
            med = lambda a: float(np.median(a)) / 1e3
            line += (f" | medians after dep: loads_issued={med(r[:, 5] - d):5.2f} accum={med(r[:, 6] - d):5.2f} parked={med(r[:, 7] - d):5.2f} "
                     f"cluster_bar={med(r[:, 8] - d):5.2f} reduced={med(r[:, 9] - d):5.2f} start={med(r[:, 2] - d):6.2f}")
        if kid == 5:
            d = dep.astype(np.float64)
            med = lambda a: float(np.median(a)) / 1e3
            line += (f" | medians after dep: loaded={med(r[:, 5] - d):5.2f} softmax={med(r[:, 6] - d):5.2f} top_p={med(r[:, 7] - d):5.2f} "
                     f"sampled={med(r[:, 8] - d):5.2f} exit={med(r[:, 4] - d):5.2f}")
        print(line)
    print(timeline.critical_path(rec, args.iters))


if __name__ == "__main__":
    main()
