#!/usr/bin/env python
"""Decode-iteration time of the 830M model at small batches (BASELINE's batch 1 / 8 points): 10 s prompt, CFG rows, iterations
`--skip` .. `--skip + --iters` of the roll-out, CUDA events around the graph replays.  Prints one JSON line.
    python tools/small_batch_probe.py --batch 1                  # per-GEMM chain (product path)
    SSRB_MEGA=1 python tools/small_batch_probe.py --batch 1      # experimental persistent whole-iteration kernel"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--skip", type=int, default=150)
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
    lib, st = _lib.load(), _lib.stream_ptr()
    _lib.check(lib.ssrb_lm_decode(m._h, args.skip, st), "decode")
    torch.cuda.synchronize()
    wb, kb0 = m.step_bytes()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.ssrb_lm_decode(m._h, args.iters, st), "decode")
    e1.record()
    torch.cuda.synchronize()
    _, kb1 = m.step_bytes()
    ms = e0.elapsed_time(e1) / args.iters
    byts = wb + 0.5 * (kb0 + kb1)
    print(json.dumps({"batch": B, "rows": 2 * B, "decode_path": m.decode_path(), "ms_per_iteration": ms,
                      "algorithmic_GB_per_iteration": byts / 1e9, "GBps": byts / ms / 1e6, "frac_of_6554": byts / ms / 1e6 / 6553.9}))


if __name__ == "__main__":
    main()
