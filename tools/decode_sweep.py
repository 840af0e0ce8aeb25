#!/usr/bin/env python
"""ms per decode iteration (CUDA-graph replay, mid-generation) for the bench geometry; used to sweep env-var knobs."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--skip", type=int, default=200)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--no-cfg", action="store_true")
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
    m.open_batch(xs, ys, [[[500, 500]]] * args.batch, top_k=0, top_p=0.8, stop_repetition=2, cfg_coef=1.5, cfg_stride=5,
                 aug_text=not args.no_cfg, seed=1)
    lib, st = _lib.load(), _lib.stream_ptr()
    _lib.check(lib.ssrb_lm_decode(m._h, args.skip, st), "decode")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.ssrb_lm_decode(m._h, args.iters, st), "decode")
    e1.record()
    torch.cuda.synchronize()
    wb, kb = m.step_bytes()
    ms = e0.elapsed_time(e1) / args.iters
    print(f"env={ {k: v for k, v in os.environ.items() if k.startswith('SSRB_')} } batch={args.batch} cfg={not args.no_cfg} "
          f"iter_ms={ms:.4f} bytes={(wb + kb) / 1e9:.3f}GB -> {(wb + kb) / ms / 1e6:.0f} GB/s")


if __name__ == "__main__":
    main()
