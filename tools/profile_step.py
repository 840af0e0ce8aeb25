#!/usr/bin/env python
"""Short driver for ncu: builds the bench geometry (830M, B utterances with CFG, 10 s prompt), opens a batch
(prefill) and runs a few decode iterations.  Used for the launch list and the `--set full` captures in profiles/.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python tools/profile_step.py --iters 3 --no-graph
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=4)
    ap.add_argument("--skip-iters", type=int, default=0, help="iterations run through the CUDA graph before the profiled ones")
    ap.add_argument("--lx", type=int, default=101)
    ap.add_argument("--frames", type=int, default=500)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--gemm-impl", type=int, default=0)
    args = ap.parse_args()
    if args.no_graph:
        os.environ["SSRB_NO_GRAPH"] = "1"
    from ssr_speech_b200 import _lib
    from ssr_speech_b200.config import cfg_830m
    from ssr_speech_b200.lm import SSR_Speech
    from ssr_speech_b200.synth import make_lm_state_dict
    cfg = cfg_830m()
    m = SSR_Speech(cfg.to_namespace(), precision=args.precision, gemm_impl=args.gemm_impl)
    m.load_state_dict(make_lm_state_dict(cfg, seed=0, pin_eog_bias=True))
    m.to("cuda:0")
    g = torch.Generator().manual_seed(0)
    xs = [torch.randint(0, 100, (args.lx,), generator=g) for _ in range(args.batch)]
    ys = [torch.randint(0, 2048, (args.frames, 4), generator=g) for _ in range(args.batch)]
    spans = [[[args.frames, args.frames]]] * args.batch
    m.open_batch(xs, ys, spans, top_k=0, top_p=0.8, stop_repetition=2, cfg_coef=1.5, cfg_stride=5, aug_text=True, seed=1)
    lib = _lib.load()
    st = _lib.stream_ptr()
    if args.skip_iters:
        _lib.check(lib.ssrb_lm_decode(m._h, args.skip_iters, st), "decode")
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("profiled_iterations")
    _lib.check(lib.ssrb_lm_decode(m._h, args.iters, st), "decode")
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    wb, kb = m.step_bytes()
    print(f"rows={2 * args.batch} weight_bytes/iter={wb:.0f} kv_bytes/iter={kb:.0f} launches={lib.ssrb_launch_count()}")


if __name__ == "__main__":
    main()
