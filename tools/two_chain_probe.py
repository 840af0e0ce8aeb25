#!/usr/bin/env python
"""Probe: does running the decode loop as TWO independent half-batches on two CUDA streams (one chain's latency-bound GEMM phase
under the other chain's HBM-bound attention phase) beat one chain over the whole batch?  Two engines (weights duplicated: the
worst case for the weight stream), 16 utterances x 2 CFG rows each, against one engine with 32 utterances.
    python tools/two_chain_probe.py [--iters 200] [--skip 150]
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--skip", type=int, default=150)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--chunk", type=int, default=8)
    args = ap.parse_args()
    from ssr_speech_b200 import _lib
    from ssr_speech_b200.config import cfg_830m
    from ssr_speech_b200.lm import SSR_Speech
    from ssr_speech_b200.synth import make_lm_state_dict
    cfg = cfg_830m()
    sd = make_lm_state_dict(cfg, seed=0, pin_eog_bias=True)
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    B = args.batch
    xs = [torch.randint(0, 100, (101,), generator=g) for _ in range(B)]
    ys = [torch.randint(0, 2048, (500, 4), generator=g) for _ in range(B)]
    kw = dict(top_k=0, top_p=0.8, stop_repetition=2, cfg_coef=1.5, cfg_stride=5, aug_text=True, seed=1)

    def model():
        m = SSR_Speech(cfg.to_namespace(), precision="bf16")
        m.load_state_dict(sd)
        return m.to("cuda:0")

    def sptr(s):
        return C.c_void_p(s.cuda_stream)

    # ---- one chain over the whole batch -------------------------------------------------------------------------------
    m = model()
    s0 = torch.cuda.Stream()
    with torch.cuda.stream(s0):
        m.open_batch(xs, ys, [[[500, 500]]] * B, **kw)
        _lib.check(lib.ssrb_lm_decode(m._h, args.skip, sptr(s0)), "decode")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s0)
        _lib.check(lib.ssrb_lm_decode(m._h, args.iters, sptr(s0)), "decode")
        e1.record(s0)
        torch.cuda.synchronize()
    one = e0.elapsed_time(e1) / args.iters
    print(f"one chain, {B} utterances: {one:.4f} ms / iteration")
    del m
    # ---- two chains of B/2 -----------------------------------------------------------------------------------------------
    h = B // 2
    ma, mb = model(), model()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(sa):
        ma.open_batch(xs[:h], ys[:h], [[[500, 500]]] * h, **kw)
        _lib.check(lib.ssrb_lm_decode(ma._h, args.skip, sptr(sa)), "decode")
    with torch.cuda.stream(sb):
        mb.open_batch(xs[h:], ys[h:], [[[500, 500]]] * h, **kw)
        _lib.check(lib.ssrb_lm_decode(mb._h, args.skip, sptr(sb)), "decode")
    torch.cuda.synchronize()
    # each alone
    for name, mm, ss in (("A", ma, sa), ("B", mb, sb)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ss)
        _lib.check(lib.ssrb_lm_decode(mm._h, 40, sptr(ss)), "decode")
        e1.record(ss)
        torch.cuda.synchronize()
        print(f"chain {name} alone, {h} utterances: {e0.elapsed_time(e1) / 40:.4f} ms / iteration")
    ea0, ea1, eb0, eb1 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    start = torch.cuda.Event(enable_timing=True)
    start.record()
    sa.wait_event(start)
    sb.wait_event(start)
    ea0.record(sa)
    eb0.record(sb)
    n = args.iters - 40
    done = 0
    while done < n:
        c = min(args.chunk, n - done)
        _lib.check(lib.ssrb_lm_decode(ma._h, c, sptr(sa)), "decode")
        _lib.check(lib.ssrb_lm_decode(mb._h, c, sptr(sb)), "decode")
        done += c
    ea1.record(sa)
    eb1.record(sb)
    torch.cuda.synchronize()
    ta, tb = start.elapsed_time(ea1), start.elapsed_time(eb1)
    print(f"two chains concurrently: A {ta / n:.4f}  B {tb / n:.4f}  -> {max(ta, tb) / n:.4f} ms / iteration of the whole batch "
          f"(one chain: {one:.4f}; ratio {max(ta, tb) / n / one:.3f})")


if __name__ == "__main__":
    main()
