#!/usr/bin/env python
"""Short driver for ncu: WM-Encodec encode + wmdecode of B utterances (random-init weights), for the launch list."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--sec", type=float, default=10.0)
    ap.add_argument("--chunk", type=int, default=8)
    ap.add_argument("--precision", default="bf16")
    args = ap.parse_args()
    from ssr_speech_b200.codec import WMEncodecModel
    from ssr_speech_b200.config import CodecConfig
    from ssr_speech_b200.synth import make_codec_state_dict
    cfg = CodecConfig()
    m = WMEncodecModel(cfg, max_batch_chunk=args.chunk, precision=args.precision)
    m.load_state_dict(make_codec_state_dict(cfg, seed=0))
    m.to("cuda:0")
    T = int(args.sec * 50) * 320
    wav = 0.1 * torch.randn(args.batch, 1, T, generator=torch.Generator().manual_seed(0)).cuda()
    codes, _, _ = m.encode(wav)
    m.wmdecode(codes, torch.zeros(args.batch, T // 320, dtype=torch.long, device='cuda'), wav, return_marks=False)
    marks = torch.zeros(args.batch, T // 320, dtype=torch.long, device="cuda")
    marks[:, T // 640:] = 1
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    codes, _, _ = m.encode(wav)
    e[1].record()
    out, _ = m.wmdecode(codes, marks, wav, return_marks=False)
    e[2].record()
    torch.cuda.synchronize()
    print(f"B={args.batch} sec={args.sec} encode {e[0].elapsed_time(e[1]):.1f} ms  wmdecode {e[1].elapsed_time(e[2]):.1f} ms")


if __name__ == "__main__":
    main()
