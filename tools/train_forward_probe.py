#!/usr/bin/env python
"""Timing of the training forward / loss (SSR_Speech.forward) at the 830M configuration: B utterances of Lx phonemes
and Ty audio positions, bf16 production mode, CUDA events around the whole call (host mask logic + H2D of tokens + forward + fused
masked cross entropy + D2H of the sums).  Prints one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--lx", type=int, default=100)
    ap.add_argument("--ty", type=int, default=500)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    from ssr_speech_b200.config import cfg_830m
    from ssr_speech_b200.lm import SSR_Speech
    from ssr_speech_b200.synth import make_lm_state_dict
    cfg = cfg_830m()
    m = SSR_Speech(cfg.to_namespace(), precision="bf16")
    m.load_state_dict(make_lm_state_dict(cfg, seed=0))
    m.to("cuda:0")
    g = torch.Generator().manual_seed(0)
    K, B = cfg.n_codebooks, args.batch
    x = torch.randint(0, cfg.text_vocab_size, (B, args.lx), generator=g)
    y = torch.randint(0, cfg.audio_vocab_size, (B, K, args.ty), generator=g)
    y[:, :, args.ty // 2] = cfg.mts                       # everything after the mask token enters the loss
    y[:, :, -1] = cfg.eog
    batch = {"x": x, "x_lens": torch.full((B,), args.lx), "y": y, "y_lens": torch.full((B,), args.ty)}
    out = m.forward(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        out = m.forward(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    pos = B * (args.lx + args.ty)
    flops = 2.0 * 822.6e6 * pos
    print(json.dumps({"batch": B, "text_len": args.lx, "audio_positions": args.ty, "ms_per_forward": ms, "positions_per_s": pos / ms * 1e3,
                      "model_TFLOPs_per_s": flops / ms / 1e9, "loss": float(out["loss"]), "effective_ntoken": int(out["effective_ntoken"])}))


if __name__ == "__main__":
    main()
