"""Dataset-scale WM-Encodec encoding — drop-in for reference data/encode.py:99-117.

Same CLI arguments, manifest format (json list of {"segment_id", "wav"}) and on-disk output: one `<segment_id>.txt` per
utterance under `<save_dir>/<dataset_name>/<save_tag>/`, K lines of space-separated codes (data/encode.py:53-57), cut to
`round(duration * model_code_sr)` frames (:113-115).  Like the reference, a batch is zero-padded to its longest utterance and
encoded in one call; multi-GPU: launch under torchrun and every rank takes a contiguous shard of [start, end).
"""
from __future__ import annotations

import argparse
import json
import logging
import os

import numpy as np
import torch

from .codec import AudioTokenizer, load_wav
from .dist import shard_range


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="encode the dataset using the WM-Encodec model (B200 kernels)")
    p.add_argument("--json_path", type=str, required=True)
    p.add_argument("--save_dir", type=str, required=True)
    p.add_argument("--save_tag", type=str, default="wmencodec")
    p.add_argument("--dataset_name", type=str, required=True)
    p.add_argument("--encodec_model_path", type=str, required=True)
    p.add_argument("--batch_size", type=int, default=64)
    p.add_argument("--model_sr", type=int, default=16000)
    p.add_argument("--downsample_rate", type=int, default=320)
    p.add_argument("--model_code_sr", type=int, default=50)
    p.add_argument("--start", type=int, default=0)
    p.add_argument("--end", type=int, default=500000)
    return p.parse_args(argv)


def write_code_rows(codes, path: str) -> None:
    """The on-disk format of data/encode.py:53-57: one line per codebook, codes separated by single spaces, no newline after the
    last line."""
    text = "\n".join(" ".join(str(int(c)) for c in row) for row in codes)
    with open(path, "w") as f:
        f.write(text)


def encode_items(tokenizer: AudioTokenizer, items, save_root: str, batch_size: int, model_sr: int, code_sr: int):
    os.makedirs(save_root, exist_ok=True)
    n_done = 0
    for b0 in range(0, len(items), batch_size):
        batch = items[b0:b0 + batch_size]
        audios, durs, ids = [], [], []
        for it in batch:
            wav, sr = load_wav(it["wav"]) if isinstance(it["wav"], str) else (it["wav"], model_sr)
            if sr != model_sr:
                import torchaudio
                wav = torchaudio.transforms.Resample(orig_freq=sr, new_freq=model_sr)(wav)
            a = wav.squeeze()
            if a.ndim > 1:
                a = a.mean(0)
            audios.append(a)
            durs.append(a.shape[-1] / sr)      # as the reference: samples AFTER resampling over the ORIGINAL rate (data/encode.py:86)
            ids.append(it["segment_id"])
        padded = torch.nn.utils.rnn.pad_sequence(audios, batch_first=True).unsqueeze(1)        # [B,1,T]
        codes = tokenizer.encode(padded)[0].cpu()
        for i, d in enumerate(durs):
            fn = os.path.join(save_root, ids[i] + ".txt")
            if not os.path.exists(fn):
                write_code_rows(codes[i, :, :round(d * code_sr)].tolist(), fn)
                n_done += 1
    return n_done


def main(argv=None):
    logging.basicConfig(format="%(asctime)s [%(levelname)s] %(filename)s:%(lineno)d || %(message)s", level=logging.INFO)
    args = parse_args(argv)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    with open(args.json_path) as f:
        data = json.load(f)[args.start:args.end]
    lo, hi = shard_range(len(data), rank, world)
    tok = AudioTokenizer(signature=args.encodec_model_path, device=dev)
    root = os.path.join(args.save_dir, args.dataset_name, args.save_tag)
    n = encode_items(tok, data[lo:hi], root, args.batch_size, args.model_sr, args.model_code_sr)
    logging.info(f"rank {rank}: wrote {n} code files under {root}")


if __name__ == "__main__":
    main()
