"""Multi-GPU: utterances are independent, so a batch is sharded by utterance (weights replicated, no
data-path collective) and the finished waveforms are gathered with ONE all_gather.

The reference has no inference-time parallelism at all (inference_v2.py:4 pins CUDA_VISIBLE_DEVICES=0);
this is the harness BASELINE config 5 asks for: one process per GPU, `torch.distributed` (NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of n_items for `rank` (first n%world ranks get one extra)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_waveforms(local: Sequence[torch.Tensor], device=None) -> List[torch.Tensor]:
    """All-gathers variable-length waveforms [1, T_i] from every rank; returns the global list in rank order.

    One all_gather of the padded payload [n_max, L_max] (+ one tiny all_gather of the lengths)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [w.clone() for w in local]
    world = dist.get_world_size()
    if device is None:
        device = local[0].device if len(local) else torch.device("cpu")
    meta = torch.tensor([len(local), max([w.shape[-1] for w in local], default=0)], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    n_max = int(max(m[0] for m in metas))
    l_max = int(max(m[1] for m in metas))
    payload = torch.zeros(n_max, l_max + 1, dtype=torch.float32, device=device)   # column 0 carries the length
    for i, w in enumerate(local):
        payload[i, 0] = float(w.shape[-1])
        payload[i, 1:1 + w.shape[-1]] = w.reshape(-1).to(device, torch.float32)
    out = [torch.zeros_like(payload) for _ in range(world)]
    dist.all_gather(out, payload)
    res = []
    for r in range(world):
        for i in range(int(metas[r][0])):
            n = int(out[r][i, 0].item())
            res.append(out[r][i, 1:1 + n].reshape(1, -1))
    return res
