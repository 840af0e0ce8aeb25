"""Multi-GPU: utterances are independent, so a batch is sharded by utterance (weights replicated, no
data-path collective) and the finished waveforms are gathered with ONE all_gather on the device.

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


_pinned: dict = {}


def _to_host_once(t: torch.Tensor) -> torch.Tensor:
    """One device-to-host copy into a cached page-locked buffer (a pageable destination would run at a fraction of the
    PCIe rate and a fresh cudaHostAlloc per call costs more than the copy)."""
    if t.device.type != "cuda":
        return t
    key = (t.numel(), t.dtype)
    buf = _pinned.get(key)
    if buf is None:
        _pinned.clear()
        buf = _pinned[key] = torch.empty(t.numel(), dtype=t.dtype, pin_memory=True)
    host = buf.view(t.shape)
    host.copy_(t)
    return host


def gather_waveforms(local: Sequence[torch.Tensor], device=None, to_host: bool = False) -> List[torch.Tensor]:
    """All-gathers variable-length waveforms [1, T_i] from every rank; returns the global list in rank order.

    ONE collective carries the samples: an `all_gather_into_tensor` of the padded payload [n_max, 1 + L_max] fp32 whose
    column 0 holds each row's length as an int32 bit pattern (exact for any length; no second length exchange), preceded by
    one 16-byte all_gather of (n_items, L_max) so every rank sizes the same payload.  The payload is built and gathered on
    `device` — under NCCL the current CUDA device, so waveforms that are still in HBM never bounce through the host — and
    `to_host=True` brings the whole gathered tensor back with ONE device-to-host copy; the per-utterance results are then
    views of that host buffer (no per-item synchronisation)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [(w.cpu() if to_host else w.clone()) for w in local]
    world = dist.get_world_size()
    if device is None:
        if dist.get_backend() == "nccl":      # CPU tensors cannot travel over NCCL, and an empty shard has no tensor to ask
            device = torch.device("cuda", torch.cuda.current_device())
        else:
            device = local[0].device if len(local) else torch.device("cpu")
    device = torch.device(device)
    lens = [int(w.shape[-1]) for w in local]
    meta = torch.tensor([len(local), max(lens, default=0)], dtype=torch.int64, device=device)
    metas = torch.empty(world * 2, dtype=torch.int64, device=device)      # (concatenated form: accepted by NCCL and gloo alike)
    dist.all_gather_into_tensor(metas, meta)
    metas = metas.view(world, 2).cpu()                                    # the one host sync before the payload is sized
    n_max, l_max = int(metas[:, 0].max()), int(metas[:, 1].max())
    payload = torch.zeros(n_max, l_max + 1, dtype=torch.float32, device=device)
    if local:
        payload[:len(local), 0] = torch.tensor(lens, dtype=torch.int32).view(torch.float32).to(device)
        for i, w in enumerate(local):
            payload[i, 1:1 + lens[i]] = w.reshape(-1).to(device, torch.float32)
    out = torch.empty(world * n_max, l_max + 1, dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(out, payload)
    if to_host:
        out = _to_host_once(out)                                          # one D2H for the whole job
    out = out.view(world, n_max, l_max + 1)
    all_lens = out[:, :, 0].contiguous().view(torch.int32).cpu()          # [world, n_max] (already on the host when to_host)
    res = []
    for r in range(world):
        for i in range(int(metas[r, 0])):
            res.append(out[r, i, 1:1 + int(all_lens[r, i])].reshape(1, -1))
    return res
