"""In-kernel timeline of the decode chain (there is no nsys in this image).

`ssrb_debug_timeline` arms a device buffer; thread 0 of every CTA of the decode-chain kernels then records
{kernel id, CTA id, entry, after griddepcontrol.wait, exit, aux0..4} from %globaltimer.  Because the kernels are launched
as a programmatic-dependent (PDL) chain inside a CUDA graph, the interesting quantity per kernel is the critical-path
segment  max(CTA exit) - min(dependency resolved): the time between its predecessor finishing and itself finishing.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np
import torch

from . import _lib

KERNELS = {1: "embed", 2: "layernorm", 3: "gemm", 4: "attention", 5: "sample"}
REC = 10


def capture(model, n_iters: int = 2, cap: int = 400000) -> np.ndarray:
    """Runs `n_iters` decode iterations of the currently open batch of `model` (lm.SSR_Speech) with the timeline armed.
    Returns the records [n, 10] (int64, ns)."""
    lib, st = _lib.load(), _lib.stream_ptr()
    buf = torch.zeros(cap * REC, dtype=torch.int64, device=model._device)
    idx = torch.zeros(1, dtype=torch.int32, device=model._device)
    torch.cuda.synchronize()
    _lib.check(lib.ssrb_debug_timeline(C.c_void_p(buf.data_ptr()), C.c_void_p(idx.data_ptr()), cap), "timeline")
    try:
        _lib.check(lib.ssrb_lm_decode(model._h, int(n_iters), st), "decode")
        torch.cuda.synchronize()
    finally:
        _lib.check(lib.ssrb_debug_timeline(None, None, 0), "timeline")
    n = min(int(idx.item()), cap)
    return buf[:n * REC].view(n, REC).cpu().numpy()


def launches(rec: np.ndarray):
    """Splits the records into kernel launches: same kernel id, dependency-resolved times within 3 us of each other."""
    # A kernel's griddepcontrol.wait returns only after its predecessor has completely finished, so in dependency-resolved
    # order the records of one launch are contiguous.  Split on kernel-id change; two back-to-back launches of the same
    # kernel (FFN1 -> FFN2) are split where a record's dependency time is past every exit seen so far in the group.
    key = np.where(rec[:, 3] > 0, rec[:, 3], rec[:, 2])
    order = np.argsort(key, kind="stable")
    rec, key = rec[order], key[order]
    out, start, max_exit = [], 0, 0
    for i in range(len(rec)):
        if i > start and (rec[i, 0] != rec[start, 0] or (rec[i, 0] == 3 and max_exit > 0 and key[i] >= max_exit)):
            out.append((int(rec[start, 0]), rec[start:i]))
            start, max_exit = i, 0
        max_exit = max(max_exit, int(rec[i, 4]))
    if len(rec):
        out.append((int(rec[start, 0]), rec[start:]))
    return out


def critical_path(rec: np.ndarray, n_iters: int) -> Dict[str, float]:
    """Per-class critical-path microseconds per iteration: sum over launches of max(exit) - min(dependency resolved)."""
    tot = {v: 0.0 for v in KERNELS.values()}
    cnt = {v: 0 for v in KERNELS.values()}
    t_first, t_last = None, None
    for kid, r in launches(rec):
        dep = np.where(r[:, 3] > 0, r[:, 3], r[:, 2]).min()
        ends = r[:, 4][r[:, 4] > 0]
        if len(ends) == 0:
            continue
        end = ends.max()
        tot[KERNELS[kid]] += (end - dep) / 1e3
        cnt[KERNELS[kid]] += 1
        t_first = dep if t_first is None else min(t_first, dep)
        t_last = end if t_last is None else max(t_last, end)
    out = {f"{k}_us": v / n_iters for k, v in tot.items()}
    out.update({f"{k}_launches": cnt[k] / n_iters for k in cnt})
    out["wall_us"] = (t_last - t_first) / 1e3 / n_iters if t_first is not None else None
    return out
