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
    out = []
    for kid in KERNELS:
        r = rec[rec[:, 0] == kid]
        if len(r) == 0:
            continue
        key = np.where(r[:, 3] > 0, r[:, 3], r[:, 2])
        r = r[np.argsort(key, kind="stable")]
        key = np.sort(key, kind="stable")
        cuts = np.where(np.diff(key) > 3000)[0] + 1
        for idxs in np.split(np.arange(len(r)), cuts):
            out.append((kid, r[idxs]))
    out.sort(key=lambda kr: float(np.where(kr[1][:, 3] > 0, kr[1][:, 3], kr[1][:, 2]).min()))
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
