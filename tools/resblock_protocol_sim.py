#!/usr/bin/env python
"""Discrete-event model of the persistent codec kernels' protocol (csrc/resblock_tc.cu; csrc/conv_tc.cu is the sub-case without the
second GEMM), same style as layer_protocol_sim.py / flat2_protocol_sim.py: one CTA = a TMA producer, an MMA issuer, the in-order
tensor engine and 8 epilogue warps as coroutines; mbarriers with parity + transaction bytes; data as tags.  Random schedules.

Per tile t the kernel runs   loads -> GEMM 1 -> acc1 -> epilogue 1 (h to shared memory) -> GEMM 2 -> acc2 -> epilogue 2 (stores)
and overlaps tiles: the producer and GEMM 1 of tile t+1 run under epilogue 2 of tile t.  What the model checks on every schedule:
  * no deadlock, every tile stored exactly once by every epilogue warp;
  * a ring slot is only overwritten after the MMA that reads it has EXECUTED (tcgen05.commit on the empty barrier);
  * every MMA reads the k-block of its own (tile, index); GEMM 2 reads an h that all 8 warps have written for this tile;
  * acc1 / acc2 are only overwritten (first MMA of a GEMM, accumulate = 0) after all 8 warps have read the previous tile's values
    — with the two accumulators in separate TMEM columns (C <= 256) and with acc1 aliasing acc2 (C = 512);
  * h is only rewritten (epilogue 1 of tile t+1) after GEMM 2 of tile t has executed.
`mutate=` removes one wait at a time; every mutation must be caught (tests/test_codec_protocol.py).

    python tools/resblock_protocol_sim.py [--seeds 200]
"""
from __future__ import annotations

import argparse
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from layer_protocol_sim import MBar  # noqa: E402

EPI = 8


class Sim:
    def __init__(self, n_ctas, n_tiles, nk1, kb2, stages, alias, rng, mutate=None):
        self.G, self.total, self.NK1, self.KB2, self.ST, self.alias, self.rng, self.mut = n_ctas, n_tiles, nk1, kb2, stages, alias, rng, mutate
        self.async_ops = []
        self.ctas = []
        self.threads = []
        for i in range(n_ctas):
            c = type("CTA", (), {})()
            c.id = i
            c.full = [MBar(1) for _ in range(stages)]
            c.empty = [MBar(1) for _ in range(stages)]
            c.acc1_full, c.acc2_full, c.h_ready, c.acc2_empty = MBar(1), MBar(1), MBar(EPI), MBar(EPI)
            c.slot = [None] * stages                    # {"key": (tile, i), "read": bool}
            c.acc1 = {"tile": None, "kbs": [], "reads": EPI}
            c.acc2 = {"tile": None, "kbs": [], "reads": EPI}
            c.h = {"tile": None, "writers": 0, "g2_left": 0}
            c.umma_q = []
            c.stored = [[] for _ in range(EPI)]
            c.mma_done = False
            self.ctas.append(c)
            self.threads += [self.producer(c), self.mma(c), self.engine(c)] + [self.epilogue(c, w) for w in range(EPI)]

    def tiles(self, c):
        return range(c.id, self.total, self.G)

    # ---- TMA producer (warp 0) ----
    def producer(self, c):
        it = 0
        for t in self.tiles(c):
            for i in range(self.NK1 + self.KB2):
                s = it % self.ST
                if self.mut != "no_empty_wait":
                    yield lambda s=s, it=it: c.empty[s].passed(((it // self.ST) & 1) ^ 1)
                else:
                    yield lambda: True
                c.full[s].expect_tx(1)

                def land(s=s, t=t, i=i):
                    old = c.slot[s]
                    assert old is None or old["read"], f"CTA {c.id}: ring slot {s} overwritten before the MMA of {old['key']} executed"
                    c.slot[s] = {"key": (t, i), "read": False}
                    c.full[s].complete_tx(1)
                self.async_ops.append(land)
                it += 1

    # ---- MMA issuer (warp 1): issues in program order, the engine executes in that order ----
    def mma(self, c):
        it = 0
        for tl, t in enumerate(self.tiles(c)):
            tph = tl & 1
            if self.alias and self.mut != "no_acc2_empty_wait":
                yield lambda tph=tph: c.acc2_empty.passed(tph ^ 1)
            for i in range(self.NK1):
                s = it % self.ST
                yield lambda s=s, it=it: c.full[s].passed((it // self.ST) & 1)
                c.umma_q.append(("g1", s, t, i))
                c.umma_q.append(("commit", c.empty[s]))
                it += 1
            c.umma_q.append(("commit", c.acc1_full))
            if self.mut != "no_h_wait":
                yield lambda tph=tph: c.h_ready.passed(tph)
            if not self.alias and self.mut != "no_acc2_empty_wait":
                yield lambda tph=tph: c.acc2_empty.passed(tph ^ 1)
            for j in range(self.KB2):
                s = it % self.ST
                yield lambda s=s, it=it: c.full[s].passed((it // self.ST) & 1)
                c.umma_q.append(("g2", s, t, j))
                c.umma_q.append(("commit", c.empty[s]))
                it += 1
            c.umma_q.append(("commit", c.acc2_full))
        c.mma_done = True

    # ---- the tensor engine: executes the queue in order, one operation per scheduling step ----
    def engine(self, c):
        while True:
            yield lambda: bool(c.umma_q) or c.mma_done
            if not c.umma_q:
                return
            op = c.umma_q.pop(0)
            if op[0] == "commit":
                op[1].arrive()
                continue
            kind, s, t, i = op
            sl = c.slot[s]
            want = (t, i) if kind == "g1" else (t, self.NK1 + i)
            assert sl is not None and sl["key"] == want, f"CTA {c.id}: {kind} of {want} read ring slot holding {sl}"
            sl["read"] = True
            if kind == "g1":
                if i == 0:
                    assert c.acc1["reads"] == EPI, f"CTA {c.id}: acc1 of tile {c.acc1['tile']} overwritten after {c.acc1['reads']} of {EPI} reads"
                    if self.alias:
                        assert c.acc2["reads"] == EPI, f"CTA {c.id}: acc1 aliases acc2 of tile {c.acc2['tile']}, read by {c.acc2['reads']} of {EPI} warps"
                    c.acc1 = {"tile": t, "kbs": [], "reads": 0}
                assert c.acc1["tile"] == t
                c.acc1["kbs"].append(i)
            else:
                assert c.h["tile"] == t and c.h["writers"] == EPI, f"CTA {c.id}: GEMM 2 of tile {t} read h = {c.h}"
                if i == 0:
                    assert c.acc2["reads"] == EPI, f"CTA {c.id}: acc2 of tile {c.acc2['tile']} overwritten after {c.acc2['reads']} of {EPI} reads"
                    c.acc2 = {"tile": t, "kbs": [], "reads": 0}
                assert c.acc2["tile"] == t
                c.acc2["kbs"].append(i)
                c.h["g2_left"] -= 1

    # ---- one epilogue warp ----
    def epilogue(self, c, w):
        for tl, t in enumerate(self.tiles(c)):
            tph = tl & 1
            yield lambda tph=tph: c.acc1_full.passed(tph)
            assert c.acc1["tile"] == t and c.acc1["kbs"] == list(range(self.NK1)), f"CTA {c.id} warp {w}: acc1 = {c.acc1} for tile {t}"
            c.acc1["reads"] += 1
            yield lambda: True
            # epilogue 1: this warp's part of h
            if c.h["tile"] != t:
                assert c.h["g2_left"] == 0, f"CTA {c.id} warp {w}: h of tile {c.h['tile']} rewritten with {c.h['g2_left']} GEMM-2 k-blocks outstanding"
                c.h = {"tile": t, "writers": 0, "g2_left": self.KB2}
            c.h["writers"] += 1
            yield lambda: True
            c.h_ready.arrive()
            if self.mut != "no_acc2_full_wait":
                yield lambda tph=tph: c.acc2_full.passed(tph)
            else:
                yield lambda: True
            assert c.acc2["tile"] == t and c.acc2["kbs"] == list(range(self.KB2)), f"CTA {c.id} warp {w}: acc2 = {c.acc2} for tile {t}"
            c.acc2["reads"] += 1
            yield lambda: True
            c.acc2_empty.arrive()
            yield lambda: True                       # stores
            c.stored[w].append(t)

    def run(self):
        live = []
        for t in self.threads:
            try:
                live.append([t, next(t)])
            except StopIteration:
                pass
        steps = 0
        while live or self.async_ops:
            steps += 1
            assert steps < 5_000_000
            choices = [i for i, (t, pred) in enumerate(live) if pred()]
            na = len(self.async_ops)
            if not choices and not na:
                raise AssertionError(f"DEADLOCK with {len(live)} blocked threads")
            k = self.rng.randrange(len(choices) + na)
            if k >= len(choices):
                self.async_ops.pop(self.rng.randrange(na))()
                continue
            i = choices[k]
            try:
                live[i][1] = next(live[i][0])
            except StopIteration:
                live.pop(i)
        for c in self.ctas:
            for w in range(EPI):
                assert c.stored[w] == list(self.tiles(c)), f"CTA {c.id} warp {w} stored {c.stored[w]}"
        return steps


# (CTAs, tiles, k-blocks of GEMM 1, of GEMM 2, ring stages, acc1 aliases acc2): the four channel counts of resblock_tc.cu + odd shapes
CONFIGS = [(2, 7, 3, 1, 3, False), (2, 7, 6, 1, 2, False), (1, 5, 12, 2, 4, False), (2, 5, 24, 4, 2, True), (3, 3, 3, 1, 3, False),
           (1, 1, 3, 1, 3, False), (1, 6, 5, 3, 2, True)]


def check(seeds=50, verbose=False, mutate=None):
    n = 0
    for cfg in CONFIGS:
        for seed in range(seeds):
            steps = Sim(*cfg, random.Random(seed), mutate=mutate).run()
            n += 1
        if verbose:
            print(f"CTAs {cfg[0]} tiles {cfg[1]} k-blocks {cfg[2]}+{cfg[3]} stages {cfg[4]} alias {cfg[5]}: {seeds} random schedules OK ({steps} events in the last)")
    return n


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=100)
    a = ap.parse_args()
    print(check(a.seeds, verbose=True), "schedules: no deadlock, no slot / accumulator / h hazard, every tile stored once per warp")
