#!/usr/bin/env python
"""Discrete-event model of the CTA-pair prefill GEMM's protocol (csrc/gemm_flat2.cu), same style as layer_protocol_sim.py:
producer / MMA / tensor-core / epilogue coroutines of both CTAs of every pair, mbarriers with parity + transaction bytes, TMA
loads of BOTH CTAs completing on the leader's full barrier, multicast commits releasing ring stages and publishing accumulators
in both CTAs, both epilogues arriving on the leader's accumulator-empty barrier.  Random schedules; data as tags.

    python tools/flat2_protocol_sim.py [--seeds 200]
"""
from __future__ import annotations

import argparse
import random

from layer_protocol_sim import MBar


class Sim:
    def __init__(self, n_pairs, m_tiles, n_tiles, nkb, stages, rng):
        self.NP, self.total, self.n_tiles, self.nkb, self.ST, self.rng = n_pairs, m_tiles * n_tiles, n_tiles, nkb, stages, rng
        self.async_ops = []
        self.ctas = []
        for i in range(2 * n_pairs):
            c = type("CTA", (), {})()
            c.id, c.pair, c.rank = i, i // 2, i % 2
            c.full = [MBar(1) for _ in range(stages)]
            c.empty = [MBar(1) for _ in range(stages)]
            c.tfull = [MBar(1), MBar(1)]
            c.tempty = [MBar(8), MBar(8)]
            c.slot = [None] * stages                 # (tile, kb) staged in this CTA's ring slot
            c.acc = [None, None]
            c.umma_q = []
            c.stored = []
            c.done = False
            self.ctas.append(c)
        self.threads = []
        for c in self.ctas:
            self.threads += [self.producer(c), self.epilogue(c)]
            if c.rank == 0:
                self.threads += [self.mma(c), self.umma_engine(c)]

    def leader(self, c):
        return self.ctas[2 * c.pair]

    def peer(self, c):
        return self.ctas[2 * c.pair + 1]

    def tiles(self, c):
        return range(c.pair, self.total, self.NP)

    def producer(self, c):
        it, L = 0, self.leader(c)
        for t in self.tiles(c):
            for kb in range(self.nkb):
                s = it % self.ST
                yield lambda s=s, it=it: c.empty[s].passed(((it // self.ST) & 1) ^ 1)
                if c.rank == 0:
                    c.full[s].expect_tx(4)           # A + W-half of both CTAs
                for which in ("A", "W"):
                    def complete(s=s, t=t, kb=kb, which=which):
                        cur = c.slot[s] if isinstance(c.slot[s], dict) and c.slot[s].get("key") == (t, kb) else {"key": (t, kb)}
                        cur[which] = True
                        c.slot[s] = cur
                        L.full[s].complete_tx(1)     # completes on the LEADER's barrier (peer bit cleared)
                    self.async_ops.append(complete)
                it += 1

    def mma(self, c):
        it, j, P = 0, 0, self.peer(c)
        for t in self.tiles(c):
            buf = j & 1
            yield lambda buf=buf, j=j: c.tempty[buf].passed(((j >> 1) & 1) ^ 1)
            for kb in range(self.nkb):
                s = it % self.ST
                yield lambda s=s, it=it: c.full[s].passed((it // self.ST) & 1)
                c.umma_q.append(("mma", s, buf, (t, kb), kb == 0))
                c.umma_q.append(("commit", [c.empty[s], P.empty[s]]))          # multicast 0b11
                it += 1
            c.umma_q.append(("commit", [c.tfull[buf], P.tfull[buf]]))
            j += 1

    def umma_engine(self, c):
        P = self.peer(c)
        while True:
            yield lambda: bool(c.umma_q) or (c.done and P.done)
            if not c.umma_q:
                return
            op = c.umma_q.pop(0)
            if op[0] == "mma":
                _, s, buf, want, first = op
                for cta in (c, P):                    # the pair's tensor cores read both CTAs' shared memory
                    got = cta.slot[s]
                    assert isinstance(got, dict) and got.get("key") == want and got.get("A") and got.get("W"), \
                        f"pair {c.pair}: MMA wanted {want}, CTA {cta.id} slot holds {got}"
                    if first:
                        cta.acc[buf] = []
                    cta.acc[buf].append(want)
            else:
                for b in op[1]:
                    b.arrive()

    def epilogue(self, c):
        j, L = 0, self.leader(c)
        for t in self.tiles(c):
            buf = j & 1
            yield lambda buf=buf, j=j: c.tfull[buf].passed((j >> 1) & 1)
            assert c.acc[buf] == [(t, kb) for kb in range(self.nkb)], f"CTA {c.id}: accumulator holds {c.acc[buf]}"
            c.stored.append(t)
            for w in range(4):
                yield lambda: True
                L.tempty[buf].arrive()
            j += 1
        c.done = True

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
        stored = sorted(t for c in self.ctas if c.rank == 0 for t in c.stored)
        assert stored == list(range(self.total)) and all(sorted(c.stored) == sorted(self.leader(c).stored) for c in self.ctas)
        return steps


def check(seeds=50, verbose=False):
    n = 0
    for n_pairs, m_tiles, n_tiles, nkb, st in [(3, 4, 5, 8, 6), (2, 1, 1, 1, 6), (4, 2, 2, 3, 2), (3, 7, 1, 13, 6), (8, 3, 2, 4, 3)]:
        for seed in range(seeds):
            steps = Sim(n_pairs, m_tiles, n_tiles, nkb, st, random.Random(seed)).run()
            n += 1
        if verbose:
            print(f"pairs {n_pairs} tiles {m_tiles}x{n_tiles} k-blocks {nkb} stages {st}: {seeds} random schedules OK ({steps} events in the last)")
    return n


if __name__ == "__main__":
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=100)
    a = ap.parse_args()
    print(check(a.seeds, verbose=True), "schedules: no deadlock, every MMA saw both CTAs' operands of its (tile, k-block), every tile stored once per CTA")
