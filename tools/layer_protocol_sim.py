#!/usr/bin/env python
"""Discrete-event model of the synchronisation protocol of csrc/gemm_layer.cu (persistent per-layer GEMM kernel).

The kernel cannot be run without a B200, and its failure mode is a hang, so the protocol is checked here first: every CTA's
three roles (TMA producer thread, MMA thread, epilogue warps) are Python generators that mirror the CUDA control flow line by
line; mbarriers (phase parity, arrival counts, transaction bytes), asynchronous TMA loads, asynchronous in-order UMMA
completion with tcgen05.commit, remote mbarrier arrives inside the cluster and the global grid barrier are modelled; a random
scheduler interleaves everything.  Data is modelled as tags, so the run also proves
  * every MMA reads the (phase, tile, k-block) operands it expects (no ring slot is overwritten early, no stale activation:
    the Q tile of phase p carries the version written by phase p-1 on EVERY CTA);
  * every DSMEM read sees the tile its peer parked for the same round (no early overwrite of the park buffer);
  * LayerNorm statistics / residual reads see the versions of the right phase;
  * all threads terminate (no deadlock, no parity aliasing).

    python tools/layer_protocol_sim.py [--seeds 200]
"""
from __future__ import annotations

import argparse
import random

CLUSTER = 8


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _maybe_flip(self):
        if self.pending == 0 and self.tx == 0:
            self.phase ^= 1
            self.pending = self.count

    def arrive(self):
        assert self.pending > 0, "more arrivals than the barrier expects in one phase"
        self.pending -= 1
        self._maybe_flip()

    def expect_tx(self, nbytes):          # mbarrier.arrive.expect_tx
        self.tx += nbytes
        self.arrive()

    def complete_tx(self, nbytes):
        self.tx -= nbytes                 # may go negative: bytes can land before the matching expect_tx (PTX: tx-count is signed);
        self._maybe_flip()                # the phase still needs the pending arrival that comes with the expect_tx

    def passed(self, parity):             # try_wait.parity: the phase with this parity has completed
        return self.phase != parity


def geo(p, n_clusters, n_tiles, cluster, rank):
    S = 4 if p & 1 else 8
    grp, s = divmod(rank, S)
    tpr = n_clusters * (CLUSTER // S)
    off = grp * n_clusters + cluster
    n_act = (n_tiles - off + tpr - 1) // tpr if off < n_tiles else 0
    return S, grp * S, s, n_act, tpr, off


class Sim:
    def __init__(self, n_clusters, n_tiles, kbps, M, stages, n_phases, rng):
        self.ncl, self.n_tiles, self.kbps, self.M, self.ST, self.NP, self.rng = n_clusters, n_tiles, kbps, M, stages, n_phases, rng
        self.ncta = n_clusters * CLUSTER
        self.gcount, self.ggen = 0, 0
        # global tensors: version = number of phases that have (fully) rewritten them; per-element writers tracked coarsely
        self.phase_done = [0] * n_phases                 # CTAs that finished their stores of phase p (before the grid arrive)
        self.ctas = [self.make_cta(i) for i in range(self.ncta)]
        self.async_ops = []                               # pending TMA loads: (fn)
        self.threads = []
        for c in self.ctas:
            self.threads += [self.producer(c), self.mma(c), self.umma_engine(c), self.epilogue(c)]
        self.launch_gen = 0
        self.n_carried = 0                                # weight stages requested one phase ahead (short phases)

    def make_cta(self, i):
        c = type("CTA", (), {})()
        c.id, c.cluster, c.rank = i, i // CLUSTER, i % CLUSTER
        c.full = [MBar(1) for _ in range(self.ST)]
        c.empty = [MBar(1) for _ in range(self.ST)]
        c.tfull = [MBar(1), MBar(1)]
        c.tempty = [MBar(4), MBar(4)]
        c.parked = {8: MBar(8), 4: MBar(4)}
        c.cons = {8: MBar(32), 4: MBar(16)}
        c.slotP = [None] * self.ST
        c.slotQ = [None] * self.ST
        c.acc = [None, None]                              # TMEM accumulators: list of consumed operand tags
        c.park = None
        c.umma_q = []                                     # in-order queue of issued MMAs / commits
        c.done = False
        c.gen0 = None
        return c

    def peer(self, c, rank):
        return self.ctas[c.cluster * CLUSTER + rank]

    # ---- roles -------------------------------------------------------------------------------------------------------
    def producer(self, c):
        it, gen0, carried = 0, None, 0
        for p in range(self.NP):
            S, gbase, s, n_act, tpr, off = geo(p, self.ncl, self.n_tiles[p], c.cluster, c.rank)
            kbps = self.kbps[p]
            total = n_act * kbps
            npre = min(total, self.ST)
            kb0 = s * kbps

            def tag(i, p=p, kbps=kbps, tpr=tpr, off=off, kb0=kb0):
                return (p, (i // kbps) * tpr + off, kb0 + i % kbps)
            for i in range(carried, npre):
                g = it + i
                slot = g % self.ST
                yield lambda slot=slot, g=g: c.empty[slot].passed(((g // self.ST) & 1) ^ 1)
                c.full[slot].expect_tx(2)
                self.tma(c, slot, "P", tag(i))
            carried = 0
            if total < self.ST and p + 1 < self.NP:           # next phase's first weight tiles into the stages this phase leaves free
                Sn, gbn, sn, n_actn, tprn, offn = geo(p + 1, self.ncl, self.n_tiles[p + 1], c.cluster, c.rank)
                kn = self.kbps[p + 1]
                for i in range(min(n_actn * kn, self.ST - total)):
                    g = it + total + i
                    slot = g % self.ST
                    if not c.empty[slot].passed(((g // self.ST) & 1) ^ 1):
                        break
                    c.full[slot].expect_tx(2)
                    self.tma(c, slot, "P", (p + 1, (i // kn) * tprn + offn, sn * kn + i % kn))
                    carried = i + 1
                    self.n_carried += 1
                    yield lambda: True
            if p > 0:
                if p == 1:                                # base generation: read once per CTA, by the arriving (epilogue) thread
                    yield lambda: c.gen0 is not None
                    gen0 = c.gen0
                yield lambda p=p: self.ggen - (gen0 + p) >= 0
            for i in range(npre):
                slot = (it + i) % self.ST
                self.tma(c, slot, "Q", tag(i))
            for i in range(npre, total):
                g = it + i
                slot = g % self.ST
                yield lambda slot=slot, g=g: c.empty[slot].passed(((g // self.ST) & 1) ^ 1)
                c.full[slot].expect_tx(2)
                self.tma(c, slot, "P", tag(i))
                self.tma(c, slot, "Q", tag(i))
            it += total

    def tma(self, c, slot, which, tag):
        p = tag[0]

        def complete():
            if which == "P":
                c.slotP[slot] = tag
            else:
                # the activation operand of phase p is what phase p-1 wrote on ALL CTAs
                if p > 0:
                    assert self.phase_done[p - 1] == self.ncta, f"Q tile of phase {p} loaded before phase {p - 1} finished everywhere"
                c.slotQ[slot] = tag
            c.full[slot].complete_tx(1)
        self.async_ops.append(complete)

    def mma(self, c):
        it, j = 0, 0
        for p in range(self.NP):
            S, gbase, s, n_act, tpr, off = geo(p, self.ncl, self.n_tiles[p], c.cluster, c.rank)
            kbps = self.kbps[p]
            for a in range(n_act):
                buf = j & 1
                yield lambda buf=buf, j=j: c.tempty[buf].passed(((j >> 1) & 1) ^ 1)
                for kb in range(kbps):
                    slot = it % self.ST
                    yield lambda slot=slot, it=it: c.full[slot].passed((it // self.ST) & 1)
                    want = (p, a * tpr + off, s * kbps + kb)
                    c.umma_q.append(("mma", slot, buf, want, kb == 0))
                    c.umma_q.append(("commit", c.empty[slot]))
                    it += 1
                c.umma_q.append(("commit", c.tfull[buf]))
                j += 1

    def umma_engine(self, c):
        """the tensor core executes issued MMAs asynchronously and in order; commits arrive after everything before them"""
        while True:
            yield lambda: bool(c.umma_q) or c.done
            if not c.umma_q:
                return
            op = c.umma_q.pop(0)
            if op[0] == "mma":
                _, slot, buf, want, first = op
                assert c.slotP[slot] == want and c.slotQ[slot] == want, f"CTA {c.id}: MMA expected {want}, ring holds {c.slotP[slot]} / {c.slotQ[slot]}"
                if first:
                    c.acc[buf] = []
                c.acc[buf].append(want)
            else:
                op[1].arrive()

    def epilogue(self, c):
        """the four epilogue warps advance together here (they are tied by bar.sync at every step that matters); the per-warp
        consumed-arrivals are issued as 4 separate arrivals"""
        j, n = 0, {8: 0, 4: 0}
        pend = None
        gen0 = c.gen0 = self.ggen
        for p in range(self.NP):
            S, gbase, s, n_act, tpr, off = geo(p, self.ncl, self.n_tiles[p], c.cluster, c.rank)
            kbps = self.kbps[p]
            if p > 0:
                yield lambda p=p: self.ggen - (gen0 + p) >= 0
                assert self.phase_done[p - 1] == self.ncta          # LayerNorm statistics / residual of the previous phase
            for a in range(n_act):
                tile = a * tpr + off
                buf = j & 1
                yield lambda buf=buf, j=j: c.tfull[buf].passed((j >> 1) & 1)
                if pend is not None:
                    bar, par = pend
                    yield lambda bar=bar, par=par: bar.passed(par)
                    pend = None
                assert c.acc[buf] == [(p, tile, s * kbps + kb) for kb in range(kbps)], f"CTA {c.id}: accumulator holds {c.acc[buf]}"
                c.park = (p, tile)
                for _ in range(4):
                    c.tempty[buf].arrive()
                yield lambda: True                                  # (bar.sync) let others run
                for l in range(S):
                    self.peer(c, gbase + l).parked[S].arrive()
                yield lambda S=S: c.parked[S].passed(n[S] & 1)
                for w in range(4):                                  # each warp reads its rows from every peer, then arrives
                    for l in range(S):
                        pr = self.peer(c, gbase + l)
                        assert pr.park == (p, tile), f"CTA {c.id} read peer {pr.id}: park holds {pr.park}, wanted {(p, tile)}"
                    yield lambda: True
                    for l in range(S):
                        self.peer(c, gbase + l).cons[S].arrive()
                pend = (c.cons[S], n[S] & 1)
                n[S] += 1
                j += 1
            if p + 1 < self.NP:
                self.phase_done[p] += 1
                yield lambda: True
                self.gcount += 1
                if self.gcount == self.ncta:
                    self.gcount = 0
                    self.ggen += 1
            else:
                self.phase_done[p] += 1
        if pend is not None:
            bar, par = pend
            yield lambda: bar.passed(par)
        c.done = True

    # ---- scheduler -----------------------------------------------------------------------------------------------------
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
            assert steps < 5_000_000, "simulation did not terminate"
            choices = [i for i, (t, pred) in enumerate(live) if pred()]
            n_async = len(self.async_ops)
            if not choices and not n_async:
                stuck = [(getattr(t, "__name__", "?")) for t, _ in live]
                raise AssertionError(f"DEADLOCK with {len(live)} blocked threads: {stuck[:8]}")
            k = self.rng.randrange(len(choices) + n_async)
            if k >= len(choices):                                  # complete a random outstanding TMA load
                self.async_ops.pop(self.rng.randrange(n_async))()
                continue
            i = choices[k]
            try:
                live[i][1] = next(live[i][0])
            except StopIteration:
                live.pop(i)
        assert all(c.done for c in self.ctas)
        assert all(d == self.ncta for d in self.phase_done)
        return steps


def check(seeds=60, verbose=False, big=True):
    shapes = [
        # n_clusters, (D, F), M, stages, phases
        (2, (1024, 4096), 64, 7, 4),       # several rounds per phase, ring shorter than a tile's k-blocks
        (2, (512, 2048), 2, 3, 4),         # idle clusters in the residual phases, tiny tiles
        (3, (1536, 6144), 33, 4, 4),       # cluster count that does not divide the tile counts
        (2, (1024, 4096), 16, 8, 3),       # last layer: no QKV phase
        (16, (2048, 8192), 64, 7, 4),      # the 830M geometry
    ]
    total = 0
    for ncl, (D, F), M, st, nph in shapes:
        if ncl >= 16 and not big:
            continue
        n_tiles = [D // 128, F // 128, D // 128, 3 * D // 128][:nph]
        kbps = [D // 64 // 8, D // 64 // 4, F // 64 // 8, D // 64 // 4][:nph]
        n = seeds if ncl < 16 else max(2, seeds // 20)
        carried = 0
        for seed in range(n):
            sim = Sim(ncl, n_tiles, kbps, M, st, nph, random.Random(seed))
            steps = sim.run()
            carried += sim.n_carried
            total += 1
        if verbose:
            print(f"n_clusters {ncl:>2}  D {D} F {F} M {M} stages {st} phases {nph}: {n} random schedules OK ({steps} events in the last, {carried} stages requested a phase ahead)")
    return total


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=100)
    a = ap.parse_args()
    print(check(a.seeds, verbose=True), "schedules, no deadlock, every operand / partial tile / version as expected")
