"""Host-side check of the synchronisation protocol of the experimental persistent layer kernel (csrc/gemm_layer.cu): the
discrete-event model in tools/layer_protocol_sim.py mirrors the kernel's producer / MMA / epilogue control flow, mbarrier
parities, remote arrives and grid barrier, and is run under random schedules.  No GPU involved."""
import importlib.util
import os
import random

import pytest

from conftest import ROOT


def _sim(name="layer_protocol_sim"):
    import sys
    tools = os.path.join(ROOT, "tools")
    if tools not in sys.path:
        sys.path.insert(0, tools)                          # flat2_protocol_sim imports the barrier model of layer_protocol_sim
    spec = importlib.util.spec_from_file_location(name, os.path.join(tools, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_protocol_terminates_and_reads_the_right_tiles():
    assert _sim().check(seeds=8, big=False) == 32


def test_short_phase_prefetches_the_next_phase_weights():
    m = _sim()
    D, F = 512, 2048
    sim = m.Sim(2, [D // 128, F // 128, D // 128, 3 * D // 128], [D // 64 // 8, D // 64 // 4, F // 64 // 8, D // 64 // 4], 2, 3, 4,
                random.Random(0))
    sim.run()
    assert sim.n_carried == 16            # every CTA: out-proj fills 2 of 3 stages, the third takes FFN1's first weight tile


def test_work_decomposition_covers_every_tile_slice_once():
    geo = _sim().geo
    for D, F, ncl in [(2048, 8192, 16), (512, 2048, 16), (1536, 6144, 11), (1024, 4096, 8)]:
        n_tiles = [D // 128, F // 128, D // 128, 3 * D // 128]
        for p in range(4):
            S = 4 if p & 1 else 8
            seen = set()
            for c in range(ncl):
                for r in range(8):
                    _, _, s, n_act, tpr, off = geo(p, ncl, n_tiles[p], c, r)
                    for a in range(n_act):
                        key = (a * tpr + off, s)
                        assert key not in seen and key[0] < n_tiles[p]
                        seen.add(key)
            assert len(seen) == n_tiles[p] * S


def test_model_detects_a_missing_wait():
    """The model is only worth something if it fails on a broken protocol: drop the wait that protects the park buffer."""
    src = open(os.path.join(ROOT, "tools", "layer_protocol_sim.py")).read()
    broken = src.replace("                    yield lambda bar=bar, par=par: bar.passed(par)\n                    pend = None", "                    pend = None")
    assert broken != src
    ns = {"__name__": "broken_sim"}
    exec(compile(broken, "broken_sim", "exec"), ns)
    with pytest.raises(AssertionError):
        ns["check"](seeds=6, big=False)


def test_cta_pair_gemm_protocol():
    """csrc/gemm_flat2.cu: both CTAs' TMA loads complete on the leader's barrier (possibly before the leader armed it), multicast
    commits, both epilogues release the leader's accumulator barrier."""
    assert _sim("flat2_protocol_sim").check(seeds=15) == 75


def test_cta_pair_model_detects_a_wrong_arrival_count():
    src = open(os.path.join(ROOT, "tools", "flat2_protocol_sim.py")).read()
    broken = src.replace("c.tempty = [MBar(8), MBar(8)]", "c.tempty = [MBar(4), MBar(4)]")      # forgets the peer's epilogue warps
    assert broken != src
    _sim()                                                   # puts tools/ on sys.path for the import inside the model
    ns = {"__name__": "broken_flat2"}
    exec(compile(broken, "broken_flat2", "exec"), ns)
    with pytest.raises(AssertionError):
        ns["check"](seeds=10)
