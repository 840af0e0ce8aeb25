"""Host-side check of the synchronisation protocol of the persistent codec kernels (csrc/resblock_tc.cu, and csrc/conv_tc.cu as its
sub-case): tools/resblock_protocol_sim.py mirrors the producer / MMA issuer / tensor engine / 8 epilogue warps, their mbarrier
parities and running counters, and is run under random schedules for the four channel counts' tilings (incl. the C = 512 case whose
two accumulators share TMEM columns).  Every wait of the protocol is also REMOVED once: the model must then report a hazard, i.e. the
checks are able to see the races the waits prevent.  No GPU involved."""
import importlib.util
import os
import random
import sys

import pytest

from conftest import ROOT


def _sim():
    tools = os.path.join(ROOT, "tools")
    if tools not in sys.path:
        sys.path.insert(0, tools)
    spec = importlib.util.spec_from_file_location("resblock_protocol_sim", os.path.join(tools, "resblock_protocol_sim.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_protocol_has_no_deadlock_and_no_hazard():
    m = _sim()
    assert m.check(seeds=12) == 12 * len(m.CONFIGS)


@pytest.mark.parametrize("mutation", ["no_empty_wait", "no_acc2_empty_wait", "no_h_wait", "no_acc2_full_wait"])
def test_every_wait_is_load_bearing(mutation):
    m = _sim()
    caught = 0
    for cfg in m.CONFIGS:
        for seed in range(12):
            try:
                m.Sim(*cfg, random.Random(seed), mutate=mutation).run()
            except AssertionError:
                caught += 1
    assert caught > 0, f"removing the wait '{mutation}' went unnoticed"
