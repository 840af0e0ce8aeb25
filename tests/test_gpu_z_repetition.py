"""GPU parity on the reference goldens in which the stop_repetition penalty (ssr.py:727-736) actually fires: the greedy roll-outs
of these seeds repeat one codebook-0 token for tens of steps, the fixtures declare those ids as silence tokens, and the
consecutive-repeat counter exceeds stop_repetition (1-3 rescaled steps per case, both with and without CFG rows).  fp32
parity mode: tokens, marks and intervals identical to the unmodified reference.  Kept in its own file, collected
last: these fixtures were recorded after the round's last GPU run."""
import os

import numpy as np
import pytest

from test_gpu_lm import make_model, run_case

pytestmark = pytest.mark.gpu
REP_CASES = ["tts_rep_greedy", "tts_rep_cfg_lowtemp", "tts_rep_cfg_greedy"]


@pytest.fixture(scope="module")
def model_fp32():
    return make_model("fp32")


def penalty_steps(g, eog):
    """Steps whose codebook-0 logit of the previous token is rescaled, recomputed from the reference's tokens."""
    import json
    sil, sr = g["silence"].tolist(), json.loads(str(g["kw"]))["stop_repetition"]
    prev, consec, n = None, 0, 0
    for s0 in g["ref_span_tokens"][:, 0].tolist():
        n += int(sr > 0 and prev in sil and consec > sr)
        if s0 == eog:
            break
        consec = consec + 1 if (s0 in sil and prev is not None and s0 == prev) else 0
        prev = s0
    return n


@pytest.mark.parametrize("name", REP_CASES)
def test_fp32_tokens_match_reference_when_the_repetition_penalty_fires(model_fp32, gold_dir, name):
    g = np.load(os.path.join(gold_dir, f"lm_{name}.npz"))
    assert penalty_steps(g, model_fp32.cfg.eog) >= 1            # the fixture exercises the branch
    res, marks, masks, nmi = run_case(model_fp32, g)
    assert np.array_equal(res[0].cpu().numpy(), g["ref_res"])
    assert np.array_equal(marks[0].numpy(), g["ref_marks"])
    assert np.array_equal(np.asarray(masks), g["ref_masks"]) and np.array_equal(np.asarray(nmi), g["ref_nmi"])
