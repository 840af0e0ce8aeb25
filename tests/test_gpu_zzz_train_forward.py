"""GPU parity of the training forward / loss (SSR_Speech.forward, models/ssr.py:280-379) against the golden recorded
from the UNMODIFIED reference (tests/golden/lm_train_forward.npz, oracle/gen_golden.py train) and against the oracle.
Tolerances: fp32 parity mode — loss within 1e-4 relative of the reference, token count exact, top-10 accuracy within one flipped
position per codebook; bf16 production mode — loss within 2e-2 relative of the bf16-storage oracle's.  The file sorts last so
that this newest row cannot hide older tests behind `pytest -x`."""
import os

import numpy as np
import pytest
import torch

from lm_oracle import LMOracle
from ssr_speech_b200.config import cfg_tiny
from ssr_speech_b200.lm import SSR_Speech
from ssr_speech_b200.synth import make_lm_state_dict

pytestmark = pytest.mark.gpu


def _model(precision, tag, g):
    cfg = cfg_tiny()
    ns = cfg.to_namespace()
    ns.predict_mask_token, ns.predict_all = int(g[f"{tag}_predict_mask_token"]), int(g[f"{tag}_predict_all"])
    cw = g[f"{tag}_codebook_weight"].tolist()
    ns.codebook_weight = None if tag == "all" else str(cw)
    m = SSR_Speech(ns, precision=precision)
    m.load_state_dict(make_lm_state_dict(cfg, seed=int(g["weights_seed"])))
    return cfg, m.to("cuda").eval()


def _batch(g):
    return {"x": torch.from_numpy(g["x"]), "x_lens": torch.from_numpy(g["x_lens"]), "y": torch.from_numpy(g["y"]),
            "y_lens": torch.from_numpy(g["y_lens"])}


@pytest.mark.parametrize("tag", ["default", "all"])
def test_fp32_training_forward_matches_the_reference(gold_dir, tag):
    g = np.load(os.path.join(gold_dir, "lm_train_forward.npz"))
    cfg, m = _model("fp32", tag, g)
    out = m.forward(_batch(g))
    assert int(out["effective_ntoken"]) == int(g[f"{tag}_ntoken"])
    want = float(g[f"{tag}_loss"])
    assert abs(float(out["loss"]) - want) <= 1e-4 * abs(want), (float(out["loss"]), want)
    # one flipped top-10 membership moves a codebook's term by ntokens_k / n_loss_k <= ~1.3 here
    by = np.asarray([float(v) for v in out["top10acc_by_codebook"]])
    assert np.all(np.abs(by - g[f"{tag}_top10acc_by_codebook"]) <= 1.5), (by, g[f"{tag}_top10acc_by_codebook"])
    assert abs(float(out["top10acc"]) - float(g[f"{tag}_top10acc"])) <= 3.0
    assert out["loss"].is_cuda and len(out["top10acc_by_codebook"]) == cfg.n_codebooks


def test_bf16_training_forward_is_close_to_the_bf16_storage_oracle(gold_dir):
    g = np.load(os.path.join(gold_dir, "lm_train_forward.npz"))
    cfg, m = _model("bf16", "default", g)
    out = m.forward(_batch(g))
    o = LMOracle(cfg, make_lm_state_dict(cfg, seed=int(g["weights_seed"])), round_weights_to_bf16=True, round_acts_to_bf16=True)
    b = _batch(g)
    want = o.forward_loss(b["x"], b["x_lens"], b["y"], b["y_lens"], predict_mask_token=True, predict_all=False,
                          codebook_weight=g["default_codebook_weight"].tolist())
    assert int(out["effective_ntoken"]) == want["effective_ntoken"]
    assert abs(float(out["loss"]) - want["loss"]) <= 2e-2 * abs(want["loss"]), (float(out["loss"]), want["loss"])
    assert abs(float(out["top10acc"]) - want["top10acc"]) <= 0.05 * want["effective_ntoken"]


def test_training_forward_rejects_bad_batches():
    cfg = cfg_tiny()
    m = SSR_Speech(cfg.to_namespace(), precision="fp32")
    m.load_state_dict(make_lm_state_dict(cfg, seed=7))
    m.to("cuda")
    y = torch.zeros(1, cfg.n_codebooks, 6, dtype=torch.long)
    assert m.forward({"x": torch.zeros(0, 3, dtype=torch.long), "x_lens": torch.zeros(0, dtype=torch.long), "y": y[:0], "y_lens": torch.zeros(0, dtype=torch.long)}) is None
    with pytest.raises(AssertionError):
        m.forward({"x": torch.zeros(1, 3, dtype=torch.long), "x_lens": torch.tensor([3]), "y": y[:, :2], "y_lens": torch.tensor([6])})
    with pytest.raises(IndexError):
        m.forward({"x": torch.full((1, 3), cfg.text_vocab_size + 5), "x_lens": torch.tensor([3]), "y": y, "y_lens": torch.tensor([6])})
