"""CPU: the oracle restatement reproduces the golden vectors recorded from the UNMODIFIED reference
(oracle/gen_golden.py).  This is what pins the oracle (the reference ships no golden vectors)."""
import json
import os

import numpy as np
import pytest
import torch

from codec_oracle import CodecOracle
from lm_oracle import LMOracle
from ssr_speech_b200 import seq
from ssr_speech_b200.config import CodecConfig, cfg_tiny
from ssr_speech_b200.synth import make_codec_state_dict, make_lm_state_dict

LM_CASES = ["tts_greedy", "edit_cfg_sampled", "edit2_cfg_greedy", "tts_cfg_temp_topk", "edit_head_nokv",
            "edit3_cfg_sampled", "ctx_edit_greedy", "ctx_tts_cfg_sampled",
            "tts_rep_greedy", "tts_rep_cfg_lowtemp", "tts_rep_cfg_greedy"]      # stop_repetition penalty fires (ssr.py:727-736)


@pytest.fixture(scope="module")
def lm_oracle():
    cfg = cfg_tiny()
    return cfg, LMOracle(cfg, make_lm_state_dict(cfg, seed=7))


@pytest.mark.parametrize("name", LM_CASES)
def test_lm_oracle_reproduces_reference_tokens(lm_oracle, gold_dir, name):
    cfg, oracle = lm_oracle
    g = np.load(os.path.join(gold_dir, f"lm_{name}.npz"))
    kw = json.loads(str(g["kw"]))
    kw.pop("kvcache")
    x, y, out_len = g["x"], g["y"].T.copy(), 0
    if kw.pop("aug_context", False) and sum(b - a for a, b in g["mask_interval"].tolist()) < 2 * 50:   # ssr.py:564-593
        x, y, out_len = np.concatenate([g["prompt_x"], x]), np.concatenate([g["prompt"].T, y], 1), g["prompt"].shape[0]
    prep = seq.prepare(cfg, y, g["mask_interval"].tolist(), out_len=out_len)
    uncond = torch.from_numpy(g["uncond_x"]) if kw["aug_text"] else None
    trace = []
    spans = oracle.inference(torch.from_numpy(x), torch.from_numpy(prep.prompt_tokens), prep.num_spans,
                             silence_tokens=g["silence"].tolist(), uncond_x=uncond, noise=torch.from_numpy(g["noise"]),
                             trace=trace, **kw)
    assert [len(s) for s in spans] == g["ref_span_lens"].tolist()
    assert np.array_equal(np.concatenate(spans, 0), g["ref_span_tokens"])
    res, marks, masks, nmi = seq.finalize(cfg, prep, spans)
    assert np.array_equal(res, g["ref_res"]) and np.array_equal(marks, g["ref_marks"])
    assert np.array_equal(np.asarray(masks), g["ref_masks"]) and np.array_equal(np.asarray(nmi), g["ref_nmi"])
    np.testing.assert_allclose(trace[0].raw_logits.numpy(), g["raw_logits_step0"], atol=2e-5)


def test_lm_oracle_teacher_forced_logits(lm_oracle, gold_dir):
    cfg, oracle = lm_oracle
    g = np.load(os.path.join(gold_dir, "lm_teacher_forced.npz"))
    lg = oracle.teacher_forced_logits(torch.from_numpy(g["x"]), torch.from_numpy(g["toks"]))
    np.testing.assert_allclose(lg.numpy(), g["ref_logits"], atol=2e-5)


def test_lm_oracle_incremental_equals_full(lm_oracle, gold_dir):
    """kvcache=1 and kvcache=0 give identical tokens under greedy (SURVEY §4 property i)."""
    cfg, oracle = lm_oracle
    g = np.load(os.path.join(gold_dir, "lm_tts_greedy.npz"))
    kw = json.loads(str(g["kw"]))
    kw.pop("kvcache")
    prep = seq.prepare(cfg, g["y"].T.copy(), g["mask_interval"].tolist())
    a = oracle.inference(torch.from_numpy(g["x"]), torch.from_numpy(prep.prompt_tokens), 1, incremental=True, max_steps=12, **kw)
    b = oracle.inference(torch.from_numpy(g["x"]), torch.from_numpy(prep.prompt_tokens), 1, incremental=False, max_steps=12, **kw)
    assert np.array_equal(a[0], b[0])


@pytest.fixture(scope="module")
def codec_oracle(gold_dir):
    g = np.load(os.path.join(gold_dir, "codec_small.npz"))
    cfg = CodecConfig()
    sd = make_codec_state_dict(cfg, seed=int(g["weights_seed"]), codebook_mu=g["codebook_mu"], codebook_sigma=g["codebook_sigma"])
    return cfg, CodecOracle(cfg, sd), g


def test_codec_oracle_encode(codec_oracle):
    cfg, o, g = codec_oracle
    codes, _, emb = o.encode(torch.from_numpy(g["wav"]))
    np.testing.assert_allclose(emb.numpy(), g["ref_emb"], atol=1e-6)
    assert np.array_equal(codes.numpy(), g["ref_codes"])
    assert np.array_equal(o.rvq_encode(torch.from_numpy(g["ref_emb"])).numpy(), g["ref_codes"])


def test_codec_oracle_decode_and_wmdecode(codec_oracle):
    cfg, o, g = codec_oracle
    codes = torch.from_numpy(g["ref_codes"])
    np.testing.assert_allclose(o.decode(codes).numpy(), g["ref_dec"], atol=1e-6)
    wm, ml = o.wmdecode(codes, torch.from_numpy(g["marks"]), torch.from_numpy(g["wav"]))
    np.testing.assert_allclose(wm.numpy(), g["ref_wm"], atol=2e-6)
    np.testing.assert_allclose(ml.numpy(), g["ref_mark_logits"], atol=2e-6)


def test_codec_oracle_demo_wav_roundtrip(gold_dir):
    """BASELINE configs[0]: the WHOLE demo/84_121550_000074_000000.wav (126 880 samples -> 127 040 -> 397 frames):
    encode -> RVQ -> decode, and wmdecode with marks[100:200] = 1 (wmencodec.py:324-375), against the unmodified reference."""
    g = np.load(os.path.join(gold_dir, "codec_demo.npz"))
    cfg = CodecConfig()
    sd = make_codec_state_dict(cfg, seed=int(g["weights_seed"]), codebook_mu=g["codebook_mu"], codebook_sigma=g["codebook_sigma"])
    o = CodecOracle(cfg, sd)
    assert g["wav"].shape == (1, 1, 127040) and int(g["n_samples"]) == 126880
    codes, _, emb = o.encode(torch.from_numpy(g["wav"]))
    assert codes.shape == (1, 4, 397)
    np.testing.assert_allclose(emb.numpy(), g["ref_emb"], atol=1e-6)
    assert np.array_equal(codes.numpy(), g["ref_codes"])
    np.testing.assert_allclose(o.decode(codes).numpy(), g["ref_dec"], atol=1e-6)
    wm, ml = o.wmdecode(codes, torch.from_numpy(g["marks"]), torch.from_numpy(g["wav"]))
    np.testing.assert_allclose(wm.numpy(), g["ref_wm"], atol=2e-6)
    np.testing.assert_allclose(ml.numpy(), g["ref_mark_logits"], atol=2e-6)


@pytest.mark.parametrize("name", ["tts_rep_greedy", "tts_rep_cfg_lowtemp", "tts_rep_cfg_greedy"])
def test_repetition_fixtures_exercise_the_penalty_branch(gold_dir, name):
    """ssr.py:727-736 only acts when a silence token repeats more than stop_repetition times: check on the reference's own
    tokens that these fixtures reach it (the other fixtures never do), so parity on them covers the branch."""
    from test_gpu_z_repetition import penalty_steps
    g = np.load(os.path.join(gold_dir, f"lm_{name}.npz"))
    assert penalty_steps(g, cfg_tiny().eog) >= 1
    for other in ("tts_greedy", "edit_cfg_sampled", "tts_cfg_temp_topk"):
        assert penalty_steps(np.load(os.path.join(gold_dir, f"lm_{other}.npz")), cfg_tiny().eog) == 0


@pytest.mark.parametrize("name", LM_CASES)
def test_iteration_bound_covers_the_reference_rollouts(gold_dir, name):
    """lm.SSR_Speech caps its last chunk of decode iterations at n_spans * seq.expected_steps(...) (the reference's length guard,
    ssr.py:739, bounds every span): the bound must cover every roll-out the reference produced, and be exact when the guard
    ends a single span."""
    cfg = cfg_tiny()
    g = np.load(os.path.join(gold_dir, f"lm_{name}.npz"))
    kw = json.loads(str(g["kw"]))
    x, y, out_len = g["x"], g["y"].T.copy(), 0
    if kw.get("aug_context") and sum(b - a for a, b in g["mask_interval"].tolist()) < 2 * 50:
        x, y, out_len = np.concatenate([g["prompt_x"], x]), np.concatenate([g["prompt"].T, y], 1), g["prompt"].shape[0]
    prep = seq.prepare(cfg, y, g["mask_interval"].tolist(), out_len=out_len)
    bound = prep.num_spans * seq.expected_steps(cfg, len(x), prep.prompt_tokens.shape[1])
    total = int(g["ref_span_lens"].sum())
    assert total <= bound, (total, bound)
    if prep.num_spans == 1 and g["ref_span_tokens"][-cfg.n_codebooks, 0] == cfg.eog and total == bound:
        assert g["ref_span_lens"][0] == seq.expected_steps(cfg, len(x), prep.prompt_tokens.shape[1])


@pytest.mark.parametrize("tag", ["default", "all"])
def test_lm_oracle_reproduces_reference_training_forward(lm_oracle, gold_dir, tag):
    """SSR_Speech.forward (models/ssr.py:280-379, eval mode) of the unmodified reference on a dataset-style batch
    (mask tokens, eog, empty-token delay pattern, padding) for both settings of predict_mask_token / predict_all: loss,
    top-10 accuracy (sum over codebooks of accuracy x ntokens) and effective_ntoken."""
    cfg, oracle = lm_oracle
    g = np.load(os.path.join(gold_dir, "lm_train_forward.npz"))
    got = oracle.forward_loss(torch.from_numpy(g["x"]), torch.from_numpy(g["x_lens"]), torch.from_numpy(g["y"]),
                              torch.from_numpy(g["y_lens"]), predict_mask_token=bool(g[f"{tag}_predict_mask_token"]),
                              predict_all=bool(g[f"{tag}_predict_all"]), codebook_weight=g[f"{tag}_codebook_weight"].tolist())
    assert got["effective_ntoken"] == int(g[f"{tag}_ntoken"])
    assert abs(got["loss"] - float(g[f"{tag}_loss"])) <= 1e-5 * abs(float(g[f"{tag}_loss"]))
    assert abs(got["top10acc"] - float(g[f"{tag}_top10acc"])) <= 1e-4
    assert np.allclose(got["top10acc_by_codebook"], g[f"{tag}_top10acc_by_codebook"], atol=1e-4)
