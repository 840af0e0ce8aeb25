"""Execution-mode invariance of the bf16 production path: CUDA graph vs eager launches and programmatic dependent launch
(PDL) on/off must give bit-identical tokens and logits (same kernels, different scheduling — a difference would be a race).
Kernel-variant switches (SIMT GEMM, simple attention) must agree within the bf16 tolerance."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

SNIPPET = r"""
import json, sys, os, hashlib
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import numpy as np, torch
from ssr_speech_b200.config import cfg_tiny
from ssr_speech_b200.lm import SSR_Speech
from ssr_speech_b200.synth import make_lm_state_dict
cfg = cfg_tiny(d_model=512, nhead=4, num_layers=3, audio_vocab_size=2048)
m = SSR_Speech(cfg.to_namespace(), precision="bf16")
m.load_state_dict(make_lm_state_dict(cfg, seed=31))
m.to("cuda")
g = torch.Generator().manual_seed(8)
xs = [torch.randint(0, 100, (n,), generator=g) for n in (9, 12, 7, 10, 11)]
ys = [torch.randint(0, 2048, (t, 4), generator=g) for t in (40, 70, 33, 150, 64)]
spans = [[[40, 40]], [[10, 30]], [[33, 33]], [[150, 150]], [[5, 9], [30, 40]]]
un = [torch.randint(0, 101, (x.shape[0],), generator=g) for x in xs]
m.poll_every = 1
out = m.inference_batch(xs, ys, spans, top_k=0, top_p=0.9, stop_repetition=2, cfg_coef=1.5, cfg_stride=2, aug_text=True,
                        uncond_xs=un, seed=99)
toks = [o[0].cpu().numpy().tolist() for o in out]
lg = m.last_raw_logits().numpy()
tf = m.teacher_forced_logits(xs[3], ys[3].T.contiguous()).numpy()
# decode chain vs prefill path on the same tokens (B=1 greedy TTS): logits of the last iteration, incremental vs full forward
from ssr_speech_b200 import seq
T, Lx = 30, 8
x1 = torch.randint(0, 100, (1, Lx), generator=g); y1 = torch.randint(0, 2048, (1, T, 4), generator=g)
res1 = m.inference(x1.cuda(), torch.tensor([Lx]), x1.cuda(), torch.tensor([Lx]), y1.cuda(), y1.cuda(),
                   mask_interval=torch.tensor([[[T, T]]]), top_k=1, top_p=1.0, temperature=1.0, stop_repetition=-1, kvcache=1,
                   aug_text=False)[0]
raw_last = m.last_raw_logits()[0].numpy()
r1 = res1[0].cpu().numpy()
prep = seq.prepare(cfg, y1[0].numpy().T.copy(), [[T, T]])
gen = seq.delay_pattern(np.concatenate([r1[:, T:], np.full((4, 1), cfg.eog)], 1), cfg.empty_token)
N = r1.shape[1] - T + 4
fed = np.concatenate([prep.prompt_tokens, np.full((4, 1), cfg.mts), gen[:, :N - 1]], 1)
tf1 = m.teacher_forced_logits(x1[0], torch.from_numpy(fed)).numpy()
inc_err = float(np.abs(tf1[-1] - raw_last).max())
print("RESULT" + json.dumps({"inc_err": inc_err, "tokens_sha": hashlib.sha256(json.dumps(toks).encode()).hexdigest(),
                             "logits_sha": hashlib.sha256(lg.tobytes()).hexdigest(),
                             "tf_sha": hashlib.sha256(tf.tobytes()).hexdigest(),
                             "tf_probe": tf[::7, :, ::97].tolist(), "lg_probe": lg[:, :, ::97].tolist(), "n_frames": [len(t[0]) for t in toks]}))
"""


def run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    p = subprocess.run([sys.executable, "-c", SNIPPET], cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT")][-1]
    return json.loads(line[len("RESULT"):])


@pytest.fixture(scope="module")
def base():
    return run({})


@pytest.mark.parametrize("env", [{"SSRB_NO_PDL": "1"}, {"SSRB_NO_GRAPH": "1"}, {"SSRB_NO_PDL": "1", "SSRB_NO_GRAPH": "1"},
                                 {"SSRB_ATTN_PREFETCH": "0"}])   # K/V stream started before / after griddepcontrol.wait
def test_scheduling_modes_are_bit_identical(base, env):
    other = run(env)
    assert other["n_frames"] == base["n_frames"]
    assert other["tokens_sha"] == base["tokens_sha"]
    assert other["logits_sha"] == base["logits_sha"]
    assert other["tf_sha"] == base["tf_sha"]


@pytest.mark.parametrize("env", [{"SSRB_GEMM_IMPL": "1"}, {"SSRB_ATTN_SIMPLE": "1"}, {"SSRB_PREFILL_SIMT": "1"},
                                 {"SSRB_LN_FOLD": "0"}])        # separate LayerNorm kernels instead of the folded epilogue
def test_kernel_variants_agree_within_bf16_tolerance(base, env):
    other = run(env)
    a, b = np.asarray(base["tf_probe"]), np.asarray(other["tf_probe"])
    assert np.abs(a - b).max() <= 2e-2, np.abs(a - b).max()
    # decode chain (KV cache, folded LayerNorm, swap-AB GEMMs) == prefill path on the same tokens, in every variant
    assert base["inc_err"] <= 2e-2 and other["inc_err"] <= 2e-2, (base["inc_err"], other["inc_err"])
