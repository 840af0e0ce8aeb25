"""The CTA-pair prefill GEMM (csrc/gemm_flat2.cu, tcgen05.mma.cta_group::2; the default prefill GEMM since round 2) against torch
fp32 and against the 1-CTA kernel it replaced (SSRB_FLAT_2CTA=0).  Every case runs in a child process under a timeout (the
kernel's spins trap after 2 s, a hang would otherwise cost the box)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

SNIPPET = r"""
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from ssr_speech_b200 import _lib
M, N, K, act, with_res, out_bf16 = (int(v) for v in sys.argv[1:7])
out = sys.argv[7]
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(M * 13 + N + K)
A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
b = torch.randn(N, device="cuda", generator=g)
r = torch.randn(M, N, device="cuda", generator=g) if with_res else None
Cd = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
outs = []
for rep in range(2):
    Cd.fill_(float("nan"))
    _lib.check(lib.ssrb_op_gemm(p(A), p(W), p(b), p(r), p(Cd), M, N, K, _lib.SSRB_DTYPE_BF16, act, 2, _lib.stream_ptr()), "op_gemm")
    torch.cuda.synchronize()
    outs.append(Cd.clone())
assert torch.equal(outs[0], outs[1]), "not deterministic"
want = A.float() @ W.float().t() + b
want = torch.relu(want) if act == 1 else want
if r is not None:
    want = want + r
np.savez(out, got=outs[0].cpu().numpy(), err=float((outs[0] - want).abs().max().item()))
print("OK")
"""

# (M, N, K, act, residual): full tiles, ragged M (not a multiple of 256 / 128), ragged N (2056: not a multiple of 256 or 32), one k-block,
# more tiles than CTA pairs (several rounds, both accumulator buffers reused), fewer tiles than pairs
SHAPES = [(256, 256, 64, 0, 0), (512, 512, 2048, 1, 1), (129, 2048, 2048, 0, 1), (611, 6144, 2048, 0, 0), (1000, 2056, 1024, 1, 1),
          (300, 8192, 2048, 0, 1), (20000, 1024, 512, 1, 0), (5000, 4096, 2048, 0, 1), (4097, 264, 128, 0, 0), (39104, 2048, 8192, 0, 1)]


def _run(shape, env_extra, out):
    env = dict(os.environ)
    env.update(env_extra)
    p = subprocess.run([sys.executable, "-c", SNIPPET, *map(str, shape), "0", out], cwd=ROOT, env=env, capture_output=True, text=True,
                       timeout=300)
    assert p.returncode == 0, p.stderr[-3000:]
    return np.load(out)


@pytest.mark.parametrize("shape", SHAPES, ids=["x".join(map(str, s[:3])) for s in SHAPES])
def test_cta_pair_gemm_matches_torch_and_the_1cta_kernel(shape, tmp_path):
    pair = _run(shape, {"SSRB_FLAT_2CTA": "1"}, str(tmp_path / "pair.npz"))
    one = _run(shape, {"SSRB_FLAT_2CTA": "0"}, str(tmp_path / "one.npz"))
    assert np.isfinite(pair["got"]).all()                      # every element of C was written (C starts as NaN)
    assert float(pair["err"]) <= 2e-3, float(pair["err"])      # vs torch fp32 on the same bf16 operands (test_gpu_gemm.py's bar)
    # same k-block order and the same K=16 instruction shape: expected bit-identical; 1e-6 relative allows for a different
    # in-instruction summation order of the two-SM datapath
    d = float(np.abs(pair["got"] - one["got"]).max())
    assert d <= 1e-6 * max(1.0, float(np.abs(one["got"]).max())), d


def test_engine_prefill_through_the_cta_pair_kernel():
    """Teacher-forced logits and whole roll-outs (prefill + decode) with the prefill GEMMs on CTA pairs: same tokens, logits within
    the bf16 tolerance of the default path."""
    from test_gpu_modes import run
    base, other = run({}), run({"SSRB_FLAT_2CTA": "1"})
    assert other["n_frames"] == base["n_frames"] and other["tokens_sha"] == base["tokens_sha"]
    a, b = np.asarray(base["tf_probe"]), np.asarray(other["tf_probe"])
    assert np.abs(a - b).max() <= 2e-3, np.abs(a - b).max()
