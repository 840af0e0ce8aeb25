"""GEMM kernels behind ssrb_op_gemm vs a plain torch fp32 reference of the same op."""
import ctypes as C

import numpy as np
import pytest
import torch

from ssr_speech_b200 import _lib

pytestmark = pytest.mark.gpu


def run_gemm(A, W, bias, res, act, dtype, impl):
    lib = _lib.load()
    M, K = A.shape
    N = W.shape[0]
    out = torch.empty(M, N, dtype=torch.float32, device="cuda")
    _lib.check(lib.ssrb_op_gemm(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()),
                                C.c_void_p(bias.data_ptr()) if bias is not None else None,
                                C.c_void_p(res.data_ptr()) if res is not None else None, C.c_void_p(out.data_ptr()),
                                M, N, K, dtype, act, impl, _lib.stream_ptr()), "op_gemm")
    torch.cuda.synchronize()
    return out


def ref_gemm(A, W, bias, res, act):
    y = A.float() @ W.float().t()
    if bias is not None:
        y = y + bias
    if act == 1:
        y = torch.relu(y)
    elif act == 2:
        y = torch.nn.functional.gelu(y)
    if res is not None:
        y = y + res
    return y


SHAPES = [(1, 768, 256), (2, 2056, 1024), (3, 6144, 2048), (7, 72, 32), (64, 2048, 2048), (200, 6144, 2048), (611, 8192, 2048),
          (130, 2056, 1024), (64, 2048, 8192)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("act", [0, 1, 2])
def test_simt_fp32(M, N, K, act):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g)
    torch.backends.cuda.matmul.allow_tf32 = False
    got = run_gemm(A, W, b, r, act, _lib.SSRB_DTYPE_F32, 1)
    want = ref_gemm(A.double(), W.double(), b.double(), r.double(), act).float() if False else ref_gemm(A, W, b, r, act)
    assert (got - want).abs().max().item() <= 2e-4


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_simt_bf16(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    got = run_gemm(A, W, b, None, 1, _lib.SSRB_DTYPE_BF16, 1)
    want = ref_gemm(A, W, b, None, 1)
    assert (got - want).abs().max().item() <= 2e-3


TC_SHAPES = [(1, 768, 256), (2, 6144, 2048), (16, 2048, 2048), (17, 2056, 1024), (64, 6144, 2048), (64, 2048, 8192),
             (64, 8192, 2048), (100, 4096, 2048), (128, 2048, 2048), (129, 2048, 2048), (611, 6144, 2048), (1000, 2056, 1024),
             (300, 8192, 2048), (20000, 1024, 512), (5000, 4096, 2048), (4097, 264, 128)]


@pytest.mark.parametrize("M,N,K", TC_SHAPES)
@pytest.mark.parametrize("act", [0, 1])
def test_tcgen05_bf16(M, N, K, act):
    """tcgen05/TMEM/TMA GEMM (swap-AB split-K for M<=128, flat tiles above) vs torch fp32 on the same bf16 inputs."""
    g = torch.Generator(device="cuda").manual_seed(M * 13 + N)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g)
    got = run_gemm(A, W, b, r, act, _lib.SSRB_DTYPE_BF16, 2)
    want = ref_gemm(A, W, b, r, act)
    err = (got - want).abs().max().item()
    assert err <= 2e-3, err
    # run twice: split-K tickets must have been reset and the result must be bit-identical (deterministic reduce)
    again = run_gemm(A, W, b, r, act, _lib.SSRB_DTYPE_BF16, 2)
    assert torch.equal(got, again)


LN_SHAPES = [(64, 2048, 2048, 6144), (64, 2048, 8192, 8192), (2, 2048, 2048, 4096), (5, 256, 256, 768), (128, 512, 2048, 2048),
             (100, 1024, 4096, 4096), (33, 2048, 8192, 6144), (1, 128, 128, 384)]


@pytest.mark.parametrize("M,D,K1,N", LN_SHAPES)
@pytest.mark.parametrize("act", [0, 1, 2])
def test_tcgen05_folded_layernorm(M, D, K1, N, act):
    """LayerNorm folded into the decode GEMMs: the residual GEMM emits fp32 x, bf16(x) and per-128-column {mean, M2}
    partials; the consuming GEMM multiplies bf16(x) by bf16(gamma*W) and applies rstd*(acc - mean*colsum) + (b + W.beta)
    in its epilogue.  Against (a) torch LayerNorm + linear in fp32 (tolerance 1e-2 x max|C|: operands are bf16) and (b) the same
    algebra evaluated in fp64 on the rounded operands (tolerance 2e-3: only accumulation order differs).  Rows carry a
    non-zero mean (1.5 sigma) so the mean*colsum correction is exercised."""
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M * 31 + N + act)
    A1 = torch.randn(M, K1, device="cuda", generator=g).bfloat16()
    W1 = (torch.randn(D, K1, device="cuda", generator=g) / K1 ** 0.5).bfloat16()
    b1 = torch.randn(D, device="cuda", generator=g)
    res = torch.randn(M, D, device="cuda", generator=g) + 1.5 * torch.randn(M, 1, device="cuda", generator=g)
    W2 = (torch.randn(N, D, device="cuda", generator=g) / D ** 0.5).bfloat16()
    gamma = 1.0 + 0.2 * torch.randn(D, device="cuda", generator=g)
    beta = 0.2 * torch.randn(D, device="cuda", generator=g)
    b2 = torch.randn(N, device="cuda", generator=g)
    X = torch.empty(M, D, dtype=torch.float32, device="cuda")
    Cout = torch.empty(M, N, dtype=torch.float32, device="cuda")

    def call():
        _lib.check(lib.ssrb_op_gemm_ln(C.c_void_p(A1.data_ptr()), C.c_void_p(W1.data_ptr()), C.c_void_p(b1.data_ptr()),
                                       C.c_void_p(res.data_ptr()), C.c_void_p(X.data_ptr()), M, D, K1, C.c_void_p(W2.data_ptr()),
                                       C.c_void_p(gamma.data_ptr()), C.c_void_p(beta.data_ptr()), C.c_void_p(b2.data_ptr()),
                                       C.c_void_p(Cout.data_ptr()), N, act, _lib.stream_ptr()), "op_gemm_ln")
        torch.cuda.synchronize()
        return X.clone(), Cout.clone()

    x_got, c_got = call()
    x_want = A1.float() @ W1.float().t() + b1 + res
    assert (x_got - x_want).abs().max().item() <= 2e-3
    actf = {0: lambda t: t, 1: torch.relu, 2: torch.nn.functional.gelu}[act]
    want32 = actf(torch.nn.functional.layer_norm(x_got, (D,), gamma, beta, 1e-5) @ W2.float().t() + b2)
    tol32 = 1e-2 * max(1.0, want32.abs().max().item())
    assert (c_got - want32).abs().max().item() <= tol32, ((c_got - want32).abs().max().item(), tol32)
    xd = x_got.double()
    mu, var = xd.mean(1, keepdim=True), xd.var(1, unbiased=False, keepdim=True)
    Wf = (W2.float() * gamma).bfloat16().double()
    acc = x_got.bfloat16().double() @ Wf.t()
    want_fold = actf((acc - mu * Wf.sum(1)) / torch.sqrt(var + 1e-5) + b2.double() + W2.double() @ beta.double()).float()
    assert (c_got - want_fold).abs().max().item() <= 2e-3, (c_got - want_fold).abs().max().item()
    x2, c2 = call()
    assert torch.equal(x_got, x2) and torch.equal(c_got, c2)           # fixed-order partial combination: deterministic


def _ln_chain_out(M, D, K1, N, act):
    """(x, C) of the residual GEMM -> folded-LayerNorm GEMM chain on seeded inputs (used in- and out-of-process)."""
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M * 17 + N + act)
    A1 = torch.randn(M, K1, device="cuda", generator=g).bfloat16()
    W1 = (torch.randn(D, K1, device="cuda", generator=g) / K1 ** 0.5).bfloat16()
    b1 = torch.randn(D, device="cuda", generator=g)
    res = torch.randn(M, D, device="cuda", generator=g)
    W2 = (torch.randn(N, D, device="cuda", generator=g) / D ** 0.5).bfloat16()
    gamma = 1.0 + 0.2 * torch.randn(D, device="cuda", generator=g)
    beta = 0.2 * torch.randn(D, device="cuda", generator=g)
    b2 = torch.randn(N, device="cuda", generator=g)
    X = torch.empty(M, D, dtype=torch.float32, device="cuda")
    Cout = torch.empty(M, N, dtype=torch.float32, device="cuda")
    _lib.check(lib.ssrb_op_gemm_ln(C.c_void_p(A1.data_ptr()), C.c_void_p(W1.data_ptr()), C.c_void_p(b1.data_ptr()),
                                   C.c_void_p(res.data_ptr()), C.c_void_p(X.data_ptr()), M, D, K1, C.c_void_p(W2.data_ptr()),
                                   C.c_void_p(gamma.data_ptr()), C.c_void_p(beta.data_ptr()), C.c_void_p(b2.data_ptr()),
                                   C.c_void_p(Cout.data_ptr()), N, act, _lib.stream_ptr()), "op_gemm_ln")
    torch.cuda.synchronize()
    return X.cpu().numpy(), Cout.cpu().numpy()


def _ln_chain_dump(path):
    np.savez(path, **{f"{k}{i}": a for i, c in enumerate(DEC_CASES) for k, a in zip("xc", _ln_chain_out(*c))})


DEC_CASES = [(64, 2048, 2048, 6144, 0), (64, 2048, 8192, 6144, 0), (2, 2048, 2048, 6144, 0), (16, 2048, 8192, 6144, 0),
             (33, 2048, 2048, 6144, 0), (128, 2048, 8192, 6144, 0)]


def test_dec_role_kernels_match_generic(tmp_path):
    """The compact per-role decode kernels (gemm_dec_kernel: ROLE_RES for the residual GEMM at 8-way split-K, ROLE_QKV for the
    folded-LayerNorm consumer at 4-way split-K) keep the generic kernel's algorithm and summation order.  Against a process
    that runs the generic gemm_tc_kernel (SSRB_GEMM_DEC=0): the residual stream x (adds only) must be BIT-identical; the
    LayerNorm-folded output may differ by FMA contraction of rstd*(acc - mean*colsum) + bias in two separately compiled
    bodies — tolerance 1e-5 x max|C| (measured: a few ulp).  The ReLU/bf16 role (FFN1) is covered by the 830M roll-outs of
    test_gpu_fullsize.py against the oracle."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    out = str(tmp_path / "generic.npz")
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import test_gpu_gemm as t; t._ln_chain_dump(%r)"
            % (os.path.dirname(here), here, out))
    env = dict(os.environ, SSRB_GEMM_DEC="0")
    p = subprocess.run([sys.executable, "-c", code], cwd=os.path.dirname(here), env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    generic = np.load(out)
    for i, c in enumerate(DEC_CASES):
        x, cc = _ln_chain_out(*c)
        gx, gc = generic[f"x{i}"], generic[f"c{i}"]
        assert np.isfinite(cc).all()
        err_x = float(np.abs(x - gx).max())
        assert err_x <= 1e-6 * max(1.0, float(np.abs(gx).max())), (c, err_x)
        err = float(np.abs(cc - gc).max())
        assert err <= 1e-5 * max(1.0, float(np.abs(gc).max())), (c, err)
