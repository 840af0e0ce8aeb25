"""The two PRODUCTION attention kernels of the bf16 path, called directly through the C ABI (ssrb_op_attn_decode /
ssrb_op_attn_prefill) at the bench geometry — 64 rows x 16 heads, up to 1115 cached positions, ragged lengths, finished rows —
against fp32 torch SDPA on the same bf16-rounded K/V (the arithmetic of models/modules/activation.py:634).

Tolerance: the decode kernel keeps q, scores, softmax and accumulators in fp32 and rounds only the output to bf16, so its error
budget is one bf16 rounding of the output (rel 2^-9) plus fp32 summation-order noise: |got - want| <= 2e-3 + 4e-3 * |want|
(outputs are O(0.1-1)).  The prefill kernel feeds bf16 Q' and P to the tensor cores: it is held to the same bar against a
reference that rounds those two operands where the kernel does, and to 1.5e-2 against exact fp32 SDPA.
The decode kernel's in-place KV append is checked bit-exactly (bf16(k_new), bf16(v_new) at slot seq_len[r], nothing else touched).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from ssr_speech_b200 import _lib

pytestmark = pytest.mark.gpu
H, DH = 16, 128
D = H * DH


def _sdpa_decode(q, kc, vc, n_keys):
    """q [R,H,128] fp32, kc/vc [R,H,Smax,128] fp32 (already bf16-rounded), n_keys [R] -> [R,H,128] fp32"""
    R, _, Smax, _ = kc.shape
    att = torch.einsum("rhd,rhsd->rhs", q, kc) / (DH ** 0.5)
    mask = torch.arange(Smax, device=q.device)[None, None, :] >= n_keys[:, None, None]
    att = att.masked_fill(mask, float("-inf"))
    return torch.einsum("rhs,rhsd->rhd", torch.softmax(att, -1), vc)


def _run_decode(R, Smax, seq_len, done=None, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    qkv = torch.randn(R, 3 * D, device="cuda", generator=g) * 1.5
    kc = (torch.randn(R, H, Smax, DH, device="cuda", generator=g)).to(torch.bfloat16)
    vc = (torch.randn(R, H, Smax, DH, device="cuda", generator=g)).to(torch.bfloat16)
    kc0, vc0 = kc.clone(), vc.clone()
    out = torch.full((R, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    sl = np.asarray(seq_len, dtype=np.int32)
    dn = None if done is None else np.asarray(done, dtype=np.int32)
    torch.cuda.synchronize()
    _lib.check(_lib.load().ssrb_op_attn_decode(
        C.c_void_p(qkv.data_ptr()), C.c_void_p(kc.data_ptr()), C.c_void_p(vc.data_ptr()), C.c_void_p(sl.ctypes.data),
        C.c_void_p(dn.ctypes.data) if dn is not None else None, C.c_void_p(out.data_ptr()), R, D, H, Smax,
        _lib.stream_ptr()), "op_attn_decode")
    # reference: caches with this step's bf16-rounded K/V row appended
    kref, vref = kc0.clone(), vc0.clone()
    live = torch.ones(R, dtype=torch.bool) if dn is None else torch.from_numpy(dn == 0)
    knew = qkv[:, D:2 * D].to(torch.bfloat16).view(R, H, DH)
    vnew = qkv[:, 2 * D:].to(torch.bfloat16).view(R, H, DH)
    for r in range(R):
        if live[r]:
            kref[r, :, sl[r]] = knew[r]
            vref[r, :, sl[r]] = vnew[r]
    assert torch.equal(kc, kref) and torch.equal(vc, vref), "in-place KV append differs"
    want = _sdpa_decode(qkv[:, :D].view(R, H, DH), kref.float(), vref.float(), torch.from_numpy(sl + 1).cuda()).reshape(R, D)
    got = out.float()
    lv = live.cuda()
    err = (got[lv] - want[lv]).abs()
    tol = 2e-3 + 4e-3 * want[lv].abs()
    assert torch.isfinite(got[lv]).all()
    assert bool((err <= tol).all()), float((err - tol).max())
    if not bool(lv.all()):
        assert torch.isnan(got[~lv]).all(), "rows of finished utterances must not be written"
    return float(err.max())


@pytest.mark.parametrize("S", [0, 1, 63, 64, 65, 127, 128, 852, 1114])
def test_decode_attention_uniform_lengths_bench_rows(S):
    """R = 64 rows (32 utterances x cond/uncond), every row with S cached keys (+ this step's): tile edges 63/64/65, the
    mid-generation length of the ncu capture (852) and the last iteration of BASELINE configs[2] (1114 cached + 1 = 1115)."""
    _run_decode(64, 1130, [S] * 64, seed=S)


def test_decode_attention_ragged_lengths_and_finished_rows():
    """Ragged batch: the balanced tile cuts straddle streams of different lengths; a third of the rows are finished (no tiles,
    no append, output untouched)."""
    rng = np.random.RandomState(5)
    R = 64
    sl = rng.randint(0, 1115, size=R)
    sl[:6] = [0, 1, 63, 64, 65, 1114]
    done = (rng.rand(R) < 0.33).astype(np.int32)
    done[:6] = 0
    _run_decode(R, 1130, sl, done, seed=77)


@pytest.mark.parametrize("R,Smax,S", [(1, 1130, 611), (2, 1130, 1000), (16, 600, 461), (128, 300, 255), (3, 8200, 8100)])
def test_decode_attention_other_batches(R, Smax, S):
    """Batch 1 / 2 (one stream cut into one piece per tile: the workspace-merge path), batch 8 of the edit config, 128 rows,
    and one very long stream."""
    _run_decode(R, Smax, [S - (r % 3) for r in range(R)], seed=R)


def _run_prefill(row_len, Smax, seed=0):
    n = len(row_len)
    M = int(sum(row_len))
    g = torch.Generator(device="cuda").manual_seed(seed)
    qkv = torch.randn(M, 3 * D, device="cuda", generator=g)
    kc = torch.zeros(n, H, Smax, DH, device="cuda", dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    out = torch.zeros(M, D, device="cuda", dtype=torch.bfloat16)
    rl = np.asarray(row_len, dtype=np.int32)
    torch.cuda.synchronize()
    _lib.check(_lib.load().ssrb_op_attn_prefill(
        C.c_void_p(qkv.data_ptr()), C.c_void_p(kc.data_ptr()), C.c_void_p(vc.data_ptr()), C.c_void_p(rl.ctypes.data), n,
        C.c_void_p(out.data_ptr()), D, H, Smax, _lib.stream_ptr()), "op_attn_prefill")
    o = 0
    worst = 0.0
    qs = 0.08838834764831845 * 1.4426950408889634          # 1/sqrt(128) * log2(e), folded into Q before its bf16 rounding
    for i, L in enumerate(row_len):
        k = qkv[o:o + L, D:2 * D].to(torch.bfloat16)
        v = qkv[o:o + L, 2 * D:].to(torch.bfloat16)
        assert torch.equal(kc[i, :, :L], k.view(L, H, DH).transpose(0, 1)), "prefill K cache"
        assert torch.equal(vc[i, :, :L], v.view(L, H, DH).transpose(0, 1)), "prefill V cache"
        kk = k.float().view(L, H, DH).transpose(0, 1)
        vv = v.float().view(L, H, DH).transpose(0, 1)
        q = qkv[o:o + L, :D].view(L, H, DH).transpose(0, 1)
        causal = torch.ones(L, L, dtype=torch.bool, device="cuda").triu(1)
        # (1) the bf16-storage oracle of this kernel: the tensor-core operands Q' = bf16(q * qs), K, V and P = bf16(2^(s - max))
        # are rounded exactly where the kernel rounds them, everything else (scores, softmax, row sums, accumulation) in fp32.
        # What is left is fp32 summation order, the final bf16 rounding of the output and the scale at which P is rounded (the
        # kernel rounds 2^(s - running max) tile by tile and rescales in fp32): 3e-3 + 6e-3 |want|.
        s2 = torch.matmul((q * qs).to(torch.bfloat16).float(), kk.transpose(1, 2)).masked_fill(causal, float("-inf"))
        p = torch.exp2(s2 - s2.max(-1, keepdim=True).values)
        want = torch.matmul(p.to(torch.bfloat16).float(), vv) / p.sum(-1, keepdim=True)
        want = want.transpose(0, 1).reshape(L, D)
        got = out[o:o + L].float()
        err = (got - want).abs()
        tol = 3e-3 + 6e-3 * want.abs()
        assert bool((err <= tol).all()), (i, L, float((err - tol).max()))
        # (2) against exact fp32 SDPA on the bf16-rounded K/V with UNROUNDED q (activation.py:634): the bf16 rounding of Q' and P
        # moves a score by ~2^-9 relative (scores are O(1) here, as in the model), i.e. the output by <= ~1e-2
        exact = torch.nn.functional.scaled_dot_product_attention(q[None], kk[None], vv[None], is_causal=True)[0]
        exact = exact.transpose(0, 1).reshape(L, D)
        assert float((got - exact).abs().max()) <= 1.5e-2, (i, L, float((got - exact).abs().max()))
        worst = max(worst, float(err.max()))
        o += L
    return worst


@pytest.mark.parametrize("row_len", [[1], [63, 64, 65], [611] * 4, [1115, 611, 7, 128], [200, 461, 461, 33, 1, 90]])
def test_prefill_attention_vs_sdpa(row_len):
    """Packed causal prefill incl. single-position rows, tile edges, the bench prompt (611) and a full 1115-position row
    (teacher forcing at the end-of-generation length)."""
    _run_prefill(row_len, 1130, seed=len(row_len))
