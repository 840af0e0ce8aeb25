// lm_mega.cu — one persistent launch per decode iteration for small batches (R <= 16 transformer rows: batch 1 and 8 of
// BASELINE's metric, with their CFG rows).  sm_100a.
//
// STATUS (round 2): EXPERIMENTAL, OFF by default (SSRB_MEGA=1 selects it).  Verified against the oracle on a B200
// (tests/test_gpu_mega.py: R = 1 / 2 / 6 / 10 ragged / 16, the 830M model at batch 1 and 8, <= 2e-2 on raw logits along whole
// roll-outs), but it does NOT beat the PDL chain of per-GEMM kernels yet: 0.85 ms per iteration at batch 1 (chain 0.66),
// 1.38 ms at batch 8 (chain 0.89).  Per-phase trace, what was tried and what the numbers say: profiles/r02b_mega_kernel.md.
//
// Replaces, for one iteration of the reference's `while True` loop (models/ssr.py:671-771), the 82 launches of the per-GEMM
// chain between the embedding and the sampler: 16 x { LayerNorm1 + packed QKV projection (transformer.py:58-75,
// activation.py:83-89), single-query attention against the in-place KV cache incl. the append (activation.py:626-634),
// out-proj + residual (activation.py:637, transformer.py:321-343), LayerNorm2 + FFN1 + ReLU, FFN2 + residual
// (transformer.py:386-388) }, final LayerNorm and the 4 prediction heads (ssr.py:175-179,687-689).
//
// Why: at R <= 16 the chain is latency-bound — 84 launches of ~7.8 us for 0.25 ms of weight stream (roofline 0.44 at batch 1,
// 0.59 at batch 8; profiles/r01e_summary.md).  A kernel boundary or a grid barrier costs about the same, so the gain is not
// "fewer launches" but a weight stream that NEVER waits for a dependency:
//   * grid = one CTA per SM, all co-resident.  Every GEMV phase gives each CTA a contiguous range of output features (to within
//     one feature: 2048 / 148 = 13.84 -> 13 or 14), so no split-K, no cross-CTA reduction, no cluster, no tensor memory: the
//     activations are tiny (<= 16 rows), the weights are the stream.  Weights are re-packed once at load time into the order
//     each CTA consumes them: 16 KB chunks of [<= 8 features][1024 k] bf16, one `cp.async.bulk` each.
//   * warp 0 is the PRODUCER: one lane walks the whole iteration's byte stream of its CTA — QKV weights, the K/V tiles of its
//     attention range, out-proj, FFN1, FFN2 weights, next layer ... heads — through a 9 x 16 KB shared-memory ring.  Everything it
//     touches is immutable during the iteration (weights; cache rows written by EARLIER iterations), so it never waits for a
//     phase barrier: while the consumers sit at a grid barrier the ring fills with the next phase's bytes (21 MB in flight
//     over the chip = 3 us of HBM time, about one barrier + activation round trip).
//   * warps 1-16 are the CONSUMERS: per phase they wait for the grid barrier, stage the phase's activations in shared memory
//     (LayerNorm applied on the way for the phases that consume one: bf16(LN(x)) exactly as the oracle rounds it, no folded
//     weights), run `mma.sync.m16n8k16` (16 activation rows x 8 features per instruction, k split over the 16 warps), reduce
//     the 16 partial tiles through shared memory in a fixed order and apply the epilogue (bias / ReLU / GELU / residual).
//   * the attention phase is attn_decode_tma_kernel's algorithm (balanced contiguous tile ranges over the live (row, head)
//     streams, fp32 online softmax, pieces of straddling streams merged in piece order, fused in-place KV append) on the
//     same ring: a 64-key K tile and its V tile are two ring stages.
// Grid barrier: one monotonically increasing arrival counter (zeroed by the embedding kernel that opens every iteration),
// one arrival per CTA per phase, polled with ld.acquire.gpu; every spin is bounded by a 2 s watchdog (trap, not hang).
// Data written inside the kernel (qkv, ao, x, hid, hh) is always re-read through L2 (ld.global.cg / cp.async.cg), never L1.
#include <cstring>

#include "lm_kernels.cuh"

namespace ssrb {

namespace {

constexpr int MG_CW = 16;                              // consumer warps
constexpr int MG_THREADS = (MG_CW + 1) * 32;           // warp 0 = producer
constexpr int MG_STAGE = 16384;
constexpr int MG_NS = 9;
constexpr int MG_KC = 1024;                            // k elements per chunk (chunk = <= 8 features x MG_KC)
constexpr int MG_ASLICE = 16 * MG_KC * 2;              // one activation slice: 16 rows x 1024 k bf16 = 32 KB
constexpr int MG_ABUF = 2 * MG_ASLICE;
constexpr int MG_MAXG = 8;                             // 8-feature groups per CTA and phase (N <= 64 * grid)
constexpr int MG_BAR_BYTES = 256;
constexpr size_t MG_SMEM = (size_t)MG_NS * MG_STAGE + MG_ABUF + MG_BAR_BYTES + 128;
constexpr int AT_SUB = 64, AT_ROWB = 256;              // keys per K/V tile, bytes per key row (head_dim 128, bf16)
#ifndef MG_WATCHDOG_CYCLES
#define MG_WATCHDOG_CYCLES 4000000000ull          // ~2 s at 1.9 GHz
#endif

enum { MP_QKV = MEGA_QKV, MP_OUT = MEGA_OUT, MP_FFN1 = MEGA_FFN1, MP_FFN2 = MEGA_FFN2, MP_H1 = MEGA_H1, MP_H2 = MEGA_H2 };

struct MegaPhase {
    const bf16* w;              // packed (mega_pack_kernel)
    const float* bias;
    const float* ln_g; const float* ln_b;      // LayerNorm on x forms the activations (QKV, FFN1, H1), else null
    int N, K;
};
struct MegaLayer { MegaPhase qkv, out, ffn1, ffn2; bf16* kc; bf16* vc; };
struct MegaParams {
    int R, D, H, F, L, NCB, V, Hh, Smax, rpu, G, max_pieces;
    float* x; float* qkv; bf16* ao; bf16* hid; bf16* hh; float* logits;
    const int* seq_len; const UttState* st;
    float* attn_ws; int* tickets;
    unsigned int* bar;
    const MegaLayer* layers;
    MegaPhase h1, h2;
    float ln_eps;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// watchdog on the SM's cycle counter (CS2R, a few cycles): %globaltimer is a slow chip-level register and a read of it inside
// every failed try_wait would add its latency to every wake-up
__device__ __forceinline__ void spin_guard(unsigned long long& t0) {
    const unsigned long long now = (unsigned long long)clock64();
    if (t0 == 0) t0 = now;
    else if (now - t0 > MG_WATCHDOG_CYCLES) __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    unsigned long long t0 = 0;
    for (;;) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        spin_guard(t0);
    }
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cons_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }      // the 16 consumer warps
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void lds8_bf16(uint32_t addr, float (&v)[8]) {
    const uint4 w4 = lds128(addr);
    const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int i = 0; i < 4; i++) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ void ldcg8(const float* p, float (&v)[8]) {       // data produced inside this kernel: L2, never L1
    const float4 a = __ldcg(reinterpret_cast<const float4*>(p)), b = __ldcg(reinterpret_cast<const float4*>(p + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// ---- static partition of a phase's output features over the grid -------------------------------------------------------
// GEMV phases: CTA c owns features [c*N/G, (c+1)*N/G); the grouped second head layer is dealt in 8-feature groups so that a
// group never straddles two codebooks (their activations differ).
__host__ __device__ __forceinline__ int part_lo(int n_units, int G, int c) { return (int)(((unsigned)c * (unsigned)n_units) / (unsigned)G); }   // c * n_units < 2^31 (n_units <= 64 * G)
__host__ __device__ __forceinline__ void phase_range(int kind, int N, int G, int c, int& f0, int& f1) {
    if (kind == MP_H2) { f0 = 8 * part_lo(N / 8, G, c); f1 = 8 * part_lo(N / 8, G, c + 1); }
    else { f0 = part_lo(N, G, c); f1 = part_lo(N, G, c + 1); }
}
__host__ __device__ __forceinline__ int chunk_k(int K) { return K < MG_KC ? K : MG_KC; }

// ---- ring bookkeeping shared by the producer and the consumers: chunk i of the CTA's stream lives in stage i % MG_NS ----
struct Ring {
    uint32_t base, full0, empty0;
    __device__ __forceinline__ uint32_t stage(uint32_t i) const { return base + (i % MG_NS) * MG_STAGE; }
    __device__ __forceinline__ uint32_t full(uint32_t i) const { return full0 + 8u * (i % MG_NS); }
    __device__ __forceinline__ uint32_t empty(uint32_t i) const { return empty0 + 8u * (i % MG_NS); }
    __device__ __forceinline__ uint32_t parity(uint32_t i) const { return (i / MG_NS) & 1u; }
};

// ---- attention tile bookkeeping (attn_decode_tma.cu) ---------------------------------------------------------------------
struct TilePos { int r, h, t, tiles; };
__device__ __forceinline__ TilePos locate_tile(const int* pref, int R, int H, int g) {
    int lo = 0, hi = R - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (H * pref[mid] <= g) lo = mid; else hi = mid - 1;
    }
    TilePos p;
    p.r = lo; p.tiles = pref[lo + 1] - pref[lo];
    const int rem = g - H * pref[lo];
    p.h = rem / p.tiles; p.t = rem - p.h * p.tiles;
    return p;
}
__device__ __forceinline__ void advance_stream(TilePos& p, const int* pref, int R, int H) {
    p.t = 0;
    if (++p.h == H) {
        p.h = 0;
        do { p.r++; } while (p.r < R && pref[p.r + 1] == pref[p.r]);
        p.tiles = p.r < R ? pref[p.r + 1] - pref[p.r] : 1;
    }
}

// =====================================================================================================================
// producer
// =====================================================================================================================
__device__ __forceinline__ void produce_gemv(const Ring& rg, uint32_t& ci, const MegaPhase& ph, int kind, int G, int c) {
    int f0, f1;
    phase_range(kind, ph.N, G, c, f0, f1);
    const int nfeat = f1 - f0;
    if (nfeat <= 0) return;
    const int Kc = chunk_k(ph.K), nkc = ph.K / Kc, ng = (nfeat + 7) >> 3;
    const bf16* base = ph.w + (long long)f0 * ph.K;
    for (int kc = 0; kc < nkc; kc++)
        for (int a = 0; a < ng; a++, ci++) {
            const int nf = min(8, nfeat - 8 * a);
            mbar_wait(rg.empty(ci), rg.parity(ci) ^ 1u);
            const uint32_t bytes = (uint32_t)(nf * Kc * 2);
            mbar_expect_tx(rg.full(ci), bytes);
            bulk_g2s(rg.stage(ci), base + (long long)kc * nfeat * Kc + (long long)a * 8 * Kc, bytes, rg.full(ci));
        }
}

__device__ __forceinline__ void produce_attn(const Ring& rg, uint32_t& ci, const MegaParams& P, const MegaLayer& ly, const int* pref, int c) {
    const int total = P.H * pref[P.R];
    const int per = (total + P.G - 1) / P.G;
    const int g0 = c * per, g1 = min(total, g0 + per);
    if (g0 >= g1) return;
    TilePos p = locate_tile(pref, P.R, P.H, g0);
    int n_old = P.seq_len[p.r];
    for (int i = 0; i < g1 - g0; i++) {
        const int k0 = p.t * AT_SUB;
        const int nk = min(AT_SUB, n_old - k0);
        const long long off = (((long long)p.r * P.H + p.h) * P.Smax + k0) * 128;
#pragma unroll
        for (int kv = 0; kv < 2; kv++, ci++) {
            mbar_wait(rg.empty(ci), rg.parity(ci) ^ 1u);
            if (nk > 0) {
                mbar_expect_tx(rg.full(ci), (uint32_t)nk * AT_ROWB);
                bulk_g2s(rg.stage(ci), (kv ? ly.vc : ly.kc) + off, (uint32_t)nk * AT_ROWB, rg.full(ci));
            } else {
                mbar_arrive(rg.full(ci));                    // a row without cached keys: an empty tile
            }
        }
        if (++p.t == p.tiles) {
            advance_stream(p, pref, P.R, P.H);
            if (p.r < P.R) n_old = P.seq_len[p.r];
        }
    }
}

// =====================================================================================================================
// consumers
// =====================================================================================================================
// optional trace (ssrb_debug_mega_trace): consumer thread 0 of every CTA stamps %globaltimer at fixed points of every phase
__device__ unsigned long long* g_mega_trace = nullptr;
__device__ int g_mega_trace_cap = 0;

struct ConsCtx {
    int tr_n; unsigned long long* tr;     // trace cursor / this CTA's slice of the trace buffer (null: off)
    Ring rg; uint32_t ci;                 // position in the CTA's chunk stream (must mirror the producer's)
    uint32_t abuf;
    int cw, lane, ctid;                   // consumer warp 0..15, lane, consumer thread 0..511
    unsigned int bar_target;              // arrivals that complete the NEXT grid barrier
};

__device__ __forceinline__ void trace(ConsCtx& cx) {
    if (cx.tr && cx.ctid == 0 && cx.tr_n < g_mega_trace_cap) cx.tr[cx.tr_n++] = globaltimer_ns();
}

__device__ __forceinline__ void grid_barrier(const MegaParams& P, ConsCtx& cx) {
    trace(cx);                                               // [phase end]
    cons_bar();                                              // every consumer's global stores of this phase are issued ...
    cx.bar_target += (unsigned int)P.G;
    if (cx.ctid == 0) {                                      // ... and ordered before this release (cumulativity over the CTA barrier)
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(P.bar) : "memory");
        unsigned long long t0 = 0;
        while ((int)(ld_acquire_gpu(P.bar) - cx.bar_target) < 0) spin_guard(t0);
    }
    cons_bar();
    trace(cx);                                               // [phase start]
}

// ---- activations of a LayerNorm-consuming phase: block-shared, whole K = d_model resident --------------------------------
// layout: K-chunk kc at abuf + kc * (16 * Kc * 2); row pitch Kc * 2 bytes; 16-byte units of odd rows swap 64-byte halves of every
// 128 bytes so that a quarter-warp's LDS.128 — lanes (g, t) and (g + 1, t) — hits disjoint banks
__device__ __forceinline__ uint32_t a_addr(uint32_t abuf, int kc, int row, int unit, int Kc) {
    const int u = Kc >= 64 ? (unit ^ ((row & 1) << 2)) : unit;
    return abuf + (uint32_t)(kc * 16 * Kc * 2) + (uint32_t)(row * Kc * 2 + u * 16);
}

// A = bf16(LN(x; g, b)).  Thread t owns columns 4t .. 4t+3 of EVERY row.  Two sweeps over x through L2 (R independent 16-byte
// loads per thread and sweep, nothing cached in registers between them — the kernel runs at 96 registers and local memory
// would thrash the 15 KB of L1 left beside 213 KB of shared memory): shifted one-pass statistics (shift = first element of the
// row, as in embed_step_fold_kernel), then normalise + round + store.
__device__ __forceinline__ void prep_ln(const MegaParams& P, const ConsCtx& cx, const float* __restrict__ ln_g, const float* __restrict__ ln_b, float* red /*[MG_CW][2][16] + [16] + [16]*/) {
    const int D = P.D, Kc = chunk_k(D), R = P.R, t = cx.ctid;
    const bool act = t < (D >> 2);
    float* wpart = red;                      // [MG_CW][32]: (s1, s2) per row
    float* mean = red + MG_CW * 32;          // [16]
    float* rstd = mean + 16;                 // [16]
    float4 gm = make_float4(0.f, 0.f, 0.f, 0.f), bt = gm;
    if (act) { gm = __ldg(reinterpret_cast<const float4*>(ln_g) + t); bt = __ldg(reinterpret_cast<const float4*>(ln_b) + t); }
#pragma unroll 4
    for (int r = 0; r < R; r++) {
        const float sh = __ldcg(P.x + (long long)r * D);
        float s1 = 0.f, s2 = 0.f;
        if (act) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(P.x + (long long)r * D) + t);
            const float a = v.x - sh, b = v.y - sh, c = v.z - sh, d = v.w - sh;
            s1 = (a + b) + (c + d); s2 = (a * a + b * b) + (c * c + d * d);
        }
        s1 = warp_sum(s1); s2 = warp_sum(s2);
        if (cx.lane == 0) { wpart[cx.cw * 32 + r] = s1; wpart[cx.cw * 32 + 16 + r] = s2; }
    }
    cons_bar();
    if (cx.cw < R && cx.lane < MG_CW) {
        float s1 = wpart[cx.lane * 32 + cx.cw], s2 = wpart[cx.lane * 32 + 16 + cx.cw];
#pragma unroll
        for (int o = MG_CW / 2; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0x0000ffffu, s1, o); s2 += __shfl_xor_sync(0x0000ffffu, s2, o); }
        if (cx.lane == 0) {
            const float sh = __ldcg(P.x + (long long)cx.cw * D), inv = 1.f / (float)D;
            mean[cx.cw] = sh + s1 * inv;
            rstd[cx.cw] = rsqrtf(fmaxf(s2 - s1 * s1 * inv, 0.f) * inv + P.ln_eps);
        }
    }
    cons_bar();
    if (act) {
        const int k = 4 * t, kc = k / Kc, kk = k - kc * Kc;
#pragma unroll 4
        for (int r = 0; r < R; r++) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(P.x + (long long)r * D) + t);
            const float m = mean[r], rs = rstd[r];
            const float y0 = (v.x - m) * rs * gm.x + bt.x, y1 = (v.y - m) * rs * gm.y + bt.y;
            const float y2 = (v.z - m) * rs * gm.z + bt.z, y3 = (v.w - m) * rs * gm.w + bt.w;
            __nv_bfloat162 lo = __floats2bfloat162_rn(y0, y1), hi = __floats2bfloat162_rn(y2, y3);
            const uint32_t addr = a_addr(cx.abuf, kc, r, kk >> 3, Kc) + (uint32_t)((kk & 4) * 2);
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(*reinterpret_cast<uint32_t*>(&lo)),
                         "r"(*reinterpret_cast<uint32_t*>(&hi)));
        }
    }
    cons_bar();
}

// ---- activations that are a bf16 matrix written earlier in this kernel (ao, hid, hh): WARP-PRIVATE staging -------------------
// Warp w only ever needs the k32-blocks w and w + 16 of each 1024-k chunk: rp rows x 2 blocks x 64 B.  Each warp copies exactly
// those bytes (cp.async through L2) into its own 4 KB of the activation buffer, several chunks ahead, and synchronises with
// nobody but itself.  Slot = one chunk's worth: [row][2 blocks][4 x 16 B], odd rows with the 64-byte halves swapped.
struct WarpA { uint32_t base; int rp, nslots; };
__device__ __forceinline__ WarpA warp_a(const ConsCtx& cx, int R) {
    WarpA w;
    w.rp = R <= 4 ? 4 : (R <= 8 ? 8 : 16);
    w.nslots = (MG_ABUF / MG_CW) / (w.rp * 128);             // 8 / 4 / 2
    w.base = cx.abuf + (uint32_t)cx.cw * (MG_ABUF / MG_CW);
    return w;
}
__device__ __forceinline__ void warp_issue(const MegaParams& P, const ConsCtx& cx, const WarpA& wa, const bf16* A, long long lda, int k0, int nkb, int slot) {
    // units of this warp and chunk: R rows x 2 k32-blocks x 4 (16-byte units)
    for (int i = cx.lane; i < P.R * 8; i += 32) {
        const int row = i >> 3, u = i & 7, kb = cx.cw + MG_CW * (u >> 2);
        if (kb < nkb)
            cp_async16(wa.base + (uint32_t)(slot * wa.rp * 128 + row * 128 + ((u ^ ((row & 1) << 2)) * 16)),
                       A + (long long)row * lda + k0 + kb * 32 + (u & 3) * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <bool HI>
__device__ __forceinline__ void warp_frags(const MegaParams& P, const ConsCtx& cx, const WarpA& wa, int slot, int nkb, uint4 (&alo)[2], uint4 (&ahi)[2]) {
    const int g = cx.lane >> 2, t = cx.lane & 3;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        alo[i] = make_uint4(0, 0, 0, 0); ahi[i] = make_uint4(0, 0, 0, 0);
        if (cx.cw + MG_CW * i < nkb) {
            const uint32_t sb = wa.base + (uint32_t)(slot * wa.rp * 128);
            if (g < P.R) alo[i] = lds128(sb + (uint32_t)(g * 128 + (((i * 4 + t) ^ ((g & 1) << 2)) * 16)));
            if (HI && g + 8 < P.R) ahi[i] = lds128(sb + (uint32_t)((g + 8) * 128 + (((i * 4 + t) ^ ((g & 1) << 2)) * 16)));
        }
    }
}
template <bool HI>
__device__ __forceinline__ void shared_frags(const MegaParams& P, const ConsCtx& cx, int kc, int Kc, int nkb, uint4 (&alo)[2], uint4 (&ahi)[2]) {
    const int g = cx.lane >> 2, t = cx.lane & 3;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int kb = cx.cw + MG_CW * i;
        alo[i] = make_uint4(0, 0, 0, 0); ahi[i] = make_uint4(0, 0, 0, 0);
        if (kb < nkb) {
            if (g < P.R) alo[i] = lds128(a_addr(cx.abuf, kc, g, kb * 4 + t, Kc));
            if (HI && g + 8 < P.R) ahi[i] = lds128(a_addr(cx.abuf, kc, g + 8, kb * 4 + t, Kc));
        }
    }
}

// consume chunk (kc, a): nf feature rows x Kc k in ring stage ci; warp cw owns k32-blocks cw and cw + 16
template <bool HI>
__device__ __forceinline__ void consume_chunk(ConsCtx& cx, int nf, int Kc, int nkb, const uint4 (&alo)[2], const uint4 (&ahi)[2], float (&acc)[4]) {
    const int g = cx.lane >> 2, t = cx.lane & 3;
    mbar_wait(cx.rg.full(cx.ci), cx.rg.parity(cx.ci));
    const uint32_t st = cx.rg.stage(cx.ci) + (uint32_t)(g * Kc * 2);
    const int swz = (Kc >= 64) ? ((g & 1) << 2) : 0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int kb = cx.cw + MG_CW * i;
        if (kb < nkb) {
            uint4 b = make_uint4(0, 0, 0, 0);
            if (g < nf) b = lds128(st + (uint32_t)(((kb * 4 + t) ^ swz) * 16));
            // k permutation shared by A and B: this thread's 8 consecutive k feed the (2t, 2t+1 | 2t+8, 2t+9) slots of two MMAs
            mma16816(acc, alo[i].x, HI ? ahi[i].x : 0u, alo[i].y, HI ? ahi[i].y : 0u, b.x, b.y);
            mma16816(acc, alo[i].z, HI ? ahi[i].z : 0u, alo[i].w, HI ? ahi[i].w : 0u, b.z, b.w);
        }
    }
    __syncwarp();
    if (cx.lane == 0) mbar_arrive(cx.rg.empty(cx.ci));
    cx.ci++;
}

// One GEMV phase on this CTA's feature range.  Inlined exactly ONCE, inside the kernel's flat phase loop: every GEMV phase of every
// layer runs the same instructions (the instruction cache in front of L2 holds 32 KB; one inlined copy per phase measured 300 KB
// of SASS), and all state stays in registers (a real call would put the context in local memory, which has no L1 to live in
// beside 213 KB of shared memory).
template <bool HI>
__device__ __forceinline__ void gemv_phase(const MegaParams& P, ConsCtx& cx, const MegaPhase& ph, const int kind, const int c, float* red) {
    int f0, f1;
    phase_range(kind, ph.N, P.G, c, f0, f1);
    const int nfeat = f1 - f0, ng = (nfeat + 7) >> 3;
    const int Kc = chunk_k(ph.K), nkc = ph.K / Kc, nkb = Kc >> 5;
    float acc[MG_MAXG][4];
#pragma unroll
    for (int a = 0; a < MG_MAXG; a++) { acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0.f; }
    uint4 alo[2], ahi[2];
    const bool ln = ph.ln_g != nullptr, h2 = kind == MP_H2;
    // ---- epilogue operands requested now, consumed after the chunks: bias (and the residual for out-proj / FFN2) ----
    float e_bias[2] = {0.f, 0.f}, e_res[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int i = cx.ctid + j * MG_CW * 32;
        if (i < P.R * nfeat) {
            const int row = i / nfeat, f = f0 + (i - row * nfeat);
            e_bias[j] = __ldg(ph.bias + f);
            if (kind == MP_OUT || kind == MP_FFN2) e_res[j] = __ldcg(P.x + (long long)row * P.D + f);
        }
    }
    // segments: the K-chunks of the activations, or, for the grouped second head layer, the <= 2 codebooks this CTA's groups
    // belong to (codebook cb reads hh[:, cb*Hh ...))
    int nseg = nkc, cb0 = 0;
    const bf16* A = kind == MP_OUT ? P.ao : (h2 ? P.hh : P.hid);
    const long long lda = h2 ? (long long)P.NCB * P.Hh : ph.K;
    const WarpA wa = warp_a(cx, P.R);
    if (ln) {
        prep_ln(P, cx, ph.ln_g, ph.ln_b, red);
    } else {
        if (h2) {
            nseg = 0;
            if (nfeat > 0) { cb0 = f0 / P.V; nseg = (f1 - 1) / P.V - cb0 + 1; }
        }
        // exactly nslots - 1 groups are committed ahead of the loop (empty ones past the last segment): the wait inside the loop
        // then always has nslots - 1 younger groups to leave pending
        for (int sg = 0; sg < wa.nslots - 1; sg++) {
            if (sg < nseg) warp_issue(P, cx, wa, A, lda, h2 ? (cb0 + sg) * P.Hh : sg * Kc, nkb, sg % wa.nslots);
            else asm volatile("cp.async.commit_group;" ::: "memory");
        }
    }
    trace(cx);                                               // [activations staged / requested]
#pragma unroll 1
    for (int sg = 0; sg < nseg; sg++) {
        int a_lo = 0, a_hi = ng;
        if (h2) {                                            // V % 8 == 0 and f0 % 8 == 0: groups never straddle a codebook
            a_lo = max(0, ((cb0 + sg) * P.V - f0) >> 3);
            a_hi = min(ng, ((cb0 + sg + 1) * P.V - f0) >> 3);
        }
        if (ln) {
            shared_frags<HI>(P, cx, sg, Kc, nkb, alo, ahi);
        } else {
            // keep nslots - 1 segments in flight; the slot refilled here was read (by this warp) one iteration ago
            const int nx = sg + wa.nslots - 1;
            if (nx < nseg) warp_issue(P, cx, wa, A, lda, h2 ? (cb0 + nx) * P.Hh : nx * Kc, nkb, nx % wa.nslots);
            else asm volatile("cp.async.commit_group;" ::: "memory");      // (keeps the group count uniform)
            // groups are committed one per iteration: all but the newest (nslots - 1) are complete after this wait
            if (wa.nslots == 8) asm volatile("cp.async.wait_group 7;" ::: "memory");
            else if (wa.nslots == 4) asm volatile("cp.async.wait_group 3;" ::: "memory");
            else asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
            warp_frags<HI>(P, cx, wa, sg % wa.nslots, nkb, alo, ahi);
            __syncwarp();
        }
#pragma unroll
        for (int a = 0; a < MG_MAXG; a++)
            if (a >= a_lo && a < a_hi) consume_chunk<HI>(cx, min(8, nfeat - 8 * a), Kc, nkb, alo, ahi, acc[a]);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");

    // ---- reduce the 16 warps' partial tiles in a fixed order, epilogue ----
    trace(cx);                                               // [chunks consumed]
    cons_bar();                                              // nobody reads the activation buffer any more: it becomes scratch
    {
        const int g = cx.lane >> 2, t = cx.lane & 3;
        const uint32_t sbase = cx.abuf + (uint32_t)(cx.cw * MG_MAXG * 512);
#pragma unroll
        for (int a = 0; a < MG_MAXG; a++)
            if (a < ng) {
                const uint32_t p0 = sbase + (uint32_t)(a * 512 + (g * 8 + 2 * t) * 4);
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(p0), "f"(acc[a][0]), "f"(acc[a][1]) : "memory");
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(p0 + 256u), "f"(acc[a][2]), "f"(acc[a][3]) : "memory");
            }
    }
    cons_bar();
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int i = cx.ctid + j * MG_CW * 32;
        if (i < P.R * nfeat) {
            const int row = i / nfeat, fl = i - row * nfeat;
            const int a = fl >> 3, n = fl & 7;
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < MG_CW; w++) {
                float pv;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pv) : "r"(cx.abuf + (uint32_t)(w * MG_MAXG * 512 + a * 512 + (row * 8 + n) * 4)));
                s += pv;
            }
            const int f = f0 + fl;
            const float v = s + e_bias[j];
            if (kind == MP_QKV) P.qkv[(long long)row * 3 * P.D + f] = v;
            else if (kind == MP_OUT || kind == MP_FFN2) P.x[(long long)row * P.D + f] = v + e_res[j];
            else if (kind == MP_FFN1) P.hid[(long long)row * P.F + f] = __float2bfloat16_rn(fmaxf(v, 0.f));
            else if (kind == MP_H1) P.hh[(long long)row * P.NCB * P.Hh + f] = __float2bfloat16_rn(gelu_erf(v));
            else P.logits[(long long)row * P.NCB * P.V + f] = v;
        }
    }
}

// attention phase: attn_decode_tma_kernel's algorithm on the shared ring (a 64-key K tile and its V tile are consecutive chunks);
// 16 warps x 4 keys per tile
__device__ __forceinline__ void attn_phase(const MegaParams& P, ConsCtx& cx, bf16* kcache, bf16* vcache, const int* pref, int c) {
    const int R = P.R, H = P.H, D = P.D;
    const int total = H * pref[R];
    const int per = (total + P.G - 1) / P.G;
    const int g0 = c * per, g1 = min(total, g0 + per);
    if (g0 >= g1) return;
    const int warp = cx.cw, lane = cx.lane, half = lane >> 4, dl = (lane & 15) * 8, tid = cx.ctid;
    constexpr int NST = 2 * MG_CW;                          // partial softmax states per piece: (warp, half)
    // merge scratch inside the (idle) activation buffer
    const uint32_t s_m = cx.abuf, s_l = cx.abuf + 4 * NST, s_o = cx.abuf + 8 * NST, s_last = s_o + NST * 512;
    const float scale = 0.08838834764831845f;          // 1/sqrt(128)
    TilePos p = locate_tile(pref, R, H, g0);
    int g = g0;
    uint32_t ti = 0;                                         // tiles of this CTA consumed so far (all pieces)
    const uint32_t ci0 = cx.ci;
    while (g < g1) {
        const int r = p.r, h = p.h;
        const int n_old = P.seq_len[r];
        const int t_end = min(p.tiles, p.t + (g1 - g));
        const bool has_new = (t_end == p.tiles);
        float q[8];
        ldcg8(P.qkv + (long long)r * 3 * D + h * 128 + dl, q);
#pragma unroll
        for (int e = 0; e < 8; e++) q[e] *= scale;
        float mrun = -INFINITY, lrun = 0.f, o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (has_new && warp == 0) {
            // this step's K (half 0) / V (half 1) row: rounded to bf16 exactly as later steps will read it back, stored in place
            float nv[8];
            ldcg8(P.qkv + (long long)r * 3 * D + (1 + half) * D + h * 128 + dl, nv);
            bf16* dst = (half ? vcache : kcache) + (((long long)r * H + h) * P.Smax + n_old) * 128 + dl;
            store8(dst, nv);
#pragma unroll
            for (int e = 0; e < 8; e++) nv[e] = __bfloat162float(__float2bfloat16_rn(nv[e]));
            float pn = 0.f;
#pragma unroll
            for (int e = 0; e < 8; e++) pn = fmaf(q[e], nv[e], pn);
            pn += __shfl_xor_sync(0xffffffffu, pn, 1);
            pn += __shfl_xor_sync(0xffffffffu, pn, 2);
            pn += __shfl_xor_sync(0xffffffffu, pn, 4);
            pn += __shfl_xor_sync(0xffffffffu, pn, 8);
            float vv[8];
#pragma unroll
            for (int e = 0; e < 8; e++) vv[e] = __shfl_sync(0xffffffffu, nv[e], (lane & 15) + 16);
            if (half == 0) {
                mrun = pn; lrun = 1.f;
#pragma unroll
                for (int e = 0; e < 8; e++) o[e] = vv[e];
            }
        }
#pragma unroll 1
        for (int t = p.t; t < t_end; t++, ti++) {
            const int nk = min(AT_SUB, n_old - t * AT_SUB);
            const uint32_t ck = ci0 + 2u * ti, cv = ck + 1;
            mbar_wait(cx.rg.full(ck), cx.rg.parity(ck));
            const uint32_t kt = cx.rg.stage(ck) + dl * 2;
            float sc[2];
            float mnew = mrun;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int kl = warp * 4 + j * 2 + half;
                float kk[8];
                lds8_bf16(kt + kl * AT_ROWB, kk);
                float pk = 0.f;
#pragma unroll
                for (int e = 0; e < 8; e++) pk = fmaf(q[e], kk[e], pk);
                pk += __shfl_xor_sync(0xffffffffu, pk, 1);
                pk += __shfl_xor_sync(0xffffffffu, pk, 2);
                pk += __shfl_xor_sync(0xffffffffu, pk, 4);
                pk += __shfl_xor_sync(0xffffffffu, pk, 8);
                sc[j] = kl < nk ? pk : -INFINITY;
                mnew = fmaxf(mnew, sc[j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(cx.rg.empty(ck));
            mbar_wait(cx.rg.full(cv), cx.rg.parity(cv));
            const uint32_t vt = cx.rg.stage(cv) + dl * 2;
            if (mnew > -INFINITY) {
                const float corr = __expf(mrun - mnew);
                lrun *= corr;
#pragma unroll
                for (int e = 0; e < 8; e++) o[e] *= corr;
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int kl = warp * 4 + j * 2 + half;
                    if (kl < nk) {
                        float vv[8];
                        lds8_bf16(vt + kl * AT_ROWB, vv);
                        const float pw = __expf(sc[j] - mnew);
                        lrun += pw;
#pragma unroll
                        for (int e = 0; e < 8; e++) o[e] = fmaf(pw, vv[e], o[e]);
                    }
                }
                mrun = mnew;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(cx.rg.empty(cv));
        }
        // ---- end of the piece: merge the 32 (warp, half) partial states ----
        const int slot = warp * 2 + half;
        if ((lane & 15) == 0) {
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_m + 4u * slot), "f"(mrun) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_l + 4u * slot), "f"(lrun) : "memory");
        }
#pragma unroll
        for (int e = 0; e < 8; e++) asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_o + (uint32_t)(slot * 512 + (dl + e) * 4)), "f"(o[e]) : "memory");
        cons_bar();
        const int G0 = H * pref[r] + h * p.tiles;
        const int zfirst = G0 / per, nsp = (G0 + p.tiles - 1) / per - zfirst + 1, z = c - zfirst;
        const int d = tid;
        float M = -INFINITY, L = 0.f, O = 0.f;
        if (d < 128) {
#pragma unroll 8
            for (int w = 0; w < NST; w++) { float mw; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mw) : "r"(s_m + 4u * w)); M = fmaxf(M, mw); }
#pragma unroll 4
            for (int w = 0; w < NST; w++) {
                float mw, lw, ow;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mw) : "r"(s_m + 4u * w));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(lw) : "r"(s_l + 4u * w));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(ow) : "r"(s_o + (uint32_t)(w * 512 + d * 4)));
                const float wgt = mw > -INFINITY ? __expf(mw - M) : 0.f;
                L += lw * wgt;
                O += ow * wgt;
            }
        }
        bf16* op = P.ao + (long long)r * D + h * 128 + d;
        if (nsp == 1) {
            if (d < 128) *op = __float2bfloat16_rn(O / L);
        } else {
            float* wsp = P.attn_ws + ((long long)(r * H + h) * P.max_pieces + z) * 130;
            if (d < 128) {
                wsp[2 + d] = O;
                if (d == 0) { wsp[0] = M; wsp[1] = L; }
                __threadfence();
            }
            cons_bar();
            if (tid == 0) {
                const int tk = atomicAdd(&P.tickets[r * H + h], 1);
                const int last = (tk == nsp - 1);
                if (last) P.tickets[r * H + h] = 0;
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(s_last), "r"(last) : "memory");
            }
            cons_bar();
            int last;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(last) : "r"(s_last));
            if (last && d < 128) {
                __threadfence();
                const float* wb = P.attn_ws + (long long)(r * H + h) * P.max_pieces * 130;
                float M2 = -INFINITY;
                for (int j = 0; j < nsp; j++) M2 = fmaxf(M2, __ldcg(wb + j * 130));
                float L2 = 0.f, O2 = 0.f;
                for (int j = 0; j < nsp; j++) {                    // piece order: deterministic
                    const float mj = __ldcg(wb + j * 130);
                    const float wgt = mj > -INFINITY ? __expf(mj - M2) : 0.f;
                    L2 += __ldcg(wb + j * 130 + 1) * wgt;
                    O2 += __ldcg(wb + j * 130 + 2 + d) * wgt;
                }
                *op = __float2bfloat16_rn(O2 / L2);
            }
        }
        cons_bar();                                          // the merge scratch is rewritten by the next piece
        g += t_end - p.t;
        p.t = t_end;
        if (p.t == p.tiles) advance_stream(p, pref, R, H);
    }
    cx.ci = ci0 + 2u * ti;
}

template <bool HI>
__global__ void __launch_bounds__(MG_THREADS, 1) decode_mega_kernel(const MegaParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int s_pref[20];
    __shared__ float s_red[MG_CW * 32 + 32];
    const uint32_t base = (smem_u32(smem) + 127u) & ~127u;
    Ring rg;
    rg.base = base; rg.full0 = base + MG_NS * MG_STAGE + MG_ABUF; rg.empty0 = rg.full0 + 8u * MG_NS;
    const uint32_t plan_bar = rg.empty0 + 8u * MG_NS;
    const uint32_t abuf = base + MG_NS * MG_STAGE;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, c = blockIdx.x;
    pdl_launch_dependents();
    if (tid == 0) {
        for (int s = 0; s < MG_NS; s++) { mbar_init(rg.full0 + 8u * s, 1); mbar_init(rg.empty0 + 8u * s, MG_CW); }
        mbar_init(plan_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // ===== producer: the whole iteration's byte stream of this CTA, in consumption order, never waiting for a phase =====
        if (lane == 0) {
            uint32_t ci = 0;
            bool planned = false;
            for (int l = 0; l < P.L; l++) {
                const MegaLayer ly = P.layers[l];
                produce_gemv(rg, ci, ly.qkv, MP_QKV, P.G, c);           // (layer 0: issued before the dependency — weights are immutable)
                if (!planned) { mbar_wait(plan_bar, 0); planned = true; }   // tile plan (row lengths, finished rows) from the consumers
                produce_attn(rg, ci, P, ly, s_pref, c);
                produce_gemv(rg, ci, ly.out, MP_OUT, P.G, c);
                produce_gemv(rg, ci, ly.ffn1, MP_FFN1, P.G, c);
                produce_gemv(rg, ci, ly.ffn2, MP_FFN2, P.G, c);
            }
            produce_gemv(rg, ci, P.h1, MP_H1, P.G, c);
            produce_gemv(rg, ci, P.h2, MP_H2, P.G, c);
        }
        return;
    }

    // ===== consumers =====
    ConsCtx cx;
    cx.rg = rg; cx.ci = 0; cx.abuf = abuf; cx.cw = warp - 1; cx.lane = lane; cx.ctid = tid - 32; cx.bar_target = 0;
    cx.tr_n = 0; cx.tr = g_mega_trace ? g_mega_trace + (long long)c * g_mega_trace_cap : nullptr;
    trace(cx);                                               // [kernel entry]
    pdl_wait();
    trace(cx);                                               // [dependency resolved]                                              // x of this iteration (embedding kernel); row state of the previous one
    if (cx.ctid == 0) {
        int acc = 0;
        s_pref[0] = 0;
        for (int r = 0; r < P.R; r++) {                      // tiles per live row (the same for its H heads), prefix
            const int n_old = P.seq_len[r];
            acc += P.st[r / P.rpu].done ? 0 : max(1, (n_old + AT_SUB - 1) / AT_SUB);
            s_pref[r + 1] = acc;
        }
    }
    cons_bar();
    if (cx.ctid == 0) mbar_arrive(plan_bar);

    // flat phase loop: 5 phases per layer (QKV, attention, out-proj, FFN1, FFN2) + the two head layers; ONE copy of each phase body
    const int n_phases = 5 * P.L + 2;
#pragma unroll 1
    for (int pi = 0; pi < n_phases; pi++) {
        if (pi > 0) grid_barrier(P, cx);                     // the previous phase's outputs of every CTA
        const int l = pi / 5, q = pi - 5 * l;
        if (l < P.L && q == 1) {
            attn_phase(P, cx, P.layers[l].kc, P.layers[l].vc, s_pref, c);
        } else {
            MegaPhase ph;
            int kind;
            if (l >= P.L) { ph = q == 0 ? P.h1 : P.h2; kind = q == 0 ? MP_H1 : MP_H2; }
            else {
                const MegaLayer* ly = P.layers + l;
                const MegaPhase* pp = q == 0 ? &ly->qkv : (q == 2 ? &ly->out : (q == 3 ? &ly->ffn1 : &ly->ffn2));
                ph = *pp;
                kind = q == 0 ? MP_QKV : (q == 2 ? MP_OUT : (q == 3 ? MP_FFN1 : MP_FFN2));
            }
            gemv_phase<HI>(P, cx, ph, kind, c, s_red);
        }
    }
    trace(cx);                                               // [kernel exit]
}

// W [N, K] bf16 row-major -> the order the CTAs stream it: CTA c's features [f0, f1) contiguous; inside, k-chunk major, then
// feature, then k (16-byte units of odd feature rows of a group swap 64-byte halves: bank-conflict-free B fragments)
__global__ void mega_pack_kernel(const bf16* __restrict__ W, bf16* __restrict__ out, int N, int K, int kind, int G) {
    const int Kc = chunk_k(K), upr = K >> 3;
    const long long n_units = (long long)N * upr;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_units; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i / upr), unit = (int)(i - (long long)f * upr);
        // owner of feature f
        int c = kind == MP_H2 ? (int)(((long long)(f / 8) * G) / (N / 8)) : (int)(((long long)f * G) / N);
        int f0, f1;
        for (;;) {
            if (c < 0) c = 0;
            if (c >= G) c = G - 1;
            phase_range(kind, N, G, c, f0, f1);
            if (f < f0) c--; else if (f >= f1) c++; else break;
        }
        const int nfeat = f1 - f0, fl = f - f0, g = fl & 7;
        const int k = unit * 8, kc = k / Kc, ul = (k - kc * Kc) >> 3;
        const int us = Kc >= 64 ? (ul ^ ((g & 1) << 2)) : ul;
        const long long dst = (long long)f0 * K + (long long)kc * nfeat * Kc + (long long)fl * Kc + us * 8;
        *reinterpret_cast<uint4*>(out + dst) = *reinterpret_cast<const uint4*>(W + (long long)f * K + k);
    }
}

bool phase_ok(int N, int K, int kind, int G) {
    if (K % 32 != 0 || (K > MG_KC && K % MG_KC != 0)) return false;
    if (kind == MP_H2 && N % 8 != 0) return false;
    const int units = kind == MP_H2 ? N / 8 : N;
    const int per = (units + G - 1) / G * (kind == MP_H2 ? 8 : 1);
    return (per + 7) / 8 <= MG_MAXG;
}

}  // namespace

// ---- host interface (lm_engine.cu) --------------------------------------------------------------------------------------
int mega_grid(int* G_out) {
    static int G = -1;
    if (G < 0) {
        int dev = 0, n_sm = 0, occ = 0;
        SSRB_CUDA(cudaGetDevice(&dev));
        SSRB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        SSRB_CUDA(cudaFuncSetAttribute(decode_mega_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MG_SMEM));
        SSRB_CUDA(cudaFuncSetAttribute(decode_mega_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MG_SMEM));
        SSRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, decode_mega_kernel<true>, MG_THREADS, MG_SMEM));
        G = occ >= 1 ? n_sm : 0;
    }
    *G_out = G;
    return 0;
}

bool mega_supported(int R, int D, int H, int F, int NCB, int V, int Hh) {
    int G = 0;
    if (mega_grid(&G) || G <= 0) return false;
    if (R < 1 || R > 16 || D != H * 128 || D % 128 != 0 || D > 2048) return false;
    return phase_ok(3 * D, D, MP_QKV, G) && phase_ok(D, D, MP_OUT, G) && phase_ok(F, D, MP_FFN1, G) && phase_ok(D, F, MP_FFN2, G) &&
           phase_ok(NCB * Hh, D, MP_H1, G) && phase_ok(NCB * V, Hh, MP_H2, G) && V % 8 == 0;
}

int mega_pack(const void* W, void* out, int N, int K, int kind, cudaStream_t s) {
    int G = 0;
    SSRB_TRY(mega_grid(&G));
    SSRB_CHECK(G > 0 && phase_ok(N, K, kind, G), "mega_pack: unsupported shape");
    SSRB_LAUNCH(mega_pack_kernel, 1024, 256, 0, s, (const bf16*)W, (bf16*)out, N, K, kind, G);
    return 0;
}

int launch_mega(const MegaArgs& a, cudaStream_t s) {
    int G = 0;
    SSRB_TRY(mega_grid(&G));
    SSRB_CHECK(G > 0, "decode_mega_kernel does not fit on this device");
    MegaParams P{};
    P.R = a.R; P.D = a.D; P.H = a.H; P.F = a.F; P.L = a.L; P.NCB = a.NCB; P.V = a.V; P.Hh = a.Hh; P.Smax = a.Smax; P.rpu = a.rpu;
    P.G = G; P.max_pieces = a.max_pieces;
    P.x = a.x; P.qkv = a.qkv; P.ao = (bf16*)a.ao; P.hid = (bf16*)a.hid; P.hh = (bf16*)a.hh; P.logits = a.logits;
    P.seq_len = a.seq_len; P.st = a.st; P.attn_ws = a.attn_ws; P.tickets = a.tickets; P.bar = a.bar;
    P.layers = (const MegaLayer*)a.layers_dev;
    P.h1 = MegaPhase{(const bf16*)a.h1_w, a.h1_b, a.lnf_g, a.lnf_b, a.NCB * a.Hh, a.D};
    P.h2 = MegaPhase{(const bf16*)a.h2_w, a.h2_b, nullptr, nullptr, a.NCB * a.V, a.Hh};
    P.ln_eps = 1e-5f;
    if (a.R > 8) return launch_pdl(decode_mega_kernel<true>, dim3(G), dim3(MG_THREADS), MG_SMEM, s, 1, P);
    return launch_pdl(decode_mega_kernel<false>, dim3(G), dim3(MG_THREADS), MG_SMEM, s, 1, P);
}

int mega_trace_arm(unsigned long long* dev_buf, int cap_per_cta) {
    SSRB_CUDA(cudaMemcpyToSymbol(g_mega_trace, &dev_buf, sizeof(dev_buf)));
    SSRB_CUDA(cudaMemcpyToSymbol(g_mega_trace_cap, &cap_per_cta, sizeof(cap_per_cta)));
    return 0;
}

size_t mega_layer_bytes() { return sizeof(MegaLayer); }

int mega_fill_layer(void* host_slot, const MegaLayerHost& h) {
    MegaLayer m{};
    m.qkv = MegaPhase{(const bf16*)h.wqkv, h.bqkv, h.ln1g, h.ln1b, 3 * h.D, h.D};
    m.out = MegaPhase{(const bf16*)h.wo, h.bo, nullptr, nullptr, h.D, h.D};
    m.ffn1 = MegaPhase{(const bf16*)h.w1, h.b1, h.ln2g, h.ln2b, h.F, h.D};
    m.ffn2 = MegaPhase{(const bf16*)h.w2, h.b2, nullptr, nullptr, h.D, h.F};
    m.kc = (bf16*)h.kc; m.vc = (bf16*)h.vc;
    memcpy(host_slot, &m, sizeof(m));
    return 0;
}

}  // namespace ssrb
