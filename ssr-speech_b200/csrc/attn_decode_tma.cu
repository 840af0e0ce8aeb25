// attn_decode_tma.cu — bf16 decode attention with bulk-async (TMA engine) staging of the K/V stream.
//
// One query per (row, head) against the in-place cache — replaces F.scaled_dot_product_attention at
// models/modules/activation.py:634 for tgt_len == 1, plus the KV "append" the reference performs by
// re-materialising the cache (activation.py:626-631, ssr.py:685-686).
//
// PERSISTENT, BALANCED: the launch is 2 CTAs per SM, and the whole layer's work — every 64-key K+V tile of every live
// (row, head) stream, in memory order — is cut into equal contiguous ranges, one per CTA.  A pure read stream reaches
// 7.3-7.5 TB/s on this part (tools/microbench/read_bw.cu) but only 6.7 TB/s when it is issued as 1024 CTAs of 448 KB
// (3.46 waves of cold-started rings); one long-lived ring per CTA that keeps streaming across stream boundaries removes
// both the wave tail and the per-CTA ramp.
//   * warp 8 is the producer: one lane issues cp.async.bulk copies of 64-key K and V tiles (16 KB each) into a 3-stage
//     shared-memory ring, each stage guarded by a full (tx-count) and an empty mbarrier; it needs nothing from the QKV
//     GEMM right before this kernel, so it starts before griddepcontrol.wait (see `prefetch` below);
//   * warps 0-7 consume: 8 keys per warp and tile, 16 lanes x 16 B per key row (conflict-free), fp32 online softmax;
//     at the end of a stream (or of the CTA's range) the 16 partial states meet in shared memory;
//   * a stream that lies inside one CTA's range is finished there; one that straddles a range boundary leaves
//     (max, sum, acc[128]) per piece in a workspace and the last piece to arrive (ticket) merges them in piece order;
//   * the piece that owns a stream's last tile takes this step's K/V row from the QKV GEMM output, rounds it to bf16,
//     stores it into the cache in place and scores it from registers (the bulk copies never read bytes written by
//     this kernel).
// The cut points depend on every row's length, so in the last bits a row's result depends on the batch around it;
// given the same batch state it is deterministic (fixed merge order) — the fp32 parity mode uses attn_decode_kernel.
// HBM roofline: algorithmic bytes per launch = R*H*(S+1)*2*128*2 B.
#include "lm_kernels.cuh"

namespace ssrb {

namespace {

#ifndef AT_SUB_KEYS
#define AT_SUB_KEYS 64         // keys per tile and ring depth (A/B'd: 64 x 3 against 32 x 6, profiles/r02d_summary.md)
#define AT_NSTAGES 3
#endif
constexpr int AT_SUB = AT_SUB_KEYS, AT_ROWB = 256, AT_STAGES = AT_NSTAGES;            // keys per tile, bytes per key row
constexpr int AT_JP = AT_SUB / 16;                                  // key pairs per consumer warp and tile
constexpr int AT_STAGE_BYTES = 2 * AT_SUB * AT_ROWB;                // K + V
constexpr int AT_SMEM = AT_STAGES * AT_STAGE_BYTES + 16 * AT_STAGES + 16;
constexpr int AT_CW = 8;                                            // consumer warps
constexpr int AT_THREADS = (AT_CW + 1) * 32;
constexpr int AT_MAXR = 1024;                                       // rows per launch (prefix table in shared memory)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
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
__device__ __forceinline__ void lds8_bf16(uint32_t addr, float (&v)[8]) {
    uint32_t w0, w1, w2, w3;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(addr));
    const uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

// position of a global tile index: row r, head h, tile t of that stream
struct TilePos { int r, h, t, tiles; };

__device__ __forceinline__ TilePos locate_tile(const int* pref /*[R+1], tiles before row r (per head)*/, int R, int H, int g) {
    int lo = 0, hi = R - 1;                                  // last row with H*pref[r] <= g
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (H * pref[mid] <= g) lo = mid; else hi = mid - 1;
    }
    // (a row without tiles shares its successor's offset, and "last" skips it)
    TilePos p;
    p.r = lo; p.tiles = pref[lo + 1] - pref[lo];
    const int rem = g - H * pref[lo];
    p.h = rem / p.tiles; p.t = rem - p.h * p.tiles;
    return p;
}
__device__ __forceinline__ void advance_stream(TilePos& p, const int* pref, int R, int H) {   // to tile 0 of the next live stream
    p.t = 0;
    if (++p.h == H) {
        p.h = 0;
        do { p.r++; } while (p.r < R && pref[p.r + 1] == pref[p.r]);
        p.tiles = p.r < R ? pref[p.r + 1] - pref[p.r] : 1;
    }
}

__global__ void __launch_bounds__(AT_THREADS, 2) attn_decode_tma_kernel(const float* __restrict__ qkv, int D, int H, bf16* kc, bf16* vc,
                                                                        int Smax, const int* __restrict__ seq_len,
                                                                        const UttState* __restrict__ st, int rpu, float* ws,
                                                                        int* __restrict__ tickets, bf16* __restrict__ out, int R,
                                                                        int max_pieces, int prefetch) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int s_pref[AT_MAXR + 1];
    __shared__ float sm_m[2 * AT_CW], sm_l[2 * AT_CW], sm_o[2 * AT_CW][128];
    __shared__ int sm_last;
    pdl_launch_dependents();
    const int ts = ts_begin(TSK_ATTN);
    // prefetch == 0: wait for the predecessor before reading anything.  prefetch == 1 (decode chain): the row state (done,
    // seq_len) was written by the previous iteration's sampler and the cached K/V rows [0, n_old) by earlier iterations; the
    // embed kernel that opens every iteration releases its dependents only AFTER its own griddepcontrol.wait, so no kernel
    // of this iteration — this one included — can start before the previous iteration has completed and flushed.  Only q
    // and this step's K/V row come from the QKV GEMM right before: the consumers wait for it, the producer warp starts
    // streaming the cache at once, so the K/V stream overlaps the tail of the GEMM.
    if (!prefetch) { pdl_wait(); ts_dep(ts); }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, half = lane >> 4, dl = (lane & 15) * 8;
    // ---- tiles per live row (the same for its H heads) and their prefix ----------------------------------------------
    for (int r = tid; r < R; r += AT_THREADS) {
        const int n_old = seq_len[r];
        s_pref[r + 1] = st[r / rpu].done ? 0 : max(1, (n_old + AT_SUB - 1) / AT_SUB);
    }
    const uint32_t ring = smem_u32(smem), bar0 = ring + AT_STAGES * AT_STAGE_BYTES;   // full[s] = bar0+8s, empty[s] = bar0+8*AT_STAGES+8s
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < AT_STAGES; s++) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar0 + 8 * AT_STAGES + 8 * s, AT_CW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {                                         // inclusive scan of s_pref[1..R] in chunks of 32
        int carry = 0;
        for (int base = 0; base < R; base += 32) {
            const int r = base + lane;
            int v = r < R ? s_pref[r + 1] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += u;
            }
            if (r < R) s_pref[r + 1] = v + carry;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
        if (lane == 0) s_pref[0] = 0;
    }
    __syncthreads();
    const int total = H * s_pref[R];
    const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x;
    const int g0 = (int)blockIdx.x * per, g1 = min(total, g0 + per);
    if (g0 >= g1) { if (prefetch) pdl_wait(); return; }      // every CTA waits: completion stays transitive along the chain

    if (warp == AT_CW) {
        // ===== producer: the CTA's whole tile range through one ring, across stream boundaries =====
        if (lane == 0) {
            TilePos p = locate_tile(s_pref, R, H, g0);
            int n_old = seq_len[p.r];
            for (int i = 0; i < g1 - g0; i++) {
                const int s = i % AT_STAGES;
                const uint32_t ph = (i / AT_STAGES) & 1;
                mbar_wait(bar0 + 8 * AT_STAGES + 8 * s, ph ^ 1);
                const int k0 = p.t * AT_SUB;
                const int nk = min(AT_SUB, n_old - k0);
                if (nk > 0) {
                    const uint32_t bytes = (uint32_t)nk * AT_ROWB;
                    const int64_t off = (((int64_t)p.r * H + p.h) * Smax + k0) * 128;
                    mbar_expect_tx(bar0 + 8 * s, 2 * bytes);
                    bulk_g2s(ring + s * AT_STAGE_BYTES, kc + off, bytes, bar0 + 8 * s);
                    bulk_g2s(ring + s * AT_STAGE_BYTES + AT_SUB * AT_ROWB, vc + off, bytes, bar0 + 8 * s);
                } else {
                    mbar_arrive(bar0 + 8 * s);               // a row with no cached key: an empty tile
                }
                if (++p.t == p.tiles) {
                    advance_stream(p, s_pref, R, H);
                    if (p.r < R) n_old = seq_len[p.r];
                }
            }
        }
        return;
    }

    // ===== consumers (warps 0..AT_CW-1) =====
    if (prefetch) { pdl_wait(); ts_dep(ts); }
    const float scale = 0.08838834764831845f;          // 1/sqrt(128)
    TilePos p = locate_tile(s_pref, R, H, g0);
    int g = g0, i = 0;
    while (g < g1) {
        const int r = p.r, h = p.h;
        const int n_old = seq_len[r];
        const int t_end = min(p.tiles, p.t + (g1 - g));                       // this piece: tiles [p.t, t_end) of stream (r, h)
        const bool has_new = (t_end == p.tiles);
        float q[8];
        load8(qkv + (int64_t)r * 3 * D + h * 128 + dl, q);
#pragma unroll
        for (int e = 0; e < 8; e++) q[e] *= scale;
        float mrun = -INFINITY, lrun = 0.f, o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (has_new && warp == 0) {
            // this step's K (half 0) / V (half 1) row: round to bf16 exactly as later steps will read it back
            float nv[8];
            load8(qkv + (int64_t)r * 3 * D + (1 + half) * D + h * 128 + dl, nv);
            bf16* dst = (half ? vc : kc) + (((int64_t)r * H + h) * Smax + n_old) * 128 + dl;
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
        for (int t = p.t; t < t_end; t++, i++) {
            const int s = i % AT_STAGES;
            const uint32_t ph = (i / AT_STAGES) & 1;
            const int nk = min(AT_SUB, n_old - t * AT_SUB);
            mbar_wait(bar0 + 8 * s, ph);
            const uint32_t kt = ring + s * AT_STAGE_BYTES + dl * 2, vt = kt + AT_SUB * AT_ROWB;
            float sc[AT_JP];
            float mnew = mrun;
#pragma unroll
            for (int j = 0; j < AT_JP; j++) {
                const int kl = warp * (AT_SUB / AT_CW) + j * 2 + half;
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
            if (mnew > -INFINITY) {
                const float corr = __expf(mrun - mnew);
                lrun *= corr;
#pragma unroll
                for (int e = 0; e < 8; e++) o[e] *= corr;
#pragma unroll
                for (int j = 0; j < AT_JP; j++) {
                    const int kl = warp * (AT_SUB / AT_CW) + j * 2 + half;
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
            if (lane == 0) mbar_arrive(bar0 + 8 * AT_STAGES + 8 * s);     // this warp is done with the stage
        }
        // ---- end of the piece: merge the 16 (warp, half) partial states (named barrier: the producer warp is elsewhere)
        const int slot = warp * 2 + half;
        if ((lane & 15) == 0) { sm_m[slot] = mrun; sm_l[slot] = lrun; }
#pragma unroll
        for (int e = 0; e < 8; e++) sm_o[slot][dl + e] = o[e];
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // pieces of this stream: the CTA ranges its global tiles [G0, G0 + tiles) intersect
        const int G0 = H * s_pref[r] + h * p.tiles;
        const int zfirst = G0 / per, nsp = (G0 + p.tiles - 1) / per - zfirst + 1, z = (int)blockIdx.x - zfirst;
        const int d = tid;
        float M = -INFINITY, L = 0.f, O = 0.f;
        if (d < 128) {
#pragma unroll
            for (int w = 0; w < 2 * AT_CW; w++) M = fmaxf(M, sm_m[w]);
#pragma unroll
            for (int w = 0; w < 2 * AT_CW; w++) {
                const float wgt = __expf(sm_m[w] - M);
                L += sm_l[w] * wgt;
                O += sm_o[w][d] * wgt;
            }
        }
        bf16* op = out + (int64_t)r * D + h * 128 + d;
        if (nsp == 1) {
            if (d < 128) *op = __float2bfloat16_rn(O / L);
        } else {
            float* wsp = ws + ((int64_t)(r * H + h) * max_pieces + z) * 130;
            if (d < 128) {
                wsp[2 + d] = O;
                if (d == 0) { wsp[0] = M; wsp[1] = L; }
                __threadfence();
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid == 0) {
                const int tk = atomicAdd(&tickets[r * H + h], 1);
                sm_last = (tk == nsp - 1);
                if (sm_last) tickets[r * H + h] = 0;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (sm_last && d < 128) {
                __threadfence();
                // the partials are requested 8 pieces at a time (independent loads: one L2 round trip per 8 pieces, not one per
                // piece); an absent piece reads as (m, l, o) = (-inf, 0, 0) and adds exact zeros
                const float* wb = ws + (int64_t)(r * H + h) * max_pieces * 130;
                float M2 = -INFINITY;
                for (int j0 = 0; j0 < nsp; j0 += 8) {
                    float mm[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) mm[j] = j0 + j < nsp ? __ldcg(wb + (j0 + j) * 130) : -INFINITY;
#pragma unroll
                    for (int j = 0; j < 8; j++) M2 = fmaxf(M2, mm[j]);
                }
                float L2 = 0.f, O2 = 0.f;
                for (int j0 = 0; j0 < nsp; j0 += 8) {              // piece order: deterministic
                    float mm[8], ll[8], oo[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const bool ok = j0 + j < nsp;
                        const float* pp = wb + (ok ? j0 + j : j0) * 130;
                        mm[j] = ok ? __ldcg(pp) : -INFINITY;
                        ll[j] = ok ? __ldcg(pp + 1) : 0.f;
                        oo[j] = ok ? __ldcg(pp + 2 + d) : 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float wgt = __expf(mm[j] - M2);
                        L2 += ll[j] * wgt;
                        O2 += oo[j] * wgt;
                    }
                }
                *op = __float2bfloat16_rn(O2 / L2);
            }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");      // sm_* are rewritten by the next piece
        g += t_end - p.t;
        p.t = t_end;
        if (p.t == p.tiles) advance_stream(p, s_pref, R, H);
    }
    ts_end(ts);
}

}  // namespace

// a stream is cut into at most one piece per tile (tiny batches: one tile per CTA)
int attn_decode_tma_max_nsplit(int Smax) { return cdiv(Smax, AT_SUB) + 1; }

int launch_attn_decode_tma(const float* qkv, int R, int D, int H, void* kcache, void* vcache, int Smax, const int* seq_len,
                           const UttState* st, int rpu, float* ws, int* tickets, void* out, int prefetch, cudaStream_t s) {
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        SSRB_CUDA(cudaGetDevice(&dev));
        SSRB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        SSRB_CUDA(cudaFuncSetAttribute(attn_decode_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    }
    SSRB_CHECK(R <= AT_MAXR, "attn_decode: too many rows for one launch");
    // SSRB_ATTN_CTAS_PER_SM=1: one persistent CTA per SM instead of two (leaves room for another stream's GEMM CTAs; probe switch)
    static const int per_sm = [] { const char* e = getenv("SSRB_ATTN_CTAS_PER_SM"); return (e && e[0] == '1') ? 1 : 2; }();
    return launch_pdl(attn_decode_tma_kernel, dim3(per_sm * n_sm), dim3(AT_THREADS), AT_SMEM, s, 1, qkv, D, H, (bf16*)kcache, (bf16*)vcache, Smax,
                      seq_len, st, rpu, ws, tickets, (bf16*)out, R, attn_decode_tma_max_nsplit(Smax), prefetch);
}

int ts_arm_attn_tma(const TsBuf& t) { return ts_arm_tu(t); }

}  // namespace ssrb
