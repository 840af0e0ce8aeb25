// resblock_tc.cu — a whole SEANetResnetBlock as ONE depth-fused tcgen05 kernel (bf16 decoder / skip-encoder path).
//
//        y = x + conv_k1( ELU( conv_k3( ELU(x) ) ) )                 audiocraft/modules/seanet.py:16-60 (true_skip, 1 residual layer)
//
// Per tile: 128 time steps of one utterance, ALL channels (persistent CTAs walk the tile list).  The hidden activation h = ELU(conv_k3(ELU(x)) + b1) never leaves the SM:
//   GEMM 1  [128 x 3C] . W1'^T -> acc1 [128 x C/2] in tensor memory     (three tap-GEMMs over row-shifted TMA views, as conv_tc.cu)
//   epilogue 1: acc1 + b1 -> ELU -> bf16 -> shared memory, written directly in the K-major 128-byte-swizzled layout UMMA reads
//   GEMM 2  h [128 x C/2] (shared memory) . W2^T -> acc2 [128 x C]      (W2 streamed through the same TMA ring)
//   epilogue 2: acc2 + b2 + x (raw skip) -> raw and / or ELU'd bf16 channels-last stores
// Against two conv_tc launches this removes the HBM round trip of h and one kernel boundary on the critical path; the persistent
// tile loop pays barrier init / TMEM allocation / pipeline fill once per CTA instead of once per tile (they dominated these short-K
// layers: ncu 9-18 % warps active with one tile per CTA, profiles/r02c_ncu_full_codec_kernels.csv).
#include <cuda.h>

#include <mutex>

#include "conv_tc.cuh"

namespace ssrb {

namespace {

constexpr int BK = 64, ROWS = 128, A_BYTES = ROWS * BK * 2;

struct RbMaps { CUtensorMap a, w1, w2; };
struct RbParams {
    int T, tiles_r, tiles_total;                            // tile t -> row tile t % tiles_r of utterance t / tiles_r
    const float* b1; const float* b2;
    const bf16* x_raw; long long x_bstride, x_off;          // skip: x_raw[b * x_bstride + x_off + row * C + c]
    bf16* out_raw; bf16* out_act; long long out_bstride, out_off;
};

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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((unsigned long long)map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((unsigned long long)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {     // K-major SWIZZLE_128B (see gemm_tc.cu)
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float (&v)[16]) {      // caller waits (tcgen05.wait::ld) once per batch
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
                   "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
                 : "r"(taddr));
}

constexpr int EPI_WARPS = 8;                  // two per TMEM lane group, each takes half of the columns
constexpr int THREADS = (2 + EPI_WARPS) * 32;

template <int C> struct RbCfg {
    static constexpr int CH = C / 2;
    static constexpr int N1 = CH < 64 ? 64 : CH;             // hidden channels, padded to one k-block (zero weight rows / columns)
    static constexpr int KB1 = C / BK, NK1 = 3 * KB1;        // k-blocks per tap / of GEMM 1
    static constexpr int KB2 = N1 / BK;                      // k-blocks of GEMM 2
    static constexpr int N2H = C > 256 ? 256 : C;            // UMMA N of GEMM 2 (two instructions side by side when C = 512)
    static constexpr int STAGE1 = A_BYTES + N1 * BK * 2, STAGE2 = C * BK * 2;
    static constexpr int STAGE = STAGE1 > STAGE2 ? STAGE1 : STAGE2;
    static constexpr int STAGES = C == 64 ? 3 : (C == 256 ? 4 : 2);
    static constexpr int H_BYTES = ROWS * N1 * 2;
    // tensor memory: acc2 [128 x C] at column 0; acc1 [128 x N1] next to it when both fit in 512 columns, so that GEMM 1 of the
    // next tile runs under epilogue 2 of this one; C = 512 has to alias them (GEMM 1 then waits for epilogue 2's reads)
    static constexpr bool ALIAS = C + N1 > 512;
    static constexpr int ACC1_OFF = ALIAS ? 0 : C;
    static constexpr int TMEM_NEED = ALIAS ? (C > N1 ? C : N1) : C + N1;
    static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
    static constexpr int CTAS_PER_SM = C <= 128 ? 2 : 1;
    // epilogue 2: a warp owns 32 rows x C/2 columns, in chunks of CHUNK columns through a staging slab (row = CHUNK bf16 + 16 B pad)
    static constexpr int CW2 = C / 2, CHUNK = (C == 64 || C == 512) ? 32 : 64, NCH = CW2 / CHUNK, LPR = CHUNK / 8;
    static constexpr int SLAB_ROWB = CHUNK * 2 + 16, SLAB_BYTES = 32 * SLAB_ROWB;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE + H_BYTES + EPI_WARPS * SLAB_BYTES + 1024 + 256;
};

// PERSISTENT: CTAs walk the tile list (tile = 128 time steps of one utterance) with stride gridDim.x; the TMA producer, the MMA
// issuer and the 8 epilogue warps each keep running counters, so the loads and GEMM 1 of tile i+1 run under epilogue 2 of tile i.
template <int C>
__global__ void __launch_bounds__(THREADS, RbCfg<C>::CTAS_PER_SM) resblock_tc_kernel(const __grid_constant__ RbMaps maps, const RbParams prm) {
    using Cfg = RbCfg<C>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t hbase = base + Cfg::STAGES * Cfg::STAGE;                      // 1024-aligned: STAGE is a multiple of 1024
    const uint32_t slab_base = hbase + Cfg::H_BYTES;
    const uint32_t bar_base = slab_base + EPI_WARPS * Cfg::SLAB_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
    const uint32_t acc1_full = bar_base + 8u * (2 * Cfg::STAGES), h_ready = acc1_full + 8u, acc2_full = h_ready + 8u,
                   acc2_empty = acc2_full + 8u, tmem_slot = acc2_empty + 8u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(acc1_full, 1); mbar_init(acc2_full, 1); mbar_init(h_ready, EPI_WARPS); mbar_init(acc2_empty, EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: per tile the GEMM 1 stages (activation tile of tap q + W1 tile), then the GEMM 2 stages (W2 k-blocks) =====
            int it = 0;
            for (int t = blockIdx.x; t < prm.tiles_total; t += gridDim.x) {
                const int row0 = (t % prm.tiles_r) * ROWS, b = t / prm.tiles_r;
                for (int i = 0; i < Cfg::NK1 + Cfg::KB2; i++, it++) {
                    const int s = it % Cfg::STAGES;
                    mbar_wait(empty_bar(s), (uint32_t)(((it / Cfg::STAGES) & 1) ^ 1));
                    const uint32_t sp = base + s * Cfg::STAGE;
                    if (i < Cfg::NK1) {
                        const int q = i / Cfg::KB1, cb = i - q * Cfg::KB1;
                        mbar_expect_tx(full_bar(s), Cfg::STAGE1);
                        tma_load_3d(sp, &maps.a, full_bar(s), cb * BK, row0 + q, b);
                        tma_load_2d(sp + A_BYTES, &maps.w1, full_bar(s), cb * BK, q * Cfg::N1);
                    } else {
                        const int j = i - Cfg::NK1;
                        mbar_expect_tx(full_bar(s), Cfg::STAGE2);
                        tma_load_2d(sp, &maps.w2, full_bar(s), j * BK, 0);
                        if (C > 256) tma_load_2d(sp + 256 * BK * 2, &maps.w2, full_bar(s), j * BK, 256);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Cfg::N1 >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Cfg::N2H >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
            const uint32_t acc1 = tmem_base + Cfg::ACC1_OFF, acc2 = tmem_base;
            int it = 0, tl = 0;
            for (int t = blockIdx.x; t < prm.tiles_total; t += gridDim.x, tl++) {
                const uint32_t tph = (uint32_t)(tl & 1);
                if (Cfg::ALIAS) {                                   // acc1 shares columns with acc2: the previous tile's epilogue 2 must have read them
                    mbar_wait(acc2_empty, tph ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                for (int i = 0; i < Cfg::NK1; i++, it++) {
                    const int s = it % Cfg::STAGES;
                    mbar_wait(full_bar(s), (uint32_t)((it / Cfg::STAGES) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sp = base + s * Cfg::STAGE;
                    const uint64_t da = make_desc(sp), db = make_desc(sp + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; k++) umma_bf16(acc1, da + 2 * k, db + 2 * k, idesc1, (i > 0 || k > 0) ? 1u : 0u);
                    umma_commit(empty_bar(s));
                }
                umma_commit(acc1_full);
                mbar_wait(h_ready, tph);                            // h is in shared memory (and acc1 has been read out)
                if (!Cfg::ALIAS) mbar_wait(acc2_empty, tph ^ 1);    // the previous tile's epilogue 2 has read acc2
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int j = 0; j < Cfg::KB2; j++, it++) {
                    const int s = it % Cfg::STAGES;
                    mbar_wait(full_bar(s), (uint32_t)((it / Cfg::STAGES) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sp = base + s * Cfg::STAGE;
                    const uint64_t da = make_desc(hbase + j * A_BYTES), db = make_desc(sp);
#pragma unroll
                    for (int k = 0; k < BK / 16; k++) {
                        umma_bf16(acc2, da + 2 * k, db + 2 * k, idesc2, (j > 0 || k > 0) ? 1u : 0u);
                        if (C > 256) umma_bf16(acc2 + 256, da + 2 * k, make_desc(sp + 256 * BK * 2) + 2 * k, idesc2, (j > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(s));
                }
                umma_commit(acc2_full);
            }
        }
    } else {
        // ===== epilogue warps: thread = one time step (row); warps 2..9 cover every (lane group, column half) pair once =====
        const int ew = warp - 2, lg = warp & 3, half = ew >> 2, rl = lg * 32 + lane;
        uint8_t* slab = smem_raw + (slab_base - smem_u32(smem_raw)) + ew * Cfg::SLAB_BYTES;
        const int rsub = lane / Cfg::LPR, piece = lane % Cfg::LPR;            // staged copies: LPR lanes cover one staging row
        constexpr int RSTEP = 32 / Cfg::LPR;
        int tl = 0;
        for (int t = blockIdx.x; t < prm.tiles_total; t += gridDim.x, tl++) {
            const int row0 = (t % prm.tiles_r) * ROWS, b = t / prm.tiles_r;
            const uint32_t tph = (uint32_t)(tl & 1);
            const int row = row0 + rl;
            const bool rok = row < prm.T;
            const uint32_t tlane = tmem_base + ((uint32_t)(lg * 32) << 16);
            // ---- epilogue 1: h = ELU(acc1 + b1) -> bf16 -> shared memory in the K-major SWIZZLE_128B layout (16-byte chunk c of row r
            // sits at chunk c ^ (r & 7) of the row's 128 bytes), k-block jb at hbase + jb * 16 KB; this warp's half of the N1 columns ----
            mbar_wait(acc1_full, tph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c0 = half * (Cfg::N1 / 2); c0 < (half + 1) * (Cfg::N1 / 2); c0 += 16) {
                float v[16];
                tmem_ld16_issue(tlane + Cfg::ACC1_OFF + c0, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    // channels past C/2 (padding of the 32-channel hidden layer) carry zero weights and a zero bias: ELU(0) = 0
                    const float4 bv = __ldg(reinterpret_cast<const float4*>(prm.b1 + c0 + 4 * j));
                    const float a0 = elu1_bf16(v[4 * j] + bv.x), a1 = elu1_bf16(v[4 * j + 1] + bv.y);
                    const float a2 = elu1_bf16(v[4 * j + 2] + bv.z), a3 = elu1_bf16(v[4 * j + 3] + bv.w);
                    __nv_bfloat162 h01 = __floats2bfloat162_rn(rok ? a0 : 0.f, rok ? a1 : 0.f), h23 = __floats2bfloat162_rn(rok ? a2 : 0.f, rok ? a3 : 0.f);
                    pk[2 * j] = *reinterpret_cast<uint32_t*>(&h01);
                    pk[2 * j + 1] = *reinterpret_cast<uint32_t*>(&h23);
                }
                const int jb = c0 >> 6, ch = (c0 & 63) >> 3;        // k-block, first of the two 16-byte chunks
                const uint32_t rowb = hbase + (uint32_t)(jb * A_BYTES + rl * 128);
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(rowb + (uint32_t)(((ch) ^ (rl & 7)) * 16)), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(rowb + (uint32_t)(((ch + 1) ^ (rl & 7)) * 16)), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // generic-proxy stores -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) mbar_arrive(h_ready);
            // ---- epilogue 2: y = acc2 + b2 + x, this warp's C/2 columns in chunks; residual in and both outputs go through the staging
            // slab so that every global access instruction covers whole row segments ----
            mbar_wait(acc2_full, tph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long xrow0 = (long long)b * prm.x_bstride + prm.x_off + (long long)(row0 + lg * 32) * C;
            const long long orow0 = (long long)b * prm.out_bstride + prm.out_off + (long long)(row0 + lg * 32) * C;
            const bool full_rows = row0 + lg * 32 + 32 <= prm.T;
#pragma unroll 1
            for (int ch = 0; ch < Cfg::NCH; ch++) {
                const int col0 = half * Cfg::CW2 + ch * Cfg::CHUNK;
                float v[Cfg::CHUNK];
#pragma unroll
                for (int c = 0; c < Cfg::CHUNK / 16; c++) tmem_ld16_issue(tlane + col0 + c * 16, *reinterpret_cast<float(*)[16]>(&v[c * 16]));
                // the skip rows of this chunk: coalesced global -> slab
#pragma unroll
                for (int j = 0; j < Cfg::LPR; j++) {
                    const int rr = j * RSTEP + rsub;
                    uint4 w = make_uint4(0u, 0u, 0u, 0u);
                    if (full_rows || row0 + lg * 32 + rr < prm.T) w = *reinterpret_cast<const uint4*>(prm.x_raw + xrow0 + (long long)rr * C + col0 + piece * 8);
                    *reinterpret_cast<uint4*>(slab + rr * Cfg::SLAB_ROWB + piece * 16) = w;
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (ch == Cfg::NCH - 1) {                           // acc2 is in registers: the tensor core may overwrite it
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc2_empty);
                } else {
                    __syncwarp();
                }
#pragma unroll
                for (int j = 0; j < Cfg::CHUNK; j += 8) {
                    float r8[8];
                    load8(reinterpret_cast<const bf16*>(slab + lane * Cfg::SLAB_ROWB) + j, r8);
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(prm.b2 + col0 + j)), b1 = __ldg(reinterpret_cast<const float4*>(prm.b2 + col0 + j + 4));
                    v[j] += r8[0] + b0.x; v[j + 1] += r8[1] + b0.y; v[j + 2] += r8[2] + b0.z; v[j + 3] += r8[3] + b0.w;
                    v[j + 4] += r8[4] + b1.x; v[j + 5] += r8[5] + b1.y; v[j + 6] += r8[6] + b1.z; v[j + 7] += r8[7] + b1.w;
                }
                __syncwarp();
#pragma unroll
                for (int pass = 0; pass < 2; pass++) {
                    bf16* outp = pass ? prm.out_act : prm.out_raw;
                    if (!outp) continue;
#pragma unroll
                    for (int j = 0; j < Cfg::CHUNK; j += 8) {
                        float e8[8];
#pragma unroll
                        for (int e = 0; e < 8; e++) e8[e] = pass ? elu1_bf16(v[j + e]) : v[j + e];
                        store8(reinterpret_cast<bf16*>(slab + lane * Cfg::SLAB_ROWB) + j, e8);
                    }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < Cfg::LPR; j++) {
                        const int rr = j * RSTEP + rsub;
                        if (full_rows || row0 + lg * 32 + rr < prm.T)
                            *reinterpret_cast<uint4*>(outp + orow0 + (long long)rr * C + col0 + piece * 8) =
                                *reinterpret_cast<const uint4*>(slab + rr * Cfg::SLAB_ROWB + piece * 16);
                    }
                    __syncwarp();
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn_rb() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

template <int C>
int launch_rb(const RbMaps& maps, const RbParams& prm, dim3 grid, int n_sm, cudaStream_t s) {
    static bool done = false;
    if (!done) {
        SSRB_CUDA(cudaFuncSetAttribute(resblock_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RbCfg<C>::SMEM));
        done = true;
    }
    static_assert(RbCfg<C>::SMEM <= 232448, "resblock_tc: shared memory budget");
    grid.x = std::min<unsigned>(grid.x, RbCfg<C>::CTAS_PER_SM * (unsigned)n_sm);
    SSRB_LAUNCH(resblock_tc_kernel<C>, grid, THREADS, RbCfg<C>::SMEM, s, maps, prm);
    return 0;
}

}  // namespace

bool resblock_tc_supported(int C) { return C == 64 || C == 128 || C == 256 || C == 512; }

int resblock_tc(const ResblockTcArgs& a, cudaStream_t s) {
    EncodeTiledFn fn = encode_fn_rb();
    SSRB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
    SSRB_CHECK(resblock_tc_supported(a.C), "resblock_tc: channel count must be 64, 128, 256 or 512");
    const int C = a.C, N1 = C / 2 < 64 ? 64 : C / 2;
    SSRB_CHECK(a.w1_N == N1 && a.w1_Cw == C && a.w2_N == C && a.w2_Cw == N1, "resblock_tc: weight repack does not match the kernel's tiling");
    RbMaps maps;
    memset(&maps, 0, sizeof(maps));
    {
        cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)a.rows_v, (cuuint64_t)a.B};
        cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)a.x_bstride * 2};
        cuuint32_t box[3] = {BK, ROWS, 1}, estr[3] = {1, 1, 1};
        CUresult r = fn(&maps.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(a.x_act + a.x_base_off), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SSRB_CHECK(r == CUDA_SUCCESS, "resblock_tc: activation tensor map failed");
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)3 * N1};
        cuuint64_t strides[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {BK, (cuuint32_t)N1}, estr[2] = {1, 1};
        CUresult r = fn(&maps.w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(a.w1), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SSRB_CHECK(r == CUDA_SUCCESS, "resblock_tc: W1 tensor map failed");
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)N1, (cuuint64_t)C};
        cuuint64_t strides[1] = {(cuuint64_t)N1 * 2};
        cuuint32_t box[2] = {BK, (cuuint32_t)(C > 256 ? 256 : C)}, estr[2] = {1, 1};
        CUresult r = fn(&maps.w2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(a.w2), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SSRB_CHECK(r == CUDA_SUCCESS, "resblock_tc: W2 tensor map failed");
    }
    RbParams p{};
    p.T = a.T; p.b1 = a.b1; p.b2 = a.b2;
    p.x_raw = a.x_raw; p.x_bstride = a.x_bstride; p.x_off = a.x_raw_off;
    p.out_raw = a.out_raw; p.out_act = a.out_act; p.out_bstride = a.out_bstride; p.out_off = a.out_off;
    p.tiles_r = cdiv(a.T, ROWS);
    const long long total = (long long)p.tiles_r * a.B;
    SSRB_CHECK(total > 0 && total < (1ll << 31), "resblock_tc: tile count out of range");
    p.tiles_total = (int)total;
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        SSRB_CUDA(cudaGetDevice(&dev));
        SSRB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    dim3 grid((unsigned)total);                             // capped to the resident CTA count in launch_rb
    switch (C) {
        case 64: return launch_rb<64>(maps, p, grid, n_sm, s);
        case 128: return launch_rb<128>(maps, p, grid, n_sm, s);
        case 256: return launch_rb<256>(maps, p, grid, n_sm, s);
        default: return launch_rb<512>(maps, p, grid, n_sm, s);
    }
}

}  // namespace ssrb
