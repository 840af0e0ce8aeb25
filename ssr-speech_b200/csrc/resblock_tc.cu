// resblock_tc.cu — a whole SEANetResnetBlock as ONE depth-fused tcgen05 kernel (bf16 decoder / skip-encoder path).
//
//        y = x + conv_k1( ELU( conv_k3( ELU(x) ) ) )                 audiocraft/modules/seanet.py:16-60 (true_skip, 1 residual layer)
//
// Per CTA: 128 time steps of one utterance, ALL channels.  The hidden activation h = ELU(conv_k3(ELU(x)) + b1) never leaves the SM:
//   GEMM 1  [128 x 3C] . W1'^T -> acc1 [128 x C/2] in tensor memory     (three tap-GEMMs over row-shifted TMA views, as conv_tc.cu)
//   epilogue 1: acc1 + b1 -> ELU -> bf16 -> shared memory, written directly in the K-major 128-byte-swizzled layout UMMA reads
//   GEMM 2  h [128 x C/2] (shared memory) . W2^T -> acc2 [128 x C]      (W2 streamed through the same TMA ring)
//   epilogue 2: acc2 + b2 + x (raw skip) -> raw and / or ELU'd bf16 channels-last stores
// Against two conv_tc launches this removes the HBM round trip of h, one kernel boundary on the critical path and half of the
// per-CTA fixed costs (barrier init, TMEM allocation, pipeline fill) that dominate these short-K layers (ncu: 4.7-30 % tensor
// pipe active, profiles/r01e_ncu_full_conv_tc_codec.csv).
#include <cuda.h>

#include <mutex>

#include "conv_tc.cuh"

namespace ssrb {

namespace {

constexpr int BK = 64, ROWS = 128, A_BYTES = ROWS * BK * 2;

struct RbMaps { CUtensorMap a, w1, w2; };
struct RbParams {
    int T;
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

template <int C> struct RbCfg {
    static constexpr int CH = C / 2;
    static constexpr int N1 = CH < 64 ? 64 : CH;             // hidden channels, padded to one k-block (zero weight rows / columns)
    static constexpr int KB1 = C / BK, NK1 = 3 * KB1;        // k-blocks per tap / of GEMM 1
    static constexpr int KB2 = N1 / BK;                      // k-blocks of GEMM 2
    static constexpr int N2H = C > 256 ? 256 : C;            // UMMA N of GEMM 2 (two instructions side by side when C = 512)
    static constexpr int STAGE1 = A_BYTES + N1 * BK * 2, STAGE2 = C * BK * 2;
    static constexpr int STAGE = STAGE1 > STAGE2 ? STAGE1 : STAGE2;
    static constexpr int STAGES = C >= 512 ? 2 : 3;
    static constexpr int H_BYTES = ROWS * N1 * 2;
    static constexpr int TMEM_COLS = C < 64 ? 64 : C;        // max(N1, C), a power of two >= 32
    static constexpr size_t SMEM = (size_t)STAGES * STAGE + H_BYTES + 1024 + 256;
};

template <int C>
__global__ void __launch_bounds__(192) resblock_tc_kernel(const __grid_constant__ RbMaps maps, const RbParams prm) {
    using Cfg = RbCfg<C>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t hbase = base + Cfg::STAGES * Cfg::STAGE;                      // 1024-aligned: STAGE is a multiple of 1024
    const uint32_t bar_base = hbase + Cfg::H_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
    const uint32_t acc1_bar = bar_base + 8u * (2 * Cfg::STAGES), acc2_bar = acc1_bar + 8u, h_bar = acc2_bar + 8u, tmem_slot = h_bar + 8u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * ROWS, b = blockIdx.y;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(acc1_bar, 1); mbar_init(acc2_bar, 1); mbar_init(h_bar, 128);
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
            // ===== TMA producer: GEMM 1 stages (activation tile of tap q + W1 tile), then GEMM 2 stages (W2 k-blocks), one ring =====
            for (int i = 0; i < Cfg::NK1 + Cfg::KB2; i++) {
                const int s = i % Cfg::STAGES;
                const uint32_t ph = (i / Cfg::STAGES) & 1;
                mbar_wait(empty_bar(s), ph ^ 1);
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
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Cfg::N1 >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Cfg::N2H >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
            for (int i = 0; i < Cfg::NK1; i++) {
                const int s = i % Cfg::STAGES;
                mbar_wait(full_bar(s), (i / Cfg::STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sp = base + s * Cfg::STAGE;
                const uint64_t da = make_desc(sp), db = make_desc(sp + A_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; k++) umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc1, (i > 0 || k > 0) ? 1u : 0u);
                umma_commit(empty_bar(s));
            }
            umma_commit(acc1_bar);
            mbar_wait(h_bar, 0);                                 // h is in shared memory (and acc1 has been read out)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int j = 0; j < Cfg::KB2; j++) {
                const int i = Cfg::NK1 + j, s = i % Cfg::STAGES;
                mbar_wait(full_bar(s), (i / Cfg::STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sp = base + s * Cfg::STAGE;
                const uint64_t da = make_desc(hbase + j * A_BYTES), db = make_desc(sp);
#pragma unroll
                for (int k = 0; k < BK / 16; k++) {
                    umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc2, (j > 0 || k > 0) ? 1u : 0u);
                    if (C > 256) umma_bf16(tmem_base + 256, da + 2 * k, make_desc(sp + 256 * BK * 2) + 2 * k, idesc2, (j > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(empty_bar(s));
            }
            umma_commit(acc2_bar);
        }
    } else {
        // ===== epilogue warps: thread = one time step (row) =====
        const int lg = warp & 3, rl = lg * 32 + lane;
        const int row = row0 + rl;
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16);
        const bool rok = row < prm.T;
        // ---- epilogue 1: h = ELU(acc1 + b1) -> bf16 -> shared memory in the K-major SWIZZLE_128B layout (16-byte chunk c of row r
        // sits at chunk c ^ (r & 7) of the row's 128 bytes), k-block jb at hbase + jb * 16 KB ----
        mbar_wait(acc1_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c0 = 0; c0 < Cfg::N1; c0 += 16) {
            float v[16];
            tmem_ld16(taddr + c0, v);
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                // channels past C/2 (padding of the 32-channel hidden layer) carry zero weights and a zero bias: ELU(0) = 0
                const float a0 = elu1_bf16(v[2 * j] + prm.b1[c0 + 2 * j]), a1 = elu1_bf16(v[2 * j + 1] + prm.b1[c0 + 2 * j + 1]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(rok ? a0 : 0.f, rok ? a1 : 0.f);
                pk[j] = *reinterpret_cast<uint32_t*>(&h2);
            }
            const int jb = c0 >> 6, ch = (c0 & 63) >> 3;        // k-block, first of the two 16-byte chunks
            const uint32_t rowb = hbase + (uint32_t)(jb * A_BYTES + rl * 128);
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(rowb + (uint32_t)(((ch) ^ (rl & 7)) * 16)), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(rowb + (uint32_t)(((ch + 1) ^ (rl & 7)) * 16)), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // generic-proxy stores -> visible to the tensor core's reads
        mbar_arrive(h_bar);
        // ---- epilogue 2: y = acc2 + b2 + x ----
        mbar_wait(acc2_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c0 = 0; c0 < C; c0 += 16) {
            float v[16];
            tmem_ld16(taddr + c0, v);
            if (!rok) continue;
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] += prm.b2[c0 + j];
            const bf16* rp = prm.x_raw + (long long)b * prm.x_bstride + prm.x_off + (long long)row * C + c0;
            float r8[8];
            load8(rp, r8);
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] += r8[j];
            load8(rp + 8, r8);
#pragma unroll
            for (int j = 0; j < 8; j++) v[8 + j] += r8[j];
            const long long o = (long long)b * prm.out_bstride + prm.out_off + (long long)row * C + c0;
            float lo[8], hi[8];
            if (prm.out_raw) {
#pragma unroll
                for (int j = 0; j < 8; j++) { lo[j] = v[j]; hi[j] = v[8 + j]; }
                store8(prm.out_raw + o, lo);
                store8(prm.out_raw + o + 8, hi);
            }
            if (prm.out_act) {
#pragma unroll
                for (int j = 0; j < 8; j++) { lo[j] = elu1_bf16(v[j]); hi[j] = elu1_bf16(v[8 + j]); }
                store8(prm.out_act + o, lo);
                store8(prm.out_act + o + 8, hi);
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
int launch_rb(const RbMaps& maps, const RbParams& prm, dim3 grid, cudaStream_t s) {
    static bool done = false;
    if (!done) {
        SSRB_CUDA(cudaFuncSetAttribute(resblock_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RbCfg<C>::SMEM));
        done = true;
    }
    SSRB_LAUNCH(resblock_tc_kernel<C>, grid, 192, RbCfg<C>::SMEM, s, maps, prm);
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
    dim3 grid(cdiv(a.T, ROWS), a.B);
    SSRB_CHECK(grid.y <= 65535, "resblock_tc: batch too large");
    switch (C) {
        case 64: return launch_rb<64>(maps, p, grid, s);
        case 128: return launch_rb<128>(maps, p, grid, s);
        case 256: return launch_rb<256>(maps, p, grid, s);
        default: return launch_rb<512>(maps, p, grid, s);
    }
}

}  // namespace ssrb
