// conv_tc.cu — SEANet convolutions as tap-accumulated GEMMs on the tcgen05 tensor cores (bf16 in, fp32 accumulate).
//
// Activations live channels-last in HBM: X[b][G + t][c] bf16 with G zero guard rows on both sides of every
// utterance, so that every convolution of the codec becomes
//
//        out[b][row][n] = bias[n] + sum_{q < taps} sum_{c < Cw}  view(b)[row + q][c] * W'[q][n][c]
//
// where `view` is a re-interpretation of the same memory (no im2col, no copies):
//   * StreamableConv1d stride 1, kernel k   (conv.py:185-201): Cw = Cin,   taps = k, view row j = buffer row G - padL + j
//   * StreamableConv1d stride s, kernel 2s  (conv.py:185-201): Cw = s*Cin, taps = 2, view row j = s consecutive time
//     steps starting at t = j*s - padL  (W'[q][n][r*Cin + c] = W[n][c][q*s + r])
//   * StreamableConvTranspose1d s, k = 2s   (conv.py:221-243): Cw = Cin,   taps = 2 (x[i-1], x[i]), N = s*Cout with
//     n = phase*Cout + co — one GEMM row holds the s output steps produced by input step i, which IS the channels-last
//     layout of the output; the reference's trim (left = total - total//2) is an address offset plus a validity window.
// The zero guards implement the reference's zero padding (pad_mode 'constant'); TMA walks the shifted rows directly.
// Pipeline: TMA (3-D map {Cw, rows, batch}, 128B swizzle) -> 3-stage smem ring -> tcgen05.mma 128 x NT x 16 -> TMEM ->
// epilogue (bias | per-mark bias, residual add, raw and/or ELU'd bf16 store).  ELU (seanet.py:39-46) is applied by the
// PRODUCER's epilogue, so consumers read ready-made GEMM operands.
#include <cuda.h>

#include <mutex>

#include "conv_tc.cuh"

namespace ssrb {

namespace {

constexpr int BK = 64, ROWS = 128, A_BYTES = ROWS * BK * 2, STAGES = 3;

struct Maps { CUtensorMap a, w; };

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

struct KParams {
    int T_rows, N, nkb_per_tap, taps;
    const float* bias; const float* bias_alt; int bias_mod;
    const long long* marks; int marks_T, marks_rep;
    const bf16* res; long long res_bstride, res_off;
    bf16* out_raw; bf16* out_act; long long out_bstride, out_off, valid_lo, valid_hi;
};

template <int NT>
__global__ void __launch_bounds__(192) conv_tc_kernel(const __grid_constant__ Maps maps, const KParams prm) {
    constexpr int W_BYTES = NT * BK * 2, STAGE_BYTES = A_BYTES + W_BYTES, TMEM_COLS = NT;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = base + STAGES * STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    const uint32_t accum_bar = bar_base + 8u * (2 * STAGES), tmem_slot = accum_bar + 8u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * NT, row0 = blockIdx.y * ROWS, b = blockIdx.z;
    const int nk = prm.taps * prm.nkb_per_tap;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nk; i++) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                const int q = i / prm.nkb_per_tap, cb = i - q * prm.nkb_per_tap;
                mbar_wait(empty_bar(s), ph ^ 1);
                mbar_expect_tx(full_bar(s), STAGE_BYTES);
                const uint32_t sp = base + s * STAGE_BYTES;
                tma_load_3d(sp, &maps.a, full_bar(s), cb * BK, row0 + q, b);
                tma_load_2d(sp + A_BYTES, &maps.w, full_bar(s), cb * BK, q * prm.N + n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
            for (int i = 0; i < nk; i++) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(full_bar(s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sp = base + s * STAGE_BYTES;
                const uint64_t da = make_desc(sp), db = make_desc(sp + A_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; k++) umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
                umma_commit(empty_bar(s));
            }
            umma_commit(accum_bar);
        }
    } else {
        const int lg = warp & 3;
        const int row = row0 + lg * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16);
        mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool rok = row < prm.T_rows;
        const float* bias = prm.bias;
        if (prm.marks && rok) {
            int mi = row / prm.marks_rep;
            if (mi >= prm.marks_T) mi = prm.marks_T - 1;
            if (prm.marks[(long long)b * prm.marks_T + mi] != 0) bias = prm.bias_alt;
        }
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 16) {
            float v[16];
            tmem_ld16(taddr + c0, v);
            if (!rok) continue;
            const int n = n0 + c0;
            const long long f = (long long)row * prm.N + n;                   // flat index inside the GEMM output
            if (f + 16 <= prm.valid_lo || f >= prm.valid_hi) continue;
            const int nb = n % prm.bias_mod;                                  // bias_mod is a multiple of 16
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] += bias[nb + j];
            if (prm.res) {
                const bf16* rp = prm.res + (long long)b * prm.res_bstride + prm.res_off + f;
                float r8[8];
                load8(rp, r8);
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] += r8[j];
                load8(rp + 8, r8);
#pragma unroll
                for (int j = 0; j < 8; j++) v[8 + j] += r8[j];
            }
            const long long o = (long long)b * prm.out_bstride + prm.out_off + f;
            if (f >= prm.valid_lo && f + 16 <= prm.valid_hi) {
                float lo[8], hi[8];
                if (prm.out_raw) {
#pragma unroll
                    for (int j = 0; j < 8; j++) { lo[j] = v[j]; hi[j] = v[8 + j]; }
                    store8(prm.out_raw + o, lo);
                    store8(prm.out_raw + o + 8, hi);
                }
                if (prm.out_act) {
#pragma unroll
                    for (int j = 0; j < 8; j++) { lo[j] = elu1(v[j]); hi[j] = elu1(v[8 + j]); }
                    store8(prm.out_act + o, lo);
                    store8(prm.out_act + o + 8, hi);
                }
            } else {
                for (int j = 0; j < 16; j++) {
                    if (f + j < prm.valid_lo || f + j >= prm.valid_hi) continue;
                    if (prm.out_raw) prm.out_raw[o + j] = __float2bfloat16_rn(v[j]);
                    if (prm.out_act) prm.out_act[o + j] = __float2bfloat16_rn(elu1(v[j]));
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
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

template <int NT>
int launch(const Maps& maps, const KParams& prm, dim3 grid, cudaStream_t s) {
    constexpr size_t SMEM = (size_t)STAGES * (A_BYTES + NT * BK * 2) + 1024 + 256;
    static bool done = false;
    if (!done) {
        SSRB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        done = true;
    }
    SSRB_LAUNCH(conv_tc_kernel<NT>, grid, 192, SMEM, s, maps, prm);
    return 0;
}

}  // namespace

int conv_tc(const ConvTcArgs& a, cudaStream_t s) {
    EncodeTiledFn fn = encode_fn();
    SSRB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
    SSRB_CHECK(a.Cw % BK == 0 && a.N % 64 == 0 && a.bias_mod % 16 == 0, "conv_tc: channel counts must be multiples of 64");
    SSRB_CHECK(((uintptr_t)a.x & 15) == 0 && (a.x_base_off % 8) == 0 && (a.x_bstride % 8) == 0, "conv_tc: input view must be 16B aligned");
    SSRB_CHECK((a.out_off % 8) == 0 && (a.out_bstride % 8) == 0 && (a.valid_lo % 8) == 0, "conv_tc: output view must be 16B aligned");
    const int NT = (a.N % 128 == 0) ? 128 : 64;
    Maps maps;
    memset(&maps, 0, sizeof(maps));
    {
        cuuint64_t dims[3] = {(cuuint64_t)a.Cw, (cuuint64_t)a.rows_v, (cuuint64_t)a.B};
        cuuint64_t strides[2] = {(cuuint64_t)a.Cw * 2, (cuuint64_t)a.x_bstride * 2};
        cuuint32_t box[3] = {BK, ROWS, 1}, estr[3] = {1, 1, 1};
        CUresult r = fn(&maps.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(a.x + a.x_base_off), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SSRB_CHECK(r == CUDA_SUCCESS, "conv_tc: activation tensor map failed");
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)a.Cw, (cuuint64_t)a.taps * a.N};
        cuuint64_t strides[1] = {(cuuint64_t)a.Cw * 2};
        cuuint32_t box[2] = {BK, (cuuint32_t)NT}, estr[2] = {1, 1};
        CUresult r = fn(&maps.w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(a.w), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SSRB_CHECK(r == CUDA_SUCCESS, "conv_tc: weight tensor map failed");
    }
    KParams p{};
    p.T_rows = a.T_rows; p.N = a.N; p.nkb_per_tap = a.Cw / BK; p.taps = a.taps;
    p.bias = a.bias; p.bias_alt = a.bias_alt ? a.bias_alt : a.bias; p.bias_mod = a.bias_mod;
    p.marks = a.marks; p.marks_T = a.marks_T; p.marks_rep = a.marks_rep > 0 ? a.marks_rep : 1;
    p.res = a.res; p.res_bstride = a.res_bstride; p.res_off = a.res_off;
    p.out_raw = a.out_raw; p.out_act = a.out_act; p.out_bstride = a.out_bstride; p.out_off = a.out_off;
    p.valid_lo = a.valid_lo; p.valid_hi = a.valid_hi;
    dim3 grid(a.N / NT, cdiv(a.T_rows, ROWS), a.B);
    SSRB_CHECK(grid.z <= 65535 && grid.y <= 65535, "conv_tc: grid too large");
    return NT == 128 ? launch<128>(maps, p, grid, s) : launch<64>(maps, p, grid, s);
}

}  // namespace ssrb
