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
// Pipeline (persistent CTAs, see conv_tc_kernel): TMA (3-D map {Cw, rows, batch}, 128B swizzle) -> 2-stage smem ring ->
// tcgen05.mma 128 x NT x 16 -> two TMEM accumulators -> 8 epilogue warps (bias | per-mark bias, residual add, raw and/or ELU'd
// bf16, staged through shared memory into whole-line stores).  ELU (seanet.py:39-46) is applied by the PRODUCER's epilogue,
// so consumers read ready-made GEMM operands.
#include <cuda.h>

#include <mutex>

#include "conv_tc.cuh"

namespace ssrb {

namespace {

constexpr int BK = 64, ROWS = 128, A_BYTES = ROWS * BK * 2, STAGES = 2;
constexpr int EPI_WARPS = 8;                  // two per TMEM lane group, each takes half of the tile's columns
constexpr int THREADS = (2 + EPI_WARPS) * 32;
constexpr int epi_rowb(int NT) { return NT + 16; }               // staging row: NT/2 bf16 + 16 B pad (conflict-free 16 B stores per quarter warp)
constexpr int epi_bytes(int NT) { return EPI_WARPS * 32 * epi_rowb(NT); }   // one 32-row slab per epilogue warp

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
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float (&v)[16]) {      // caller waits (tcgen05.wait::ld) once for a batch
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
                   "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
                 : "r"(taddr));
}

struct KParams {
    int T_rows, N, nkb_per_tap, taps;
    int tiles_n, tiles_r, tiles_total;                 // tile t -> n-tile t % tiles_n, row-tile (t / tiles_n) % tiles_r, utterance t / (tiles_n * tiles_r)
    const float* bias; const float* bias_alt; int bias_mod;
    const long long* marks; int marks_T, marks_rep;
    const bf16* res; long long res_bstride, res_off;
    bf16* out_raw; bf16* out_act; long long out_bstride, out_off, valid_lo, valid_hi;
};

// PERSISTENT: two CTAs per SM walk the tile list with stride gridDim.x.  Three roles, each with its own running counters, so the
// TMA loads of tile i+1 and its MMAs run under the epilogue of tile i:
//   warp 0 (one lane)  TMA producer  : k-blocks (tap q, channel block cb) of A [128 rows] and W [NT rows] into a 2-stage ring
//   warp 1 (one lane)  MMA issuer    : accumulates a tile into TMEM buffer (tile & 1) (2 x NT columns), commits tfull[buf]
//   warps 2-9          epilogue      : TMEM -> registers (a thread owns a row, a warp half of the columns), releases the buffer as soon as the
//                                      tile is in registers, then bias / residual / ELU, and stores through a per-warp staging
//                                      slab so that every global store instruction writes whole 128-byte lines (a thread-per-row
//                                      store touches 32 lines per instruction and made the LSU the bottleneck).
template <int NT>
__global__ void __launch_bounds__(THREADS, 2) conv_tc_kernel(const __grid_constant__ Maps maps, const KParams prm) {
    constexpr int W_BYTES = NT * BK * 2, STAGE_BYTES = A_BYTES + W_BYTES, TMEM_COLS = 2 * NT;
    constexpr int CW = NT / 2, ROWB = epi_rowb(NT), LPR = CW / 8;      // columns per epilogue warp, staging row bytes, lanes per staging row
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t epi_base = base + STAGES * STAGE_BYTES;
    const uint32_t bar_base = epi_base + epi_bytes(NT);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = prm.taps * prm.nkb_per_tap;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; b++) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), EPI_WARPS); }
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
            int it = 0;
            for (int t = blockIdx.x; t < prm.tiles_total; t += gridDim.x) {
                const int n0 = (t % prm.tiles_n) * NT, tr = t / prm.tiles_n, row0 = (tr % prm.tiles_r) * ROWS, b = tr / prm.tiles_r;
                for (int i = 0; i < nk; i++, it++) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    const int q = i / prm.nkb_per_tap, cb = i - q * prm.nkb_per_tap;
                    mbar_wait(empty_bar(s), ph ^ 1);
                    mbar_expect_tx(full_bar(s), STAGE_BYTES);
                    const uint32_t sp = base + s * STAGE_BYTES;
                    tma_load_3d(sp, &maps.a, full_bar(s), cb * BK, row0 + q, b);
                    tma_load_2d(sp + A_BYTES, &maps.w, full_bar(s), cb * BK, q * prm.N + n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
            int it = 0, tl = 0;
            for (int t = blockIdx.x; t < prm.tiles_total; t += gridDim.x, tl++) {
                const int buf = tl & 1;
                mbar_wait(tempty_bar(buf), (uint32_t)(((tl >> 1) & 1) ^ 1));          // the epilogue has read this buffer's previous tile
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + (uint32_t)(buf * NT);
                for (int i = 0; i < nk; i++, it++) {
                    const int s = it % STAGES;
                    mbar_wait(full_bar(s), (uint32_t)((it / STAGES) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sp = base + s * STAGE_BYTES;
                    const uint64_t da = make_desc(sp), db = make_desc(sp + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; k++) umma_bf16(tacc, da + 2 * k, db + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
                    umma_commit(empty_bar(s));
                }
                umma_commit(tfull_bar(buf));
            }
        }
    } else {
        const int ew = warp - 2, lg = warp & 3, half = ew >> 2;      // warps 2..9: every (lane group, column half) pair once
        uint8_t* slab = smem_raw + (epi_base - smem_u32(smem_raw)) + ew * 32 * ROWB;      // this warp's staging slab
        // read-back / store geometry: LPR lanes cover one staging row, a store instruction writes 32 / LPR whole row segments
        const int rsub = lane / LPR, piece = lane % LPR;
        int tl = 0;
        for (int t = blockIdx.x; t < prm.tiles_total; t += gridDim.x, tl++) {
            const int n0 = (t % prm.tiles_n) * NT, tr = t / prm.tiles_n, row0 = (tr % prm.tiles_r) * ROWS, b = tr / prm.tiles_r;
            const int buf = tl & 1;
            const int ncol = n0 + half * CW;
            const int row = row0 + lg * 32 + lane;
            const float* bias = prm.bias;
            if (prm.marks && row < prm.T_rows) {
                int mi = row / prm.marks_rep;
                if (mi >= prm.marks_T) mi = prm.marks_T - 1;
                if (prm.marks[(long long)b * prm.marks_T + mi] != 0) bias = prm.bias_alt;
            }
            mbar_wait(tfull_bar(buf), (uint32_t)((tl >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * NT + half * CW);
            float v[CW];
#pragma unroll
            for (int c = 0; c < CW / 16; c++) tmem_ld16_issue(taddr + c * 16, *reinterpret_cast<float(*)[16]>(&v[c * 16]));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty_bar(buf)) : "memory");
            {
                int bi = ncol % prm.bias_mod;                                      // bias_mod is a multiple of 16
#pragma unroll
                for (int j = 0; j < CW; j += 4) {
                    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + bi));
                    v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
                    bi += 4;
                    if (bi >= prm.bias_mod) bi -= prm.bias_mod;
                }
            }
            if (prm.res && row < prm.T_rows) {                                     // unfused residual blocks only (SSRB_RESBLOCK_UNFUSED=1)
                const bf16* rp = prm.res + (long long)b * prm.res_bstride + prm.res_off + (long long)row * prm.N + ncol;
#pragma unroll
                for (int j = 0; j < CW; j += 8) {
                    float r8[8];
                    load8(rp + j, r8);
#pragma unroll
                    for (int e = 0; e < 8; e++) v[j + e] += r8[e];
                }
            }
            const long long f0 = (long long)(row0 + lg * 32 + rsub) * prm.N + ncol + piece * 8;      // flat index of this lane's first 16 B
            const long long fstep = (long long)(32 / LPR) * prm.N;
            const bool inside = row0 + ROWS <= prm.T_rows && (long long)row0 * prm.N + n0 >= prm.valid_lo &&
                                (long long)(row0 + ROWS - 1) * prm.N + n0 + NT <= prm.valid_hi;
#pragma unroll
            for (int pass = 0; pass < 2; pass++) {
                bf16* outp = pass ? prm.out_act : prm.out_raw;
                if (!outp) continue;
                outp += (long long)b * prm.out_bstride + prm.out_off;
                // a thread's row -> its staging row
#pragma unroll
                for (int j = 0; j < CW; j += 8) {
                    float e8[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) e8[e] = pass ? elu1_bf16(v[j + e]) : v[j + e];
                    store8(reinterpret_cast<bf16*>(slab + lane * ROWB) + j, e8);
                }
                __syncwarp();
                if (inside) {
#pragma unroll
                    for (int j = 0; j < LPR; j++)
                        *reinterpret_cast<uint4*>(outp + f0 + j * fstep) =
                            *reinterpret_cast<const uint4*>(slab + (j * (32 / LPR) + rsub) * ROWB + piece * 16);
                } else {
#pragma unroll 1
                    for (int j = 0; j < LPR; j++) {
                        const int grow = row0 + lg * 32 + j * (32 / LPR) + rsub;
                        const long long f = f0 + j * fstep;
                        if (grow >= prm.T_rows || f + 8 <= prm.valid_lo || f >= prm.valid_hi) continue;
                        const uint4 w = *reinterpret_cast<const uint4*>(slab + (j * (32 / LPR) + rsub) * ROWB + piece * 16);
                        if (f >= prm.valid_lo && f + 8 <= prm.valid_hi) {
                            *reinterpret_cast<uint4*>(outp + f) = w;
                        } else {
                            const bf16* we = reinterpret_cast<const bf16*>(&w);
                            for (int e = 0; e < 8; e++)
                                if (f + e >= prm.valid_lo && f + e < prm.valid_hi) outp[f + e] = we[e];
                        }
                    }
                }
                __syncwarp();
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
    constexpr size_t SMEM = (size_t)STAGES * (A_BYTES + NT * BK * 2) + epi_bytes(NT) + 1024 + 256;
    static bool done = false;
    if (!done) {
        SSRB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        done = true;
    }
    SSRB_LAUNCH(conv_tc_kernel<NT>, grid, THREADS, SMEM, s, maps, prm);
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
    p.tiles_n = a.N / NT; p.tiles_r = cdiv(a.T_rows, ROWS);
    const long long total = (long long)p.tiles_n * p.tiles_r * a.B;
    SSRB_CHECK(total > 0 && total < (1ll << 31), "conv_tc: tile count out of range");
    p.tiles_total = (int)total;
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        SSRB_CUDA(cudaGetDevice(&dev));
        SSRB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    dim3 grid((unsigned)std::min<long long>(total, 2ll * n_sm));
    return NT == 128 ? launch<128>(maps, p, grid, s) : launch<64>(maps, p, grid, s);
}

}  // namespace ssrb
