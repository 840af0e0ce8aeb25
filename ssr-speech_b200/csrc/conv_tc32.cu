// conv_tc32.cu — SEANet ENCODER convolutions on the tcgen05 tensor cores at fp32-grade accuracy (3 x TF32).
//
// The encoder's outputs are quantised by nearest-neighbour searches (core_vq.py:164-172), so its convolutions
// (seanet.py:63-153, conv.py:185-201) must stay bit-comparable with the fp32 reference to ~1e-6: bf16 operands are out.  Every fp32
// operand is split into two TF32 numbers, x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi) (22 significant bits), and
//
//        D += A_hi . W_hi  +  A_lo . W_hi  +  A_hi . W_lo            (fp32 accumulate in tensor memory)
//
// drops only the lo x lo term (2^-22 relative): the error budget of an fp32 FMA chain with another summation order, at ~1/3 of
// the TF32 tensor rate instead of the CUDA-core FMA rate (23 TFLOP/s measured for the fp32 kernels, profiles/r01e_summary.md).
//
// Same formulation as conv_tc.cu: channels-last activations X[b][G + t][c] (fp32 here, a hi and a lo array) with G zero guard rows,
// a convolution = taps tap-GEMMs accumulated in TMEM over shifted row views (stride-s convolutions view s time steps as one row of
// s*Cin channels), 3-D TMA maps {Cw, rows, batch} with 128-byte swizzle (32 fp32 per k-block), tcgen05.mma.kind::tf32 128 x NT x 8.
// The PRODUCER's epilogue writes what its consumers need: raw fp32 (residual inputs), and the ELU'd operand already split (hi, lo).
#include <cuda.h>

#include <mutex>

#include "conv_tc.cuh"

namespace ssrb {

namespace {

constexpr int BK = 32, ROWS = 128, A_BYTES = ROWS * BK * 4;

struct Maps32 { CUtensorMap ah, al, wh, wl; };

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
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
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
__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

struct K32Params {
    int T_rows, N, nkb_per_tap, taps;
    const float* bias;
    const float* res; long long res_bstride, res_off;
    float* out_raw; float* out_hi; float* out_lo; long long out_bstride, out_off;
    int elu;                                        // hi/lo hold ELU(out) (else out itself)
};

// Accumulation: the tensor core adds into its fp32 accumulator with one-sided rounding, so a long K loop drifts LINEARLY (measured:
// latents off by 6.4e-5 of full scale after the 15 encoder layers with all of K accumulated in tensor memory, 17 x the fp32
// kernels' 3.8e-6).  K is therefore accumulated in CHUNKS of CHUNK_KB k-blocks (128 products per output): two TMEM accumulators
// alternate, the epilogue warps drain each finished chunk into fp32 registers with round-to-nearest adds while the next chunk's MMAs
// run (the promotion trick fp8 GEMMs use).
constexpr int CHUNK_KB = 4;

constexpr int EPI32_WARPS = 8, THREADS32 = (2 + EPI32_WARPS) * 32;     // two epilogue warps per TMEM lane group, each takes half of the columns

template <int NT, int STAGES>
__global__ void __launch_bounds__(THREADS32, (NT == 128 && STAGES == 1) ? 2 : 1) conv_tc32_kernel(const __grid_constant__ Maps32 maps, const K32Params prm) {
    constexpr int W_BYTES = NT * BK * 4, STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES, TMEM_COLS = 2 * NT;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = base + STAGES * STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    const uint32_t tfull0 = bar_base + 8u * (2 * STAGES), tempty0 = tfull0 + 16u, tmem_slot = tempty0 + 16u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * NT, row0 = blockIdx.y * ROWS, b = blockIdx.z;
    const int nk = prm.taps * prm.nkb_per_tap;
    const int nchunks = (nk + CHUNK_KB - 1) / CHUNK_KB;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(tfull0 + 8u * i, 1); mbar_init(tempty0 + 8u * i, EPI32_WARPS); }
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
                tma_load_3d(sp, &maps.ah, full_bar(s), cb * BK, row0 + q, b);
                tma_load_3d(sp + A_BYTES, &maps.al, full_bar(s), cb * BK, row0 + q, b);
                tma_load_2d(sp + 2 * A_BYTES, &maps.wh, full_bar(s), cb * BK, q * prm.N + n0);
                tma_load_2d(sp + 2 * A_BYTES + W_BYTES, &maps.wl, full_bar(s), cb * BK, q * prm.N + n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // kind::tf32: c_format F32 (1 << 4), a_format = b_format = TF32 (2), K-major both, N >> 3, M >> 4
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
            for (int c = 0; c < nchunks; c++) {
                const int buf = c & 1;
                mbar_wait(tempty0 + 8u * buf, (uint32_t)(((c >> 1) & 1) ^ 1));       // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + (uint32_t)(buf * NT);
                const int i1 = min(nk, (c + 1) * CHUNK_KB);
                for (int i = c * CHUNK_KB; i < i1; i++) {
                    const int s = i % STAGES;
                    const uint32_t ph = (i / STAGES) & 1;
                    mbar_wait(full_bar(s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sp = base + s * STAGE_BYTES;
                    const uint64_t dah = make_desc(sp), dal = make_desc(sp + A_BYTES);
                    const uint64_t dwh = make_desc(sp + 2 * A_BYTES), dwl = make_desc(sp + 2 * A_BYTES + W_BYTES);
                    const bool first = i == c * CHUNK_KB;
#pragma unroll
                    for (int k = 0; k < BK / 8; k++) {          // K = 8 per tf32 MMA = 32 bytes = 2 descriptor units
                        umma_tf32(acc, dal + 2 * k, dwh + 2 * k, idesc, (!first || k > 0) ? 1u : 0u);   // small terms first
                        umma_tf32(acc, dah + 2 * k, dwl + 2 * k, idesc, 1u);
                        umma_tf32(acc, dah + 2 * k, dwh + 2 * k, idesc, 1u);
                    }
                    umma_commit(empty_bar(s));
                }
                umma_commit(tfull0 + 8u * buf);
            }
        }
    } else {
        constexpr int CW = NT / 2;                                   // columns of this warp
        const int lg = warp & 3, half = (warp - 2) >> 2;             // warps 2..9: every (lane group, column half) pair once
        const int row = row0 + lg * 32 + lane;
        const bool rok = row < prm.T_rows;
        float acc[CW];
#pragma unroll
        for (int j = 0; j < CW; j++) acc[j] = 0.f;
        for (int c = 0; c < nchunks; c++) {
            const int buf = c & 1;
            mbar_wait(tfull0 + 8u * buf, (uint32_t)((c >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * NT + half * CW);
#pragma unroll
            for (int c0 = 0; c0 < CW; c0 += 16) {
                float v[16];
                tmem_ld16(taddr + c0, v);
#pragma unroll
                for (int j = 0; j < 16; j++) acc[c0 + j] = __fadd_rn(acc[c0 + j], v[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty0 + 8u * buf) : "memory");
        }
        if (rok) {
#pragma unroll
            for (int c0 = 0; c0 < CW; c0 += 16) {
                const int n = n0 + half * CW + c0;
                const long long f = (long long)row * prm.N + n;
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 bv = *reinterpret_cast<const float4*>(prm.bias + n + j);
                    v[j] = acc[c0 + j] + bv.x; v[j + 1] = acc[c0 + j + 1] + bv.y; v[j + 2] = acc[c0 + j + 2] + bv.z; v[j + 3] = acc[c0 + j + 3] + bv.w;
                }
                if (prm.res) {
                    const float* rp = prm.res + (long long)b * prm.res_bstride + prm.res_off + f;
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 rv = *reinterpret_cast<const float4*>(rp + j);
                        v[j] += rv.x; v[j + 1] += rv.y; v[j + 2] += rv.z; v[j + 3] += rv.w;
                    }
                }
                const long long o = (long long)b * prm.out_bstride + prm.out_off + f;
                if (prm.out_raw) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(prm.out_raw + o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
                if (prm.out_hi) {
                    float hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const float a = prm.elu ? elu1(v[j]) : v[j];
                        hi[j] = rna_tf32(a);
                        lo[j] = rna_tf32(a - hi[j]);
                    }
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        *reinterpret_cast<float4*>(prm.out_hi + o + j) = make_float4(hi[j], hi[j + 1], hi[j + 2], hi[j + 3]);
                        *reinterpret_cast<float4*>(prm.out_lo + o + j) = make_float4(lo[j], lo[j + 1], lo[j + 2], lo[j + 3]);
                    }
                }
            }
        }
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
EncodeTiledFn encode_fn32() {
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

template <int NT, int STAGES>
int launch32(const Maps32& maps, const K32Params& prm, dim3 grid, cudaStream_t s) {
    constexpr size_t SMEM = (size_t)STAGES * (2 * A_BYTES + 2 * NT * BK * 4) + 1024 + 256;
    static bool done = false;
    if (!done) {
        SSRB_CUDA(cudaFuncSetAttribute(conv_tc32_kernel<NT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        done = true;
    }
    SSRB_LAUNCH((conv_tc32_kernel<NT, STAGES>), grid, THREADS32, SMEM, s, maps, prm);
    return 0;
}

// first SEANet conv (1 -> C channels, kernel k; seanet.py:118-122) -> channels-last fp32: raw, and ELU(raw) split into hi / lo
__global__ void __launch_bounds__(256) cl32_first_conv_kernel(const float* __restrict__ wav, int T, const float* __restrict__ W,
                                                              const float* __restrict__ bias, int C, int k, float* __restrict__ out_raw,
                                                              float* __restrict__ out_hi, float* __restrict__ out_lo) {
    __shared__ float xs[64 + 16];
    __shared__ float ws[64 * 16];
    const int b = blockIdx.y, t0 = blockIdx.x * 64, padL = (k - 1) - (k - 1) / 2;
    for (int e = threadIdx.x; e < 64 + k - 1; e += 256) {
        const int g = t0 + e - padL;
        xs[e] = (g >= 0 && g < T) ? wav[(int64_t)b * T + g] : 0.f;
    }
    for (int e = threadIdx.x; e < C * k; e += 256) ws[e] = W[e];
    __syncthreads();
    const int c = threadIdx.x % 64, tg = threadIdx.x / 64;
    if (c >= C) return;
    const float bv = bias[c];
    const int64_t base = (int64_t)b * (T + 2 * CL_GUARD) * C;
    for (int i = 0; i < 16; i++) {
        const int tl = tg + 4 * i, t = t0 + tl;
        if (t >= T) break;
        // same operation order as the fp32 reference kernel (conv1d: bias first, taps in order, fused multiply-add)
        float acc = bv;
        for (int j = 0; j < k; j++) acc = fmaf(ws[c * k + j], xs[tl + j], acc);
        const int64_t o = base + (int64_t)(CL_GUARD + t) * C + c;
        out_raw[o] = acc;
        const float a = elu1(acc), hi = rna_tf32(a);
        out_hi[o] = hi;
        out_lo[o] = rna_tf32(a - hi);
    }
}

__global__ void cl32_zero_guards_kernel(float* __restrict__ p, int T, int C) {
    const int b = blockIdx.x, side = blockIdx.y;
    float* g = p + ((int64_t)b * (T + 2 * CL_GUARD) + (side ? CL_GUARD + T : 0)) * C;
    const int n = CL_GUARD * C;                               // multiple of 4
    for (int e = threadIdx.x * 4; e < n; e += blockDim.x * 4) *reinterpret_cast<float4*>(g + e) = make_float4(0.f, 0.f, 0.f, 0.f);
}

// channels-last fp32 [B][G+T+G][C] -> channels-first [B,C,T]
__global__ void cl32_to_cf32_kernel(const float* __restrict__ in, int C, int T, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int t = t0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && t < T) ? in[((int64_t)b * (T + 2 * CL_GUARD) + CL_GUARD + t) * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, t = t0 + threadIdx.x;
        if (c < C && t < T) out[((int64_t)b * C + c) * T + t] = tile[threadIdx.x][i];
    }
}

// x -> (hi, lo) TF32 pair, elementwise (n % 4 == 0)
__global__ void split_tf32_kernel(const float* __restrict__ in, float* __restrict__ hi, float* __restrict__ lo, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(in)[i];
        float4 h, l;
        h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
        l.x = rna_tf32(v.x - h.x); l.y = rna_tf32(v.y - h.y); l.z = rna_tf32(v.z - h.z); l.w = rna_tf32(v.w - h.w);
        reinterpret_cast<float4*>(hi)[i] = h;
        reinterpret_cast<float4*>(lo)[i] = l;
    }
}

}  // namespace

int launch_split_tf32(const float* in, float* hi, float* lo, long long n, cudaStream_t s) {
    SSRB_CHECK(n % 4 == 0 && ((uintptr_t)in & 15) == 0 && ((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0, "split_tf32: 16B alignment");
    const long long n4 = n / 4;
    const int grid = (int)std::min<long long>((n4 + 255) / 256, 148 * 8);
    SSRB_LAUNCH(split_tf32_kernel, grid, 256, 0, s, in, hi, lo, n4);
    return 0;
}

int conv_tc32(const ConvTc32Args& a, cudaStream_t s) {
    EncodeTiledFn fn = encode_fn32();
    SSRB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
    SSRB_CHECK(a.Cw % BK == 0 && a.N % 32 == 0, "conv_tc32: channel counts must be multiples of 32");
    SSRB_CHECK(((uintptr_t)a.x_hi & 15) == 0 && ((uintptr_t)a.x_lo & 15) == 0 && (a.x_base_off % 4) == 0 && (a.x_bstride % 4) == 0,
               "conv_tc32: input view must be 16B aligned");
    SSRB_CHECK((a.out_off % 4) == 0 && (a.out_bstride % 4) == 0, "conv_tc32: output view must be 16B aligned");
    const int NT = (a.N % 128 == 0) ? 128 : ((a.N % 64 == 0) ? 64 : 32);
    Maps32 maps;
    memset(&maps, 0, sizeof(maps));
    for (int h = 0; h < 2; h++) {
        cuuint64_t dims[3] = {(cuuint64_t)a.Cw, (cuuint64_t)a.rows_v, (cuuint64_t)a.B};
        cuuint64_t strides[2] = {(cuuint64_t)a.Cw * 4, (cuuint64_t)a.x_bstride * 4};
        cuuint32_t box[3] = {BK, ROWS, 1}, estr[3] = {1, 1, 1};
        CUresult r = fn(h ? &maps.al : &maps.ah, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>((h ? a.x_lo : a.x_hi) + a.x_base_off), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SSRB_CHECK(r == CUDA_SUCCESS, "conv_tc32: activation tensor map failed");
    }
    for (int h = 0; h < 2; h++) {
        cuuint64_t dims[2] = {(cuuint64_t)a.Cw, (cuuint64_t)a.taps * a.N};
        cuuint64_t strides[1] = {(cuuint64_t)a.Cw * 4};
        cuuint32_t box[2] = {BK, (cuuint32_t)NT}, estr[2] = {1, 1};
        CUresult r = fn(h ? &maps.wl : &maps.wh, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(h ? a.w_lo : a.w_hi), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SSRB_CHECK(r == CUDA_SUCCESS, "conv_tc32: weight tensor map failed");
    }
    K32Params p{};
    p.T_rows = a.T_rows; p.N = a.N; p.nkb_per_tap = a.Cw / BK; p.taps = a.taps;
    p.bias = a.bias;
    p.res = a.res; p.res_bstride = a.res_bstride; p.res_off = a.res_off;
    p.out_raw = a.out_raw; p.out_hi = a.out_hi; p.out_lo = a.out_lo; p.out_bstride = a.out_bstride; p.out_off = a.out_off;
    p.elu = a.elu ? 1 : 0;
    SSRB_CHECK(!a.out_hi == !a.out_lo, "conv_tc32: hi and lo outputs go together");
    dim3 grid(a.N / NT, cdiv(a.T_rows, ROWS), a.B);
    SSRB_CHECK(grid.z <= 65535 && grid.y <= 65535, "conv_tc32: grid too large");
#ifndef C32_STAGES_128
#define C32_STAGES_128 3          // measured on 32 x 10 s: 3 stages 38.1 ms, 2 stages 39.5 ms, 1 stage with two CTAs per SM 38.5 ms
#endif
    if (NT == 128) return launch32<128, C32_STAGES_128>(maps, p, grid, s);
    if (NT == 64) return launch32<64, 2>(maps, p, grid, s);
    return launch32<32, 2>(maps, p, grid, s);
}

int launch_cl32_first_conv(const float* wav, int B, int T, const float* W, const float* bias, int C, int k, float* out_raw, float* out_hi,
                           float* out_lo, cudaStream_t s) {
    SSRB_CHECK(C <= 64 && k <= 16, "cl32_first_conv: unsupported shape");
    dim3 grid(cdiv(T, 64), B);
    SSRB_LAUNCH(cl32_first_conv_kernel, grid, 256, 0, s, wav, T, W, bias, C, k, out_raw, out_hi, out_lo);
    return 0;
}
int launch_cl32_zero_guards(float* p, int B, int T, int C, cudaStream_t s) {
    SSRB_CHECK(C % 4 == 0, "cl32_zero_guards: C must be a multiple of 4");
    dim3 grid(B, 2);
    SSRB_LAUNCH(cl32_zero_guards_kernel, grid, 256, 0, s, p, T, C);
    return 0;
}
int launch_cl32_to_cf32(const float* in, int B, int C, int T, float* out, cudaStream_t s) {
    dim3 grid(cdiv(T, 32), cdiv(C, 32), B), block(32, 8);
    SSRB_LAUNCH(cl32_to_cf32_kernel, grid, block, 0, s, in, C, T, out);
    return 0;
}

}  // namespace ssrb
