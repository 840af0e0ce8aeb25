// gemm_simt.cu — CUDA-core GEMM  C[M,N] = act(A[M,K] . W[N,K]^T + bias) (+ residual).
//
// Replaces the F.linear call sites of the reference's decoder layer and heads
// (models/modules/activation.py:86,637; models/modules/transformer.py:386-388; models/ssr.py:175-179,688)
// for (a) the fp32 parity mode (fp32 weights and activations, fp32 FMA accumulation: the arithmetic of
// the reference's own fp32 inference, inference_v2.py:203-204) and (b) small-M bf16 decode steps where a
// 128-bit-load GEMV is already HBM-bound.  The tensor-core path lives in gemm_tc.cu.
#include "common.cuh"
#include "../../include/ssr_b200.h"

namespace ssrb {

template <typename TC>
__device__ __forceinline__ void epilogue_store(const GemmArgs& g, int grp, int m, int n, float v) {
    if (g.bias) v += g.bias[grp * g.bias_gs + n];
    if (g.act == ACT_RELU) v = fmaxf(v, 0.f);
    else if (g.act == ACT_GELU) v = gelu_erf(v);
    if (g.residual) v += g.residual[(int64_t)m * g.ldr + n];
    reinterpret_cast<TC*>(g.C)[grp * g.c_gs + (int64_t)m * g.ldc + n] = from_f32<TC>(v);
}

// ---- skinny path: M <= MR rows, one warp per output column, 128-bit weight loads -----------------------
template <typename T, typename TC, int MR>
__global__ void __launch_bounds__(256) gemv_rows_kernel(GemmArgs g) {
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = blockIdx.z;
    const int n = blockIdx.x * 8 + warp;
    if (n >= g.N) return;
    const T* W = reinterpret_cast<const T*>(g.W) + grp * g.w_gs + (int64_t)n * g.ldw;
    const T* A = reinterpret_cast<const T*>(g.A) + grp * g.a_gs;
    float acc[MR];
#pragma unroll
    for (int r = 0; r < MR; r++) acc[r] = 0.f;
    for (int k = lane * 8; k < g.K; k += 256) {
        float w[8];
        load8(W + k, w);
#pragma unroll
        for (int r = 0; r < MR; r++) {
            if (r < g.M) {
                float a[8];
                load8(A + (int64_t)r * g.lda + k, a);
#pragma unroll
                for (int i = 0; i < 8; i++) acc[r] = fmaf(a[i], w[i], acc[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < MR; r++) {
        float v = warp_sum(acc[r]);
        if (lane == 0 && r < g.M) epilogue_store<TC>(g, grp, r, n, v);
    }
}

// ---- tiled path: 64x64x16 tiles, 256 threads, 4x4 outputs per thread -----------------------------------
template <typename T, typename TC>
__global__ void __launch_bounds__(256) gemm_tile_kernel(GemmArgs g) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Ws[2][BK][BN + PAD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int grp = blockIdx.z;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const T* A = reinterpret_cast<const T*>(g.A) + grp * g.a_gs;
    const T* W = reinterpret_cast<const T*>(g.W) + grp * g.w_gs;
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const bool a_ok = (m0 + lrow) < g.M, w_ok = (n0 + lrow) < g.N;
    const T* Ap = A + (int64_t)(m0 + lrow) * g.lda + lk;
    const T* Wp = W + (int64_t)(n0 + lrow) * g.ldw + lk;
    float acc[4][4] = {};
    float ra[4] = {0, 0, 0, 0}, rw[4] = {0, 0, 0, 0};
    if (a_ok) load4(Ap, ra);
    if (w_ok) load4(Wp, rw);
    const int nk = g.K / BK;
    for (int kt = 0; kt < nk; kt++) {
        const int buf = kt & 1;
#pragma unroll
        for (int i = 0; i < 4; i++) { As[buf][lk + i][lrow] = ra[i]; Ws[buf][lk + i][lrow] = rw[i]; }
        __syncthreads();
        if (kt + 1 < nk) {
            if (a_ok) load4(Ap + (kt + 1) * BK, ra);
            if (w_ok) load4(Wp + (kt + 1) * BK, rw);
        }
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 w4 = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int n = n0 + tx * 4 + j;
            if (n < g.N) epilogue_store<TC>(g, grp, m, n, acc[i][j]);
        }
    }
}

template <typename T, typename TC>
static int launch_simt(const GemmArgs& g, cudaStream_t s) {
    SSRB_CHECK(g.K % 16 == 0, "gemm_simt: K must be a multiple of 16");
    SSRB_CHECK(g.lda % 8 == 0 && g.ldw % 8 == 0, "gemm_simt: leading dims must be multiples of 8 elements");
    if (g.M <= 4 && g.K % 256 == 0) {
        dim3 grid(cdiv(g.N, 8), 1, g.groups);
        if (g.M <= 1) SSRB_LAUNCH_PDL((gemv_rows_kernel<T, TC, 1>), grid, 256, 0, s, g);
        else if (g.M <= 2) SSRB_LAUNCH_PDL((gemv_rows_kernel<T, TC, 2>), grid, 256, 0, s, g);
        else SSRB_LAUNCH_PDL((gemv_rows_kernel<T, TC, 4>), grid, 256, 0, s, g);
        return 0;
    }
    dim3 grid(cdiv(g.N, 64), cdiv(g.M, 64), g.groups);
    SSRB_LAUNCH_PDL((gemm_tile_kernel<T, TC>), grid, 256, 0, s, g);
    return 0;
}

int gemm_simt(const GemmArgs& g, cudaStream_t s) {
    if (g.M <= 0 || g.N <= 0) return 0;
    if (g.ab_dtype == SSRB_DTYPE_F32) {
        SSRB_CHECK(g.c_dtype == SSRB_DTYPE_F32, "gemm_simt: fp32 inputs require fp32 output");
        return launch_simt<float, float>(g, s);
    }
    if (g.c_dtype == SSRB_DTYPE_F32) return launch_simt<bf16, float>(g, s);
    return launch_simt<bf16, bf16>(g, s);
}

}  // namespace ssrb
