// codec_kernels.cu — WM-Encodec kernels: SEANet Conv1d / ConvTranspose1d with fused ELU-in, bias and
// residual-add epilogues, the persistent LSTM recurrence, RVQ encode/decode and the watermark concat.
// All arithmetic is fp32 FMA (the reference codec runs fp32; RVQ indices must be reproducible).
// Reference call sites are cited per kernel (paths relative to audiocraft/audiocraft/).
#include "codec_kernels.cuh"

namespace ssrb {

// =================================================================================================
// Conv1d as an implicit GEMM on CUDA cores.           modules/conv.py:185-201 (StreamableConv1d)
//   grid (ceil(Tout/64), ceil(Cout/BCO), B), 256 threads: ty = co group (TCO channels), tx + 16j = time.
//   The input strip is staged in shared memory de-interleaved by stride phase so that strided taps
//   read consecutive banks; ELU (modules/seanet.py:39-46,132,144) is applied while staging.
// =================================================================================================
template <int TCO>
__global__ void __launch_bounds__(256) conv1d_kernel(const float* __restrict__ in, int Cin, int Tin,
                                                     const float* __restrict__ W, const float* __restrict__ bias,
                                                     int Cout, int ksz, int stride, int padL, int Tout, int elu_in,
                                                     const float* __restrict__ res, float* __restrict__ out, int CI) {
    constexpr int BCO = 16 * TCO, BT = 64;
    extern __shared__ __align__(16) float smem[];
    const int Qw = BT + (ksz - 1) / stride;
    const int strip = stride * Qw;                 // floats per input channel
    float* in_s = smem;                            // [CI][stride][Qw]
    float* w_s = smem + ((CI * strip + 3) & ~3);   // [CI][ksz][BCO]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int t0 = blockIdx.x * BT, c0 = blockIdx.y * BCO, b = blockIdx.z;
    const float* inb = in + (int64_t)b * Cin * Tin;
    const int gbase = t0 * stride - padL;
    float acc[TCO][4];
#pragma unroll
    for (int i = 0; i < TCO; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (int ci0 = 0; ci0 < Cin; ci0 += CI) {
        const int nci = min(CI, Cin - ci0);
        __syncthreads();
        for (int e = tid; e < nci * strip; e += 256) {
            const int ci = e / strip, r = e - ci * strip;
            const int g = gbase + r;
            float v = 0.f;
            if (g >= 0 && g < Tin) {
                v = inb[(int64_t)(ci0 + ci) * Tin + g];
                if (elu_in) v = elu1(v);
            }
            const int p = r % stride, q = r / stride;
            in_s[ci * strip + p * Qw + q] = v;
        }
        const int wrun = nci * ksz;                // contiguous (ci,k) run per output channel
        for (int e = tid; e < BCO * wrun; e += 256) {
            const int co = e / wrun, r = e - co * wrun;
            float v = 0.f;
            if (c0 + co < Cout) v = W[((int64_t)(c0 + co) * Cin + ci0) * ksz + r];
            w_s[r * BCO + co] = v;                 // r = ci*ksz + k
        }
        __syncthreads();
        for (int ci = 0; ci < nci; ci++) {
            const float* xs = in_s + ci * strip;
            const float* ws = w_s + ci * ksz * BCO + ty * TCO;
            int p = 0, qo = 0;
            for (int k = 0; k < ksz; k++) {
                float w[TCO];
#pragma unroll
                for (int i = 0; i < TCO; i += 2) {
                    const float2 t2 = *reinterpret_cast<const float2*>(ws + k * BCO + i);
                    w[i] = t2.x; w[i + 1] = t2.y;
                }
                const float* xr = xs + p * Qw + qo + tx;
                float x[4];
#pragma unroll
                for (int j = 0; j < 4; j++) x[j] = xr[16 * j];
#pragma unroll
                for (int i = 0; i < TCO; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = fmaf(w[i], x[j], acc[i][j]);
                if (++p == stride) { p = 0; qo++; }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < TCO; i++) {
        const int co = c0 + ty * TCO + i;
        if (co >= Cout) continue;
        const float bv = bias ? bias[co] : 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int t = t0 + tx + 16 * j;
            if (t < Tout) {
                const int64_t o = ((int64_t)b * Cout + co) * Tout + t;
                float v = acc[i][j] + bv;
                if (res) v += res[o];
                out[o] = v;
            }
        }
    }
}

// =================================================================================================
// Conv1d, register-tiled version for the layer shapes of the SEANet stacks (compile-time kernel size and stride).
//   CTA tile = (16*TCO) output channels x 128 output steps, 256 threads; thread (ty, tx) owns TCO channels x
//   two groups of 4 consecutive steps (tx*4 + {0,64}) -> every shared-memory access is a conflict-free LDS.128:
//   one x window per (input channel, stride phase) serves all taps of that phase from registers, weights come
//   pre-transposed as Wt[Cin][k][Cout].  Per-output accumulation order (ci outer, k inner) is the same as in
//   conv1d_kernel above, so both kernels produce identical bits.
// =================================================================================================
template <int KSZ, int STRIDE, int TCO>
__global__ void __launch_bounds__(256, 2) conv1d_v2_kernel(const float* __restrict__ in, int Cin, int Tin,
                                                           const float* __restrict__ Wt, const float* __restrict__ bias,
                                                           int Cout, int padL, int Tout, int elu_in,
                                                           const float* __restrict__ res, float* __restrict__ out, int CI) {
    constexpr int BCO = 16 * TCO, BT = 128;
    constexpr int QO_MAX = (KSZ - 1) / STRIDE;            // largest tap shift inside one stride phase
    constexpr int NW = (4 + QO_MAX + 3) / 4;              // float4 per x window
    constexpr int QW = 124 + 4 * NW;                      // strip length per phase (covers the last window)
    constexpr int STRIP = STRIDE * QW;
    extern __shared__ __align__(16) float smem[];
    float* in_s = smem;                                    // [CI][STRIDE][QW]
    float* w_s = smem + CI * STRIP;                        // [CI][KSZ][BCO]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int t0 = blockIdx.x * BT, c0 = blockIdx.y * BCO, b = blockIdx.z;
    const float* inb = in + (int64_t)b * Cin * Tin;
    const int gbase = t0 * STRIDE - padL;
    float acc[TCO][8];
#pragma unroll
    for (int i = 0; i < TCO; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
    for (int ci0 = 0; ci0 < Cin; ci0 += CI) {
        const int nci = min(CI, Cin - ci0);
        __syncthreads();
        for (int e = tid; e < nci * STRIP; e += 256) {
            const int ci = e / STRIP, r = e - ci * STRIP;
            const int g = gbase + r;
            float v = 0.f;
            if (g >= 0 && g < Tin) {
                v = inb[(int64_t)(ci0 + ci) * Tin + g];
                if (elu_in) v = elu1(v);
            }
            in_s[ci * STRIP + (r % STRIDE) * QW + r / STRIDE] = v;
        }
        {
            const float* wsrc = Wt + (int64_t)ci0 * KSZ * Cout + c0;
            constexpr int V = BCO / 4;                     // float4 per (ci,k) row
            for (int e = tid; e < nci * KSZ * V; e += 256) {
                const int r = e / V, c4 = (e - r * V) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 + c4 + 3 < Cout) v = *reinterpret_cast<const float4*>(wsrc + (int64_t)r * Cout + c4);
                else {
                    float t[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int z = 0; z < 4; z++) if (c0 + c4 + z < Cout) t[z] = wsrc[(int64_t)r * Cout + c4 + z];
                    v = make_float4(t[0], t[1], t[2], t[3]);
                }
                *reinterpret_cast<float4*>(w_s + r * BCO + c4) = v;
            }
        }
        __syncthreads();
        for (int ci = 0; ci < nci; ci++) {
#pragma unroll
            for (int p = 0; p < STRIDE; p++) {
                if (p >= KSZ) break;
                float xw[2][4 * NW];
#pragma unroll
                for (int g = 0; g < 2; g++)
#pragma unroll
                    for (int n = 0; n < NW; n++) {
                        const float4 t4 = *reinterpret_cast<const float4*>(in_s + ci * STRIP + p * QW + g * 64 + tx * 4 + 4 * n);
                        xw[g][4 * n] = t4.x; xw[g][4 * n + 1] = t4.y; xw[g][4 * n + 2] = t4.z; xw[g][4 * n + 3] = t4.w;
                    }
#pragma unroll
                for (int k = p; k < KSZ; k += STRIDE) {
                    const int qo = k / STRIDE;
                    float w[TCO];
                    const float* wp = w_s + (ci * KSZ + k) * BCO + ty * TCO;
                    if (TCO >= 4) {
#pragma unroll
                        for (int i = 0; i < TCO; i += 4) {
                            const float4 t4 = *reinterpret_cast<const float4*>(wp + i);
                            w[i] = t4.x; w[i + 1] = t4.y; w[i + 2] = t4.z; w[i + 3] = t4.w;
                        }
                    } else {
                        const float2 t2 = *reinterpret_cast<const float2*>(wp);
                        w[0] = t2.x; w[1] = t2.y;
                    }
#pragma unroll
                    for (int i = 0; i < TCO; i++)
#pragma unroll
                        for (int g = 0; g < 2; g++)
#pragma unroll
                            for (int j = 0; j < 4; j++) acc[i][g * 4 + j] = fmaf(w[i], xw[g][j + qo], acc[i][g * 4 + j]);
                }
            }
        }
    }
    const bool vec = (Tout & 3) == 0;
#pragma unroll
    for (int i = 0; i < TCO; i++) {
        const int co = c0 + ty * TCO + i;
        if (co >= Cout) continue;
        const float bv = bias ? bias[co] : 0.f;
#pragma unroll
        for (int g = 0; g < 2; g++) {
            const int t = t0 + g * 64 + tx * 4;
            if (t >= Tout) continue;
            const int64_t o = ((int64_t)b * Cout + co) * Tout + t;
            if (vec) {                                      // Tout % 4 == 0 and t % 4 == 0: the float4 is in range and aligned
                float4 v = make_float4(acc[i][g * 4] + bv, acc[i][g * 4 + 1] + bv, acc[i][g * 4 + 2] + bv, acc[i][g * 4 + 3] + bv);
                if (res) { const float4 r4 = *reinterpret_cast<const float4*>(res + o); v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w; }
                *reinterpret_cast<float4*>(out + o) = v;
            } else {
                for (int j = 0; j < 4 && t + j < Tout; j++) {
                    float v = acc[i][g * 4 + j] + bv;
                    if (res) v += res[o + j];
                    out[o + j] = v;
                }
            }
        }
    }
}

__global__ void conv_w_transpose_kernel(const float* __restrict__ W, float* __restrict__ Wt, int Cout, int Cin, int ksz) {
    const int64_t n = (int64_t)Cout * Cin * ksz;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int co = (int)(e % Cout);
        const int64_t r = e / Cout;                        // r = ci*ksz + k
        Wt[e] = W[(int64_t)co * Cin * ksz + r];
    }
}
int launch_conv_w_transpose(const float* W, float* Wt, int Cout, int Cin, int ksz, cudaStream_t s) {
    SSRB_LAUNCH(conv_w_transpose_kernel, 256, 256, 0, s, W, Wt, Cout, Cin, ksz);
    return 0;
}

template <int KSZ, int STRIDE, int TCO>
static int launch_conv1d_v2(const float* in, int B, int Cin, int Tin, const float* Wt, const float* bias, int Cout, int padL,
                            int Tout, bool elu_in, const float* res, float* out, cudaStream_t s) {
    constexpr int QO_MAX = (KSZ - 1) / STRIDE, NW = (4 + QO_MAX + 3) / 4, QW = 124 + 4 * NW, BCO = 16 * TCO;
    constexpr int per_ci = STRIDE * QW + KSZ * BCO;
    int CI = 24000 / per_ci;                               // ~96 KB per CTA -> two CTAs per SM
    if (CI > 16) CI = 16;
    if (CI > Cin) CI = Cin;
    SSRB_CHECK(CI >= 1, "conv1d: kernel too large for the shared-memory tile");
    const size_t smem = (size_t)CI * per_ci * 4;
    static bool attr_done = false;
    if (!attr_done) {
        SSRB_CUDA(cudaFuncSetAttribute(conv1d_v2_kernel<KSZ, STRIDE, TCO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 24000 * 4));
        attr_done = true;
    }
    dim3 grid(cdiv(Tout, 128), cdiv(Cout, BCO), B);
    SSRB_LAUNCH((conv1d_v2_kernel<KSZ, STRIDE, TCO>), grid, 256, smem, s, in, Cin, Tin, Wt, bias, Cout, padL, Tout, (int)elu_in, res, out, CI);
    return 0;
}

// single-output-channel conv (final decoder layer 64 -> 1, k7): memory bound, one thread per sample
__global__ void __launch_bounds__(256) conv_cout1_kernel(const float* __restrict__ in, int Cin, int Tin,
                                                         const float* __restrict__ W, const float* __restrict__ bias,
                                                         int ksz, int padL, int Tout, int elu_in, float* __restrict__ out) {
    extern __shared__ float w_s[];
    for (int e = threadIdx.x; e < Cin * ksz; e += 256) w_s[e] = W[e];
    __syncthreads();
    const int b = blockIdx.y;
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= Tout) return;
    const float* inb = in + (int64_t)b * Cin * Tin;
    float acc = 0.f;
    for (int ci = 0; ci < Cin; ci++) {
        const float* xr = inb + (int64_t)ci * Tin;
        for (int k = 0; k < ksz; k++) {
            const int g = t + k - padL;
            if (g >= 0 && g < Tin) {
                float v = xr[g];
                if (elu_in) v = elu1(v);
                acc = fmaf(w_s[ci * ksz + k], v, acc);
            }
        }
    }
    out[(int64_t)b * Tout + t] = acc + (bias ? bias[0] : 0.f);
}

int launch_conv1d(const float* in, int B, int Cin, int Tin, const float* W, const float* bias, int Cout, int ksz,
                  int stride, int padL, int Tout, bool elu_in, const float* res, float* out, cudaStream_t s, const float* Wt) {
    if (Cout == 1 && stride == 1 && !res) {
        dim3 grid(cdiv(Tout, 256), B);
        SSRB_LAUNCH(conv_cout1_kernel, grid, 256, Cin * ksz * 4, s, in, Cin, Tin, W, bias, ksz, padL, Tout, (int)elu_in, out);
        return 0;
    }
    const int TCO = Cout >= 128 ? 8 : (Cout >= 64 ? 4 : 2);
    static const bool v1 = [] { const char* e = getenv("SSRB_CONV_V1"); return e && e[0] == '1'; }();
    if (Wt && !v1 && Cout % 4 == 0) {
#define SSRB_CONV_V2(K, S, T)                                                                                           \
    if (ksz == K && stride == S && TCO == T)                                                                            \
        return launch_conv1d_v2<K, S, T>(in, B, Cin, Tin, Wt, bias, Cout, padL, Tout, elu_in, res, out, s);
        SSRB_CONV_V2(7, 1, 4) SSRB_CONV_V2(7, 1, 8) SSRB_CONV_V2(3, 1, 2) SSRB_CONV_V2(3, 1, 4) SSRB_CONV_V2(3, 1, 8)
        SSRB_CONV_V2(1, 1, 4) SSRB_CONV_V2(1, 1, 8) SSRB_CONV_V2(4, 2, 8) SSRB_CONV_V2(8, 4, 8) SSRB_CONV_V2(10, 5, 8)
        SSRB_CONV_V2(16, 8, 8)
#undef SSRB_CONV_V2
    }
    const int BCO = 16 * TCO;
    const int Qw = 64 + (ksz - 1) / stride;
    const int per_ci = stride * Qw + ksz * BCO;
    int CI = 11800 / per_ci;
    if (CI > 16) CI = 16;
    if (CI > Cin) CI = Cin;
    SSRB_CHECK(CI >= 1, "conv1d: kernel too large for the shared-memory tile");
    const size_t smem = ((size_t)((CI * stride * Qw + 3) & ~3) + (size_t)CI * ksz * BCO) * 4;
    dim3 grid(cdiv(Tout, 64), cdiv(Cout, BCO), B);
    if (TCO == 8) SSRB_LAUNCH(conv1d_kernel<8>, grid, 256, smem, s, in, Cin, Tin, W, bias, Cout, ksz, stride, padL, Tout, (int)elu_in, res, out, CI);
    else if (TCO == 4) SSRB_LAUNCH(conv1d_kernel<4>, grid, 256, smem, s, in, Cin, Tin, W, bias, Cout, ksz, stride, padL, Tout, (int)elu_in, res, out, CI);
    else SSRB_LAUNCH(conv1d_kernel<2>, grid, 256, smem, s, in, Cin, Tin, W, bias, Cout, ksz, stride, padL, Tout, (int)elu_in, res, out, CI);
    return 0;
}

// =================================================================================================
// ConvTranspose1d (ksz = 2*stride) decomposed by output phase.   modules/conv.py:221-243
//   full[j] = sum_ci W[ci,co,j%s] x[ci,j/s] + W[ci,co,j%s+s] x[ci,j/s-1];  out[t'] = full[t'+padL]
//   grid (ceil(NI/64), ceil(Cout/BCO), B*stride): one phase p per CTA, tile = BCO channels x 64 input steps.
// =================================================================================================
template <int TCO>
__global__ void __launch_bounds__(256) convtr1d_kernel(const float* __restrict__ in, int Cin, int Tin,
                                                       const float* __restrict__ W, const float* __restrict__ bias,
                                                       int Cout, int ksz, int stride, int padL, int Tout, int elu_in,
                                                       float* __restrict__ out, int CI, int i_first) {
    constexpr int BCO = 16 * TCO, BI = 64;
    extern __shared__ __align__(16) float smem[];
    float* in_s = smem;                        // [CI][BI+1]  (index 0 = i0-1 of the first column)
    float* w_s = smem + ((CI * (BI + 1) + 3) & ~3);   // [CI][2][BCO]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int b = blockIdx.z / stride, p = blockIdx.z - b * stride;
    const int i0 = i_first + blockIdx.x * BI, c0 = blockIdx.y * BCO;
    const float* inb = in + (int64_t)b * Cin * Tin;
    float acc[TCO][4];
#pragma unroll
    for (int i = 0; i < TCO; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (int ci0 = 0; ci0 < Cin; ci0 += CI) {
        const int nci = min(CI, Cin - ci0);
        __syncthreads();
        for (int e = tid; e < nci * (BI + 1); e += 256) {
            const int ci = e / (BI + 1), r = e - ci * (BI + 1);
            const int g = i0 - 1 + r;
            float v = 0.f;
            if (g >= 0 && g < Tin) {
                v = inb[(int64_t)(ci0 + ci) * Tin + g];
                if (elu_in) v = elu1(v);
            }
            in_s[e] = v;
        }
        for (int e = tid; e < nci * 2 * BCO; e += 256) {
            const int co = e % BCO, r = e / BCO, tap = r & 1, ci = r >> 1;
            float v = 0.f;
            if (c0 + co < Cout) v = W[((int64_t)(ci0 + ci) * Cout + c0 + co) * ksz + p + tap * stride];
            w_s[e] = v;                        // [(ci*2+tap)][co]
        }
        __syncthreads();
        for (int ci = 0; ci < nci; ci++) {
            const float* xs = in_s + ci * (BI + 1) + tx;
            const float* ws = w_s + ci * 2 * BCO + ty * TCO;
            float w0[TCO], w1[TCO], x0[4], x1[4];
#pragma unroll
            for (int i = 0; i < TCO; i += 2) {
                const float2 a = *reinterpret_cast<const float2*>(ws + i);
                const float2 c = *reinterpret_cast<const float2*>(ws + BCO + i);
                w0[i] = a.x; w0[i + 1] = a.y; w1[i] = c.x; w1[i + 1] = c.y;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) { x1[j] = xs[16 * j]; x0[j] = xs[16 * j + 1]; }
#pragma unroll
            for (int i = 0; i < TCO; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(w1[i], x1[j], fmaf(w0[i], x0[j], acc[i][j]));
        }
    }
#pragma unroll
    for (int i = 0; i < TCO; i++) {
        const int co = c0 + ty * TCO + i;
        if (co >= Cout) continue;
        const float bv = bias ? bias[co] : 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int ii = i0 + tx + 16 * j;
            const int t = ii * stride + p - padL;
            if (t >= 0 && t < Tout) out[((int64_t)b * Cout + co) * Tout + t] = acc[i][j] + bv;
        }
    }
}

int launch_convtr1d(const float* in, int B, int Cin, int Tin, const float* W, const float* bias, int Cout, int ksz,
                    int stride, int padL, int Tout, bool elu_in, float* out, cudaStream_t s) {
    SSRB_CHECK(ksz == 2 * stride, "convtr1d: kernel_size must be 2*stride");
    const int TCO = Cout >= 128 ? 8 : (Cout >= 64 ? 4 : 2);
    const int BCO = 16 * TCO;
    const int per_ci = 65 + 2 * BCO;
    int CI = 11800 / per_ci;
    if (CI > 32) CI = 32;
    if (CI > Cin) CI = Cin;
    const int i_first = padL / stride;                         // first input step touching out[0]
    const int i_last = (padL + Tout - 1) / stride;
    const int NI = i_last - i_first + 1;
    const size_t smem = ((size_t)((CI * 65 + 3) & ~3) + (size_t)CI * 2 * BCO) * 4;
    dim3 grid(cdiv(NI, 64), cdiv(Cout, BCO), B * stride);
    if (TCO == 8) SSRB_LAUNCH(convtr1d_kernel<8>, grid, 256, smem, s, in, Cin, Tin, W, bias, Cout, ksz, stride, padL, Tout, (int)elu_in, out, CI, i_first);
    else if (TCO == 4) SSRB_LAUNCH(convtr1d_kernel<4>, grid, 256, smem, s, in, Cin, Tin, W, bias, Cout, ksz, stride, padL, Tout, (int)elu_in, out, CI, i_first);
    else SSRB_LAUNCH(convtr1d_kernel<2>, grid, 256, smem, s, in, Cin, Tin, W, bias, Cout, ksz, stride, padL, Tout, (int)elu_in, out, CI, i_first);
    return 0;
}

// =================================================================================================
// layout shuffles around the LSTM (modules/lstm.py:20-25: permute(2,0,1) ... + skip ... permute(1,2,0))
// =================================================================================================
__global__ void bct_to_tbc_kernel(const float* __restrict__ in, int B, int C, int T, float* __restrict__ out, bf16* __restrict__ out_bf16) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, t = t0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && t < T) ? in[((int64_t)b * C + c) * T + t] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int t = t0 + i, c = c0 + threadIdx.x;
        if (t < T && c < C) {
            out[((int64_t)t * B + b) * C + c] = tile[threadIdx.x][i];
            if (out_bf16) out_bf16[((int64_t)t * B + b) * C + c] = __float2bfloat16_rn(tile[threadIdx.x][i]);
        }
    }
}
__global__ void tbc_to_bct_add_kernel(const float* __restrict__ seq, const float* __restrict__ skip, int B, int C, int T,
                                      float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int t = t0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && t < T) ? seq[((int64_t)t * B + b) * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, t = t0 + threadIdx.x;
        if (c < C && t < T) {
            const int64_t o = ((int64_t)b * C + c) * T + t;
            out[o] = tile[threadIdx.x][i] + (skip ? skip[o] : 0.f);
        }
    }
}
int launch_bct_to_tbc(const float* in, int B, int C, int T, float* out, cudaStream_t s, bf16* out_bf16) {
    dim3 grid(cdiv(T, 32), cdiv(C, 32), B), block(32, 8);
    SSRB_LAUNCH(bct_to_tbc_kernel, grid, block, 0, s, in, B, C, T, out, out_bf16);
    return 0;
}
int launch_tbc_to_bct_add(const float* seq, const float* skip, int B, int C, int T, float* out, cudaStream_t s) {
    dim3 grid(cdiv(T, 32), cdiv(C, 32), B), block(32, 8);
    SSRB_LAUNCH(tbc_to_bct_add_kernel, grid, block, 0, s, seq, skip, B, C, T, out);
    return 0;
}
__global__ void bct_to_btc_kernel(const float* __restrict__ in, int B, int C, int T, float* __restrict__ out) {
    const int64_t n = (int64_t)B * C * T;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); const int64_t r = i / C; const int t = (int)(r % T); const int b = (int)(r / T);
        out[i] = in[((int64_t)b * C + c) * T + t];
    }
}
int launch_bct_to_btc(const float* in, int B, int C, int T, float* out, cudaStream_t s) {
    SSRB_LAUNCH(bct_to_btc_kernel, 256, 256, 0, s, in, B, C, T, out);
    return 0;
}

// =================================================================================================
// LSTM recurrence, persistent + cooperative.          modules/lstm.py:17 (nn.LSTM(dim, dim, 2)); gates i,f,g,o
//   Each CTA owns 8 hidden units: its 32 rows of W_hh stay resident in shared memory (fp32) for all T
//   steps; h_{t-1} is exchanged through L2 with one grid-wide barrier per step.  h_{t-1} is staged 8 batch
//   entries per pass through two shared-memory buffers (cp.async): pass p+1 arrives while pass p is multiplied.
// =================================================================================================
constexpr int LSTM_UPB = 8;     // hidden units per CTA
constexpr int LSTM_BC = 8;      // batch entries staged per pass
size_t lstm_smem_bytes(int C) { return ((size_t)4 * LSTM_UPB * C + (size_t)2 * LSTM_BC * C + 4 * LSTM_UPB * 32) * 4; }

__device__ __forceinline__ void cp_async16_zfill(float* dst_smem, const float* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src), "r"(src_bytes) : "memory");
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256) lstm_layer_kernel(const float* __restrict__ pre, const float* __restrict__ Whh,
                                                         float* __restrict__ hseq, float* hbuf, unsigned int* bar,
                                                         int T, int B, int C, bf16* __restrict__ hseq_bf16) {
    extern __shared__ __align__(16) float smem[];
    float* w_s = smem;                               // [32 rows = gate*8+unit][C]
    float* h_s0 = w_s + 4 * LSTM_UPB * C;            // [2][LSTM_BC][C]
    float* g_s = h_s0 + 2 * LSTM_BC * C;             // [32 rows][32 batch]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int u0 = blockIdx.x * LSTM_UPB;
    const unsigned int G = gridDim.x;
    for (int e = tid; e < 4 * LSTM_UPB * C; e += 256) {
        const int rr = e / C, k = e - rr * C;
        const int gate = rr / LSTM_UPB, unit = rr - gate * LSTM_UPB;
        w_s[e] = Whh[((int64_t)gate * C + u0 + unit) * C + k];
    }
    // thread -> (unit, batch) for the cell update; c state lives in a register for the whole sequence
    const int cu = tid % LSTM_UPB, cb = tid / LSTM_UPB;
    float c_state = 0.f;
    float pre_v[4] = {0.f, 0.f, 0.f, 0.f};
    if (cb < B) {
        const float* pr = pre + (size_t)cb * 4 * C + u0 + cu;
        pre_v[0] = pr[0]; pre_v[1] = pr[C]; pre_v[2] = pr[2 * C]; pre_v[3] = pr[3 * C];
    }
    __syncthreads();
    const int npass = (B + LSTM_BC - 1) / LSTM_BC;
    // one pass of h_{t-1}: LSTM_BC batch entries x C, 16 bytes per request (C % 4 == 0: a request never straddles rows); entries
    // past the batch are zero-filled (src-size 0).  cp.async.cg reads through L2, where the other CTAs' h_t stores were made
    // visible by the grid barrier.
    auto issue_pass = [&](const float* hprev, int b0, float* dst) {
        const int nb = min(LSTM_BC, B - b0);
        for (int e = tid * 4; e < LSTM_BC * C; e += 1024) {
            const int bb = e / C;
            const bool ok = bb < nb;
            cp_async16_zfill(dst + e, hprev + (size_t)(b0 + (ok ? bb : 0)) * C + (e - bb * C), ok ? 16 : 0);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int t = 0; t < T; t++) {
        const float* hprev = hbuf + (size_t)(t & 1) * B * C;
        float* hnext = hbuf + (size_t)((t + 1) & 1) * B * C;
        issue_pass(hprev, 0, h_s0);
        for (int p = 0; p < npass; p++) {
            const int b0 = p * LSTM_BC;
            const int nb = min(LSTM_BC, B - b0);
            const float* h_s = h_s0 + (p & 1) * LSTM_BC * C;
            if (p + 1 < npass) {
                issue_pass(hprev, b0 + LSTM_BC, h_s0 + ((p + 1) & 1) * LSTM_BC * C);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncthreads();
            // 4 gate rows x 8 batch entries per warp; each lane owns 4 consecutive k per 128-wide slab (LDS.128, no bank
            // conflicts), so one loop trip issues 12 shared loads for 128 independent FMAs
            float acc[32];
#pragma unroll
            for (int i = 0; i < 32; i++) acc[i] = 0.f;
            for (int k = lane * 4; k < C; k += 128) {
                float4 w[4], h[LSTM_BC];
#pragma unroll
                for (int i = 0; i < 4; i++) w[i] = *reinterpret_cast<const float4*>(w_s + (warp * 4 + i) * C + k);
#pragma unroll
                for (int j = 0; j < LSTM_BC; j++) h[j] = *reinterpret_cast<const float4*>(h_s + j * C + k);
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < LSTM_BC; j++) {
                        float a = acc[i * LSTM_BC + j];
                        a = fmaf(w[i].x, h[j].x, a); a = fmaf(w[i].y, h[j].y, a);
                        a = fmaf(w[i].z, h[j].z, a); a = fmaf(w[i].w, h[j].w, a);
                        acc[i * LSTM_BC + j] = a;
                    }
            }
            // halving butterfly: 31 shuffles reduce the 32 accumulators over the 32 lanes; lane L ends with accumulator L
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < off; i++) {
                    const float send = up ? acc[i] : acc[i + off];
                    const float keep = up ? acc[i + off] : acc[i];
                    acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            {
                const int i = lane / LSTM_BC, j = lane % LSTM_BC;
                if (j < nb) g_s[(warp * 4 + i) * 32 + ((b0 + j) & 31)] = acc[0];
            }
            __syncthreads();             // g_s complete; this pass's buffer may be refilled two passes on
        }
        if (cb < B) {
            const float gi = g_s[(0 * LSTM_UPB + cu) * 32 + cb] + pre_v[0];
            const float gf = g_s[(1 * LSTM_UPB + cu) * 32 + cb] + pre_v[1];
            const float gg = g_s[(2 * LSTM_UPB + cu) * 32 + cb] + pre_v[2];
            const float go = g_s[(3 * LSTM_UPB + cu) * 32 + cb] + pre_v[3];
            c_state = sigmoidf_(gf) * c_state + sigmoidf_(gi) * tanhf(gg);
            const float hv = sigmoidf_(go) * tanhf(c_state);
            hnext[(size_t)cb * C + u0 + cu] = hv;
            hseq[((size_t)t * B + cb) * C + u0 + cu] = hv;
            if (hseq_bf16) hseq_bf16[((size_t)t * B + cb) * C + u0 + cu] = __float2bfloat16_rn(hv);
            if (t + 1 < T) {                                  // input-projection terms of the next step: off the critical path
                const float* pr = pre + ((size_t)(t + 1) * B + cb) * 4 * C + u0 + cu;
                pre_v[0] = pr[0]; pre_v[1] = pr[C]; pre_v[2] = pr[2 * C]; pre_v[3] = pr[3 * C];
            }
        }
        // grid barrier (monotonic counter; all CTAs are co-resident: cooperative launch).  release/acquire at gpu scope
        // orders the h_t stores of every CTA before the h_t loads of every other CTA.
        __syncthreads();
        if (tid == 0) {
            const unsigned int target = G * (unsigned int)(t + 1);
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
            unsigned int v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            } while (v < target);
        }
        __syncthreads();
    }
}

int launch_lstm_layer(const float* pre, const float* Whh, float* hseq, float* hbuf, unsigned int* bar, int T, int B,
                      int C, cudaStream_t s, bf16* hseq_bf16) {
    SSRB_CHECK(C % LSTM_UPB == 0, "lstm: hidden size must be a multiple of 8");
    SSRB_CHECK(B <= 32, "lstm: batch chunk must be <= 32");
    const size_t smem = lstm_smem_bytes(C);
    SSRB_CUDA(cudaFuncSetAttribute(lstm_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SSRB_CUDA(cudaMemsetAsync(hbuf, 0, (size_t)2 * B * C * 4, s));
    SSRB_CUDA(cudaMemsetAsync(bar, 0, 4, s));
    void* args[] = {(void*)&pre, (void*)&Whh, (void*)&hseq, (void*)&hbuf, (void*)&bar, (void*)&T, (void*)&B, (void*)&C, (void*)&hseq_bf16};
    SSRB_CUDA(cudaLaunchCooperativeKernel((void*)lstm_layer_kernel, dim3(C / LSTM_UPB), dim3(256), args, smem, s));
    g_launch_count++;
    return 0;
}

// =================================================================================================
// LSTM recurrence on the tensor cores (bf16 decoder path only; the encoder keeps the fp32 kernel above because its
// output decides RVQ indices).  Same decomposition — one cooperative CTA per 8 hidden units, one grid barrier per
// step — but h_{t-1} . W_hh^T is D[32 gate rows x 32 batch] = W[32 x C] . h^T[C x 32] as bf16 mma.sync m16n8k16
// with fp32 accumulation: warp w owns the k-slice [w*C/8, (w+1)*C/8); its W fragments are loaded ONCE and stay in
// registers for all T steps, its h fragments are read straight from L2 (16-byte loads; the k order inside a slice
// is permuted identically for both operands so that four lanes read 64 contiguous bytes), the 8 partial tiles meet in
// shared memory.  Cell state and gate math stay fp32.
// =================================================================================================
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int KS16>   // k-steps of 16 per warp: C = 8 warps x KS16 x 16
__global__ void __launch_bounds__(256, 1) lstm_layer_mma_kernel(const float* __restrict__ pre, const float* __restrict__ Whh,
                                                                float* __restrict__ hseq, bf16* hbuf /*[2][32][C]*/,
                                                                unsigned int* bar, int T, int B, int C,
                                                                bf16* __restrict__ hseq_bf16) {
    __shared__ float red[8][4][32][LSTM_UPB];          // [warp][gate][batch][unit] partial pre-activations
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int u0 = blockIdx.x * LSTM_UPB;
    const unsigned int G = gridDim.x;
    const int kbase = warp * KS16 * 16;
    // W fragments: tile row r = gate*8 + unit  <->  W_hh row gate*C + u0 + unit
    uint32_t afr[2][KS16][4];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int ks = 0; ks < KS16; ks++) {
            const int kp = kbase + (ks >> 1) * 32 + q * 8 + (ks & 1) * 4;
            const float* plo = Whh + ((size_t)(2 * mt) * C + u0 + g) * C + kp;        // row mt*16 + g     : gate 2mt,   unit g
            const float* phi = Whh + ((size_t)(2 * mt + 1) * C + u0 + g) * C + kp;    // row mt*16 + g + 8 : gate 2mt+1, unit g
            afr[mt][ks][0] = pack_bf16x2(plo[0], plo[1]);
            afr[mt][ks][1] = pack_bf16x2(phi[0], phi[1]);
            afr[mt][ks][2] = pack_bf16x2(plo[2], plo[3]);
            afr[mt][ks][3] = pack_bf16x2(phi[2], phi[3]);
        }
    const int cu = tid % LSTM_UPB, cb = tid / LSTM_UPB;
    float c_state = 0.f;
    float pre_v[4] = {0.f, 0.f, 0.f, 0.f};
    if (cb < B) {
        const float* pr = pre + (size_t)cb * 4 * C + u0 + cu;
        pre_v[0] = pr[0]; pre_v[1] = pr[C]; pre_v[2] = pr[2 * C]; pre_v[3] = pr[3 * C];
    }
    for (int t = 0; t < T; t++) {
        const bf16* hprev = hbuf + (size_t)(t & 1) * 32 * C;
        bf16* hnext = hbuf + (size_t)((t + 1) & 1) * 32 * C;
        uint4 bv[4][KS16 / 2];
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < KS16 / 2; i++)
                bv[j][i] = __ldcg(reinterpret_cast<const uint4*>(hprev + (size_t)(j * 8 + g) * C + kbase + i * 32 + q * 8));
        float acc[2][4][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int z = 0; z < 4; z++) acc[mt][j][z] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS16; ks++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint4 v = bv[j][ks >> 1];
                const uint32_t b0 = (ks & 1) ? v.z : v.x, b1 = (ks & 1) ? v.w : v.y;
#pragma unroll
                for (int mt = 0; mt < 2; mt++) mma_bf16_16816(acc[mt][j], afr[mt][ks], b0, b1);
            }
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int n0 = j * 8 + 2 * q;
                red[warp][2 * mt][n0][g] = acc[mt][j][0];
                red[warp][2 * mt][n0 + 1][g] = acc[mt][j][1];
                red[warp][2 * mt + 1][n0][g] = acc[mt][j][2];
                red[warp][2 * mt + 1][n0 + 1][g] = acc[mt][j][3];
            }
        __syncthreads();
        if (cb < B) {
            float gs[4];
#pragma unroll
            for (int gate = 0; gate < 4; gate++) {
                float a = 0.f;
#pragma unroll
                for (int w = 0; w < 8; w++) a += red[w][gate][cb][cu];      // fixed order: deterministic
                gs[gate] = a + pre_v[gate];
            }
            c_state = sigmoidf_(gs[1]) * c_state + sigmoidf_(gs[0]) * tanhf(gs[2]);
            const float hv = sigmoidf_(gs[3]) * tanhf(c_state);
            hnext[(size_t)cb * C + u0 + cu] = __float2bfloat16_rn(hv);
            hseq[((size_t)t * B + cb) * C + u0 + cu] = hv;
            if (hseq_bf16) hseq_bf16[((size_t)t * B + cb) * C + u0 + cu] = __float2bfloat16_rn(hv);
            if (t + 1 < T) {
                const float* pr = pre + ((size_t)(t + 1) * B + cb) * 4 * C + u0 + cu;
                pre_v[0] = pr[0]; pre_v[1] = pr[C]; pre_v[2] = pr[2 * C]; pre_v[3] = pr[3 * C];
            }
        }
        __syncthreads();
        if (tid == 0) {
            const unsigned int target = G * (unsigned int)(t + 1);
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
            unsigned int v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            } while (v < target);
        }
        __syncthreads();
    }
}

bool lstm_mma_supported(int C) { return C == 256 || C == 512 || C == 1024; }

int launch_lstm_layer_mma(const float* pre, const float* Whh, float* hseq, bf16* hbuf, unsigned int* bar, int T, int B,
                          int C, cudaStream_t s, bf16* hseq_bf16) {
    SSRB_CHECK(lstm_mma_supported(C), "lstm (tensor-core): hidden size must be 256, 512 or 1024");
    SSRB_CHECK(B <= 32, "lstm: batch chunk must be <= 32");
    SSRB_CUDA(cudaMemsetAsync(hbuf, 0, (size_t)2 * 32 * C * 2, s));
    SSRB_CUDA(cudaMemsetAsync(bar, 0, 4, s));
    void* args[] = {(void*)&pre, (void*)&Whh, (void*)&hseq, (void*)&hbuf, (void*)&bar, (void*)&T, (void*)&B, (void*)&C, (void*)&hseq_bf16};
    const void* fn = C == 1024 ? (const void*)lstm_layer_mma_kernel<8> : C == 512 ? (const void*)lstm_layer_mma_kernel<4> : (const void*)lstm_layer_mma_kernel<2>;
    SSRB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(C / LSTM_UPB), dim3(256), args, 0, s));
    g_launch_count++;
    return 0;
}

// =================================================================================================
// RVQ encode: per stage  ind = argmax(-(|r|^2 - 2 r.E^T + |E|^2)), r -= E[ind]      quantization/core_vq.py:164-172,382-392
//   one CTA per 8 frames; each thread scans bins/256 codes; first-index tie-break like torch.max.
// =================================================================================================
constexpr int RVQ_F = 8;
__global__ void __launch_bounds__(256) rvq_stage_kernel(float* __restrict__ resid /*[N,Dm]*/, int N, int Dm,
                                                        const float* __restrict__ E, const float* __restrict__ Esq,
                                                        int bins, long long* __restrict__ codes, int B, int T, int n_q, int q) {
    __shared__ float xs[RVQ_F][128];
    __shared__ float xx[RVQ_F];
    __shared__ float bestv[RVQ_F][8];
    __shared__ int besti[RVQ_F][8];
    const int n0 = blockIdx.x * RVQ_F, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < RVQ_F * Dm; e += 256) {
        const int f = e / Dm, d = e - f * Dm;
        xs[f][d] = (n0 + f) < N ? resid[(int64_t)(n0 + f) * Dm + d] : 0.f;
    }
    __syncthreads();
    if (tid < RVQ_F) {
        float s = 0.f;
        for (int d = 0; d < Dm; d++) s = fmaf(xs[tid][d], xs[tid][d], s);
        xx[tid] = s;
    }
    __syncthreads();
    float bv[RVQ_F]; int bi[RVQ_F];
#pragma unroll
    for (int f = 0; f < RVQ_F; f++) { bv[f] = -INFINITY; bi[f] = 0x7fffffff; }
    for (int c = tid; c < bins; c += 256) {
        const float* er = E + (int64_t)c * Dm;
        float dot[RVQ_F];
#pragma unroll
        for (int f = 0; f < RVQ_F; f++) dot[f] = 0.f;
        for (int d = 0; d < Dm; d += 4) {
            const float4 e4 = *reinterpret_cast<const float4*>(er + d);
#pragma unroll
            for (int f = 0; f < RVQ_F; f++) {
                dot[f] = fmaf(xs[f][d], e4.x, dot[f]); dot[f] = fmaf(xs[f][d + 1], e4.y, dot[f]);
                dot[f] = fmaf(xs[f][d + 2], e4.z, dot[f]); dot[f] = fmaf(xs[f][d + 3], e4.w, dot[f]);
            }
        }
        const float esq = Esq[c];
#pragma unroll
        for (int f = 0; f < RVQ_F; f++) {
            const float dist = -(__fadd_rn(__fsub_rn(xx[f], __fmul_rn(2.f, dot[f])), esq));
            if (dist > bv[f]) { bv[f] = dist; bi[f] = c; }      // c ascending per thread: first index kept
        }
    }
#pragma unroll
    for (int f = 0; f < RVQ_F; f++) {
        float v = bv[f]; int i = bi[f];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
        }
        if (lane == 0) { bestv[f][warp] = v; besti[f][warp] = i; }
    }
    __syncthreads();
    __shared__ int win[RVQ_F];
    if (tid < RVQ_F) {
        float v = bestv[tid][0]; int i = besti[tid][0];
        for (int w = 1; w < 8; w++)
            if (bestv[tid][w] > v || (bestv[tid][w] == v && besti[tid][w] < i)) { v = bestv[tid][w]; i = besti[tid][w]; }
        win[tid] = i;
        const int n = n0 + tid;
        if (n < N) { const int b = n / T, t = n - b * T; codes[((int64_t)b * n_q + q) * T + t] = i; }
    }
    __syncthreads();
    for (int e = tid; e < RVQ_F * Dm; e += 256) {
        const int f = e / Dm, d = e - f * Dm;
        if (n0 + f < N) resid[(int64_t)(n0 + f) * Dm + d] = xs[f][d] - E[(int64_t)win[f] * Dm + d];
    }
}

int launch_rvq_encode(const float* emb, int B, int Dm, int T, const float* codebooks, const float* cb_sq, int n_q,
                      int bins, float* residual_ws, long long* codes, cudaStream_t s) {
    SSRB_CHECK(Dm <= 128 && Dm % 4 == 0, "rvq: dimension must be <= 128 and a multiple of 4");
    SSRB_TRY(launch_bct_to_btc(emb, B, Dm, T, residual_ws, s));   // [B,Dm,T] -> [B*T, Dm]
    const int N = B * T;
    for (int q = 0; q < n_q; q++)
        SSRB_LAUNCH(rvq_stage_kernel, cdiv(N, RVQ_F), 256, 0, s, residual_ws, N, Dm, codebooks + (size_t)q * bins * Dm,
                    cb_sq + (size_t)q * bins, bins, codes, B, T, n_q, q);
    return 0;
}

// RVQ decode: sum_q E_q[codes_q]  (core_vq.py:394-400 — starts from tensor(0.0) and adds in order) -> [B,Dm,T]
__global__ void rvq_decode_kernel(const long long* __restrict__ codes, int n_q, int T, const float* __restrict__ E,
                                  int bins, int Dm, float* __restrict__ out) {
    const int b = blockIdx.y, t = blockIdx.x * 32 + (threadIdx.x & 31);
    if (t >= T) return;
    for (int d = threadIdx.x >> 5; d < Dm; d += 8) {
        float v = 0.f;
        for (int q = 0; q < n_q; q++) {
            const long long c = codes[((int64_t)b * n_q + q) * T + t];
            v = __fadd_rn(v, E[((int64_t)q * bins + c) * Dm + d]);
        }
        out[((int64_t)b * Dm + d) * T + t] = v;
    }
}
int launch_rvq_decode(const long long* codes, int B, int n_q, int T, const float* codebooks, int bins, int Dm,
                      float* out, cudaStream_t s) {
    dim3 grid(cdiv(T, 32), B);
    SSRB_LAUNCH(rvq_decode_kernel, grid, 256, 0, s, codes, n_q, T, codebooks, bins, Dm, out);
    return 0;
}

// watermark conditioning input: cat([skip, wm_embed(repeat_interleave(marks, rep))^T], dim=1)   modules/seanet.py:563-590
__global__ void concat_marks_kernel(const float* __restrict__ skip, int C, int T, const long long* __restrict__ marks,
                                    int Tm, int rep, const float* __restrict__ wm_embed, int E, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int64_t n = (int64_t)(C + E) * T;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i / T), t = (int)(i - (int64_t)c * T);
        float v;
        if (c < C) v = skip[((int64_t)b * C + c) * T + t];
        else {
            const long long m = marks[(int64_t)b * Tm + t / rep];
            v = wm_embed[m * E + (c - C)];
        }
        out[(int64_t)b * n + i] = v;
    }
}
int launch_concat_marks(const float* skip, int B, int C, int T, const long long* marks, int Tm, int rep,
                        const float* wm_embed, int E, float* out, cudaStream_t s) {
    int64_t nb = ((int64_t)(C + E) * T + 1023) / 1024;
    dim3 grid((unsigned)(nb > 4096 ? 4096 : (nb < 1 ? 1 : nb)), B);
    SSRB_LAUNCH(concat_marks_kernel, grid, 256, 0, s, skip, C, T, marks, Tm, rep, wm_embed, E, out);
    return 0;
}

}  // namespace ssrb
