// codec_cl.cu — boundary kernels of the channels-last bf16 codec path (conv_tc.cu): the 1-channel first / last
// convolutions (memory-bound, CUDA cores) and the layout converters around the frame-rate fp32 region (LSTM).
#include "conv_tc.cuh"

namespace ssrb {

// SEANet first conv (channels=1 -> C, kernel k, stride 1; seanet.py:118-122): wav fp32 [B,1,T] -> cl [B][G+T+G][C]
__global__ void __launch_bounds__(256) cl_first_conv_kernel(const float* __restrict__ wav, int T, const float* __restrict__ W,
                                                            const float* __restrict__ bias, int C, int k, bf16* __restrict__ out_raw,
                                                            bf16* __restrict__ out_act) {
    __shared__ float xs[64 + 16];
    __shared__ float ws[64 * 16];
    const int b = blockIdx.y, t0 = blockIdx.x * 64, padL = (k - 1) - (k - 1) / 2;
    for (int e = threadIdx.x; e < 64 + k - 1; e += 256) {
        const int g = t0 + e - padL;
        xs[e] = (g >= 0 && g < T) ? wav[(int64_t)b * T + g] : 0.f;
    }
    for (int e = threadIdx.x; e < C * k; e += 256) { const int cc = e / k, j = e - cc * k; ws[j * 64 + cc] = W[e]; }      // [tap][channel]: conflict-free pairs
    __syncthreads();
    // a thread owns two adjacent channels (one 4-byte store per output array and time step; a warp writes one whole 128-byte row when
    // C = 64) and every 8th time step of the tile
    const int c = 2 * (threadIdx.x % 32), tg = threadIdx.x / 32;
    if (c >= C) return;
    const bool two = c + 1 < C;
    const float b0 = bias[c], b1 = two ? bias[c + 1] : 0.f;
    const int64_t base = (int64_t)b * (T + 2 * CL_GUARD) * C;
    for (int i = 0; i < 8; i++) {
        const int tl = tg + 8 * i, t = t0 + tl;
        if (t >= T) break;
        float a0 = b0, a1 = b1;
        for (int j = 0; j < k; j++) {
            a0 = fmaf(ws[j * 64 + c], xs[tl + j], a0);
            if (two) a1 = fmaf(ws[j * 64 + c + 1], xs[tl + j], a1);
        }
        const int64_t o = base + (int64_t)(CL_GUARD + t) * C + c;
        if (two && (C & 1) == 0) {
            if (out_raw) *reinterpret_cast<__nv_bfloat162*>(out_raw + o) = __floats2bfloat162_rn(a0, a1);
            if (out_act) *reinterpret_cast<__nv_bfloat162*>(out_act + o) = __floats2bfloat162_rn(elu1_bf16(a0), elu1_bf16(a1));
        } else {
            if (out_raw) { out_raw[o] = __float2bfloat16_rn(a0); if (two) out_raw[o + 1] = __float2bfloat16_rn(a1); }
            if (out_act) { out_act[o] = __float2bfloat16_rn(elu1_bf16(a0)); if (two) out_act[o + 1] = __float2bfloat16_rn(elu1_bf16(a1)); }
        }
    }
}
int launch_cl_first_conv(const float* wav, int B, int T, const float* W, const float* bias, int C, int k, bf16* out_raw,
                         bf16* out_act, cudaStream_t s) {
    SSRB_CHECK(C <= 64 && k <= 16, "cl_first_conv: unsupported shape");
    dim3 grid(cdiv(T, 64), B);
    SSRB_LAUNCH(cl_first_conv_kernel, grid, 256, 0, s, wav, T, W, bias, C, k, out_raw, out_act);
    return 0;
}

// SEANet last conv (C -> 1, kernel k; seanet.py:240-245) on an already ELU'd cl input -> wav fp32 [B,1,T]
__global__ void __launch_bounds__(128) cl_last_conv_kernel(const bf16* __restrict__ in_act, int T, int C, const float* __restrict__ W,
                                                           const float* __restrict__ bias, int k, float* __restrict__ wav) {
    extern __shared__ float sm[];
    const int pitch = C + 1;
    float* xs = sm;                                   // [(128 + k - 1)][C + 1]
    float* ws = sm + (128 + k - 1) * pitch;           // [k][C]
    const int b = blockIdx.y, t0 = blockIdx.x * 128, padL = (k - 1) - (k - 1) / 2;
    const bf16* src = in_act + (int64_t)b * (T + 2 * CL_GUARD) * C + (int64_t)(CL_GUARD + t0 - padL) * C;   // guards supply the zeros
    const int nrow = 128 + k - 1;
    for (int e = threadIdx.x * 8; e < nrow * C; e += 128 * 8) {
        const int r = e / C, c = e - r * C;
        float v[8];
        if (t0 - padL + r < T + CL_GUARD) load8(src + e, v);
        else {
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) xs[r * pitch + c + j] = v[j];
    }
    for (int e = threadIdx.x; e < C * k; e += 128) { const int c = e / k, j = e - c * k; ws[j * C + c] = W[e]; }
    __syncthreads();
    const int t = t0 + threadIdx.x;
    if (t >= T) return;
    float acc = bias[0];
    for (int j = 0; j < k; j++) {
        const float* xr = xs + (threadIdx.x + j) * pitch;
        const float* wr = ws + j * C;
        for (int c = 0; c < C; c++) acc = fmaf(wr[c], xr[c], acc);
    }
    wav[(int64_t)b * T + t] = acc;
}
int launch_cl_last_conv(const bf16* in_act, int B, int T, int C, const float* W, const float* bias, int k, float* wav, cudaStream_t s) {
    SSRB_CHECK(C % 8 == 0 && k <= 16, "cl_last_conv: unsupported shape");
    const size_t smem = ((size_t)(128 + k - 1) * (C + 1) + (size_t)k * C) * 4;
    SSRB_CUDA(cudaFuncSetAttribute(cl_last_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(cdiv(T, 128), B);
    SSRB_LAUNCH(cl_last_conv_kernel, grid, 128, smem, s, in_act, T, C, W, bias, k, wav);
    return 0;
}

// fp32 channels-first [B,C,T] -> bf16 channels-last with guards (optionally ELU'd)
__global__ void cf32_to_cl_kernel(const float* __restrict__ in, int C, int T, int elu, bf16* __restrict__ out) {
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
            float v = tile[threadIdx.x][i];
            if (elu) v = elu1_bf16(v);
            out[((int64_t)b * (T + 2 * CL_GUARD) + CL_GUARD + t) * C + c] = __float2bfloat16_rn(v);
        }
    }
}
int launch_cf32_to_cl(const float* in, int B, int C, int T, bool elu, bf16* out, cudaStream_t s) {
    dim3 grid(cdiv(T, 32), cdiv(C, 32), B), block(32, 8);
    SSRB_LAUNCH(cf32_to_cl_kernel, grid, block, 0, s, in, C, T, (int)elu, out);
    return 0;
}
__global__ void cl_to_cf32_kernel(const bf16* __restrict__ in, int C, int T, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int t = t0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && t < T) ? __bfloat162float(in[((int64_t)b * (T + 2 * CL_GUARD) + CL_GUARD + t) * C + c]) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, t = t0 + threadIdx.x;
        if (c < C && t < T) out[((int64_t)b * C + c) * T + t] = tile[threadIdx.x][i];
    }
}
int launch_cl_to_cf32(const bf16* in, int B, int C, int T, float* out, cudaStream_t s) {
    dim3 grid(cdiv(T, 32), cdiv(C, 32), B), block(32, 8);
    SSRB_LAUNCH(cl_to_cf32_kernel, grid, block, 0, s, in, C, T, out);
    return 0;
}

// zero only the guard rows of a channels-last tensor (rows [0,T) are fully written by the producing epilogue)
__global__ void cl_zero_guards_kernel(bf16* __restrict__ p, int T, int C) {
    const int b = blockIdx.x, side = blockIdx.y;
    bf16* g = p + ((int64_t)b * (T + 2 * CL_GUARD) + (side ? CL_GUARD + T : 0)) * C;
    const int n = CL_GUARD * C;                               // multiple of 8
    for (int e = threadIdx.x * 8; e < n; e += blockDim.x * 8) *reinterpret_cast<uint4*>(g + e) = make_uint4(0, 0, 0, 0);
}
int launch_cl_zero_guards(bf16* p, int B, int T, int C, cudaStream_t s) {
    SSRB_CHECK(C % 8 == 0, "cl_zero_guards: C must be a multiple of 8");
    dim3 grid(B, 2);
    SSRB_LAUNCH(cl_zero_guards_kernel, grid, 256, 0, s, p, T, C);
    return 0;
}

}  // namespace ssrb
