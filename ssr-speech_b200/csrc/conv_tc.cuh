// conv_tc.cuh — tensor-core (tcgen05) convolution as a tap-accumulated GEMM over channels-last bf16 activations.
#pragma once
#include "common.cuh"

namespace ssrb {

constexpr int CL_GUARD = 40;     // zero guard rows on both sides of every utterance (multiple of every stride: 2,4,5,8)

struct ConvTcArgs {
    // A operand: view(b)[row][c] = x[b * x_bstride + x_base_off + row * Cw + c]
    const bf16* x = nullptr; int B = 0; long long x_bstride = 0, x_base_off = 0; int Cw = 0, rows_v = 0;
    // W' [taps * N, Cw] bf16, K-major
    const bf16* w = nullptr; int taps = 0, N = 0;
    int T_rows = 0;                                   // GEMM rows per utterance
    const float* bias = nullptr; int bias_mod = 0;    // bias index = n % bias_mod
    const float* bias_alt = nullptr;                  // used instead of `bias` on rows whose mark is 1
    const long long* marks = nullptr; int marks_T = 0, marks_rep = 1;
    const bf16* res = nullptr; long long res_bstride = 0, res_off = 0;     // residual[b][res_off + row*N + n] (raw bf16)
    bf16* out_raw = nullptr; bf16* out_act = nullptr;                      // out[b][out_off + row*N + n], ELU'd copy
    long long out_bstride = 0, out_off = 0, valid_lo = 0, valid_hi = 0;    // store iff valid_lo <= row*N+n < valid_hi
};
int conv_tc(const ConvTcArgs& a, cudaStream_t s);

// ---- a whole SEANetResnetBlock, depth-fused (resblock_tc.cu): y = x + conv_k1(ELU(conv_k3(ELU(x)))), hidden activation on chip ----
struct ResblockTcArgs {
    int B = 0, T = 0, C = 0;
    const bf16* x_act = nullptr; long long x_bstride = 0, x_base_off = 0; int rows_v = 0;   // ELU(x), channels-last, view shifted by the k3 padding
    const bf16* x_raw = nullptr; long long x_raw_off = 0;                                   // x for the skip: x_raw[b*x_bstride + x_raw_off + row*C + c]
    const bf16* w1 = nullptr; int w1_N = 0, w1_Cw = 0; const float* b1 = nullptr;           // [3][w1_N][w1_Cw] (conv_tc repack)
    const bf16* w2 = nullptr; int w2_N = 0, w2_Cw = 0; const float* b2 = nullptr;           // [1][w2_N][w2_Cw]
    bf16* out_raw = nullptr; bf16* out_act = nullptr; long long out_bstride = 0, out_off = 0;
};
bool resblock_tc_supported(int C);
int resblock_tc(const ResblockTcArgs& a, cudaStream_t s);

// ---- encoder path: fp32 channels-last, every operand split into two TF32 numbers (conv_tc32.cu) -------------------------------
struct ConvTc32Args {
    // A operand (hi and lo arrays, same geometry): view(b)[row][c] = x[b * x_bstride + x_base_off + row * Cw + c]
    const float* x_hi = nullptr; const float* x_lo = nullptr; int B = 0; long long x_bstride = 0, x_base_off = 0; int Cw = 0, rows_v = 0;
    // W' [taps * N, Cw] fp32 (hi, lo), K-major
    const float* w_hi = nullptr; const float* w_lo = nullptr; int taps = 0, N = 0;
    int T_rows = 0;
    const float* bias = nullptr;                                              // [N]
    const float* res = nullptr; long long res_bstride = 0, res_off = 0;       // fp32 residual[b][res_off + row*N + n]
    float* out_raw = nullptr; float* out_hi = nullptr; float* out_lo = nullptr;   // out[b][out_off + row*N + n]; hi/lo = split of (ELU'd) out
    long long out_bstride = 0, out_off = 0;
    bool elu = true;
};
int conv_tc32(const ConvTc32Args& a, cudaStream_t s);
int launch_cl32_first_conv(const float* wav, int B, int T, const float* W /*[C][1][k]*/, const float* bias, int C, int k, float* out_raw,
                           float* out_hi, float* out_lo, cudaStream_t s);
int launch_cl32_zero_guards(float* p, int B, int T, int C, cudaStream_t s);
int launch_cl32_to_cf32(const float* in, int B, int C, int T, float* out, cudaStream_t s);
int launch_split_tf32(const float* in, float* hi, float* lo, long long n, cudaStream_t s);   // x = hi + lo, both TF32-representable

// channels-last helpers (codec_cl.cu).  cl tensors: [B][CL_GUARD + T + CL_GUARD][C] bf16.
int launch_cl_first_conv(const float* wav, int B, int T, const float* W /*[C][1][k]*/, const float* bias, int C, int k,
                         bf16* out_raw, bf16* out_act, cudaStream_t s);
int launch_cl_last_conv(const bf16* in_act, int B, int T, int C, const float* W /*[1][C][k]*/, const float* bias, int k,
                        float* wav, cudaStream_t s);
int launch_cf32_to_cl(const float* in, int B, int C, int T, bool elu, bf16* out, cudaStream_t s);
int launch_cl_to_cf32(const bf16* in, int B, int C, int T, float* out, cudaStream_t s);
int launch_cl_zero_guards(bf16* p, int B, int T, int C, cudaStream_t s);

}  // namespace ssrb
