// codec_kernels.cuh — launch wrappers of the WM-Encodec kernels (codec_kernels.cu).
#pragma once
#include "common.cuh"

namespace ssrb {

// out[b,co,t] = bias[co] + sum_{ci,k} W[co,ci,k] * f(in[b,ci,t*s+k-padL]) (+ res[b,co,t]);  f = ELU if elu_in.
// Zero padding outside [0,Tin) (audiocraft/modules/conv.py:185-201).  W layout [Cout][Cin][ksz].
int launch_conv1d(const float* in, int B, int Cin, int Tin, const float* W, const float* bias, int Cout, int ksz,
                  int stride, int padL, int Tout, bool elu_in, const float* res, float* out, cudaStream_t s,
                  const float* Wt = nullptr);
// Wt [Cin][ksz][Cout] <- W [Cout][Cin][ksz]: operand layout of the register-tiled kernel (conv1d_v2_kernel)
int launch_conv_w_transpose(const float* W, float* Wt, int Cout, int Cin, int ksz, cudaStream_t s);
// transposed conv with the reference's trim (conv.py:221-243).  W layout [Cin][Cout][ksz], ksz == 2*stride.
int launch_convtr1d(const float* in, int B, int Cin, int Tin, const float* W, const float* bias, int Cout, int ksz,
                    int stride, int padL, int Tout, bool elu_in, float* out, cudaStream_t s);
// [B,C,T] -> [T,B,C]
int launch_bct_to_tbc(const float* in, int B, int C, int T, float* out, cudaStream_t s, bf16* out_bf16 = nullptr);
// out[B,C,T] = seq[T,B,C] + skip[B,C,T]        (lstm.py:21-25)
int launch_tbc_to_bct_add(const float* seq, const float* skip, int B, int C, int T, float* out, cudaStream_t s);
// persistent recurrent pass of one LSTM layer (lstm.py:17 nn.LSTM): pre [T,B,4C] = x.Wih^T + b_ih + b_hh
int launch_lstm_layer(const float* pre, const float* Whh, float* hseq, float* hbuf, unsigned int* bar, int T, int B,
                      int C, cudaStream_t s, bf16* hseq_bf16 = nullptr);
size_t lstm_smem_bytes(int C);
// the same recurrence with bf16 W_hh / h on mma.sync (fp32 accumulate, fp32 cell state); hbuf = [2][32][C] bf16
bool lstm_mma_supported(int C);
int launch_lstm_layer_mma(const float* pre, const float* Whh, float* hseq, bf16* hbuf, unsigned int* bar, int T, int B,
                          int C, cudaStream_t s, bf16* hseq_bf16 = nullptr);
// RVQ (core_vq.py:164-193,382-400).  emb [B,Dm,T] fp32; codebooks [n_q][bins][Dm]; cb_sq [n_q][bins]
int launch_rvq_encode(const float* emb, int B, int Dm, int T, const float* codebooks, const float* cb_sq, int n_q,
                      int bins, float* residual_ws, long long* codes, cudaStream_t s);
int launch_rvq_decode(const long long* codes, int B, int n_q, int T, const float* codebooks, int bins, int Dm,
                      float* out, cudaStream_t s);
// cat([skip[B,C,T], wm_embed(marks repeated x rep)[B,E,T]]) -> out [B,C+E,T]   (seanet.py:577-590)
int launch_concat_marks(const float* skip, int B, int C, int T, const long long* marks, int Tm, int rep,
                        const float* wm_embed, int E, float* out, cudaStream_t s);
// [B,C,T] -> [B,T,C]
int launch_bct_to_btc(const float* in, int B, int C, int T, float* out, cudaStream_t s);

}  // namespace ssrb
