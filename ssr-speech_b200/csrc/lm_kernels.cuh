// lm_kernels.cuh — launch wrappers of the LM kernels (lm_kernels.cu), used by lm_engine.cu.
#pragma once
#include "common.cuh"
#include "../../include/ssr_b200.h"

namespace ssrb {

// per-utterance decode state — the Python locals of the reference's span loop (models/ssr.py:647-652)
// plus batch bookkeeping.  Lives on the device; the host only polls `done`.
struct UttState {
    int num_gen, num_eog, cfg_tag, prev_token, consec_silence;   // ssr.py:647-652
    int span_idx, n_spans, done;
    int x_len;        // text length (length guard ssr.py:739)
    int y_len;        // audio positions already in the KV cache
    int n_tok;        // iterations recorded so far (all spans)
    int span_len[SSRB_MAX_SPANS];
    int rng_id;       // Philox stream of this utterance: its index in a plain batch, the request index under continuous batching
};

struct SampleParams {
    int K, V;                       // codebooks, audio classes
    int rpu;                        // rows per utterance (2 with CFG)
    int empty_token, eog, eos, sos, mts, max_n_spans;
    int top_k; float top_p; float temperature;
    int stop_repetition; int n_silence; int silence[SSRB_MAX_SILENCE];
    float cfg_coef; int cfg_stride;
    unsigned long long seed;
    int max_steps;
    int n_utt;
    int utt0;         // first utterance handled by this launch (blockIdx.x == 0); 0 for the decode loop
    int count_iter;   // 1: this launch is one iteration of the loop (bumps the iteration counter)
};

// packed prompt position descriptor for the prefill embedding (host-built)
struct PosDesc { int text_tok; int a0, a1, a2, a3; int pe_idx; };   // text_tok >= 0 -> text position

int launch_embed_prefill(const PosDesc* desc, int M, int D, const float* text_emb, const float* audio_emb, int V,
                         const float* pe, float alpha_t, float alpha_a, float* x, cudaStream_t s);
// zero_word (optional): a device word the kernel clears — the grid-barrier counter of the persistent small-batch kernel
int launch_embed_step(const int* next_tok, const UttState* st, int R, int rpu, int K, int D, const float* audio_emb,
                      int V, const float* pe, float alpha_a, float* x, cudaStream_t s, unsigned int* zero_word = nullptr);
// decode step with LayerNorm folded into the GEMMs: also writes bf16(x) and the per-128-column {mean, M2} partials
int launch_embed_step_fold(const int* next_tok, const UttState* st, int R, int rpu, int K, int D, const float* audio_emb,
                           int V, const float* pe, float alpha_a, float* x, void* xb, float2* part, int part_ld,
                           cudaStream_t s);
// Wf = bf16(gamma . W) (row-wise), colsum[n] = sum_k Wf[n,k], biasf[n] = bias[n] + sum_k beta_k W[n,k]
int launch_fold_ln(const void* W, int N, int Kd, const float* gamma, const float* beta, const float* bias, void* Wf,
                   float* colsum, float* biasf, cudaStream_t s);
// y[m] = LN(x[idx ? idx[m] : m]); out dtype SSRB_DTYPE_*
int launch_layernorm(const float* x, const int* idx, int M, int D, const float* w, const float* b, void* out,
                     int out_dtype, cudaStream_t s);
// scatter K/V of qkv[M,3D] into cache[(kv, r, h, slot, d)]; rows/slots null => r = m, slot = seq_len[m]
int launch_kv_append(const float* qkv, int M, int D, int H, const int* rows, const int* slots, const int* seq_len,
                     void* kcache, void* vcache, int cache_dtype, int Smax, cudaStream_t s);
// single-query attention against the cache (decode); out [R, D] in act dtype
int launch_attn_decode(const float* qkv, int R, int D, int H, const void* kcache, const void* vcache, int cache_dtype,
                       int Smax, const int* seq_len, const UttState* st, int rpu, float* ws, int* tickets,
                       void* out, int out_dtype, int prefetch, cudaStream_t s);
size_t attn_decode_ws_floats(int R, int H, int Smax);
// bf16 production path (attn_decode_tma.cu): bulk-async staged K/V tiles; also appends this step's K/V row in place
int launch_attn_decode_tma(const float* qkv, int R, int D, int H, void* kcache, void* vcache, int Smax, const int* seq_len,
                           const UttState* st, int rpu, float* ws, int* tickets, void* out, int prefetch, cudaStream_t s);
int attn_decode_tma_max_nsplit(int Smax);
int attn_decode_nsplit(int Smax);
// causal attention over packed prompt rows (prefill); K/V read from the cache
int launch_attn_prefill(const float* qkv, int D, int H, const void* kcache, const void* vcache, int cache_dtype,
                        int Smax, int n_rows, const int* row_ids, const int* row_start, const int* row_len,
                        int max_len, void* out, int out_dtype, cudaStream_t s);
// bf16 tensor-core prefill attention (attn_prefill_mma.cu)
int launch_attn_prefill_mma(const float* qkv, int D, int H, const void* kcache, const void* vcache, int Smax, int n_rows,
                            const int* row_ids, const int* row_start, const int* row_len, int max_len, void* out, cudaStream_t s);
// CFG + logit rules + top-k/top-p + sample + state machine (models/ssr.py:690-754)
// training forward / loss: masked per-codebook cross entropy + top-10 accuracy of teacher-forced logits (lm_kernels.cu)
int launch_masked_ce(const float* logits, const int* audio, const unsigned char* flags, int Ty, int K, int V, float* nll,
                     unsigned char* hit, double* out, cudaStream_t s);
int launch_sample(const float* logits, UttState* st, int* seq_len, int* next_tok, int* gen_tok, const float* noise,
                  int* iter_counter, const SampleParams& p, cudaStream_t s, int only_utt = -1);

// ---- persistent whole-iteration decode kernel for R <= 16 rows (lm_mega.cu) ------------------------------------------------
struct MegaLayerHost {            // one decoder layer: PACKED bf16 weights (mega_pack), fp32 biases / LayerNorm parameters, cache
    int D, F;
    const void *wqkv, *wo, *w1, *w2;
    const float *bqkv, *bo, *b1, *b2, *ln1g, *ln1b, *ln2g, *ln2b;
    void *kc, *vc;
};
struct MegaArgs {
    int R, D, H, F, L, NCB, V, Hh, Smax, rpu, max_pieces;
    float* x; float* qkv; void* ao; void* hid; void* hh; float* logits;
    const int* seq_len; const UttState* st; float* attn_ws; int* tickets; unsigned int* bar;
    const void* layers_dev;       // device array of L records built with mega_fill_layer
    const void *h1_w, *h2_w; const float *h1_b, *h2_b, *lnf_g, *lnf_b;
};
bool mega_supported(int R, int D, int H, int F, int NCB, int V, int Hh);
enum { MEGA_QKV = 0, MEGA_OUT = 1, MEGA_FFN1 = 2, MEGA_FFN2 = 3, MEGA_H1 = 4, MEGA_H2 = 5 };
int mega_pack(const void* W, void* out, int N, int K, int kind, cudaStream_t s);     // W [N, K] bf16 -> streaming order
size_t mega_layer_bytes();
int mega_fill_layer(void* host_slot, const MegaLayerHost& h);
int launch_mega(const MegaArgs& a, cudaStream_t s);
int mega_trace_arm(unsigned long long* dev_buf, int cap_per_cta);   // debug: per-CTA %globaltimer stamps (null disarms)
int mega_grid(int* G_out);

}  // namespace ssrb
