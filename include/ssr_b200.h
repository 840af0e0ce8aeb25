/* ssr_b200.h — C ABI of libssr_b200.so: the B200 (sm_100a) implementation of SSR-Speech's inference
 * hot path.  Plain C types only; no torch types, no exceptions, no ownership transfer of caller memory.
 *
 * The reference (WangHelin1997/SSR-Speech) is 100 % Python/PyTorch and has NO FFI/plugin interface
 * (SURVEY.md §0 #10, §8b); the drop-in boundary is therefore Python-level and this library is bound by
 * the thin ctypes layer in ssr-speech_b200/_lib.py.  Each entry point cites the reference interface it
 * replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; ssrb_last_error() returns the message
 *     of the last failure on the calling thread's context (static storage, do not free).
 *   - "dev" pointers are device pointers on the context's device, "host" pointers are host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is enqueued
 *     asynchronously on it unless the function is documented as synchronising.
 *   - a context is bound to one device and is not thread-safe.
 */
#ifndef SSR_B200_H
#define SSR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSRB_DTYPE_F32  0
#define SSRB_DTYPE_BF16 1

#define SSRB_MAX_SPANS     3
#define SSRB_MAX_SILENCE   8

const char* ssrb_last_error(void);
int ssrb_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches claim) */
uint64_t ssrb_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Autoregressive LM  —  replaces models/ssr.py::SSR_Speech (inference path) and the modules under it
 * (models/modules/transformer.py, activation.py, embedding.py).
 * ---------------------------------------------------------------------------------------------- */
typedef struct ssrb_lm ssrb_lm;

typedef struct {
    /* model (the fields SSR_Speech.__init__ reads from ckpt["config"], models/ssr.py:113-179) */
    int d_model, n_head, n_layer, ffn_dim;
    int n_codebooks, n_audio_tokens, n_text_tokens, head_hidden;
    int empty_token, eog, eos, sos, mts, max_n_spans;
    /* engine capacity */
    int max_rows;            /* transformer rows resident at once = utterances x (2 if CFG else 1)   */
    int max_seq;             /* max cached positions per row (text + audio)                           */
    int max_prefill_tokens;  /* packed prompt positions processed per prefill chunk                   */
    int max_steps;           /* max decode iterations recorded per utterance                          */
    int weight_dtype;        /* SSRB_DTYPE_F32: fp32 weights/KV/activations (parity mode)
                                SSRB_DTYPE_BF16: bf16 weights/KV/GEMM operands, fp32 accumulate (production) */
    int gemm_impl;           /* 0 = auto, 1 = SIMT kernels only, 2 = tcgen05 kernels (bf16 only)      */
} ssrb_lm_config;

/* SSR_Speech(config) + .to(device)                                        — models/ssr.py:104-179 */
int ssrb_lm_create(const ssrb_lm_config* cfg, int device, ssrb_lm** out);
void ssrb_lm_destroy(ssrb_lm* lm);

/* load_state_dict(ckpt["model"]) — one call per state_dict entry, `name` is the reference key
 * (inference_v2.py:198-202).  `host` is fp32, row-major, `shape[ndim]`.
 * The extra key "pe_table" [n_pos, d_model] carries the sinusoid table of
 * models/modules/embedding.py:69-92 (computed by the caller so it matches the reference bit for bit).
 * Synchronises. */
int ssrb_lm_load_tensor(ssrb_lm* lm, const char* name, const float* host, const int64_t* shape, int ndim);
/* returns 0 when every tensor the engine needs has been loaded, else 1 (message lists missing keys) */
int ssrb_lm_check_loaded(ssrb_lm* lm);

/* sampling / CFG arguments of SSR_Speech.inference                       — models/ssr.py:504-524 */
typedef struct {
    int   top_k;             /* <=0: disabled                                  (ssr.py:38-45)          */
    float top_p;             /* >=1: disabled                                  (ssr.py:47-67)          */
    float temperature;       /*                                                (ssr.py:80-81)          */
    int   stop_repetition;   /* <=0: disabled                                  (ssr.py:725-730)        */
    int   n_silence;
    int   silence_tokens[SSRB_MAX_SILENCE];
    float cfg_coef;          /*                                                (ssr.py:691-696)        */
    int   cfg_stride;
    int   aug_text;          /* 1: every utterance owns 2 rows (cond, uncond)  (ssr.py:571-577)        */
    uint64_t seed;           /* counter-based RNG seed when `noise` is NULL                            */
} ssrb_sampling;

/* one batch of utterances, host memory, all int32 */
typedef struct {
    int n_utt;
    const int32_t* text;        /* [n_rows, text_stride] phoneme ids per ROW (row = utt*rpu + j)      */
    int            text_stride;
    const int32_t* text_len;    /* [n_utt]  (cond and uncond rows share the length, ssr.py:574)       */
    const int32_t* prompt;      /* [n_utt, n_codebooks, prompt_stride] audio tokens before the first
                                   generation slot = seq.prepare().prompt_tokens (ssr.py:604-626)     */
    int            prompt_stride;
    const int32_t* prompt_len;  /* [n_utt]                                                            */
    const int32_t* n_spans;     /* [n_utt] number of masked spans to generate, 1..max_n_spans         */
} ssrb_lm_batch;

/* Prologue + first dec_forward of SSR_Speech.inference (ssr.py:596-689): embeds text/audio prompt,
 * runs the prompt through the decoder (filling the in-place KV cache), samples iteration 1.
 * `noise_dev`: optional device fp32 [max_steps, n_utt, n_codebooks, n_audio_tokens] Exp(1) variates
 * consumed as sample = argmax(p / noise) — the identity torch.multinomial(num_samples=1) uses — so a
 * caller can reproduce the reference's sampled tokens exactly; NULL = in-kernel Philox.  Asynchronous. */
int ssrb_lm_begin(ssrb_lm* lm, const ssrb_lm_batch* batch, const ssrb_sampling* sp,
                  const float* noise_dev, void* stream);

/* Runs up to `n_steps` iterations of the `while True` loop (ssr.py:671-771) for all unfinished
 * utterances without host synchronisation (CUDA-graph replay).  Asynchronous. */
int ssrb_lm_decode(ssrb_lm* lm, int n_steps, void* stream);

/* Synchronises `stream`, returns how many utterances have finished all their spans and the number of
 * loop iterations executed so far (1 = only the prefill sample). */
int ssrb_lm_poll(ssrb_lm* lm, void* stream, int* n_done, int* n_iter);

/* Continuous batching (the reference decodes one utterance at a time, inference_v2.py:331-333): replaces the
 * FINISHED utterance in slot `utt` of the open batch by a new one — prologue + first dec_forward + first sample of
 * SSR_Speech.inference (ssr.py:596-689) for that utterance only, into the slot's KV rows; the other slots keep decoding
 * from where they are.  text [rows_per_utt, text_len] (cond row, then the uncond row when aug_text), prompt
 * [n_codebooks, prompt_len] as in ssrb_lm_batch.  `rng_stream` selects the utterance's Philox stream (its index in a plain
 * batch), so its samples do not depend on the slot or the moment it was admitted.  Capacity (max_seq, max_steps,
 * max_prefill_tokens) is that of the engine; batches begun with injected noise cannot admit.  Synchronises once (reads
 * the slot state), the prefill itself is asynchronous. */
int ssrb_lm_admit(ssrb_lm* lm, int utt, const int32_t* text, int text_len, const int32_t* prompt, int prompt_len,
                  int n_spans, int rng_stream, void* stream);
/* Like ssrb_lm_poll, per slot: done_flags[n_utt] (1 = all spans finished).  Synchronises. */
int ssrb_lm_poll_flags(ssrb_lm* lm, void* stream, int32_t* done_flags, int* n_iter);

/* Sampled tokens of utterance `utt`: out[n, n_codebooks] int32 for n = 0..*n_tokens-1 (every iteration
 * of every span, including each span's EOG tail, concatenated); span_len[max_n_spans] = iterations per
 * span.  Capacity `cap` iterations.  Synchronises. */
int ssrb_lm_read_tokens(ssrb_lm* lm, void* stream, int utt, int32_t* out, int cap, int* n_tokens, int32_t* span_len);

/* Which kernels the decode iterations of the open batch run through (test / bench introspection): 0 = per-GEMM chain with separate
 * LayerNorm kernels (fp32 parity mode, SIMT), 1 = per-GEMM chain with folded LayerNorm (bf16, rows <= 128), 2 = experimental
 * per-layer kernel, 3 = the persistent whole-iteration kernel for <= 16 rows (csrc/lm_mega.cu); -1 = no open batch. */
int ssrb_lm_decode_path(ssrb_lm* lm);

/* Test hook: raw head outputs of the most recent iteration, fp32 [n_rows, n_codebooks, n_audio_tokens]
 * (the tensor `logits` of ssr.py:688 before CFG / rules).  Synchronises. */
int ssrb_lm_read_logits(ssrb_lm* lm, void* stream, float* host_out);

/* Test hook (teacher forcing; incremental == full-forward check of SURVEY §4): text[Lx], audio tokens
 * [n_codebooks, Ty]; writes fp32 logits [Ty, n_codebooks, n_audio_tokens] for every audio position,
 * computed by the prefill path.  Synchronises. */
int ssrb_lm_teacher_forced(ssrb_lm* lm, const int32_t* text, int Lx, const int32_t* audio, int Ty,
                           float* host_logits, void* stream);

/* Training forward / loss of ONE utterance at its own length (replaces SSR_Speech.forward, models/ssr.py:280-379, in eval mode;
 * the reference masks padded key positions, so a padded batch is the sum over its utterances): text[Lx], audio tokens
 * [n_codebooks, Ty] as the dataset prepared them (mask tokens, eog, empty-token delay pattern), flags [n_codebooks, Ty-1] for the
 * targets audio[:, 1:] — bit 0: the position enters the cross entropy / top-10 accuracy (tmp_masks, ssr.py:339-345), bit 1: it
 * counts as a token (masks, ssr.py:333-337).  out [n_codebooks][4] (host, double): sum of -log p(target), positions in the loss,
 * top-10 hits, token count.  Forward through the prefill path, fused masked cross entropy on the device.  Synchronises. */
int ssrb_lm_forward_loss(ssrb_lm* lm, const int32_t* text, int Lx, const int32_t* audio, int Ty, const uint8_t* flags,
                         double* out, void* stream);

/* algorithmic HBM bytes of one decode iteration for the current batch state:
 * weights streamed once + KV read for every active row + KV written.  Synchronises. */
int ssrb_lm_step_bytes(ssrb_lm* lm, void* stream, double* weight_bytes, double* kv_bytes);

/* Profiling hook (bench.py roofline): runs `n_steps` decode iterations WITHOUT the CUDA graph, with CUDA
 * events (on `stream`) around every kernel class, and returns the average milliseconds per iteration spent in
 * ms_by_class[4] = {attention, GEMMs, layernorm/embed/kv-append, sampling} and in total.  The iterations are
 * real (they advance the batch state).  Synchronises. */
int ssrb_lm_profile_steps(ssrb_lm* lm, int n_steps, void* stream, double* ms_by_class, double* total_ms);

/* ------------------------------------------------------------------------------------------------
 * WM-Encodec  —  replaces audiocraft/models/wmencodec.py::WMEncodecModel.{encode,decode,wmdecode}
 * and below it audiocraft/modules/{seanet,conv,lstm}.py, audiocraft/quantization/{vq,core_vq}.py.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ssrb_codec ssrb_codec;

typedef struct {
    int channels, dimension, n_filters;
    int n_ratios; int ratios[8];          /* decoder order, e.g. 8,5,4,2 (encoder reverses them)       */
    int kernel_size, residual_kernel_size, last_kernel_size, compress, lstm_layers;
    int n_q, bins;
    int max_batch_chunk;                  /* utterances processed per internal pass (bounds workspace) */
    int tensor_cores;                     /* 0: fp32 CUDA-core kernels everywhere (parity mode; encode is always fp32 because the
                                             RVQ indices must be reproducible); 1: decode / wmdecode run their SEANet
                                             convolutions as bf16 tcgen05 GEMMs over channels-last activations           */
} ssrb_codec_config;

int ssrb_codec_create(const ssrb_codec_config* cfg, int device, ssrb_codec** out);
void ssrb_codec_destroy(ssrb_codec* c);

/* load_state_dict(state['best_state']['model']) — `name` is the reference key (SURVEY Appendix D;
 * audiocraft/solvers/wmcompression.py:302-312).  weight_g/weight_v pairs are folded (w = g*v/||v||)
 * when both halves have arrived.  Synchronises. */
int ssrb_codec_load_tensor(ssrb_codec* c, const char* name, const float* host, const int64_t* shape, int ndim);
int ssrb_codec_check_loaded(ssrb_codec* c);

/* WMEncodecModel.encode (wmencodec.py:324-339): wav_dev fp32 [B,1,T] (T multiple of hop) ->
 * codes_dev int64 [B,n_q,T/hop], emb_dev fp32 [B,dimension,T/hop] (may be NULL). */
int ssrb_codec_encode(ssrb_codec* c, const float* wav_dev, int B, int T, int64_t* codes_dev, float* emb_dev, void* stream);
/* RVQ only: EuclideanCodebook.quantize on given latents (core_vq.py:164-172, :382-392) */
int ssrb_codec_quantize(ssrb_codec* c, const float* emb_dev, int B, int Tf, int64_t* codes_dev, void* stream);
/* WMEncodecModel.decode (wmencodec.py:341-356): codes int64 [B,n_q,Tf] -> wav fp32 [B,1,Tf*hop] */
int ssrb_codec_decode(ssrb_codec* c, const int64_t* codes_dev, int B, int Tf, float* wav_dev, void* stream);
/* WMEncodecModel.wmdecode (wmencodec.py:358-375; WMSEANetDecoder.forward seanet.py:555-600):
 * marks int64 [B,Tf], wav_in fp32 [B,1,Tf*hop] -> wav_out fp32 [B,1,Tf*hop];
 * mark_logits_dev fp32 [B,Tf,2] or NULL (the reference caller discards it, data/tokenizer.py:133). */
int ssrb_codec_wmdecode(ssrb_codec* c, const int64_t* codes_dev, const int64_t* marks_dev, const float* wav_in_dev,
                        int B, int Tf, float* wav_out_dev, float* mark_logits_dev, void* stream);

/* WMEncodecModel.detect_watermark (wmencodec.py:377-382): wav fp32 [B,1,T] -> per-frame watermark logits fp32 [B,T/hop,2]
 * = wm_predictor(wm_encoder(wav)).  (The reference then takes argmax over the TIME axis, a bug — SURVEY §0; the Python
 * mirror exposes both behaviours.) */
int ssrb_codec_detect_watermark(ssrb_codec* c, const float* wav_dev, int B, int T, float* mark_logits_dev, void* stream);

/* Debug: arms (dev_buf != NULL) or disarms the in-kernel timeline of the decode-chain kernels.  dev_buf holds cap x 4 u64
 * records {kernel id, CTA id, globaltimer ns at entry, at exit}; *dev_idx counts records (tools/timeline.py). */
int ssrb_debug_timeline(unsigned long long* dev_buf, unsigned int* dev_idx, unsigned int cap);

/* Debug: arms (dev_buf != NULL) or disarms the per-CTA trace of the persistent small-batch decode kernel (csrc/lm_mega.cu):
 * consumer thread 0 of CTA c writes %globaltimer (ns) stamps into dev_buf[c * cap_per_cta ...] at kernel entry, dependency
 * resolved, then per phase {start (grid barrier passed), activations staged, chunks consumed (GEMV phases only), end}.
 * *n_ctas receives the kernel's grid size (tools/mega_trace.py). */
int ssrb_debug_mega_trace(unsigned long long* dev_buf, int cap_per_cta, int* n_ctas);

/* ------------------------------------------------------------------------------------------------
 * Stand-alone op hooks used by the unit tests (each runs exactly the kernel the engines use).
 * ---------------------------------------------------------------------------------------------- */
/* C[M,N] = act(A[M,K] . W[N,K]^T + bias) (+ residual); dtype of A/W = `dtype`, C fp32.
 * impl: 1 = SIMT, 2 = tcgen05 (bf16 only).  act: 0 none, 1 relu, 2 gelu(erf). */
int ssrb_op_gemm(const void* A_dev, const void* W_dev, const float* bias_dev, const float* residual_dev,
                 float* C_dev, int M, int N, int K, int dtype, int act, int impl, void* stream);

/* The folded-LayerNorm GEMM pair exactly as the bf16 decode chain runs it (stands in for transformer.py:58-75 LayerNorm
 * followed by the F.linear of activation.py:86 / transformer.py:386 / ssr.py:688):
 *   stage 1  X = A1 . W1^T + bias1 + residual            (fp32 X_out [M,D]; the kernel also emits bf16(X) and the
 *                                                          per-128-column {mean, M2} partials of every row)
 *   stage 2  C = act(LayerNorm(X; gamma, beta, 1e-5) . W2^T + bias2)     LayerNorm applied in the GEMM epilogue
 * A1 [M,K1], W1 [D,K1], W2 [N,D] bf16; M <= 128, D % 128 == 0, D <= 2048. */
int ssrb_op_gemm_ln(const void* A1_dev, const void* W1_dev, const float* bias1_dev, const float* residual_dev,
                    float* X_out_dev, int M, int D, int K1, const void* W2_dev, const float* gamma_dev,
                    const float* beta_dev, const float* bias2_dev, float* C_out_dev, int N, int act, void* stream);

/* EXPERIMENTAL.  The GEMM chain of one decoder layer between two decode-attention kernels (out-proj + residual, LayerNorm 2,
 * FFN1 + ReLU, FFN2 + residual, and optionally the NEXT layer's LayerNorm 1 + QKV projection when wqkv != NULL;
 * transformer.py:321-343,386-388, activation.py:83-89,637) on caller-provided device buffers:
 *   impl 0 = the four per-GEMM launches the decode chain uses today, impl 1 = the persistent per-layer kernel
 *   (csrc/gemm_layer.cu, SSRB_LAYER_KERNEL=1 in the engine).  Both must produce the same bits.
 * ao [M,D] bf16, x_inout [M,D] fp32 (residual stream, updated), weights bf16 [out,in], hid_out [M,F] bf16, qkv_out [M,3D] fp32.
 * impl 1 needs D % 512 == 0 and F % 512 == 0. */
int ssrb_op_layer_chain(const void* ao_dev, float* x_inout_dev, const void* wo_dev, const float* bo_dev, const void* w1_dev,
                        const float* b1_dev, const float* gamma2_dev, const float* beta2_dev, const void* w2_dev,
                        const float* b2_dev, const void* wqkv_dev, const float* bqkv_dev, const float* gamma1n_dev,
                        const float* beta1n_dev, void* hid_out_dev, float* qkv_out_dev, int M, int D, int F, int impl,
                        void* stream);

/* Single-query attention against the in-place bf16 cache exactly as the decode chain runs it (replaces
 * F.scaled_dot_product_attention at models/modules/activation.py:634 for tgt_len == 1 plus the cache re-materialisation of
 * activation.py:626-631): qkv_dev fp32 [R, 3D] (q | k | v of this step), caches bf16 [R][H][Smax][128] holding seq_len[r]
 * keys per row; the kernel appends this step's K/V row at slot seq_len[r] and writes out_dev bf16 [R, D].  seq_len_host /
 * done_host (may be NULL) are HOST arrays [R]; rows with done != 0 are skipped.  Synchronises. */
int ssrb_op_attn_decode(const float* qkv_dev, void* kcache_dev, void* vcache_dev, const int32_t* seq_len_host,
                        const int32_t* done_host, void* out_dev, int R, int D, int H, int Smax, void* stream);

/* Causal prefill attention over packed rows exactly as the bf16 prefill runs it (activation.py:634 for the first
 * dec_forward, models/ssr.py:673-684): qkv_dev fp32 [sum(row_len), 3D]; K/V are first scattered into the caches
 * (row i -> cache row i, slot = position), then every position attends to positions <= its own.  out_dev bf16 [sum, D]. */
int ssrb_op_attn_prefill(const float* qkv_dev, void* kcache_dev, void* vcache_dev, const int32_t* row_len_host, int n_rows,
                         void* out_dev, int D, int H, int Smax, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSR_B200_H */
