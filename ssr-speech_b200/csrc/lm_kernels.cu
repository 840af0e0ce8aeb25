// lm_kernels.cu — non-GEMM kernels of the SSR-Speech decoder: embedding + sinusoidal PE, LayerNorm,
// in-place KV append, single-query (decode) and causal (prefill) attention, and the fused
// CFG + logit-rule + top-k/top-p + sampling + per-utterance state machine kernel.
// Reference call sites are cited on each kernel (paths relative to the reference root).
#include "lm_kernels.cuh"

namespace ssrb {

// =================================================================================================
// Embedding + PE   (models/modules/embedding.py:22-48,94-98; models/ssr.py:191-198,596-600,756-763)
//   out = sum_k E_k[tok_k] * 1.0 + alpha * pe[pos]   — the reference rounds alpha*pe before the add.
// =================================================================================================
__global__ void __launch_bounds__(256) embed_prefill_kernel(const PosDesc* __restrict__ desc, int D,
                                                            const float* __restrict__ text_emb,
                                                            const float* __restrict__ audio_emb, int V,
                                                            const float* __restrict__ pe, float alpha_t, float alpha_a,
                                                            float* __restrict__ x) {
    pdl_launch_dependents();
    pdl_wait();
    const int m = blockIdx.x;
    const PosDesc pd = desc[m];
    const float* per = pe + (int64_t)pd.pe_idx * D;
    float* xo = x + (int64_t)m * D;
    if (pd.text_tok >= 0) {
        const float* e = text_emb + (int64_t)pd.text_tok * D;
        for (int d = threadIdx.x; d < D; d += blockDim.x) xo[d] = __fadd_rn(e[d], __fmul_rn(alpha_t, per[d]));
    } else {
        const int64_t tbl = (int64_t)V * D;
        const float* e0 = audio_emb + (int64_t)pd.a0 * D;
        const float* e1 = audio_emb + tbl + (int64_t)pd.a1 * D;
        const float* e2 = audio_emb + 2 * tbl + (int64_t)pd.a2 * D;
        const float* e3 = audio_emb + 3 * tbl + (int64_t)pd.a3 * D;
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            float e = __fadd_rn(__fadd_rn(__fadd_rn(e0[d], e1[d]), e2[d]), e3[d]);
            xo[d] = __fadd_rn(e, __fmul_rn(alpha_a, per[d]));
        }
    }
}

__global__ void __launch_bounds__(256) embed_step_kernel(const int* __restrict__ next_tok, const UttState* __restrict__ st,
                                                         int rpu, int K, int D, const float* __restrict__ audio_emb, int V,
                                                         const float* __restrict__ pe, float alpha_a, float* __restrict__ x,
                                                         unsigned int* __restrict__ zero_word) {
    const int ts = ts_begin(TSK_EMBED);
    pdl_wait();
    if (zero_word && blockIdx.x == 0 && threadIdx.x == 0) *zero_word = 0u;   // grid-barrier counter of decode_mega_kernel (lm_mega.cu)
    // first kernel of a decode iteration: dependents are released only once the previous iteration (sampler included) has
    // completed, so kernels further down the chain may read the row state / cached K/V before their own griddepcontrol.wait
    pdl_launch_dependents();
    ts_dep(ts);
    const int r = blockIdx.x, u = r / rpu;
    const int* tk = next_tok + u * K;
    const int pos = st[u].y_len;
    const int64_t tbl = (int64_t)V * D;
    const float* e0 = audio_emb + (int64_t)tk[0] * D;
    const float* e1 = audio_emb + tbl + (int64_t)tk[1] * D;
    const float* e2 = audio_emb + 2 * tbl + (int64_t)tk[2] * D;
    const float* e3 = audio_emb + 3 * tbl + (int64_t)tk[3] * D;
    const float* per = pe + (int64_t)pos * D;
    float* xo = x + (int64_t)r * D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float e = __fadd_rn(__fadd_rn(__fadd_rn(e0[d], e1[d]), e2[d]), e3[d]);
        xo[d] = __fadd_rn(e, __fmul_rn(alpha_a, per[d]));
    }
    ts_end(ts);
}

// Decode step with LayerNorm folded into the consuming GEMMs (gemm_tc.cu): besides the fp32 residual stream the kernel
// writes bf16(x) — the GEMM operand — and the {mean, M2} partial of every 128-column block of the row.  Same per-element
// arithmetic as embed_step_kernel; each thread owns 8 consecutive columns, 16 lanes own one block.
__global__ void __launch_bounds__(256) embed_step_fold_kernel(const int* __restrict__ next_tok, const UttState* __restrict__ st,
                                                              int rpu, int K, int D, const float* __restrict__ audio_emb, int V,
                                                              const float* __restrict__ pe, float alpha_a, float* __restrict__ x,
                                                              bf16* __restrict__ xb, float2* __restrict__ part, int part_ld) {
    const int ts = ts_begin(TSK_EMBED);
    pdl_wait();
    pdl_launch_dependents();                      // see embed_step_kernel: released only after the previous iteration completed
    ts_dep(ts);
    const int r = blockIdx.x, u = r / rpu;
    const int* tk = next_tok + u * K;
    const int pos = st[u].y_len;
    const int64_t tbl = (int64_t)V * D;
    const int d0 = threadIdx.x * 8;
    const bool valid = d0 < D;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = 0.f;
    if (valid) {
        float e0[8], e1[8], e2[8], e3[8], pr[8];
        load8(audio_emb + (int64_t)tk[0] * D + d0, e0);
        load8(audio_emb + tbl + (int64_t)tk[1] * D + d0, e1);
        load8(audio_emb + 2 * tbl + (int64_t)tk[2] * D + d0, e2);
        load8(audio_emb + 3 * tbl + (int64_t)tk[3] * D + d0, e3);
        load8(pe + (int64_t)pos * D + d0, pr);
#pragma unroll
        for (int i = 0; i < 8; i++)
            v[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(e0[i], e1[i]), e2[i]), e3[i]), __fmul_rn(alpha_a, pr[i]));
        store8(x + (int64_t)r * D + d0, v);
        store8(xb + (int64_t)r * D + d0, v);
    }
    float s = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / 128.f);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) { const float d = v[i] - mean; q += d * d; }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (valid && (threadIdx.x & 15) == 0) part[(int64_t)(threadIdx.x >> 4) * part_ld + r] = make_float2(mean, q);
    ts_end(ts);
}

// one-time weight preparation for the folded LayerNorm: Wf[n,k] = bf16(gamma_k * W[n,k]), colsum[n] = sum_k Wf[n,k],
// biasf[n] = bias[n] + sum_k beta_k * W[n,k]   (one warp per output feature)
__global__ void __launch_bounds__(256) fold_ln_kernel(const bf16* __restrict__ W, int N, int Kd, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, const float* __restrict__ bias,
                                                      bf16* __restrict__ Wf, float* __restrict__ colsum, float* __restrict__ biasf) {
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    float cs = 0.f, bs = 0.f;
    for (int k = lane; k < Kd; k += 32) {
        const float w = __bfloat162float(W[(int64_t)n * Kd + k]);
        const bf16 wf = __float2bfloat16_rn(w * gamma[k]);
        Wf[(int64_t)n * Kd + k] = wf;
        cs += __bfloat162float(wf);
        bs += beta[k] * w;
    }
    cs = warp_sum(cs); bs = warp_sum(bs);
    if (lane == 0) { colsum[n] = cs; biasf[n] = (bias ? bias[n] : 0.f) + bs; }
}

int launch_fold_ln(const void* W, int N, int Kd, const float* gamma, const float* beta, const float* bias, void* Wf,
                   float* colsum, float* biasf, cudaStream_t s) {
    SSRB_LAUNCH(fold_ln_kernel, cdiv(N, 8), 256, 0, s, reinterpret_cast<const bf16*>(W), N, Kd, gamma, beta, bias,
                reinterpret_cast<bf16*>(Wf), colsum, biasf);
    return 0;
}

int launch_embed_step_fold(const int* next_tok, const UttState* st, int R, int rpu, int K, int D, const float* audio_emb,
                           int V, const float* pe, float alpha_a, float* x, void* xb, float2* part, int part_ld,
                           cudaStream_t s) {
    SSRB_CHECK(K == 4 && D % 128 == 0 && D <= 2048, "embed_step_fold: K must be 4, d_model a multiple of 128 <= 2048");
    const int threads = ((D / 8) + 31) / 32 * 32;
    SSRB_LAUNCH_PDL(embed_step_fold_kernel, R, threads, 0, s, next_tok, st, rpu, K, D, audio_emb, V, pe, alpha_a, x,
                    reinterpret_cast<bf16*>(xb), part, part_ld);
    return 0;
}

int launch_embed_prefill(const PosDesc* desc, int M, int D, const float* text_emb, const float* audio_emb, int V,
                         const float* pe, float alpha_t, float alpha_a, float* x, cudaStream_t s) {
    if (M <= 0) return 0;
    SSRB_LAUNCH_PDL(embed_prefill_kernel, M, 256, 0, s, desc, D, text_emb, audio_emb, V, pe, alpha_t, alpha_a, x);
    return 0;
}
int launch_embed_step(const int* next_tok, const UttState* st, int R, int rpu, int K, int D, const float* audio_emb,
                      int V, const float* pe, float alpha_a, float* x, cudaStream_t s, unsigned int* zero_word) {
    SSRB_CHECK(K == 4, "embed_step: K must be 4");
    SSRB_LAUNCH_PDL(embed_step_kernel, R, 256, 0, s, next_tok, st, rpu, K, D, audio_emb, V, pe, alpha_a, x, zero_word);
    return 0;
}

// =================================================================================================
// LayerNorm (eps 1e-5, affine)            models/modules/transformer.py:58-75 (F.layer_norm)
// one CTA per row, two-pass statistics in fp32 registers
// =================================================================================================
template <typename TO>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const int* __restrict__ idx, int D,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        TO* __restrict__ out) {
    pdl_launch_dependents();
    const int ts = ts_begin(TSK_LN);
    pdl_wait();
    ts_dep(ts);
    __shared__ float red[8];
    __shared__ float stat[2];
    const int m = blockIdx.x;
    const float* xr = x + (int64_t)(idx ? idx[m] : m) * D;
    constexpr int MAXPT = 8;   // D <= 2048
    float v[MAXPT];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXPT; i++) {
        const int d = threadIdx.x + i * 256;
        v[i] = d < D ? xr[d] : 0.f;
        s += v[i];
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; i++) t += red[i];
        stat[0] = t / (float)D;
    }
    __syncthreads();
    const float mean = stat[0];
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXPT; i++) {
        const int d = threadIdx.x + i * 256;
        const float c = d < D ? v[i] - mean : 0.f;
        q += c * c;
    }
    q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; i++) t += red[i];
        stat[1] = rsqrtf(t / (float)D + 1e-5f);
    }
    __syncthreads();
    const float rstd = stat[1];
    TO* o = out + (int64_t)m * D;
#pragma unroll
    for (int i = 0; i < MAXPT; i++) {
        const int d = threadIdx.x + i * 256;
        if (d < D) o[d] = from_f32<TO>((v[i] - mean) * rstd * w[d] + b[d]);
    }
    ts_end(ts);
}

int launch_layernorm(const float* x, const int* idx, int M, int D, const float* w, const float* b, void* out,
                     int out_dtype, cudaStream_t s) {
    if (M <= 0) return 0;
    SSRB_CHECK(D <= 2048, "layernorm: d_model > 2048 not supported");
    if (out_dtype == SSRB_DTYPE_F32) SSRB_LAUNCH_PDL(layernorm_kernel<float>, M, 256, 0, s, x, idx, D, w, b, (float*)out);
    else SSRB_LAUNCH_PDL(layernorm_kernel<bf16>, M, 256, 0, s, x, idx, D, w, b, (bf16*)out);
    return 0;
}

// =================================================================================================
// KV append — in place.  Replaces the reference's per-step re-materialisation of the whole cache
// (models/modules/activation.py:626-631 torch.stack/cat; transformer.py:486; ssr.py:685-686).
// cache layout per layer: [row][head][slot][128]  (K and V separate)
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) kv_append_kernel(const float* __restrict__ qkv, int D, int H,
                                                        const int* __restrict__ rows, const int* __restrict__ slots,
                                                        const int* __restrict__ seq_len, T* __restrict__ kc,
                                                        T* __restrict__ vc, int Smax) {
    pdl_launch_dependents();
    pdl_wait();
    const int m = blockIdx.x;
    const int r = rows ? rows[m] : m;
    const int slot = slots ? slots[m] : seq_len[m];
    const float* src = qkv + (int64_t)m * 3 * D;
    for (int c = threadIdx.x * 8; c < 2 * D; c += 256 * 8) {
        const int kv = c >= D;
        const int cc = c - kv * D;
        const int h = cc >> 7, d = cc & 127;
        float v[8];
        load8(src + D + c, v);
        T* dst = (kv ? vc : kc) + (((int64_t)r * H + h) * Smax + slot) * 128 + d;
        store8(dst, v);
    }
}

int launch_kv_append(const float* qkv, int M, int D, int H, const int* rows, const int* slots, const int* seq_len,
                     void* kcache, void* vcache, int cache_dtype, int Smax, cudaStream_t s) {
    if (M <= 0) return 0;
    if (cache_dtype == SSRB_DTYPE_F32)
        SSRB_LAUNCH_PDL(kv_append_kernel<float>, M, 256, 0, s, qkv, D, H, rows, slots, seq_len, (float*)kcache, (float*)vcache, Smax);
    else
        SSRB_LAUNCH_PDL(kv_append_kernel<bf16>, M, 256, 0, s, qkv, D, H, rows, slots, seq_len, (bf16*)kcache, (bf16*)vcache, Smax);
    return 0;
}

// =================================================================================================
// Decode attention: one query per (row, head) against the in-place cache.
// Replaces F.scaled_dot_product_attention(q,k,v,attn_mask) at models/modules/activation.py:634 for
// tgt_len == 1 (the causal mask row is all-visible: models/ssr.py:227-237,263-269).
// grid (H, R, NSPLIT); flash-decoding split over keys, last-arriving CTA merges the partials.
// =================================================================================================
constexpr int ATT_CHUNK = 256;    // keys per split
int attn_decode_nsplit(int Smax) { return cdiv(Smax, ATT_CHUNK); }
size_t attn_decode_ws_floats(int R, int H, int Smax) { return (size_t)R * H * (size_t)attn_decode_tma_max_nsplit(Smax) * 130; }   // the finer of the two splits

template <typename T, typename TO>
__global__ void __launch_bounds__(128) attn_decode_kernel(const float* __restrict__ qkv, int D, int H,
                                                          const T* kc, const T* vc, int Smax,
                                                          const int* __restrict__ seq_len, const UttState* __restrict__ st,
                                                          int rpu, float* __restrict__ ws, int* __restrict__ tickets,
                                                          TO* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int h = blockIdx.x, r = blockIdx.y, z = blockIdx.z, nz = gridDim.z;
    if (st[r / rpu].done) return;
    const int n_keys = seq_len[r] + 1;
    const int nsplit = (n_keys + ATT_CHUNK - 1) / ATT_CHUNK;
    if (z >= nsplit) return;
    const int s0 = z * ATT_CHUNK, s1 = min(n_keys, s0 + ATT_CHUNK);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, half = lane >> 4, dl = (lane & 15) * 8;
    const float scale = 0.08838834764831845f;   // 1/sqrt(128)
    float q[8];
    load8(qkv + (int64_t)r * 3 * D + h * 128 + dl, q);
#pragma unroll
    for (int i = 0; i < 8; i++) q[i] *= scale;
    // fused in-place KV append: the CTA whose key range ends at the new position stores this step's K/V row
    // (the reference re-materialises the whole cache instead: activation.py:626-631, ssr.py:685-686)
    if (s1 == n_keys) {
        if (warp == 0) {
            float nv[8];
            load8(qkv + (int64_t)r * 3 * D + (1 + half) * D + h * 128 + dl, nv);
            T* dst = const_cast<T*>(half ? vc : kc) + (((int64_t)r * H + h) * Smax + (n_keys - 1)) * 128 + dl;
            store8(dst, nv);
        }
        __syncthreads();
    }
    const T* kb = kc + ((int64_t)r * H + h) * Smax * 128 + dl;
    const T* vb = vc + ((int64_t)r * H + h) * Smax * 128 + dl;
    float mrun = -INFINITY, lrun = 0.f, o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    constexpr int U = 4;
    for (int kq = s0 + warp * 2; kq < s1; kq += 8 * U) {   // warp-uniform trip count (full-mask shuffles inside)
        const int kp0 = kq + half;
        float kk[U][8], sc[U];
#pragma unroll
        for (int j = 0; j < U; j++) {
            const int kp = kp0 + j * 8;
            if (kp < s1) load8(kb + (int64_t)kp * 128, kk[j]);
        }
        float mnew = mrun;
#pragma unroll
        for (int j = 0; j < U; j++) {
            const int kp = kp0 + j * 8;
            float p = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) p = fmaf(q[i], kk[j][i], p);
            p += __shfl_xor_sync(0xffffffffu, p, 1);
            p += __shfl_xor_sync(0xffffffffu, p, 2);
            p += __shfl_xor_sync(0xffffffffu, p, 4);
            p += __shfl_xor_sync(0xffffffffu, p, 8);
            sc[j] = kp < s1 ? p : -INFINITY;
            mnew = fmaxf(mnew, sc[j]);
        }
        if (mnew > -INFINITY) {                      // this half-warp has seen at least one key
            const float corr = __expf(mrun - mnew);   // mrun = -inf on the first pass -> 0
            lrun *= corr;
#pragma unroll
            for (int i = 0; i < 8; i++) o[i] *= corr;
#pragma unroll
            for (int j = 0; j < U; j++) {
                const int kp = kp0 + j * 8;
                if (kp < s1) {
                    float vv[8];
                    load8(vb + (int64_t)kp * 128, vv);
                    const float p = __expf(sc[j] - mnew);
                    lrun += p;
#pragma unroll
                    for (int i = 0; i < 8; i++) o[i] = fmaf(p, vv[i], o[i]);
                }
            }
            mrun = mnew;
        }
    }
    // merge the 8 (warp, half) partial states of this CTA
    __shared__ float sm_m[8], sm_l[8], sm_o[8][128];
    __shared__ int sm_last;
    const int slot = warp * 2 + half;
    if ((lane & 15) == 0) { sm_m[slot] = mrun; sm_l[slot] = lrun; }
#pragma unroll
    for (int i = 0; i < 8; i++) sm_o[slot][dl + i] = o[i];
    __syncthreads();
    const int d = threadIdx.x;
    float M = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; i++) M = fmaxf(M, sm_m[i]);
    float L = 0.f, O = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float w = __expf(sm_m[i] - M);
        L += sm_l[i] * w;
        O += sm_o[i][d] * w;
    }
    TO* op = out + (int64_t)r * D + h * 128 + d;
    if (nsplit == 1) { *op = from_f32<TO>(O / L); return; }
    float* wsp = ws + ((int64_t)(r * H + h) * nz + z) * 130;
    wsp[2 + d] = O;
    if (d == 0) { wsp[0] = M; wsp[1] = L; }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(&tickets[r * H + h], 1);
        sm_last = (t == nsplit - 1);
        if (sm_last) tickets[r * H + h] = 0;
    }
    __syncthreads();
    if (!sm_last) return;
    __threadfence();
    const float* wb = ws + (int64_t)(r * H + h) * nz * 130;
    float M2 = -INFINITY;
    for (int i = 0; i < nsplit; i++) M2 = fmaxf(M2, __ldcg(wb + i * 130));
    float L2 = 0.f, O2 = 0.f;
    for (int i = 0; i < nsplit; i++) {
        const float w = __expf(__ldcg(wb + i * 130) - M2);
        L2 += __ldcg(wb + i * 130 + 1) * w;
        O2 += __ldcg(wb + i * 130 + 2 + d) * w;
    }
    *op = from_f32<TO>(O2 / L2);
}

int launch_attn_decode(const float* qkv, int R, int D, int H, const void* kcache, const void* vcache, int cache_dtype,
                       int Smax, const int* seq_len, const UttState* st, int rpu, float* ws, int* tickets,
                       void* out, int out_dtype, int prefetch, cudaStream_t s) {
    dim3 grid(H, R, attn_decode_nsplit(Smax));
    if (cache_dtype == SSRB_DTYPE_F32) {
        SSRB_CHECK(out_dtype == SSRB_DTYPE_F32, "attn_decode: fp32 cache implies fp32 activations");
        SSRB_LAUNCH_PDL((attn_decode_kernel<float, float>), grid, 128, 0, s, qkv, D, H, (const float*)kcache,
                    (const float*)vcache, Smax, seq_len, st, rpu, ws, tickets, (float*)out);
    } else {
        SSRB_CHECK(out_dtype == SSRB_DTYPE_BF16, "attn_decode: bf16 cache implies bf16 activations");
        static const bool simple = [] { const char* e = getenv("SSRB_ATTN_SIMPLE"); return e && e[0] == '1'; }();
        if (!simple)
            return launch_attn_decode_tma(qkv, R, D, H, const_cast<void*>(kcache), const_cast<void*>(vcache), Smax, seq_len, st,
                                          rpu, ws, tickets, out, prefetch, s);
        SSRB_LAUNCH_PDL((attn_decode_kernel<bf16, bf16>), grid, 128, 0, s, qkv, D, H, (const bf16*)kcache,
                    (const bf16*)vcache, Smax, seq_len, st, rpu, ws, tickets, (bf16*)out);
    }
    return 0;
}

// =================================================================================================
// Prefill attention: causal over the packed prompt [text ; audio] of each row
// (models/ssr.py:227-257 builds exactly triu(ones(S,S),1) — SURVEY §0; SDPA at activation.py:634).
// grid (ceil(max_len/32), H, n_rows); 128 threads = 32 queries x 4 head-dim quarters;
// K/V tiles of 32 keys staged in shared memory as fp32, online softmax in registers.
// =================================================================================================
constexpr int PF_Q = 32, PF_K = 32, PF_PITCH = 4 * 36;   // quarter q at float offset q*36 (bank-conflict free)

template <typename T, typename TO>
__global__ void __launch_bounds__(128) attn_prefill_kernel(const float* __restrict__ qkv, int D, int H,
                                                           const T* __restrict__ kc, const T* __restrict__ vc, int Smax,
                                                           const int* __restrict__ row_ids, const int* __restrict__ row_start,
                                                           const int* __restrict__ row_len, TO* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ __align__(16) float Ks[PF_K][PF_PITCH];
    __shared__ __align__(16) float Vs[PF_K][PF_PITCH];
    const int qt = blockIdx.x, h = blockIdx.y, ri = blockIdx.z;
    const int len = row_len[ri];
    if (qt * PF_Q >= len) return;
    const int r = row_ids[ri], base = row_start[ri];
    const int tid = threadIdx.x, ql = tid >> 2, dq = tid & 3;
    const int qpos = qt * PF_Q + ql;
    const bool qok = qpos < len;
    const float scale = 0.08838834764831845f;
    float q[32], o[32];
    {
        const float* qp = qkv + (int64_t)(base + (qok ? qpos : 0)) * 3 * D + h * 128 + dq * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            float4 t = *reinterpret_cast<const float4*>(qp + i);
            q[i] = t.x * scale; q[i + 1] = t.y * scale; q[i + 2] = t.z * scale; q[i + 3] = t.w * scale;
        }
#pragma unroll
        for (int i = 0; i < 32; i++) o[i] = 0.f;
    }
    float mrun = -INFINITY, lrun = 0.f;
    const int kend = min(len, (qt + 1) * PF_Q);
    const T* kb = kc + ((int64_t)r * H + h) * Smax * 128;
    const T* vb = vc + ((int64_t)r * H + h) * Smax * 128;
    for (int k0 = 0; k0 < kend; k0 += PF_K) {
        __syncthreads();
        // stage 32 keys x 128 dims of K and V: 512 chunks of 8 elements each, 4 per thread per tensor
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int c = tid + i * 128;
            const int key = c >> 4, d8 = (c & 15) * 8;
            float kv8[8], vv8[8];
            if (k0 + key < kend) {
                load8(kb + (int64_t)(k0 + key) * 128 + d8, kv8);
                load8(vb + (int64_t)(k0 + key) * 128 + d8, vv8);
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) { kv8[j] = 0.f; vv8[j] = 0.f; }
            }
            const int off = (d8 >> 5) * 36 + (d8 & 31);
            *reinterpret_cast<float4*>(&Ks[key][off]) = make_float4(kv8[0], kv8[1], kv8[2], kv8[3]);
            *reinterpret_cast<float4*>(&Ks[key][off + 4]) = make_float4(kv8[4], kv8[5], kv8[6], kv8[7]);
            *reinterpret_cast<float4*>(&Vs[key][off]) = make_float4(vv8[0], vv8[1], vv8[2], vv8[3]);
            *reinterpret_cast<float4*>(&Vs[key][off + 4]) = make_float4(vv8[4], vv8[5], vv8[6], vv8[7]);
        }
        __syncthreads();
        const int nk = min(PF_K, kend - k0);
        for (int j0 = 0; j0 < nk; j0 += 4) {
            float sc[4];
            float mnew = mrun;
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                const int j = j0 + jj;
                const float* kr = &Ks[j < PF_K ? j : 0][dq * 36];
                float p = 0.f;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(kr + i);
                    p = fmaf(q[i], t.x, p); p = fmaf(q[i + 1], t.y, p); p = fmaf(q[i + 2], t.z, p); p = fmaf(q[i + 3], t.w, p);
                }
                p += __shfl_xor_sync(0xffffffffu, p, 1);
                p += __shfl_xor_sync(0xffffffffu, p, 2);
                sc[jj] = (j < nk && (k0 + j) <= qpos) ? p : -INFINITY;
                mnew = fmaxf(mnew, sc[jj]);
            }
            if (mnew == -INFINITY) continue;   // nothing visible yet for this query (only when !qok)
            const float corr = __expf(mrun - mnew);
            lrun *= corr;
#pragma unroll
            for (int i = 0; i < 32; i++) o[i] *= corr;
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                const int j = j0 + jj;
                const float p = __expf(sc[jj] - mnew);   // 0 for masked
                lrun += p;
                const float* vr = &Vs[j < PF_K ? j : 0][dq * 36];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(vr + i);
                    o[i] = fmaf(p, t.x, o[i]); o[i + 1] = fmaf(p, t.y, o[i + 1]);
                    o[i + 2] = fmaf(p, t.z, o[i + 2]); o[i + 3] = fmaf(p, t.w, o[i + 3]);
                }
            }
            mrun = mnew;
        }
    }
    if (qok) {
        const float inv = 1.f / lrun;
        TO* op = out + (int64_t)(base + qpos) * D + h * 128 + dq * 32;
#pragma unroll
        for (int i = 0; i < 32; i++) op[i] = from_f32<TO>(o[i] * inv);
    }
}

int launch_attn_prefill(const float* qkv, int D, int H, const void* kcache, const void* vcache, int cache_dtype,
                        int Smax, int n_rows, const int* row_ids, const int* row_start, const int* row_len,
                        int max_len, void* out, int out_dtype, cudaStream_t s) {
    if (n_rows <= 0) return 0;
    dim3 grid(cdiv(max_len, PF_Q), H, n_rows);
    if (cache_dtype == SSRB_DTYPE_F32)
        SSRB_LAUNCH_PDL((attn_prefill_kernel<float, float>), grid, 128, 0, s, qkv, D, H, (const float*)kcache,
                    (const float*)vcache, Smax, row_ids, row_start, row_len, (float*)out);
    else {
        static const bool simt = [] { const char* e = getenv("SSRB_PREFILL_SIMT"); return e && e[0] == '1'; }();
        if (!simt) return launch_attn_prefill_mma(qkv, D, H, kcache, vcache, Smax, n_rows, row_ids, row_start, row_len, max_len, out, s);
        SSRB_LAUNCH_PDL((attn_prefill_kernel<bf16, bf16>), grid, 128, 0, s, qkv, D, H, (const bf16*)kcache,
                    (const bf16*)vcache, Smax, row_ids, row_start, row_len, (bf16*)out);
    }
    (void)out_dtype;
    return 0;
}

// =================================================================================================
// Sampling head: CFG mix (ssr.py:690-696), logit rules (:698-730), temperature / top-k / top-p
// (:26-86), sample = argmax(p / Exp(1)) (what torch.multinomial(n=1) evaluates), EOG bookkeeping and the
// per-span state machine (:709-754,646-660).  One CTA per (utterance, codebook), a cluster of 4 CTAs per utterance.
// No host synchronisation: all Python-side control flow of the reference loop lives in UttState.
// =================================================================================================
__device__ __forceinline__ uint32_t fkey(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ uint32_t philox_u32(unsigned long long seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

constexpr int SMP_T = 256;   // threads per CTA
constexpr int SMP_NPT = 9;   // values per thread: 256 threads x 9 >= 2056 classes
constexpr int SMP_W = SMP_T / 32;

struct BlockRed {
    float f[2][SMP_W][4];
    int i[2][SMP_W][4];
};

// block-wide sums of N independent values with ONE __syncthreads (double-buffered scratch, fixed summation order)
template <int N> __device__ __forceinline__ void block_sum(float (&v)[N], BlockRed& g, int& phase, int warp, int lane) {
#pragma unroll
    for (int n = 0; n < N; n++) v[n] = warp_sum(v[n]);
    if (lane == 0) {
#pragma unroll
        for (int n = 0; n < N; n++) g.f[phase][warp][n] = v[n];
    }
    __syncthreads();
#pragma unroll
    for (int n = 0; n < N; n++) {
        float r = 0.f;
#pragma unroll
        for (int w = 0; w < SMP_W; w++) r += g.f[phase][w][n];
        v[n] = r;
    }
    phase ^= 1;
}
template <int N> __device__ __forceinline__ void block_isum(int (&v)[N], BlockRed& g, int& phase, int warp, int lane) {
#pragma unroll
    for (int n = 0; n < N; n++) v[n] = __reduce_add_sync(0xffffffffu, v[n]);
    if (lane == 0) {
#pragma unroll
        for (int n = 0; n < N; n++) g.i[phase][warp][n] = v[n];
    }
    __syncthreads();
#pragma unroll
    for (int n = 0; n < N; n++) {
        int r = 0;
#pragma unroll
        for (int w = 0; w < SMP_W; w++) r += g.i[phase][w][n];
        v[n] = r;
    }
    phase ^= 1;
}
__device__ __forceinline__ float block_max(float v, BlockRed& g, int& phase, int warp, int lane) {
    v = warp_max(v);
    if (lane == 0) g.f[phase][warp][0] = v;
    __syncthreads();
    float r = g.f[phase][0][0];
#pragma unroll
    for (int w = 1; w < SMP_W; w++) r = fmaxf(r, g.f[phase][w][0]);
    phase ^= 1;
    return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_cluster_u32(uint32_t local_addr, uint32_t rank, int v) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ra), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// grid (n_utt, 4), cluster (1, 4, 1): CTA y of a cluster samples codebook y of utterance x; the four samples meet in the
// shared memory of cluster rank 0 (DSMEM stores), whose thread 0 runs the per-utterance state machine.
__global__ void __launch_bounds__(SMP_T) sample_kernel(const float* __restrict__ logits, UttState* __restrict__ st,
                                                       int* __restrict__ seq_len, int* __restrict__ next_tok,
                                                       int* __restrict__ gen_tok, const float* __restrict__ noise,
                                                       int* __restrict__ iter_counter, SampleParams p) {
    pdl_launch_dependents();
    const int ts = ts_begin(TSK_SAMPLE);
    pdl_wait();
    ts_dep(ts);
    __shared__ BlockRed red;
    __shared__ int s_samples[4];
    __shared__ int s_argmax0;
    __shared__ float s_bestv[SMP_W], s_amv[SMP_W];
    __shared__ int s_besti[SMP_W], s_ami[SMP_W];
    const int u = p.utt0 + blockIdx.x, k = blockIdx.y;
    if (p.count_iter && blockIdx.x == 0 && k == 0 && threadIdx.x == 0) atomicAdd(iter_counter, 1);
    const UttState S = st[u];
    if (S.done) return;                                    // uniform over the cluster
    const int K = p.K, V = p.V;
    const int t = threadIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int phase = 0;
    const int row0 = u * p.rpu;
    const bool use_cfg = (p.rpu == 2) && (S.cfg_tag == p.cfg_stride);
    const float c1 = p.cfg_coef, c2 = (float)(1.0 - (double)p.cfg_coef);
    const float* l0 = logits + ((int64_t)row0 * K + k) * V;
    const float* l1 = logits + ((int64_t)(row0 + (use_cfg ? 1 : 0)) * K + k) * V;
    // all loads first (independent, clamped), then the rules
    float l[SMP_NPT], lu[SMP_NPT];
#pragma unroll
    for (int i = 0; i < SMP_NPT; i++) {
        const int v = min(t + SMP_T * i, V - 1);
        l[i] = l0[v];
        lu[i] = l1[v];
    }
    bool prev_sil = false;
    for (int i = 0; i < p.n_silence; i++) prev_sil |= (S.prev_token == p.silence[i]);
    const bool rep_rule = (S.num_eog == 0) && (k == 0) && p.stop_repetition > 0 && prev_sil &&
                          S.consec_silence > p.stop_repetition;
    const float rep_f = (float)(S.consec_silence - (p.stop_repetition - 1));
#pragma unroll
    for (int i = 0; i < SMP_NPT; i++) {
        const int v = t + SMP_T * i;
        float a = l[i];
        if (use_cfg) a = __fadd_rn(__fmul_rn(c1, a), __fmul_rn(c2, lu[i]));
        if (v == p.eos || v == p.sos || (v >= p.mts && v < p.mts + p.max_n_spans)) a = -10000.f;
        if (S.num_gen < K - 1 && k >= S.num_gen + 1 && v == p.empty_token) a = 10000.f;
        if (S.num_eog > 0) {
            if (k >= S.num_eog + 1 && (v == p.eog || v == p.empty_token)) a = -10000.f;
        } else {
            if (k >= 1 && v == p.eog) a = -10000.f;
            if (rep_rule && v == S.prev_token) a = a < 0.f ? a * rep_f : a / rep_f;
        }
        if (p.temperature != 1.0f) a = a / p.temperature;
        l[i] = (v < V) ? a : -INFINITY;
    }
    ts_aux(ts, 0);
    uint32_t key[SMP_NPT];
#pragma unroll
    for (int i = 0; i < SMP_NPT; i++) key[i] = fkey(l[i]);
    // ---- top-k: threshold = k-th largest value.  Radix descent on the order-preserving keys, two bits per round
    // (three candidate thresholds counted at once; identical to a bit-by-bit descent) ---------------------------------
    if (p.top_k > 0) {
        const int keff = min(max(p.top_k, 1), V);
        uint32_t T = 0;
        for (int bit = 30; bit >= 0; bit -= 2) {
            const uint32_t c1k = T | (1u << bit), c2k = T | (2u << bit), c3k = T | (3u << bit);
            int c[3] = {0, 0, 0};
#pragma unroll
            for (int i = 0; i < SMP_NPT; i++) {
                const bool in = t + SMP_T * i < V;
                c[0] += in && key[i] >= c1k; c[1] += in && key[i] >= c2k; c[2] += in && key[i] >= c3k;
            }
            block_isum(c, red, phase, warp, lane);
            T = c[2] >= keff ? c3k : c[1] >= keff ? c2k : c[0] >= keff ? c1k : T;
        }
#pragma unroll
        for (int i = 0; i < SMP_NPT; i++) if (key[i] < T) { l[i] = -INFINITY; key[i] = fkey(-INFINITY); }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < SMP_NPT; i++) mx = fmaxf(mx, l[i]);
    mx = block_max(mx, red, phase, warp, lane);
    float e[SMP_NPT];
    float zs[1] = {0.f};
#pragma unroll
    for (int i = 0; i < SMP_NPT; i++) { e[i] = (l[i] == -INFINITY) ? 0.f : expf(l[i] - mx); zs[0] += e[i]; }
    block_sum(zs, red, phase, warp, lane);
    float Z = zs[0];
    ts_aux(ts, 1);
    // ---- top-p: keep token i iff the probability mass ranked strictly above it is <= top_p ---------------
    if (p.top_p < 1.0f) {
        const float lim = p.top_p * Z;
        uint32_t T = 0;
        for (int bit = 30; bit >= 0; bit -= 2) {
            const uint32_t c1k = T | (1u << bit), c2k = T | (2u << bit), c3k = T | (3u << bit);
            float ms[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < SMP_NPT; i++) {
                ms[0] += key[i] >= c1k ? e[i] : 0.f; ms[1] += key[i] >= c2k ? e[i] : 0.f; ms[2] += key[i] >= c3k ? e[i] : 0.f;
            }
            block_sum(ms, red, phase, warp, lane);
            T = ms[2] > lim ? c3k : ms[1] > lim ? c2k : ms[0] > lim ? c1k : T;
        }
        zs[0] = 0.f;
#pragma unroll
        for (int i = 0; i < SMP_NPT; i++) {
            if (key[i] < T) { l[i] = -INFINITY; e[i] = 0.f; }
            zs[0] += e[i];
        }
        block_sum(zs, red, phase, warp, lane);
        Z = zs[0];
    }
    ts_aux(ts, 2);
    // ---- sample: argmax_i (e_i / Z) / q_i, q ~ Exp(1); first index wins ties; also argmax of logits ----
    float bestv = -1.f; int besti = 0x7fffffff;
    float amv = -INFINITY; int ami = 0x7fffffff;
    const float* nz = noise ? noise + (((int64_t)S.n_tok * p.n_utt + u) * K + k) * V : nullptr;
    float qn[SMP_NPT];
#pragma unroll
    for (int i = 0; i < SMP_NPT; i++) {
        const int v = min(t + SMP_T * i, V - 1);
        if (nz) qn[i] = nz[v];
        else {
            const uint32_t rb = philox_u32(p.seed, (uint32_t)v, (uint32_t)k, (uint32_t)S.n_tok, (uint32_t)S.rng_id);
            qn[i] = fmaxf(-logf(((float)rb + 0.5f) * 2.3283064365386963e-10f), 1e-30f);
        }
    }
#pragma unroll
    for (int i = 0; i < SMP_NPT; i++) {
        const int v = t + SMP_T * i;
        if (v < V) {
            const float sc = (e[i] / Z) / qn[i];
            if (sc > bestv) { bestv = sc; besti = v; }     // ascending v within a thread: first index kept
            if (l[i] > amv) { amv = l[i]; ami = v; }
        }
    }
    // reduce (value desc, index asc) over the block
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bestv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ov > bestv || (ov == bestv && oi < besti)) { bestv = ov; besti = oi; }
        const float av = __shfl_xor_sync(0xffffffffu, amv, o);
        const int ai = __shfl_xor_sync(0xffffffffu, ami, o);
        if (av > amv || (av == amv && ai < ami)) { amv = av; ami = ai; }
    }
    if (lane == 0) { s_bestv[warp] = bestv; s_besti[warp] = besti; s_amv[warp] = amv; s_ami[warp] = ami; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float bv = s_bestv[0]; int bi = s_besti[0];
        float av = s_amv[0]; int ai = s_ami[0];
        for (int w = 1; w < SMP_W; w++) {
            if (s_bestv[w] > bv || (s_bestv[w] == bv && s_besti[w] < bi)) { bv = s_bestv[w]; bi = s_besti[w]; }
            if (s_amv[w] > av || (s_amv[w] == av && s_ami[w] < ai)) { av = s_amv[w]; ai = s_ami[w]; }
        }
        st_cluster_u32(smem_u32(&s_samples[k]), 0, bi);
        if (k == 0) s_argmax0 = ai;
    }
    cluster_barrier();                                     // the four samples are in rank 0's shared memory
    if (k != 0 || threadIdx.x != 0) return;
    ts_aux(ts, 3);
    // ---- state machine (single thread) ------------------------------------------------------------------
    UttState N = S;
    int smp[4] = {s_samples[0], s_samples[1], s_samples[2], s_samples[3]};
    if (p.rpu == 2) N.cfg_tag = use_cfg ? 1 : S.cfg_tag + 1;
    if (S.num_eog > 0) {
        for (int kk = 0; kk < S.num_eog; kk++) smp[kk] = p.empty_token;
        smp[S.num_eog] = p.eog;
        N.num_eog = S.num_eog + 1;
    } else {
        const int Ty = S.y_len + 1;                       // y_input.shape[1] of ssr.py:739
        if (smp[0] == p.eog || s_argmax0 == p.eog || Ty > S.x_len * 10) { smp[0] = p.eog; N.num_eog = 1; }
        bool sil = false;
        for (int i = 0; i < p.n_silence; i++) sil |= (smp[0] == p.silence[i]);
        N.consec_silence = (sil && smp[0] == S.prev_token) ? S.consec_silence + 1 : 0;
        N.prev_token = smp[0];
    }
    N.num_gen = S.num_gen + 1;
    int* gt = gen_tok + ((int64_t)u * p.max_steps + S.n_tok) * K;
    for (int kk = 0; kk < K; kk++) gt[kk] = smp[kk];
    N.n_tok = S.n_tok + 1;
    N.span_len[S.span_idx] = S.span_len[S.span_idx] + 1;
    if (N.num_eog == K) {                                  // span finished; last samples are not fed back
        N.span_idx = S.span_idx + 1;
        if (N.span_idx >= S.n_spans) N.done = 1;
        else {
            for (int kk = 0; kk < K; kk++) next_tok[u * K + kk] = p.mts + N.span_idx;   // ssr.py:654-660
            N.num_gen = 0; N.num_eog = 0; N.cfg_tag = 1; N.prev_token = -1; N.consec_silence = 0;
        }
    } else {
        for (int kk = 0; kk < K; kk++) next_tok[u * K + kk] = smp[kk];
    }
    if (N.n_tok >= p.max_steps) N.done = 1;
    N.y_len = S.y_len + 1;
    for (int j = 0; j < p.rpu; j++) seq_len[row0 + j] += 1;
    st[u] = N;
    ts_end(ts);
}

// =================================================================================================
// Training forward / loss (models/ssr.py:326-379): per-codebook masked cross entropy and top-10 accuracy of the
// teacher-forced logits.  logits [Ty][K][V] fp32; the target of position t is the input token of position t+1 (ssr.py:330-331);
// flags [K][Ty-1]: bit 0 = the position enters the loss / accuracy (tmp_mask), bit 1 = it counts as a token (mask).
//   masked_ce_kernel        one CTA per (position, codebook): nll = logsumexp(l) - l[target]; hit = fewer than 10 logits exceed l[target]
//   masked_ce_reduce_kernel one CTA per codebook: sums in a fixed order (double) -> {sum nll, n_loss, hits, n_tokens}
// =================================================================================================
__global__ void __launch_bounds__(256) masked_ce_kernel(const float* __restrict__ logits, const int* __restrict__ audio,
                                                        const unsigned char* __restrict__ flags, int Ty, int K, int V,
                                                        float* __restrict__ nll, unsigned char* __restrict__ hit) {
    const int t = blockIdx.x, k = blockIdx.y, o = k * (Ty - 1) + t;
    if (!(flags[o] & 1)) {
        if (threadIdx.x == 0) { nll[o] = 0.f; hit[o] = 0; }
        return;
    }
    __shared__ float red[8];
    __shared__ int redi[8];
    const int target = audio[k * Ty + t + 1];
    const float* row = logits + ((size_t)t * K + k) * V;
    const float lt = row[target];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float m = -INFINITY;
    int above = 0;
    for (int v = threadIdx.x; v < V; v += 256) {
        const float l = row[v];
        m = fmaxf(m, l);
        above += l > lt ? 1 : 0;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
        above += __shfl_xor_sync(0xffffffffu, above, off);
    }
    if (lane == 0) { red[warp] = m; redi[warp] = above; }
    __syncthreads();
    m = red[0]; above = redi[0];
#pragma unroll
    for (int w = 1; w < 8; w++) { m = fmaxf(m, red[w]); above += redi[w]; }
    __syncthreads();
    float sum = 0.f;
    for (int v = threadIdx.x; v < V; v += 256) sum += expf(row[v] - m);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < 8; w++) tot += red[w];                      // fixed order
        nll[o] = (m + logf(tot)) - lt;
        hit[o] = above < 10 ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) masked_ce_reduce_kernel(const float* __restrict__ nll, const unsigned char* __restrict__ hit,
                                                               const unsigned char* __restrict__ flags, int n, double* __restrict__ out) {
    const int k = blockIdx.x;
    __shared__ double s_l[256];
    __shared__ int s_n[256], s_h[256], s_c[256];
    double l = 0.0;
    int nl = 0, nh = 0, nc = 0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const unsigned char f = flags[k * n + i];
        if (f & 1) { l += (double)nll[k * n + i]; nl++; nh += hit[k * n + i]; }
        if (f & 2) nc++;
    }
    s_l[threadIdx.x] = l; s_n[threadIdx.x] = nl; s_h[threadIdx.x] = nh; s_c[threadIdx.x] = nc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            s_l[threadIdx.x] += s_l[threadIdx.x + w]; s_n[threadIdx.x] += s_n[threadIdx.x + w];
            s_h[threadIdx.x] += s_h[threadIdx.x + w]; s_c[threadIdx.x] += s_c[threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[k * 4 + 0] = s_l[0]; out[k * 4 + 1] = s_n[0]; out[k * 4 + 2] = s_h[0]; out[k * 4 + 3] = s_c[0]; }
}

int launch_masked_ce(const float* logits, const int* audio, const unsigned char* flags, int Ty, int K, int V, float* nll,
                     unsigned char* hit, double* out, cudaStream_t s) {
    SSRB_CHECK(Ty >= 2 && K >= 1 && V >= 10, "masked_ce: needs at least two positions and ten classes");
    SSRB_LAUNCH(masked_ce_kernel, dim3(Ty - 1, K), 256, 0, s, logits, audio, flags, Ty, K, V, nll, hit);
    SSRB_LAUNCH(masked_ce_reduce_kernel, K, 256, 0, s, nll, hit, flags, Ty - 1, out);
    return 0;
}

int launch_sample(const float* logits, UttState* st, int* seq_len, int* next_tok, int* gen_tok, const float* noise,
                  int* iter_counter, const SampleParams& p_in, cudaStream_t s, int only_utt) {
    SampleParams p = p_in;
    SSRB_CHECK(p.K == 4, "sample: n_codebooks must be 4");
    SSRB_CHECK(p.V <= SMP_T * SMP_NPT, "sample: audio vocabulary too large for the sampling kernel");
    // only_utt >= 0: the first sample of ONE freshly admitted utterance (continuous batching); not a loop iteration
    p.utt0 = only_utt >= 0 ? only_utt : 0;
    p.count_iter = only_utt >= 0 ? 0 : 1;
    SSRB_TRY(launch_pdl(sample_kernel, dim3(only_utt >= 0 ? 1 : p.n_utt, 4), dim3(SMP_T), 0, s, /*cluster_y=*/4, logits, st, seq_len,
                        next_tok, gen_tok, noise, iter_counter, p));
    return 0;
}

int ts_arm_lm_kernels(const TsBuf& t) { return ts_arm_tu(t); }

}  // namespace ssrb
