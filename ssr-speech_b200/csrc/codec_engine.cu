// codec_engine.cu — WM-Encodec (SEANet encoder / RVQ / SEANet + watermark decoder) behind the C ABI.
//
// Mirrors audiocraft/models/wmencodec.py::WMEncodecModel.{encode,decode,wmdecode} with
// renormalize=False, n_residual_layers=1, non-causal zero padding.  Weight-norm is
// folded once at load time (the reference recomputes g*v/||v|| on every call, modules/conv.py:21-30).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "codec_kernels.cuh"
#include "conv_tc.cuh"
#include "../../include/ssr_b200.h"

using namespace ssrb;

namespace {

struct ConvW {
    float* w = nullptr; float* b = nullptr; int d0 = 0, d1 = 0, k = 0; bool has_w = false, has_b = false;
    mutable float* wt = nullptr;        // [Cin][k][Cout] copy for conv1d_v2_kernel (built on first use)
    std::vector<float> hw, hb;          // host copies (folded weight, bias) for the tensor-core repack
};
// tensor-core operand of one convolution: W' [taps][N][Cw] bf16 (conv_tc.cu), fp32 bias (+ per-mark bias for wm_proj)
struct TcW { bf16* w = nullptr; float* bias = nullptr; float* bias_alt = nullptr; int taps = 0, N = 0, Cw = 0, bias_mod = 0; };
struct ClT { bf16* raw = nullptr; bf16* act = nullptr; int C = 0, T = 0; };   // channels-last tensor [B][G+T+G][C]
// encoder path (conv_tc32.cu): fp32 channels-last; weights and ELU'd activations split into two TF32 numbers (hi, lo)
struct Tc32W { float* wh = nullptr; float* wl = nullptr; float* bias = nullptr; int taps = 0, N = 0, Cw = 0; };
struct Cl32 { float* raw = nullptr; float* hi = nullptr; float* lo = nullptr; int C = 0, T = 0; };
struct LstmW {
    float *wih[4] = {}, *whh[4] = {}, *bsum[4] = {};
    bf16* wih_bf16[4] = {};            // tensor-core input projection (decoder-side LSTMs only)
    float *wih_hi[4] = {}, *wih_lo[4] = {};   // 3 x TF32 input projection (encoder)
    std::vector<float> bih[4], bhh[4];
    bool has_wih[4] = {}, has_whh[4] = {}, has_b[4] = {};
    int C = 0;
};
struct Pending { std::vector<float> g, v; std::vector<int64_t> vshape; bool has_g = false, has_v = false; };

struct Arena {
    char* base = nullptr; size_t cap = 0, off = 0; bool dry = true;
    float* f(size_t n) { size_t o = off; off += ((n * 4 + 255) / 256) * 256; return dry ? nullptr : (float*)(base + o); }
};

struct Tensor { float* p; int C; int T; };

}  // namespace

struct ssrb_codec {
    ssrb_codec_config cfg;
    int device = 0;
    int hop = 1;
    std::map<std::string, ConvW> convs;
    std::map<std::string, LstmW> lstms;
    std::map<std::string, Pending> pending;
    float* codebooks = nullptr; float* cb_sq = nullptr; bool has_cb[16] = {};
    float* wm_embed = nullptr; bool has_wm_embed = false;
    unsigned int* bar = nullptr;
    Arena arena;
    std::map<std::string, TcW> tcw;
    std::map<std::string, Tc32W> tc32w;
    std::vector<float> wm_embed_host;   // renormalised rows
    bool use_tc = false;
    bool derived_stale = false;         // a tensor was (re)loaded: repacked copies below are dropped before the next compute call
};

// the repacked / re-typed weight copies built lazily by the compute paths (tap-major bf16, TF32 hi/lo pairs, bf16 W_ih)
static void drop_derived(ssrb_codec* c) {
    cudaDeviceSynchronize();
    for (auto& kv : c->tcw) { cudaFree(kv.second.w); cudaFree(kv.second.bias); cudaFree(kv.second.bias_alt); }
    for (auto& kv : c->tc32w) { cudaFree(kv.second.wh); cudaFree(kv.second.wl); cudaFree(kv.second.bias); }
    c->tcw.clear(); c->tc32w.clear();
    for (auto& kv : c->lstms)
        for (int l = 0; l < 4; l++) {
            cudaFree(kv.second.wih_bf16[l]); cudaFree(kv.second.wih_hi[l]); cudaFree(kv.second.wih_lo[l]);
            kv.second.wih_bf16[l] = nullptr; kv.second.wih_hi[l] = kv.second.wih_lo[l] = nullptr;
        }
    c->derived_stale = false;
}

static int dalloc(void** p, size_t bytes) { SSRB_CUDA(cudaMalloc(p, bytes ? bytes : 16)); return 0; }

int ssrb_codec_create(const ssrb_codec_config* c, int device, ssrb_codec** out) {
    SSRB_CHECK(c && out, "null argument");
    SSRB_CHECK(c->n_ratios == 4, "codec: exactly 4 ratios supported (WMSEANetDecoder.forward hard-codes 4 skips)");
    SSRB_CHECK(c->channels == 1, "codec: mono only");
    SSRB_CHECK(c->lstm_layers >= 0 && c->lstm_layers <= 4, "codec: lstm_layers out of range");
    SSRB_CHECK(c->n_q <= 16 && c->dimension <= 128 && c->dimension % 16 == 0, "codec: unsupported RVQ geometry");
    SSRB_CUDA(cudaSetDevice(device));
    ssrb_codec* cd = new ssrb_codec();
    struct Guard { ssrb_codec* p; ~Guard() { if (p) ssrb_codec_destroy(p); } } guard{cd};   // an allocation failure below frees what exists
    cd->cfg = *c; cd->device = device;
    cd->use_tc = c->tensor_cores != 0;
    if (cd->cfg.max_batch_chunk <= 0 || cd->cfg.max_batch_chunk > 32) cd->cfg.max_batch_chunk = 8;
    cd->hop = 1;
    for (int i = 0; i < c->n_ratios; i++) cd->hop *= c->ratios[i];
    SSRB_TRY(dalloc((void**)&cd->codebooks, (size_t)c->n_q * c->bins * c->dimension * 4));
    SSRB_TRY(dalloc((void**)&cd->cb_sq, (size_t)c->n_q * c->bins * 4));
    SSRB_TRY(dalloc((void**)&cd->wm_embed, 2 * (c->dimension / 16) * 4));
    SSRB_TRY(dalloc((void**)&cd->bar, 4));
    guard.p = nullptr;
    *out = cd;
    return 0;
}

void ssrb_codec_destroy(ssrb_codec* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto& kv : c->convs) { cudaFree(kv.second.w); cudaFree(kv.second.b); cudaFree(kv.second.wt); }
    drop_derived(c);
    for (auto& kv : c->lstms) for (int l = 0; l < 4; l++) { cudaFree(kv.second.wih[l]); cudaFree(kv.second.whh[l]); cudaFree(kv.second.bsum[l]); }
    cudaFree(c->codebooks); cudaFree(c->cb_sq); cudaFree(c->wm_embed); cudaFree(c->bar); cudaFree(c->arena.base);
    delete c;
}

static bool ends_with(const std::string& s, const char* suf) {
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

static int upload_new(float** dst, const float* host, size_t n) {
    if (*dst) cudaFree(*dst);
    *dst = nullptr;
    SSRB_TRY(dalloc((void**)dst, n * 4));
    SSRB_CUDA(cudaMemcpy(*dst, host, n * 4, cudaMemcpyHostToDevice));
    return 0;
}

int ssrb_codec_load_tensor(ssrb_codec* c, const char* name_c, const float* host, const int64_t* shape, int ndim) {
    SSRB_CHECK(c && name_c && host, "null argument");
    SSRB_CUDA(cudaSetDevice(c->device));
    const std::string name(name_c);
    int64_t n = 1;
    for (int i = 0; i < ndim; i++) n *= shape[i];
    c->derived_stale = true;
    // RVQ codebooks: quantizer.vq.layers.{q}._codebook.embed [bins, dim]
    if (name.rfind("quantizer.vq.layers.", 0) == 0) {
        if (!ends_with(name, "._codebook.embed")) return 0;          // inited / cluster_size / embed_avg: training state
        const int q = atoi(name.c_str() + 20);
        SSRB_CHECK(q >= 0 && q < c->cfg.n_q && n == (int64_t)c->cfg.bins * c->cfg.dimension, "bad codebook");
        SSRB_CUDA(cudaMemcpy(c->codebooks + (size_t)q * n, host, n * 4, cudaMemcpyHostToDevice));
        std::vector<float> sq(c->cfg.bins);
        for (int i = 0; i < c->cfg.bins; i++) {
            double s = 0;
            for (int d = 0; d < c->cfg.dimension; d++) { const double v = host[(size_t)i * c->cfg.dimension + d]; s += v * v; }
            sq[i] = (float)s;
        }
        SSRB_CUDA(cudaMemcpy(c->cb_sq + (size_t)q * c->cfg.bins, sq.data(), c->cfg.bins * 4, cudaMemcpyHostToDevice));
        c->has_cb[q] = true;
        return 0;
    }
    if (name == "wmdecoder.wm_embed.weight") {      // nn.Embedding(2, dim//16, max_norm=True): rows renormalised to norm<=1
        const int E = c->cfg.dimension / 16;
        SSRB_CHECK(n == 2 * E, "bad wm_embed shape");
        std::vector<float> w(host, host + n);
        for (int r = 0; r < 2; r++) {
            float nr = 0.f;
            for (int d = 0; d < E; d++) nr += w[r * E + d] * w[r * E + d];
            nr = std::sqrt(nr);
            if (nr > 1.0f) { const float sc = 1.0f / (nr + 1e-7f); for (int d = 0; d < E; d++) w[r * E + d] *= sc; }
        }
        SSRB_CUDA(cudaMemcpy(c->wm_embed, w.data(), n * 4, cudaMemcpyHostToDevice));
        c->wm_embed_host = w;
        c->has_wm_embed = true;
        return 0;
    }
    // LSTM: "<prefix>lstm.{weight_ih,weight_hh,bias_ih,bias_hh}_l{n}"
    const size_t lp = name.find("lstm.");
    if (lp != std::string::npos) {
        const std::string prefix = name.substr(0, lp), sub = name.substr(lp + 5);
        const int l = sub.back() - '0';
        SSRB_CHECK(l >= 0 && l < 4, "bad lstm layer index");
        LstmW& L = c->lstms[prefix];
        if (sub.rfind("weight_ih", 0) == 0) { L.C = (int)shape[1]; SSRB_TRY(upload_new(&L.wih[l], host, n)); L.has_wih[l] = true; }
        else if (sub.rfind("weight_hh", 0) == 0) { L.C = (int)shape[1]; SSRB_TRY(upload_new(&L.whh[l], host, n)); L.has_whh[l] = true; }
        else if (sub.rfind("bias_ih", 0) == 0) L.bih[l].assign(host, host + n);
        else if (sub.rfind("bias_hh", 0) == 0) L.bhh[l].assign(host, host + n);
        if (!L.bih[l].empty() && !L.bhh[l].empty() && L.bih[l].size() == L.bhh[l].size()) {      // (re)built whenever either half changes
            std::vector<float> s(L.bih[l].size());
            for (size_t i = 0; i < s.size(); i++) s[i] = L.bih[l][i] + L.bhh[l][i];
            SSRB_TRY(upload_new(&L.bsum[l], s.data(), s.size()));
            L.has_b[l] = true;
        }
        return 0;
    }
    // convs: "<prefix>{weight_g, weight_v, weight, bias}" with prefix ending in "conv.conv." or "convtr.convtr."
    const size_t dot = name.rfind('.');
    SSRB_CHECK(dot != std::string::npos, "unrecognised tensor name");
    const std::string prefix = name.substr(0, dot + 1), leaf = name.substr(dot + 1);
    if (prefix.find("conv.conv.") == std::string::npos && prefix.find("convtr.convtr.") == std::string::npos) return 0;
    ConvW& W = c->convs[prefix];
    if (leaf == "bias") { SSRB_TRY(upload_new(&W.b, host, n)); W.hb.assign(host, host + n); W.has_b = true; return 0; }
    if (leaf == "weight") {
        SSRB_CHECK(ndim == 3, "conv weight must be 3-D");
        W.d0 = (int)shape[0]; W.d1 = (int)shape[1]; W.k = (int)shape[2];
        if (W.wt) { cudaFree(W.wt); W.wt = nullptr; }
        SSRB_TRY(upload_new(&W.w, host, n)); W.hw.assign(host, host + n); W.has_w = true; return 0;
    }
    if (leaf == "weight_g" || leaf == "weight_v") {
        Pending& P = c->pending[prefix];
        if (leaf == "weight_g") { P.g.assign(host, host + n); P.has_g = true; }
        else { SSRB_CHECK(ndim == 3, "weight_v must be 3-D"); P.v.assign(host, host + n); P.vshape.assign(shape, shape + 3); P.has_v = true; }
        if (P.has_g && P.has_v) {
            // legacy torch.nn.utils.weight_norm(dim=0): w = v * (g / ||v||), norm over dims (1,2)
            const int64_t d0 = P.vshape[0], inner = P.vshape[1] * P.vshape[2];
            SSRB_CHECK((int64_t)P.g.size() == d0, "weight_g / weight_v mismatch");
            std::vector<float> w(P.v.size());
            for (int64_t i = 0; i < d0; i++) {
                double s = 0;
                for (int64_t j = 0; j < inner; j++) { const double v = P.v[i * inner + j]; s += v * v; }
                const float sc = P.g[i] / (float)std::sqrt(s);
                for (int64_t j = 0; j < inner; j++) w[i * inner + j] = P.v[i * inner + j] * sc;
            }
            W.d0 = (int)P.vshape[0]; W.d1 = (int)P.vshape[1]; W.k = (int)P.vshape[2];
            if (W.wt) { cudaFree(W.wt); W.wt = nullptr; }
        SSRB_TRY(upload_new(&W.w, w.data(), w.size())); W.hw = w; W.has_w = true;
            c->pending.erase(prefix);
        }
        return 0;
    }
    return 0;
}

// ---- graph helpers ------------------------------------------------------------------------------------
struct Ctx { ssrb_codec* c; cudaStream_t s; int B; };

static int get_conv(ssrb_codec* c, const std::string& key, const ConvW** out) {
    auto it = c->convs.find(key);
    if (it == c->convs.end() || !it->second.has_w || !it->second.has_b) { set_error("codec tensor missing: " + key); return 1; }
    *out = &it->second;
    return 0;
}

// StreamableConv1d (conv.py:185-201); W [Cout][Cin][k]
static int conv(Ctx& x, const std::string& key, Tensor in, int stride, bool elu_in, const float* res, Tensor* out) {
    const ConvW* W;
    SSRB_TRY(get_conv(x.c, key, &W));
    SSRB_CHECK(W->d1 == in.C, ("conv input channels mismatch at " + key).c_str());
    const int k = W->k, total = k - stride, right = total / 2, left = total - right;
    const int Tout = (in.T + stride - 1) / stride;
    out->C = W->d0; out->T = Tout; out->p = x.c->arena.f((size_t)x.B * W->d0 * Tout);
    if (x.c->arena.dry) return 0;
    if (!W->wt) {
        SSRB_TRY(dalloc((void**)&W->wt, (size_t)W->d0 * W->d1 * k * 4));
        SSRB_TRY(launch_conv_w_transpose(W->w, W->wt, W->d0, W->d1, k, x.s));
    }
    return launch_conv1d(in.p, x.B, in.C, in.T, W->w, W->b, W->d0, k, stride, left, Tout, elu_in, res, out->p, x.s, W->wt);
}
// StreamableConvTranspose1d (conv.py:221-243); W [Cin][Cout][k]
static int convtr(Ctx& x, const std::string& key, Tensor in, int stride, bool elu_in, Tensor* out) {
    const ConvW* W;
    SSRB_TRY(get_conv(x.c, key, &W));
    SSRB_CHECK(W->d0 == in.C, ("convtr input channels mismatch at " + key).c_str());
    const int k = W->k, total = k - stride, right = total / 2, left = total - right;
    const int Tout = (in.T - 1) * stride + k - total;
    out->C = W->d1; out->T = Tout; out->p = x.c->arena.f((size_t)x.B * W->d1 * Tout);
    if (x.c->arena.dry) return 0;
    return launch_convtr1d(in.p, x.B, in.C, in.T, W->w, W->b, W->d1, k, stride, left, Tout, elu_in, out->p, x.s);
}
// SEANetResnetBlock (seanet.py:16-60): x + conv_k1(ELU(conv_k3(ELU(x))))
static int resblock(Ctx& x, const std::string& prefix, Tensor in, Tensor* out) {
    Tensor h;
    SSRB_TRY(conv(x, prefix + "block.1.conv.conv.", in, 1, true, nullptr, &h));
    return conv(x, prefix + "block.3.conv.conv.", h, 1, true, in.p, out);
}
template <typename T>
__global__ void f32_to_T_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = from_f32<T>(src[i]);
}
// StreamableLSTM (lstm.py:10-25).  tc: input projections x.W_ih^T as bf16 tcgen05 GEMMs (decoder-side LSTMs of the tensor-core path)
// tc: bf16 tensor-core input projection + mma.sync recurrence (decoder side).  tc32: the input projection as a 3 x TF32 tcgen05 GEMM
// (conv_tc32 with one tap: fp32-grade), the recurrence stays the fp32 kernel (encoder: its output decides RVQ indices).
static int lstm(Ctx& x, const std::string& prefix, Tensor in, Tensor* out, bool tc = false, bool tc32 = false) {
    auto it = x.c->lstms.find(prefix);
    if (it == x.c->lstms.end()) { set_error("codec lstm missing: " + prefix); return 1; }
    LstmW& L = it->second;
    const int C = in.C, T = in.T, B = x.B, nl = x.c->cfg.lstm_layers;
    SSRB_CHECK(L.C == C, "lstm width mismatch");
    tc = tc && (C % 64 == 0);
    tc32 = tc32 && !tc && (C % 32 == 0);
    Arena& A = x.c->arena;
    float* seq = A.f((size_t)T * B * C);
    float* pre = A.f((size_t)T * B * 4 * C);
    float* hs[2] = {A.f((size_t)T * B * C), A.f((size_t)T * B * C)};
    float* hbuf = A.f((size_t)2 * 32 * C);            // fp32 kernel: [2][B][C] fp32; tensor-core kernel: [2][32][C] bf16
    bf16* seq16 = tc ? (bf16*)A.f(((size_t)T * B * C + 1) / 2) : nullptr;
    bf16* hs16 = tc ? (bf16*)A.f(((size_t)T * B * C + 1) / 2) : nullptr;
    float* x_hi = tc32 ? A.f((size_t)T * B * C) : nullptr;
    float* x_lo = tc32 ? A.f((size_t)T * B * C) : nullptr;
    out->C = C; out->T = T; out->p = A.f((size_t)B * C * T);
    if (A.dry) return 0;
    SSRB_TRY(launch_bct_to_tbc(in.p, B, C, T, seq, x.s, seq16));
    const float* cur = seq;
    const bf16* cur16 = seq16;
    for (int l = 0; l < nl; l++) {
        SSRB_CHECK(L.has_wih[l] && L.has_whh[l] && L.has_b[l], "lstm layer weights missing");
        GemmArgs g;
        g.bias = L.bsum[l]; g.C = pre; g.ldc = 4 * C; g.lda = C; g.ldw = C;
        g.M = T * B; g.N = 4 * C; g.K = C; g.c_dtype = SSRB_DTYPE_F32;
        if (tc) {
            if (!L.wih_bf16[l]) {
                SSRB_TRY(dalloc((void**)&L.wih_bf16[l], (size_t)4 * C * C * 2));
                SSRB_LAUNCH(f32_to_T_kernel<bf16>, 512, 256, 0, x.s, L.wih[l], L.wih_bf16[l], (int64_t)4 * C * C);
            }
            g.A = cur16; g.W = L.wih_bf16[l]; g.ab_dtype = SSRB_DTYPE_BF16;
            SSRB_TRY(gemm_tc(g, nullptr, 0, x.s));
        } else if (tc32) {
            if (!L.wih_hi[l]) {
                SSRB_TRY(dalloc((void**)&L.wih_hi[l], (size_t)4 * C * C * 4));
                SSRB_TRY(dalloc((void**)&L.wih_lo[l], (size_t)4 * C * C * 4));
                SSRB_TRY(launch_split_tf32(L.wih[l], L.wih_hi[l], L.wih_lo[l], (long long)4 * C * C, x.s));
            }
            SSRB_TRY(launch_split_tf32(cur, x_hi, x_lo, (long long)T * B * C, x.s));
            ConvTc32Args a;                      // pre[T*B, 4C] = seq[T*B, C] . W_ih^T + (b_ih + b_hh): one "utterance" of T*B rows, one tap
            a.x_hi = x_hi; a.x_lo = x_lo; a.B = 1; a.x_bstride = (long long)T * B * C; a.x_base_off = 0; a.Cw = C; a.rows_v = T * B;
            a.w_hi = L.wih_hi[l]; a.w_lo = L.wih_lo[l]; a.taps = 1; a.N = 4 * C; a.T_rows = T * B; a.bias = L.bsum[l];
            a.out_raw = pre; a.out_bstride = 0; a.out_off = 0; a.elu = false;
            SSRB_TRY(conv_tc32(a, x.s));
        } else {
            g.A = cur; g.W = L.wih[l]; g.ab_dtype = SSRB_DTYPE_F32;
            SSRB_TRY(gemm_simt(g, x.s));
        }
        static const bool lstm_simt = [] { const char* e = getenv("SSRB_LSTM_SIMT"); return e && e[0] == '1'; }();
        if (tc && lstm_mma_supported(C) && !lstm_simt)
            SSRB_TRY(launch_lstm_layer_mma(pre, L.whh[l], hs[l & 1], (bf16*)hbuf, x.c->bar, T, B, C, x.s, l + 1 < nl ? hs16 : nullptr));
        else
            SSRB_TRY(launch_lstm_layer(pre, L.whh[l], hs[l & 1], hbuf, x.c->bar, T, B, C, x.s, (tc && l + 1 < nl) ? hs16 : nullptr));
        cur = hs[l & 1];
        cur16 = hs16;
    }
    return launch_tbc_to_bct_add(cur, in.p, B, C, T, out->p, x.s);
}

// frame-rate k7 convolution (128 <-> 1024 channels) through the tensor-core path: fp32 channels-first in and out
static int conv_frame_tc(Ctx& x, const std::string& key, Tensor in, bool elu_in, const float* res, Tensor* out);

// SEANetEncoder as the 5 slices WMSEANetDecoder.forward uses (seanet.py:559-574)
static int encoder_stage(Ctx& x, const std::string& p, int stage, Tensor in, Tensor* out) {
    const int* r = x.c->cfg.ratios;              // decoder order; the encoder reverses (seanet.py:101)
    const int er[4] = {r[3], r[2], r[1], r[0]};
    Tensor a, b;
    if (stage == 0) {
        SSRB_TRY(conv(x, p + "model.0.conv.conv.", in, 1, false, nullptr, &a));
        return resblock(x, p + "model.1.", a, out);
    }
    if (stage <= 3) {
        const int ci = 3 * stage;               // 3, 6, 9
        SSRB_TRY(conv(x, p + "model." + std::to_string(ci) + ".conv.conv.", in, er[stage - 1], true, nullptr, &a));
        return resblock(x, p + "model." + std::to_string(ci + 1) + ".", a, out);
    }
    SSRB_TRY(conv(x, p + "model.12.conv.conv.", in, er[3], true, nullptr, &a));
    if (x.c->cfg.lstm_layers > 0) { SSRB_TRY(lstm(x, p + "model.13.", a, &b)); } else b = a;
    return conv(x, p + "model.15.conv.conv.", b, 1, true, nullptr, out);
}
static int encoder(Ctx& x, const std::string& p, Tensor in, Tensor* out) {
    Tensor cur = in, nxt;
    for (int st = 0; st < 5; st++) { SSRB_TRY(encoder_stage(x, p, st, cur, &nxt)); cur = nxt; }
    *out = cur;
    return 0;
}
// SEANetDecoder as the slices model[:4], [4:7], [7:10], [10:] (seanet.py:577-591)
static int decoder_stage(Ctx& x, const std::string& p, int stage, Tensor in, Tensor* out) {
    const int* r = x.c->cfg.ratios;
    Tensor a, b;
    if (stage == 0) {
        SSRB_TRY(conv(x, p + "model.0.conv.conv.", in, 1, false, nullptr, &a));
        if (x.c->cfg.lstm_layers > 0) { SSRB_TRY(lstm(x, p + "model.1.", a, &b)); } else b = a;
        return convtr(x, p + "model.3.convtr.convtr.", b, r[0], true, out);
    }
    if (stage <= 2) {
        const int ri = 3 * stage + 1;           // 4, 7
        SSRB_TRY(resblock(x, p + "model." + std::to_string(ri) + ".", in, &a));
        return convtr(x, p + "model." + std::to_string(ri + 2) + ".convtr.convtr.", a, r[stage], true, out);
    }
    SSRB_TRY(resblock(x, p + "model.10.", in, &a));
    SSRB_TRY(convtr(x, p + "model.12.convtr.convtr.", a, r[3], true, &b));
    SSRB_TRY(resblock(x, p + "model.13.", b, &a));
    return conv(x, p + "model.15.conv.conv.", a, 1, true, nullptr, out);
}
static int decoder(Ctx& x, const std::string& p, Tensor in, Tensor* out) {
    Tensor cur = in, nxt;
    for (int st = 0; st < 4; st++) { SSRB_TRY(decoder_stage(x, p, st, cur, &nxt)); cur = nxt; }
    *out = cur;
    return 0;
}


// =====================================================================================================================
// Tensor-core path (cfg.tensor_cores): decode / wmdecode convolutions as bf16 tap-GEMMs over channels-last activations
// (conv_tc.cu).  The frame-rate region (first decoder conv, LSTMs, wm_proj0, last encoder conv) stays on the fp32 kernels.
// =====================================================================================================================
static inline int pad64(int c) { return (c + 63) / 64 * 64; }
static inline unsigned short f2bf(float f) {          // round-to-nearest-even fp32 -> bf16 (host)
    unsigned int u; memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (unsigned short)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (unsigned short)(u >> 16);
}
static inline float eluf(float x) { return x > 0.f ? x : std::expm1(x); }

enum TcKind { TC_CONV = 0, TC_CONVS = 1, TC_CONVTR = 2, TC_WMPROJ = 3 };

// builds (once) the repacked operand of convolution `key`
static int get_tcw(ssrb_codec* c, const std::string& key, int kind, int stride, const TcW** out) {
    auto it = c->tcw.find(key);
    if (it != c->tcw.end()) { *out = &it->second; return 0; }
    const ConvW* W;
    SSRB_TRY(get_conv(c, key, &W));
    SSRB_CHECK(!W->hw.empty() && !W->hb.empty(), ("host weights missing for " + key).c_str());
    TcW t;
    std::vector<unsigned short> wp;
    std::vector<float> bias, bias_alt;
    const int k = W->k;
    if (kind == TC_CONV || kind == TC_WMPROJ) {             // W [Cout][Cin][k], stride 1
        const int Cout = W->d0, Cin_all = W->d1;
        const int E = kind == TC_WMPROJ ? c->cfg.dimension / 16 : 0;
        const int Cin = Cin_all - E;
        t.taps = k; t.N = pad64(Cout); t.Cw = pad64(Cin); t.bias_mod = t.N;
        wp.assign((size_t)t.taps * t.N * t.Cw, 0);
        for (int q = 0; q < k; q++)
            for (int n = 0; n < Cout; n++)
                for (int ci = 0; ci < Cin; ci++)
                    wp[((size_t)q * t.N + n) * t.Cw + ci] = f2bf(W->hw[((size_t)n * Cin_all + ci) * k + q]);
        bias.assign(t.N, 0.f);
        for (int n = 0; n < Cout; n++) bias[n] = W->hb[n];
        if (kind == TC_WMPROJ) {                            // cat([skip, emb]) -> ELU -> 1x1 conv: fold the embedding rows into 2 biases
            SSRB_CHECK((int)c->wm_embed_host.size() == 2 * E && k == 1, "wm_proj repack: bad wm_embed");
            bias_alt = bias;
            for (int n = 0; n < Cout; n++)
                for (int m = 0; m < 2; m++) {
                    float acc = 0.f;
                    for (int e = 0; e < E; e++) acc += W->hw[(size_t)n * Cin_all + Cin + e] * eluf(c->wm_embed_host[m * E + e]);
                    (m ? bias_alt : bias)[n] += acc;
                }
        }
    } else if (kind == TC_CONVS) {                          // W [Cout][Cin][2s]: view rows of s time steps
        const int Cout = W->d0, Cin = W->d1, s = stride;
        SSRB_CHECK(k == 2 * s && Cin % 64 == 0, "strided conv repack: unsupported shape");
        t.taps = 2; t.N = pad64(Cout); t.Cw = s * Cin; t.bias_mod = t.N;
        wp.assign((size_t)t.taps * t.N * t.Cw, 0);
        for (int q = 0; q < 2; q++)
            for (int n = 0; n < Cout; n++)
                for (int r = 0; r < s; r++)
                    for (int ci = 0; ci < Cin; ci++)
                        wp[((size_t)q * t.N + n) * t.Cw + r * Cin + ci] = f2bf(W->hw[((size_t)n * Cin + ci) * k + q * s + r]);
        bias.assign(t.N, 0.f);
        for (int n = 0; n < Cout; n++) bias[n] = W->hb[n];
    } else {                                                // TC_CONVTR: W [Cin][Cout][2s]; tap 0 = x[i-1] (k = p+s), tap 1 = x[i] (k = p)
        const int Cin = W->d0, Cout = W->d1, s = stride;
        SSRB_CHECK(k == 2 * s && Cin % 64 == 0 && Cout % 16 == 0, "convtr repack: unsupported shape");
        t.taps = 2; t.N = s * Cout; t.Cw = Cin; t.bias_mod = Cout;
        SSRB_CHECK(t.N % 64 == 0, "convtr repack: s*Cout must be a multiple of 64");
        wp.assign((size_t)t.taps * t.N * t.Cw, 0);
        for (int p = 0; p < s; p++)
            for (int co = 0; co < Cout; co++)
                for (int ci = 0; ci < Cin; ci++) {
                    wp[((size_t)0 * t.N + p * Cout + co) * t.Cw + ci] = f2bf(W->hw[((size_t)ci * Cout + co) * k + p + s]);
                    wp[((size_t)1 * t.N + p * Cout + co) * t.Cw + ci] = f2bf(W->hw[((size_t)ci * Cout + co) * k + p]);
                }
        bias.assign(W->hb.begin(), W->hb.begin() + Cout);
    }
    SSRB_TRY(dalloc((void**)&t.w, wp.size() * 2));
    SSRB_CUDA(cudaMemcpy(t.w, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice));
    SSRB_TRY(upload_new(&t.bias, bias.data(), bias.size()));
    if (!bias_alt.empty()) SSRB_TRY(upload_new(&t.bias_alt, bias_alt.data(), bias_alt.size()));
    c->tcw[key] = t;
    *out = &c->tcw[key];
    return 0;
}

static ClT cl_alloc(Ctx& x, int C, int T, bool raw, bool act) {
    ClT t; t.C = C; t.T = T;
    const size_t n = (size_t)x.B * (T + 2 * CL_GUARD) * C;           // bf16 elements = n*2 bytes = n/2 floats
    if (raw) t.raw = (bf16*)x.c->arena.f((n + 1) / 2);
    if (act) t.act = (bf16*)x.c->arena.f((n + 1) / 2);
    if (!x.c->arena.dry) {                                            // only the guard rows need zeroing: producers write every row
        if (raw) launch_cl_zero_guards(t.raw, x.B, T, C, x.s);
        if (act) launch_cl_zero_guards(t.act, x.B, T, C, x.s);
    }
    return t;
}

// generic tensor-core convolution on a channels-last input (its ELU'd copy is the operand)
static int tc_conv(Ctx& x, const std::string& key, int kind, int stride, const ClT& in, const bf16* in_buf, const bf16* residual,
                   bool want_raw, bool want_act, const long long* marks, int marks_T, int marks_rep, ClT* out) {
    const TcW* W;
    SSRB_TRY(get_tcw(x.c, key, kind, stride, &W));
    const int G = CL_GUARD, C = in.C, T = in.T;
    ConvTcArgs a;
    a.x = in_buf; a.B = x.B; a.x_bstride = (long long)(T + 2 * G) * C; a.w = W->w; a.taps = W->taps; a.N = W->N; a.Cw = W->Cw;
    a.bias = W->bias; a.bias_alt = W->bias_alt; a.bias_mod = W->bias_mod; a.marks = marks; a.marks_T = marks_T; a.marks_rep = marks_rep;
    int Tout, Cout;
    if (kind == TC_CONV || kind == TC_WMPROJ) {
        SSRB_CHECK(W->Cw == C, ("tc conv input width mismatch at " + key).c_str());
        const int k = W->taps, total = k - 1, padL = total - total / 2;
        Tout = T; Cout = W->N;
        a.x_base_off = (long long)(G - padL) * C; a.rows_v = T + G + padL; a.T_rows = T;
        a.out_off = (long long)G * Cout; a.valid_lo = 0; a.valid_hi = (long long)T * Cout;
    } else if (kind == TC_CONVS) {
        SSRB_CHECK(W->Cw == stride * C && T % stride == 0, ("tc strided conv shape mismatch at " + key).c_str());
        const int padL = stride - stride / 2;
        Tout = T / stride; Cout = W->N;
        a.x_base_off = (long long)(G - padL) * C;
        a.rows_v = (int)((a.x_bstride - a.x_base_off) / W->Cw); a.T_rows = Tout;
        a.out_off = (long long)G * Cout; a.valid_lo = 0; a.valid_hi = (long long)Tout * Cout;
    } else {
        SSRB_CHECK(W->Cw == C, ("tc convtr input width mismatch at " + key).c_str());
        const int padL = stride - stride / 2;
        Tout = T * stride; Cout = W->bias_mod;
        a.x_base_off = (long long)(G - 1) * C; a.rows_v = T + G + 1; a.T_rows = T + 1;
        a.out_off = (long long)(G - padL) * Cout; a.valid_lo = (long long)padL * Cout; a.valid_hi = a.valid_lo + (long long)Tout * Cout;
    }
    *out = cl_alloc(x, Cout, Tout, want_raw, want_act);
    a.out_raw = out->raw; a.out_act = out->act; a.out_bstride = (long long)(Tout + 2 * G) * Cout;
    if (residual) { a.res = residual; a.res_bstride = a.out_bstride; a.res_off = (long long)G * Cout; }
    if (x.c->arena.dry) return 0;
    return conv_tc(a, x.s);
}
static int conv_frame_tc(Ctx& x, const std::string& key, Tensor in, bool elu_in, const float* res, Tensor* out) {
    const ConvW* W;
    SSRB_TRY(get_conv(x.c, key, &W));
    if (W->d0 % 64 != 0 || W->d1 % 64 != 0 || res) return conv(x, key, in, 1, elu_in, res, out);     // not tensor-core shaped
    ClT a = cl_alloc(x, in.C, in.T, false, true), o;
    if (!x.c->arena.dry) SSRB_TRY(launch_cf32_to_cl(in.p, x.B, in.C, in.T, elu_in, a.act, x.s));
    SSRB_TRY(tc_conv(x, key, TC_CONV, 1, a, a.act, nullptr, true, false, nullptr, 0, 1, &o));
    out->C = o.C; out->T = o.T; out->p = x.c->arena.f((size_t)x.B * o.C * o.T);
    if (x.c->arena.dry) return 0;
    return launch_cl_to_cf32(o.raw, x.B, o.C, o.T, out->p, x.s);
}
// SEANetResnetBlock on the tensor cores: y = x + conv_k1(ELU(conv_k3(ELU(x)))); only ELU(y) (and optionally y) is kept
static int tc_resblock(Ctx& x, const std::string& prefix, const ClT& in, bool want_raw, ClT* out) {
    static const bool unfused = [] { const char* e = getenv("SSRB_RESBLOCK_UNFUSED"); return e && e[0] == '1'; }();   // A/B switch: two conv_tc launches
    if (!unfused && resblock_tc_supported(in.C) && in.raw && in.act) {
        // depth-fused: one launch, the hidden activation stays in shared memory (resblock_tc.cu)
        const TcW *W1, *W2;
        SSRB_TRY(get_tcw(x.c, prefix + "block.1.conv.conv.", TC_CONV, 1, &W1));
        SSRB_TRY(get_tcw(x.c, prefix + "block.3.conv.conv.", TC_CONV, 1, &W2));
        SSRB_CHECK(W1->taps == 3 && W2->taps == 1, "fused resblock: kernel sizes must be 3 and 1");
        const int G = CL_GUARD, C = in.C, T = in.T;
        *out = cl_alloc(x, C, T, want_raw, true);
        if (x.c->arena.dry) return 0;
        ResblockTcArgs a;
        a.B = x.B; a.T = T; a.C = C;
        a.x_act = in.act; a.x_bstride = (long long)(T + 2 * G) * C; a.x_base_off = (long long)(G - 1) * C; a.rows_v = T + G + 1;
        a.x_raw = in.raw; a.x_raw_off = (long long)G * C;
        a.w1 = W1->w; a.w1_N = W1->N; a.w1_Cw = W1->Cw; a.b1 = W1->bias;
        a.w2 = W2->w; a.w2_N = W2->N; a.w2_Cw = W2->Cw; a.b2 = W2->bias;
        a.out_raw = out->raw; a.out_act = out->act; a.out_bstride = a.x_bstride; a.out_off = (long long)G * C;
        return resblock_tc(a, x.s);
    }
    ClT h;
    SSRB_TRY(tc_conv(x, prefix + "block.1.conv.conv.", TC_CONV, 1, in, in.act, nullptr, false, true, nullptr, 0, 1, &h));
    return tc_conv(x, prefix + "block.3.conv.conv.", TC_CONV, 1, h, h.act, in.raw, want_raw, true, nullptr, 0, 1, out);
}

// decoder from the first transposed conv to the waveform; `frame` = ELU'able LSTM output [B,1024,Tf] fp32 channels-first.
// skips (wm path): ELU'd channels-last skip tensors for stages 1..3, else null.
static int decoder_tail_tc(Ctx& x, const std::string& p, Tensor frame, const ClT* skips, const long long* marks, int Tf, float* wav_out) {
    const int* r = x.c->cfg.ratios;
    ClT cur = cl_alloc(x, frame.C, frame.T, false, true);
    if (!x.c->arena.dry) SSRB_TRY(launch_cf32_to_cl(frame.p, x.B, frame.C, frame.T, true, cur.act, x.s));
    int rep = 1;
    for (int st = 0; st < 4; st++) {
        ClT up, y;
        const int tr_idx = 3 + 3 * st;                                  // model.3, 6, 9, 12
        SSRB_TRY(tc_conv(x, p + "model." + std::to_string(tr_idx) + ".convtr.convtr.", TC_CONVTR, r[st], cur, cur.act, nullptr, true, true,
                         nullptr, 0, 1, &up));
        rep *= r[st];
        if (skips && st < 3) {                                          // out = wm_proj_{st+1}(ELU(cat(skip, emb))) + x   (seanet.py:581-591)
            ClT o;
            SSRB_TRY(tc_conv(x, "wmdecoder.wm_proj" + std::to_string(st + 1) + ".1.conv.conv.", TC_WMPROJ, 1, skips[st], skips[st].act, up.raw,
                             true, true, marks, Tf, rep, &o));
            up = o;
        }
        SSRB_TRY(tc_resblock(x, p + "model." + std::to_string(tr_idx + 1) + ".", up, false, &y));
        cur = y;
    }
    const ConvW* Wl;
    SSRB_TRY(get_conv(x.c, p + "model.15.conv.conv.", &Wl));
    if (x.c->arena.dry) return 0;
    return launch_cl_last_conv(cur.act, x.B, cur.T, cur.C, Wl->w, Wl->b, Wl->k, wav_out, x.s);
}

// skip encoder of WMSEANetDecoder on the tensor cores: returns ELU'd skips for stages 1..3 (512@T/40, 256@T/8, 128@T/2) and the
// frame-rate skip3 (fp32 channels-first) for wm_proj0.
static int skip_encoder_tc(Ctx& x, const std::string& p, const float* wav, int T, ClT* skips /*[3]*/, Tensor* skip3) {
    const int* r = x.c->cfg.ratios;
    const int er[4] = {r[3], r[2], r[1], r[0]};
    const ConvW* W0;
    SSRB_TRY(get_conv(x.c, p + "model.0.conv.conv.", &W0));
    ClT z0 = cl_alloc(x, pad64(W0->d0), T, true, true), cur;
    if (!x.c->arena.dry) SSRB_TRY(launch_cl_first_conv(wav, x.B, T, W0->w, W0->b, W0->d0, W0->k, z0.raw, z0.act, x.s));
    SSRB_TRY(tc_resblock(x, p + "model.1.", z0, false, &cur));
    for (int st = 1; st <= 3; st++) {
        ClT d, y;
        const int ci = 3 * st;
        SSRB_TRY(tc_conv(x, p + "model." + std::to_string(ci) + ".conv.conv.", TC_CONVS, er[st - 1], cur, cur.act, nullptr, true, true, nullptr, 0, 1, &d));
        SSRB_TRY(tc_resblock(x, p + "model." + std::to_string(ci + 1) + ".", d, false, &y));
        skips[3 - st] = y;                                               // st=1 -> skips[2] (128@T/2) ... st=3 -> skips[0] (512@T/40)
        cur = y;
    }
    ClT d8;
    SSRB_TRY(tc_conv(x, p + "model.12.conv.conv.", TC_CONVS, er[3], cur, cur.act, nullptr, true, false, nullptr, 0, 1, &d8));
    Tensor a{x.c->arena.f((size_t)x.B * d8.C * d8.T), d8.C, d8.T}, b;
    if (!x.c->arena.dry) SSRB_TRY(launch_cl_to_cf32(d8.raw, x.B, d8.C, d8.T, a.p, x.s));
    if (x.c->cfg.lstm_layers > 0) { SSRB_TRY(lstm(x, p + "model.13.", a, &b, true)); } else b = a;
    return conv_frame_tc(x, p + "model.15.conv.conv.", b, true, nullptr, skip3);
}

// =====================================================================================================================
// Encoder on the tensor cores at fp32-grade accuracy (cfg.tensor_cores, conv_tc32.cu): every convolution between the first
// (1-channel) one and the LSTM as 3 x TF32 tap-GEMMs over channels-last fp32 activations.  The RVQ indices downstream must stay
// reproducible, so nothing here is bf16.
// =====================================================================================================================
static inline float host_rna_tf32(float x) {                 // cvt.rna.tf32.f32: nearest, ties away from zero, low 13 bits cleared
    unsigned int u; memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;
    u += 0x1000u; u &= 0xffffe000u;
    float r; memcpy(&r, &u, 4);
    return r;
}
static int get_tc32w(ssrb_codec* c, const std::string& key, int kind, int stride, const Tc32W** out) {
    auto it = c->tc32w.find(key);
    if (it != c->tc32w.end()) { *out = &it->second; return 0; }
    const ConvW* W;
    SSRB_TRY(get_conv(c, key, &W));
    SSRB_CHECK(!W->hw.empty() && !W->hb.empty(), ("host weights missing for " + key).c_str());
    Tc32W t;
    std::vector<float> wp;
    const int k = W->k, Cout = W->d0, Cin = W->d1;
    SSRB_CHECK(Cout % 32 == 0, "tc32 repack: output channels must be a multiple of 32");
    if (kind == TC_CONV) {                                   // W [Cout][Cin][k], stride 1: W'[q][n][ci]
        SSRB_CHECK(Cin % 32 == 0, "tc32 repack: input channels must be a multiple of 32");
        t.taps = k; t.N = Cout; t.Cw = Cin;
        wp.assign((size_t)t.taps * t.N * t.Cw, 0.f);
        for (int q = 0; q < k; q++)
            for (int n = 0; n < Cout; n++)
                for (int ci = 0; ci < Cin; ci++) wp[((size_t)q * t.N + n) * t.Cw + ci] = W->hw[((size_t)n * Cin + ci) * k + q];
    } else {                                                 // TC_CONVS: W [Cout][Cin][2s] -> W'[q][n][r * Cin + ci] = W[n][ci][q * s + r]
        const int s = stride;
        SSRB_CHECK(k == 2 * s && (s * Cin) % 32 == 0, "tc32 strided conv repack: unsupported shape");
        t.taps = 2; t.N = Cout; t.Cw = s * Cin;
        wp.assign((size_t)t.taps * t.N * t.Cw, 0.f);
        for (int q = 0; q < 2; q++)
            for (int n = 0; n < Cout; n++)
                for (int r = 0; r < s; r++)
                    for (int ci = 0; ci < Cin; ci++) wp[((size_t)q * t.N + n) * t.Cw + r * Cin + ci] = W->hw[((size_t)n * Cin + ci) * k + q * s + r];
    }
    std::vector<float> hi(wp.size()), lo(wp.size());
    for (size_t i = 0; i < wp.size(); i++) { hi[i] = host_rna_tf32(wp[i]); lo[i] = host_rna_tf32(wp[i] - hi[i]); }
    SSRB_TRY(upload_new(&t.wh, hi.data(), hi.size()));
    SSRB_TRY(upload_new(&t.wl, lo.data(), lo.size()));
    SSRB_TRY(upload_new(&t.bias, W->hb.data(), W->hb.size()));
    c->tc32w[key] = t;
    *out = &c->tc32w[key];
    return 0;
}
static Cl32 cl32_alloc(Ctx& x, int C, int T, bool raw, bool act) {
    Cl32 t; t.C = C; t.T = T;
    const size_t n = (size_t)x.B * (T + 2 * CL_GUARD) * C;
    if (raw) t.raw = x.c->arena.f(n);
    if (act) { t.hi = x.c->arena.f(n); t.lo = x.c->arena.f(n); }
    if (!x.c->arena.dry && act) {                            // only operands are read through the guards (zero padding, conv.py:185-201)
        launch_cl32_zero_guards(t.hi, x.B, T, C, x.s);
        launch_cl32_zero_guards(t.lo, x.B, T, C, x.s);
    }
    return t;
}
static int tc32_conv(Ctx& x, const std::string& key, int kind, int stride, const Cl32& in, const float* residual, bool want_raw, bool want_act,
                     Cl32* out) {
    const Tc32W* W;
    SSRB_TRY(get_tc32w(x.c, key, kind, stride, &W));
    const int G = CL_GUARD, C = in.C, T = in.T;
    ConvTc32Args a;
    a.x_hi = in.hi; a.x_lo = in.lo; a.B = x.B; a.x_bstride = (long long)(T + 2 * G) * C; a.w_hi = W->wh; a.w_lo = W->wl;
    a.taps = W->taps; a.N = W->N; a.Cw = W->Cw; a.bias = W->bias;
    int Tout;
    if (kind == TC_CONV) {
        SSRB_CHECK(W->Cw == C, ("tc32 conv input width mismatch at " + key).c_str());
        const int total = W->taps - 1, padL = total - total / 2;
        Tout = T;
        a.x_base_off = (long long)(G - padL) * C; a.rows_v = T + G + padL; a.T_rows = T;
    } else {
        SSRB_CHECK(W->Cw == stride * C && T % stride == 0, ("tc32 strided conv shape mismatch at " + key).c_str());
        const int padL = stride - stride / 2;
        Tout = T / stride;
        a.x_base_off = (long long)(G - padL) * C;
        a.rows_v = (int)((a.x_bstride - a.x_base_off) / W->Cw); a.T_rows = Tout;
    }
    *out = cl32_alloc(x, W->N, Tout, want_raw, want_act);
    a.out_raw = out->raw; a.out_hi = out->hi; a.out_lo = out->lo;
    a.out_bstride = (long long)(Tout + 2 * G) * W->N; a.out_off = (long long)G * W->N;
    if (residual) { a.res = residual; a.res_bstride = a.out_bstride; a.res_off = a.out_off; }
    if (x.c->arena.dry) return 0;
    return conv_tc32(a, x.s);
}
// SEANetResnetBlock: y = x + conv_k1(ELU(conv_k3(ELU(x)))); the next layer consumes ELU(y) only
static int tc32_resblock(Ctx& x, const std::string& prefix, const Cl32& in, Cl32* out) {
    Cl32 h;
    SSRB_TRY(tc32_conv(x, prefix + "block.1.conv.conv.", TC_CONV, 1, in, nullptr, false, true, &h));
    return tc32_conv(x, prefix + "block.3.conv.conv.", TC_CONV, 1, h, in.raw, false, true, out);
}
static bool encoder_tc32_supported(ssrb_codec* c, int T) {
    if (!c->use_tc || c->cfg.n_filters % 64 != 0 || c->cfg.n_filters > 64 || c->cfg.residual_kernel_size != 3 || c->cfg.compress != 2) return false;
    if (T % c->hop != 0) return false;
    static const bool off = [] { const char* e = getenv("SSRB_ENC_FP32"); return e && e[0] == '1'; }();   // SSRB_ENC_FP32=1: keep the CUDA-core kernels
    return !off;
}
// SEANetEncoder (seanet.py:63-153) -> latents [B, dimension, T / hop] fp32 channels-first
static int encoder_tc32(Ctx& x, const std::string& p, const float* wav, int T, Tensor* out) {
    const int* r = x.c->cfg.ratios;
    const int er[4] = {r[3], r[2], r[1], r[0]};
    const ConvW* W0;
    SSRB_TRY(get_conv(x.c, p + "model.0.conv.conv.", &W0));
    Cl32 z0 = cl32_alloc(x, W0->d0, T, true, true), cur;
    if (!x.c->arena.dry) SSRB_TRY(launch_cl32_first_conv(wav, x.B, T, W0->w, W0->b, W0->d0, W0->k, z0.raw, z0.hi, z0.lo, x.s));
    SSRB_TRY(tc32_resblock(x, p + "model.1.", z0, &cur));
    for (int st = 1; st <= 3; st++) {
        Cl32 d, y;
        const int ci = 3 * st;
        SSRB_TRY(tc32_conv(x, p + "model." + std::to_string(ci) + ".conv.conv.", TC_CONVS, er[st - 1], cur, nullptr, true, true, &d));
        SSRB_TRY(tc32_resblock(x, p + "model." + std::to_string(ci + 1) + ".", d, &y));
        cur = y;
    }
    Cl32 d8;
    SSRB_TRY(tc32_conv(x, p + "model.12.conv.conv.", TC_CONVS, er[3], cur, nullptr, true, false, &d8));
    Tensor a{x.c->arena.f((size_t)x.B * d8.C * d8.T), d8.C, d8.T}, b;
    if (!x.c->arena.dry) SSRB_TRY(launch_cl32_to_cf32(d8.raw, x.B, d8.C, d8.T, a.p, x.s));
    if (x.c->cfg.lstm_layers > 0) { SSRB_TRY(lstm(x, p + "model.13.", a, &b, false, true)); } else b = a;
    return conv(x, p + "model.15.conv.conv.", b, 1, true, nullptr, out);
}

// runs `fn` twice: a dry pass to size the arena, then for real
template <typename F>
static int run_planned(ssrb_codec* c, F fn) {
    c->arena.dry = true; c->arena.off = 0;
    SSRB_TRY(fn());
    const size_t need = c->arena.off + 256;
    if (need > c->arena.cap) {
        if (c->arena.base) { SSRB_CUDA(cudaDeviceSynchronize()); cudaFree(c->arena.base); }
        c->arena.base = nullptr; c->arena.cap = 0;
        SSRB_TRY(dalloc((void**)&c->arena.base, need));
        c->arena.cap = need;
    }
    c->arena.dry = false; c->arena.off = 0;
    return fn();
}

int ssrb_codec_check_loaded(ssrb_codec* c) {
    SSRB_CHECK(c, "null argument");
    if (c->derived_stale) drop_derived(c);
    if (!c->pending.empty()) { set_error("weight_g/weight_v pair incomplete for " + c->pending.begin()->first); return 1; }
    for (int q = 0; q < c->cfg.n_q; q++) if (!c->has_cb[q]) { set_error("codebook missing: " + std::to_string(q)); return 1; }
    return 0;
}

int ssrb_codec_quantize(ssrb_codec* c, const float* emb, int B, int Tf, int64_t* codes, void* stream) {
    SSRB_CHECK(c && emb && codes, "null argument");
    SSRB_CUDA(cudaSetDevice(c->device));
    SSRB_TRY(ssrb_codec_check_loaded(c));
    cudaStream_t s = (cudaStream_t)stream;
    return run_planned(c, [&]() -> int {
        float* ws = c->arena.f((size_t)B * Tf * c->cfg.dimension);
        if (c->arena.dry) return 0;
        return launch_rvq_encode(emb, B, c->cfg.dimension, Tf, c->codebooks, c->cb_sq, c->cfg.n_q, c->cfg.bins, ws,
                                 (long long*)codes, s);
    });
}

int ssrb_codec_encode(ssrb_codec* c, const float* wav, int B, int T, int64_t* codes, float* emb_out, void* stream) {
    SSRB_CHECK(c && wav && codes, "null argument");
    SSRB_CHECK(T > 0, "encode: empty waveform");
    SSRB_CUDA(cudaSetDevice(c->device));
    SSRB_TRY(ssrb_codec_check_loaded(c));
    cudaStream_t s = (cudaStream_t)stream;
    // frames = ceil chain of the strided convs (StreamableConv1d pads the tail: conv.py:47-53); = T/hop when hop | T
    int Tf = T;
    for (int i = c->cfg.n_ratios - 1; i >= 0; i--) Tf = (Tf + c->cfg.ratios[i] - 1) / c->cfg.ratios[i];
    const int Dm = c->cfg.dimension, nq = c->cfg.n_q;
    for (int b0 = 0; b0 < B; b0 += c->cfg.max_batch_chunk) {
        const int nb = std::min(c->cfg.max_batch_chunk, B - b0);
        SSRB_TRY(run_planned(c, [&]() -> int {
            Ctx x{c, s, nb};
            Tensor in{const_cast<float*>(wav) + (size_t)b0 * T, 1, T}, emb;
            if (encoder_tc32_supported(c, T)) { SSRB_TRY(encoder_tc32(x, "encoder.", in.p, T, &emb)); }
            else SSRB_TRY(encoder(x, "encoder.", in, &emb));
            float* ws = c->arena.f((size_t)nb * Tf * Dm);
            if (c->arena.dry) return 0;
            SSRB_CHECK(emb.C == Dm && emb.T == Tf, "encoder output shape mismatch");
            if (emb_out) SSRB_CUDA(cudaMemcpyAsync(emb_out + (size_t)b0 * Dm * Tf, emb.p, (size_t)nb * Dm * Tf * 4, cudaMemcpyDeviceToDevice, s));
            return launch_rvq_encode(emb.p, nb, Dm, Tf, c->codebooks, c->cb_sq, nq, c->cfg.bins, ws,
                                     (long long*)codes + (size_t)b0 * nq * Tf, s);
        }));
    }
    return 0;
}

int ssrb_codec_decode(ssrb_codec* c, const int64_t* codes, int B, int Tf, float* wav, void* stream) {
    SSRB_CHECK(c && codes && wav && Tf > 0, "null argument");
    SSRB_CUDA(cudaSetDevice(c->device));
    SSRB_TRY(ssrb_codec_check_loaded(c));
    cudaStream_t s = (cudaStream_t)stream;
    const int Dm = c->cfg.dimension, nq = c->cfg.n_q, T = Tf * c->hop;
    for (int b0 = 0; b0 < B; b0 += c->cfg.max_batch_chunk) {
        const int nb = std::min(c->cfg.max_batch_chunk, B - b0);
        SSRB_TRY(run_planned(c, [&]() -> int {
            Ctx x{c, s, nb};
            Tensor z{c->arena.f((size_t)nb * Dm * Tf), Dm, Tf}, out;
            if (!c->arena.dry)
                SSRB_TRY(launch_rvq_decode((const long long*)codes + (size_t)b0 * nq * Tf, nb, nq, Tf, c->codebooks, c->cfg.bins, Dm, z.p, s));
            if (c->use_tc) {
                Tensor a, b;
                SSRB_TRY(conv_frame_tc(x, "decoder.model.0.conv.conv.", z, false, nullptr, &a));
                if (c->cfg.lstm_layers > 0) { SSRB_TRY(lstm(x, "decoder.model.1.", a, &b, true)); } else b = a;
                return decoder_tail_tc(x, "decoder.", b, nullptr, nullptr, Tf, wav + (size_t)b0 * T);
            }
            SSRB_TRY(decoder(x, "decoder.", z, &out));
            if (c->arena.dry) return 0;
            SSRB_CHECK(out.C == 1 && out.T == T, "decoder output shape mismatch");
            SSRB_CUDA(cudaMemcpyAsync(wav + (size_t)b0 * T, out.p, (size_t)nb * T * 4, cudaMemcpyDeviceToDevice, s));
            return 0;
        }));
    }
    return 0;
}

int ssrb_codec_wmdecode(ssrb_codec* c, const int64_t* codes, const int64_t* marks, const float* wav_in, int B, int Tf,
                        float* wav_out, float* mark_logits, void* stream) {
    SSRB_CHECK(c && codes && marks && wav_in && wav_out && Tf > 0, "null argument");
    SSRB_CUDA(cudaSetDevice(c->device));
    SSRB_TRY(ssrb_codec_check_loaded(c));
    SSRB_CHECK(c->has_wm_embed, "wm_embed missing");
    cudaStream_t s = (cudaStream_t)stream;
    const int Dm = c->cfg.dimension, nq = c->cfg.n_q, T = Tf * c->hop, E = Dm / 16;
    const int* r = c->cfg.ratios;
    for (int b0 = 0; b0 < B; b0 += c->cfg.max_batch_chunk) {
        const int nb = std::min(c->cfg.max_batch_chunk, B - b0);
        SSRB_TRY(run_planned(c, [&]() -> int {
            Ctx x{c, s, nb};
            Arena& A = c->arena;
            const long long* mk = (const long long*)marks + (size_t)b0 * Tf;
            Tensor lat{A.f((size_t)nb * Dm * Tf), Dm, Tf};
            if (!A.dry)
                SSRB_TRY(launch_rvq_decode((const long long*)codes + (size_t)b0 * nq * Tf, nb, nq, Tf, c->codebooks, c->cfg.bins, Dm, lat.p, s));
            if (c->use_tc && !mark_logits) {
                ClT sk[3];
                Tensor skip3, cat{nullptr, Dm + E, Tf}, o0, a, b;
                SSRB_TRY(skip_encoder_tc(x, "wmdecoder.skip_encoder.", wav_in + (size_t)b0 * T, T, sk, &skip3));
                cat.p = A.f((size_t)nb * (Dm + E) * Tf);
                if (!A.dry) SSRB_TRY(launch_concat_marks(skip3.p, nb, Dm, Tf, mk, Tf, 1, c->wm_embed, E, cat.p, s));
                SSRB_TRY(conv(x, "wmdecoder.wm_proj0.1.conv.conv.", cat, 1, true, lat.p, &o0));      // + x (seanet.py:577-578)
                SSRB_TRY(conv_frame_tc(x, "wmdecoder.model.0.conv.conv.", o0, false, nullptr, &a));
                if (c->cfg.lstm_layers > 0) { SSRB_TRY(lstm(x, "wmdecoder.model.1.", a, &b, true)); } else b = a;
                return decoder_tail_tc(x, "wmdecoder.", b, sk, mk, Tf, wav_out + (size_t)b0 * T);
            }
            // skip encoder over the (partly zeroed) original waveform (seanet.py:559-574)
            Tensor z{const_cast<float*>(wav_in) + (size_t)b0 * T, 1, T}, nx, skips[4];
            SSRB_TRY(encoder_stage(x, "wmdecoder.skip_encoder.", 0, z, &nx)); z = nx;
            for (int st = 1; st <= 4; st++) { SSRB_TRY(encoder_stage(x, "wmdecoder.skip_encoder.", st, z, &nx)); z = nx; skips[st - 1] = z; }
            const int reps[4] = {1, r[0], r[0] * r[1], r[0] * r[1] * r[2]};     // label repeats, popped last-first
            Tensor cur = lat;
            for (int i = 0; i < 4; i++) {
                const Tensor sk = skips[3 - i];
                SSRB_CHECK(A.dry || (sk.C == cur.C && sk.T == cur.T), "skip / decoder shape mismatch");
                Tensor cat{A.f((size_t)nb * (sk.C + E) * sk.T), sk.C + E, sk.T}, pr, st;
                if (!A.dry) SSRB_TRY(launch_concat_marks(sk.p, nb, sk.C, sk.T, mk, Tf, reps[i], c->wm_embed, E, cat.p, s));
                SSRB_TRY(conv(x, "wmdecoder.wm_proj" + std::to_string(i) + ".1.conv.conv.", cat, 1, true, cur.p, &pr));
                SSRB_TRY(decoder_stage(x, "wmdecoder.", i, pr, &st));
                cur = st;
            }
            Tensor m, mp;
            if (mark_logits) {
                SSRB_TRY(encoder(x, "wmdecoder.wm_encoder.", cur, &m));
                SSRB_TRY(conv(x, "wmdecoder.wm_predictor.1.conv.conv.", m, 1, true, nullptr, &mp));
            }
            if (A.dry) return 0;
            SSRB_CHECK(cur.C == 1 && cur.T == T, "wmdecoder output shape mismatch");
            SSRB_CUDA(cudaMemcpyAsync(wav_out + (size_t)b0 * T, cur.p, (size_t)nb * T * 4, cudaMemcpyDeviceToDevice, s));
            if (mark_logits) SSRB_TRY(launch_bct_to_btc(mp.p, nb, 2, Tf, mark_logits + (size_t)b0 * Tf * 2, s));
            return 0;
        }));
    }
    return 0;
}

// WMEncodecModel.detect_watermark (wmencodec.py:377-382): m = wm_predictor(wm_encoder(x)) -> logits [B,Tf,2] (fp32 path).
int ssrb_codec_detect_watermark(ssrb_codec* c, const float* wav, int B, int T, float* mark_logits, void* stream) {
    SSRB_CHECK(c && wav && mark_logits, "null argument");
    SSRB_CHECK(T > 0 && T % c->hop == 0, "detect_watermark: T must be a positive multiple of the hop length");
    SSRB_CUDA(cudaSetDevice(c->device));
    SSRB_TRY(ssrb_codec_check_loaded(c));
    cudaStream_t s = (cudaStream_t)stream;
    const int Tf = T / c->hop;
    for (int b0 = 0; b0 < B; b0 += c->cfg.max_batch_chunk) {
        const int nb = std::min(c->cfg.max_batch_chunk, B - b0);
        SSRB_TRY(run_planned(c, [&]() -> int {
            Ctx x{c, s, nb};
            Tensor in{const_cast<float*>(wav) + (size_t)b0 * T, 1, T}, m, mp;
            SSRB_TRY(encoder(x, "wmdecoder.wm_encoder.", in, &m));
            SSRB_TRY(conv(x, "wmdecoder.wm_predictor.1.conv.conv.", m, 1, true, nullptr, &mp));
            if (c->arena.dry) return 0;
            return launch_bct_to_btc(mp.p, nb, 2, Tf, mark_logits + (size_t)b0 * Tf * 2, s);
        }));
    }
    return 0;
}
