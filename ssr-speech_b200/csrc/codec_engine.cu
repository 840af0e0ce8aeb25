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
#include "../../include/ssr_b200.h"

using namespace ssrb;

namespace {

struct ConvW { float* w = nullptr; float* b = nullptr; int d0 = 0, d1 = 0, k = 0; bool has_w = false, has_b = false; };
struct LstmW {
    float *wih[4] = {}, *whh[4] = {}, *bsum[4] = {};
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
};

static int dalloc(void** p, size_t bytes) { SSRB_CUDA(cudaMalloc(p, bytes ? bytes : 16)); return 0; }

int ssrb_codec_create(const ssrb_codec_config* c, int device, ssrb_codec** out) {
    SSRB_CHECK(c && out, "null argument");
    SSRB_CHECK(c->n_ratios == 4, "codec: exactly 4 ratios supported (WMSEANetDecoder.forward hard-codes 4 skips)");
    SSRB_CHECK(c->channels == 1, "codec: mono only");
    SSRB_CHECK(c->lstm_layers >= 0 && c->lstm_layers <= 4, "codec: lstm_layers out of range");
    SSRB_CHECK(c->n_q <= 16 && c->dimension <= 128 && c->dimension % 16 == 0, "codec: unsupported RVQ geometry");
    SSRB_CUDA(cudaSetDevice(device));
    ssrb_codec* cd = new ssrb_codec();
    cd->cfg = *c; cd->device = device;
    if (cd->cfg.max_batch_chunk <= 0 || cd->cfg.max_batch_chunk > 32) cd->cfg.max_batch_chunk = 8;
    cd->hop = 1;
    for (int i = 0; i < c->n_ratios; i++) cd->hop *= c->ratios[i];
    SSRB_TRY(dalloc((void**)&cd->codebooks, (size_t)c->n_q * c->bins * c->dimension * 4));
    SSRB_TRY(dalloc((void**)&cd->cb_sq, (size_t)c->n_q * c->bins * 4));
    SSRB_TRY(dalloc((void**)&cd->wm_embed, 2 * (c->dimension / 16) * 4));
    SSRB_TRY(dalloc((void**)&cd->bar, 4));
    *out = cd;
    return 0;
}

void ssrb_codec_destroy(ssrb_codec* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto& kv : c->convs) { cudaFree(kv.second.w); cudaFree(kv.second.b); }
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
        if (!L.bih[l].empty() && !L.bhh[l].empty() && !L.has_b[l]) {
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
    if (leaf == "bias") { SSRB_TRY(upload_new(&W.b, host, n)); W.has_b = true; return 0; }
    if (leaf == "weight") {
        SSRB_CHECK(ndim == 3, "conv weight must be 3-D");
        W.d0 = (int)shape[0]; W.d1 = (int)shape[1]; W.k = (int)shape[2];
        SSRB_TRY(upload_new(&W.w, host, n)); W.has_w = true; return 0;
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
            SSRB_TRY(upload_new(&W.w, w.data(), w.size())); W.has_w = true;
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
    return launch_conv1d(in.p, x.B, in.C, in.T, W->w, W->b, W->d0, k, stride, left, Tout, elu_in, res, out->p, x.s);
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
// StreamableLSTM (lstm.py:10-25)
static int lstm(Ctx& x, const std::string& prefix, Tensor in, Tensor* out) {
    auto it = x.c->lstms.find(prefix);
    if (it == x.c->lstms.end()) { set_error("codec lstm missing: " + prefix); return 1; }
    LstmW& L = it->second;
    const int C = in.C, T = in.T, B = x.B, nl = x.c->cfg.lstm_layers;
    SSRB_CHECK(L.C == C, "lstm width mismatch");
    Arena& A = x.c->arena;
    float* seq = A.f((size_t)T * B * C);
    float* pre = A.f((size_t)T * B * 4 * C);
    float* hs[2] = {A.f((size_t)T * B * C), A.f((size_t)T * B * C)};
    float* hbuf = A.f((size_t)2 * B * C);
    out->C = C; out->T = T; out->p = A.f((size_t)B * C * T);
    if (A.dry) return 0;
    SSRB_TRY(launch_bct_to_tbc(in.p, B, C, T, seq, x.s));
    const float* cur = seq;
    for (int l = 0; l < nl; l++) {
        SSRB_CHECK(L.has_wih[l] && L.has_whh[l] && L.has_b[l], "lstm layer weights missing");
        GemmArgs g;
        g.A = cur; g.lda = C; g.W = L.wih[l]; g.ldw = C; g.bias = L.bsum[l]; g.C = pre; g.ldc = 4 * C;
        g.M = T * B; g.N = 4 * C; g.K = C; g.ab_dtype = SSRB_DTYPE_F32; g.c_dtype = SSRB_DTYPE_F32;
        SSRB_TRY(gemm_simt(g, x.s));
        SSRB_TRY(launch_lstm_layer(pre, L.whh[l], hs[l & 1], hbuf, x.c->bar, T, B, C, x.s));
        cur = hs[l & 1];
    }
    return launch_tbc_to_bct_add(cur, in.p, B, C, T, out->p, x.s);
}

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
    SSRB_CHECK(T > 0 && T % c->hop == 0, "encode: T must be a positive multiple of the hop length");
    SSRB_CUDA(cudaSetDevice(c->device));
    SSRB_TRY(ssrb_codec_check_loaded(c));
    cudaStream_t s = (cudaStream_t)stream;
    const int Tf = T / c->hop, Dm = c->cfg.dimension, nq = c->cfg.n_q;
    for (int b0 = 0; b0 < B; b0 += c->cfg.max_batch_chunk) {
        const int nb = std::min(c->cfg.max_batch_chunk, B - b0);
        SSRB_TRY(run_planned(c, [&]() -> int {
            Ctx x{c, s, nb};
            Tensor in{const_cast<float*>(wav) + (size_t)b0 * T, 1, T}, emb;
            SSRB_TRY(encoder(x, "encoder.", in, &emb));
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
