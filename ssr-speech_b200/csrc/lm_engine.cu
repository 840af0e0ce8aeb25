// lm_engine.cu — the SSR-Speech decoder engine behind the C ABI (include/ssr_b200.h).
//
// Owns the packed weights, the in-place KV cache [layer][K|V][row][head][slot][128], the fp32 residual
// stream and the per-utterance decode state.  One decode iteration of the reference's `while True`
// loop (models/ssr.py:671-771) is enqueued as
//     embed -> 16 x (LN1, QKV GEMM, attention (+ in-place KV append), out-proj GEMM(+res), LN2, FFN1 GEMM(ReLU),
//     FFN2 GEMM(+res)) -> final LN -> head GEMM(GELU) -> head GEMM -> sample/state kernel
// and replayed as a CUDA graph, with no host synchronisation inside the loop.
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "lm_kernels.cuh"

namespace ssrb {

static thread_local std::string g_err;
void set_error(const std::string& m) { g_err = m; }
unsigned long long g_launch_count = 0;
bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("SSRB_NO_PDL"); return !(e && e[0] == '1'); }();
    return on;
}

// SSRB_ATTN_PREFETCH=0: the decode attention waits for the QKV GEMM before it starts streaming the cache
// SSRB_LN_FOLD=0: keep the separate LayerNorm kernels in the decode chain (33 more launches per iteration)
static bool ln_fold_enabled() {
    static const bool on = [] { const char* e = getenv("SSRB_LN_FOLD"); return !(e && e[0] == '0'); }();
    return on;
}
// SSRB_LAYER_KERNEL=1 (experimental, default off): out-proj -> FFN1 -> FFN2 -> next QKV as ONE persistent launch per layer
// (gemm_layer.cu) instead of four
static bool layer_kernel_enabled() {
    static const bool on = [] { const char* e = getenv("SSRB_LAYER_KERNEL"); return e && e[0] == '1'; }();
    return on;
}
// SSRB_MEGA=1 (experimental, default off): batches of <= 16 rows decode through the persistent whole-iteration kernel (lm_mega.cu)
// instead of the per-GEMM chain.  Read at every engine creation (tests switch it inside one process).
static bool mega_enabled() {
    const char* e = getenv("SSRB_MEGA");
    return e && e[0] == '1';
}
static bool attn_prefetch_enabled() {
    static const bool on = [] { const char* e = getenv("SSRB_ATTN_PREFETCH"); return !(e && e[0] == '0'); }();
    return on;
}

template <typename T>
__global__ void convert_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = from_f32<T>(src[i]);
}

struct LayerW {
    void *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;
    float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr;
    float *ln1w = nullptr, *ln1b = nullptr, *ln2w = nullptr, *ln2b = nullptr;
    // LayerNorm folded into the decode GEMMs: gamma-scaled copies of the matrices that consume a LayerNorm, their column
    // sums and beta-shifted biases (fold_ln_kernel)
    void *wqkv_f = nullptr, *w1_f = nullptr;
    float *cqkv = nullptr, *bqkv_f = nullptr, *c1 = nullptr, *b1_f = nullptr;
};

}  // namespace ssrb

using namespace ssrb;

struct ssrb_lm {
    ssrb_lm_config cfg;
    int device = 0;
    int D, H, L, F, K, V, Vt, Hh;
    int wdt;                 // weight / activation-operand / KV dtype
    size_t esz;              // its element size
    std::vector<LayerW> layers;
    float *text_emb = nullptr, *audio_emb = nullptr, *pe = nullptr, *lnfw = nullptr, *lnfb = nullptr;
    int n_pos = 0;
    float alpha_t = 1.f, alpha_a = 1.f;
    void *hw1 = nullptr, *hw2 = nullptr;       // [K*Hh, D], [K][V, Hh]
    float *hb1 = nullptr, *hb2 = nullptr;
    void* hw1_f = nullptr; float *hc1 = nullptr, *hb1_f = nullptr;   // final LayerNorm folded into the first head layer
    bool fold_ok = false;      // the configuration supports the folded chain (bf16 tensor-core GEMMs, d_model <= 2048)
    bool fold_dirty = true;    // weights changed since the folded copies were built
    bool fold = false;         // the open batch decodes through the folded chain (R <= 128)
    bool layer_kernel = false; // ... with the persistent per-layer GEMM kernel (gemm_layer.cu)
    unsigned int* gbar = nullptr;   // its grid barrier {count, generation}
    // persistent whole-iteration kernel for R <= 16 rows (lm_mega.cu): weights re-packed into streaming order on first use
    bool mega_ok = false;      // configuration supports it (bf16, shapes) and SSRB_MEGA != 0
    bool mega = false;         // the open batch decodes through it
    bool mega_dirty = true;    // weights changed since the packed copies were built
    std::vector<void*> mega_w; // packed matrices: 4 per layer + 2 heads
    void* mega_layers = nullptr;    // device array of per-layer records
    unsigned int* mega_bar = nullptr;
    float2* ln_part = nullptr; // [max_rows][d_model / 128] {mean, M2} partials of the residual stream
    std::map<std::string, bool> loaded;
    // workspace
    int Mmax = 0;
    float *x = nullptr, *qkv = nullptr, *logits = nullptr;
    void *hn = nullptr, *ao = nullptr, *hid = nullptr, *hlast = nullptr, *hh = nullptr;
    void *kcache = nullptr, *vcache = nullptr;
    size_t kv_layer_elems = 0;
    float* attn_ws = nullptr; int* tickets = nullptr;
    char* tf_buf = nullptr; size_t tf_cap = 0;      // grow-only scratch of the teacher-forcing / training-loss entry points
    void* tc_ws = nullptr; size_t tc_ws_bytes = 0;
    PosDesc* d_desc = nullptr; int *d_rows = nullptr, *d_slots = nullptr;
    int *d_row_ids = nullptr, *d_row_start = nullptr, *d_row_len = nullptr, *d_last_idx = nullptr;
    UttState* d_state = nullptr; int *d_seq_len = nullptr, *d_next_tok = nullptr, *d_gen_tok = nullptr, *d_iter = nullptr;
    float* staging = nullptr; size_t staging_elems = 0;
    // batch
    int n_utt = 0, rpu = 1, R = 0;
    SampleParams sp{};
    const float* noise = nullptr;
    cudaGraphExec_t graph = nullptr;
    bool use_graph = true;
    unsigned long long graph_kernels = 0;       // kernels captured in one decode-step graph
    // profiling (ssrb_lm_profile_steps): CUDA events around every kernel class of un-graphed steps
    bool prof = false;
    std::vector<std::tuple<int, cudaEvent_t, cudaEvent_t>> prof_ev;
};

enum ProfClass { PC_ATTN = 0, PC_GEMM = 1, PC_SMALL = 2, PC_SAMPLE = 3, PC_COUNT = 4 };
struct ProfScope {
    ssrb_lm* lm; cudaStream_t s; cudaEvent_t e0 = nullptr, e1 = nullptr; int cls;
    ProfScope(ssrb_lm* l, int c, cudaStream_t st) : lm(l), s(st), cls(c) {
        if (lm->prof) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, s); }
    }
    ~ProfScope() { if (lm->prof) { cudaEventRecord(e1, s); lm->prof_ev.emplace_back(cls, e0, e1); } }
};

static int dev_alloc(void** p, size_t bytes) {
    SSRB_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    return 0;
}

const char* ssrb_last_error(void) { return g_err.c_str(); }
int ssrb_version(void) { return 100; }
uint64_t ssrb_launch_count(void) { return g_launch_count; }

int ssrb_lm_create(const ssrb_lm_config* c, int device, ssrb_lm** out) {
    SSRB_CHECK(c && out, "null argument");
    SSRB_CHECK(c->d_model % 128 == 0 && c->d_model / c->n_head == 128, "head_dim must be 128");
    SSRB_CHECK(c->d_model <= 2048, "d_model > 2048 not supported");
    SSRB_CHECK(c->n_codebooks == 4, "n_codebooks must be 4");
    // UttState::span_len[] and ssrb_lm_read_tokens hold SSRB_MAX_SPANS entries: a checkpoint with more spans must fail here, not
    // write past the array in sample_kernel
    SSRB_CHECK(c->max_n_spans >= 1 && c->max_n_spans <= SSRB_MAX_SPANS, "max_n_spans must be 1..SSRB_MAX_SPANS (3)");
    SSRB_CHECK(c->weight_dtype == SSRB_DTYPE_F32 || c->weight_dtype == SSRB_DTYPE_BF16, "bad weight_dtype");
    SSRB_CHECK(c->max_rows > 0 && c->max_seq > 0 && c->max_prefill_tokens > 0 && c->max_steps > 0, "bad capacity");
    SSRB_CUDA(cudaSetDevice(device));
    ssrb_lm* lm = new ssrb_lm();
    struct Guard { ssrb_lm* p; ~Guard() { if (p) ssrb_lm_destroy(p); } } guard{lm};    // an allocation failure below frees what exists
    lm->cfg = *c; lm->device = device;
    lm->D = c->d_model; lm->H = c->n_head; lm->L = c->n_layer; lm->F = c->ffn_dim; lm->K = c->n_codebooks;
    lm->V = c->n_audio_tokens; lm->Vt = c->n_text_tokens; lm->Hh = c->head_hidden;
    lm->wdt = c->weight_dtype; lm->esz = c->weight_dtype == SSRB_DTYPE_F32 ? 4 : 2;
    const char* ng = getenv("SSRB_NO_GRAPH");
    lm->use_graph = !(ng && ng[0] == '1');
    const int D = lm->D, F = lm->F, K = lm->K, V = lm->V, Hh = lm->Hh, L = lm->L;
    const size_t e = lm->esz;
    lm->layers.resize(L);
    for (int n = 0; n < L; n++) {
        LayerW& w = lm->layers[n];
        SSRB_TRY(dev_alloc(&w.wqkv, (size_t)3 * D * D * e)); SSRB_TRY(dev_alloc(&w.wo, (size_t)D * D * e));
        SSRB_TRY(dev_alloc(&w.w1, (size_t)F * D * e)); SSRB_TRY(dev_alloc(&w.w2, (size_t)D * F * e));
        SSRB_TRY(dev_alloc((void**)&w.bqkv, 3 * D * 4)); SSRB_TRY(dev_alloc((void**)&w.bo, D * 4));
        SSRB_TRY(dev_alloc((void**)&w.b1, F * 4)); SSRB_TRY(dev_alloc((void**)&w.b2, D * 4));
        SSRB_TRY(dev_alloc((void**)&w.ln1w, D * 4)); SSRB_TRY(dev_alloc((void**)&w.ln1b, D * 4));
        SSRB_TRY(dev_alloc((void**)&w.ln2w, D * 4)); SSRB_TRY(dev_alloc((void**)&w.ln2b, D * 4));
    }
    SSRB_TRY(dev_alloc((void**)&lm->text_emb, (size_t)lm->Vt * D * 4));
    SSRB_TRY(dev_alloc((void**)&lm->audio_emb, (size_t)K * V * D * 4));
    SSRB_TRY(dev_alloc((void**)&lm->lnfw, D * 4)); SSRB_TRY(dev_alloc((void**)&lm->lnfb, D * 4));
    SSRB_TRY(dev_alloc(&lm->hw1, (size_t)K * Hh * D * e)); SSRB_TRY(dev_alloc(&lm->hw2, (size_t)K * V * Hh * e));
    SSRB_TRY(dev_alloc((void**)&lm->hb1, K * Hh * 4)); SSRB_TRY(dev_alloc((void**)&lm->hb2, K * V * 4));
    lm->fold_ok = ln_fold_enabled() && c->weight_dtype == SSRB_DTYPE_BF16 && c->gemm_impl != 1 && D % 128 == 0 &&
                  D <= 2048 && F % 128 == 0 && (K * Hh) % 4 == 0;
    if (lm->fold_ok) {
        for (int n = 0; n < L; n++) {
            LayerW& w = lm->layers[n];
            SSRB_TRY(dev_alloc(&w.wqkv_f, (size_t)3 * D * D * e)); SSRB_TRY(dev_alloc(&w.w1_f, (size_t)F * D * e));
            SSRB_TRY(dev_alloc((void**)&w.cqkv, 3 * D * 4)); SSRB_TRY(dev_alloc((void**)&w.bqkv_f, 3 * D * 4));
            SSRB_TRY(dev_alloc((void**)&w.c1, F * 4)); SSRB_TRY(dev_alloc((void**)&w.b1_f, F * 4));
        }
        SSRB_TRY(dev_alloc(&lm->hw1_f, (size_t)K * Hh * D * e));
        SSRB_TRY(dev_alloc((void**)&lm->hc1, K * Hh * 4)); SSRB_TRY(dev_alloc((void**)&lm->hb1_f, K * Hh * 4));
        SSRB_TRY(dev_alloc((void**)&lm->ln_part, (size_t)c->max_rows * (D / 128) * sizeof(float2)));
    }
    // workspace
    const int R = c->max_rows;
    lm->Mmax = c->max_prefill_tokens > R ? c->max_prefill_tokens : R;
    const size_t M = lm->Mmax;
    SSRB_TRY(dev_alloc((void**)&lm->x, M * D * 4)); SSRB_TRY(dev_alloc((void**)&lm->qkv, M * 3 * D * 4));
    SSRB_TRY(dev_alloc(&lm->hn, M * D * e)); SSRB_TRY(dev_alloc(&lm->ao, M * D * e));
    SSRB_TRY(dev_alloc(&lm->hid, M * F * e));
    SSRB_TRY(dev_alloc(&lm->hlast, (size_t)R * D * e)); SSRB_TRY(dev_alloc(&lm->hh, (size_t)R * K * Hh * e));
    SSRB_TRY(dev_alloc((void**)&lm->logits, (size_t)R * K * V * 4));
    lm->kv_layer_elems = (size_t)R * lm->H * c->max_seq * 128;
    SSRB_TRY(dev_alloc(&lm->kcache, lm->kv_layer_elems * L * e));
    SSRB_TRY(dev_alloc(&lm->vcache, lm->kv_layer_elems * L * e));
    SSRB_TRY(dev_alloc((void**)&lm->attn_ws, attn_decode_ws_floats(R, lm->H, c->max_seq) * 4));
    SSRB_TRY(dev_alloc((void**)&lm->tickets, (size_t)R * lm->H * 4));
    SSRB_CUDA(cudaMemset(lm->tickets, 0, (size_t)R * lm->H * 4));
    lm->tc_ws_bytes = gemm_tc_workspace_bytes(R, 3 * D > F ? 3 * D : F);
    SSRB_TRY(dev_alloc(&lm->tc_ws, lm->tc_ws_bytes));
    SSRB_CUDA(cudaMemset(lm->tc_ws, 0, lm->tc_ws_bytes ? lm->tc_ws_bytes : 16));
    SSRB_TRY(dev_alloc((void**)&lm->d_desc, M * sizeof(PosDesc)));
    SSRB_TRY(dev_alloc((void**)&lm->d_rows, M * 4)); SSRB_TRY(dev_alloc((void**)&lm->d_slots, M * 4));
    SSRB_TRY(dev_alloc((void**)&lm->d_row_ids, R * 4)); SSRB_TRY(dev_alloc((void**)&lm->d_row_start, R * 4));
    SSRB_TRY(dev_alloc((void**)&lm->d_row_len, R * 4)); SSRB_TRY(dev_alloc((void**)&lm->d_last_idx, M * 4));
    SSRB_TRY(dev_alloc((void**)&lm->d_state, R * sizeof(UttState)));
    SSRB_TRY(dev_alloc((void**)&lm->d_seq_len, R * 4)); SSRB_TRY(dev_alloc((void**)&lm->d_next_tok, R * K * 4));
    SSRB_TRY(dev_alloc((void**)&lm->d_gen_tok, (size_t)R * c->max_steps * K * 4));
    SSRB_TRY(dev_alloc((void**)&lm->d_iter, 4));
    SSRB_TRY(dev_alloc((void**)&lm->gbar, 64 * 4));
    SSRB_CUDA(cudaMemset(lm->gbar, 0, 64 * 4));
    SSRB_TRY(dev_alloc((void**)&lm->mega_bar, 64 * 4));
    SSRB_CUDA(cudaMemset(lm->mega_bar, 0, 64 * 4));
    lm->mega_ok = mega_enabled() && c->weight_dtype == SSRB_DTYPE_BF16 && c->gemm_impl != 1 &&
                  mega_supported(1, D, lm->H, F, K, V, Hh);
    guard.p = nullptr;
    *out = lm;
    return 0;
}

void ssrb_lm_destroy(ssrb_lm* lm) {
    if (!lm) return;
    cudaSetDevice(lm->device);
    cudaDeviceSynchronize();
    if (lm->graph) cudaGraphExecDestroy(lm->graph);
    cudaFree(lm->tf_buf);
    for (auto& w : lm->layers) {
        void* ps[] = {w.wqkv, w.wo, w.w1, w.w2, w.bqkv, w.bo, w.b1, w.b2, w.ln1w, w.ln1b, w.ln2w, w.ln2b,
                      w.wqkv_f, w.w1_f, w.cqkv, w.bqkv_f, w.c1, w.b1_f};
        for (void* p : ps) cudaFree(p);
    }
    void* ps[] = {lm->text_emb, lm->audio_emb, lm->pe, lm->lnfw, lm->lnfb, lm->hw1, lm->hw2, lm->hb1, lm->hb2, lm->x,
                  lm->qkv, lm->logits, lm->hn, lm->ao, lm->hid, lm->hlast, lm->hh, lm->kcache, lm->vcache, lm->attn_ws,
                  lm->tickets, lm->tc_ws, lm->d_desc, lm->d_rows, lm->d_slots, lm->d_row_ids, lm->d_row_start,
                  lm->d_row_len, lm->d_last_idx, lm->d_state, lm->d_seq_len, lm->d_next_tok, lm->d_gen_tok, lm->d_iter,
                  lm->staging, lm->hw1_f, lm->hc1, lm->hb1_f, lm->ln_part, lm->gbar, lm->mega_layers, lm->mega_bar};
    for (void* p : ps) cudaFree(p);
    for (void* p : lm->mega_w) cudaFree(p);
    delete lm;
}

// ---- weight loading -------------------------------------------------------------------------------
static int upload(ssrb_lm* lm, const float* host, int64_t n, void* dst, int dst_dtype) {
    if (dst_dtype == SSRB_DTYPE_F32) {
        SSRB_CUDA(cudaMemcpy(dst, host, n * 4, cudaMemcpyHostToDevice));
        return 0;
    }
    if ((size_t)n > lm->staging_elems) {
        if (lm->staging) cudaFree(lm->staging);
        lm->staging = nullptr; lm->staging_elems = 0;
        SSRB_TRY(dev_alloc((void**)&lm->staging, n * 4));
        lm->staging_elems = n;
    }
    SSRB_CUDA(cudaMemcpy(lm->staging, host, n * 4, cudaMemcpyHostToDevice));
    SSRB_LAUNCH(convert_kernel<bf16>, 1024, 256, 0, 0, lm->staging, (bf16*)dst, n);
    SSRB_CUDA(cudaDeviceSynchronize());
    return 0;
}

static bool starts_with(const std::string& s, const char* p) { return s.rfind(p, 0) == 0; }

int ssrb_lm_load_tensor(ssrb_lm* lm, const char* name_c, const float* host, const int64_t* shape, int ndim) {
    SSRB_CHECK(lm && name_c && host, "null argument");
    SSRB_CUDA(cudaSetDevice(lm->device));
    const std::string name(name_c);
    int64_t n = 1;
    for (int i = 0; i < ndim; i++) n *= shape[i];
    const int D = lm->D, F = lm->F, K = lm->K, V = lm->V, Hh = lm->Hh;
    const int wdt = lm->wdt;
    auto expect = [&](int64_t want) -> bool { return n == want; };
#define SSRB_EXPECT(want) SSRB_CHECK(expect(want), ("shape mismatch for " + name).c_str())
    if (name == "pe_table") {
        SSRB_CHECK(ndim == 2 && shape[1] == D, "pe_table must be [n_pos, d_model]");
        if (lm->pe) cudaFree(lm->pe);
        SSRB_TRY(dev_alloc((void**)&lm->pe, n * 4));
        lm->n_pos = (int)shape[0];
        SSRB_TRY(upload(lm, host, n, lm->pe, SSRB_DTYPE_F32));
    } else if (name == "text_embedding.word_embeddings.weight") {
        SSRB_EXPECT((int64_t)lm->Vt * D); SSRB_TRY(upload(lm, host, n, lm->text_emb, SSRB_DTYPE_F32));
    } else if (starts_with(name, "audio_embedding.")) {
        const int k = atoi(name.c_str() + 16);
        SSRB_CHECK(k >= 0 && k < K, "bad codebook index"); SSRB_EXPECT((int64_t)V * D);
        SSRB_TRY(upload(lm, host, n, lm->audio_emb + (size_t)k * V * D, SSRB_DTYPE_F32));
    } else if (name == "text_positional_embedding.alpha") { lm->alpha_t = host[0];
    } else if (name == "audio_positional_embedding.alpha") { lm->alpha_a = host[0];
    } else if (name == "decoder.norm.weight") { SSRB_EXPECT(D); SSRB_TRY(upload(lm, host, n, lm->lnfw, SSRB_DTYPE_F32));
    } else if (name == "decoder.norm.bias") { SSRB_EXPECT(D); SSRB_TRY(upload(lm, host, n, lm->lnfb, SSRB_DTYPE_F32));
    } else if (starts_with(name, "decoder.layers.")) {
        const int nl = atoi(name.c_str() + 15);
        SSRB_CHECK(nl >= 0 && nl < lm->L, "bad layer index");
        LayerW& w = lm->layers[nl];
        const std::string sub = name.substr(name.find('.', 15) + 1);
        if (sub == "self_attn.in_proj_weight") { SSRB_EXPECT((int64_t)3 * D * D); SSRB_TRY(upload(lm, host, n, w.wqkv, wdt)); }
        else if (sub == "self_attn.in_proj_bias") { SSRB_EXPECT(3 * D); SSRB_TRY(upload(lm, host, n, w.bqkv, 0)); }
        else if (sub == "self_attn.out_proj.weight") { SSRB_EXPECT((int64_t)D * D); SSRB_TRY(upload(lm, host, n, w.wo, wdt)); }
        else if (sub == "self_attn.out_proj.bias") { SSRB_EXPECT(D); SSRB_TRY(upload(lm, host, n, w.bo, 0)); }
        else if (sub == "linear1.weight") { SSRB_EXPECT((int64_t)F * D); SSRB_TRY(upload(lm, host, n, w.w1, wdt)); }
        else if (sub == "linear1.bias") { SSRB_EXPECT(F); SSRB_TRY(upload(lm, host, n, w.b1, 0)); }
        else if (sub == "linear2.weight") { SSRB_EXPECT((int64_t)D * F); SSRB_TRY(upload(lm, host, n, w.w2, wdt)); }
        else if (sub == "linear2.bias") { SSRB_EXPECT(D); SSRB_TRY(upload(lm, host, n, w.b2, 0)); }
        else if (sub == "norm1.weight") { SSRB_EXPECT(D); SSRB_TRY(upload(lm, host, n, w.ln1w, 0)); }
        else if (sub == "norm1.bias") { SSRB_EXPECT(D); SSRB_TRY(upload(lm, host, n, w.ln1b, 0)); }
        else if (sub == "norm2.weight") { SSRB_EXPECT(D); SSRB_TRY(upload(lm, host, n, w.ln2w, 0)); }
        else if (sub == "norm2.bias") { SSRB_EXPECT(D); SSRB_TRY(upload(lm, host, n, w.ln2b, 0)); }
        else return 0;
    } else if (starts_with(name, "predict_layer.")) {
        const int k = atoi(name.c_str() + 14);
        SSRB_CHECK(k >= 0 && k < K, "bad head index");
        const std::string sub = name.substr(name.find('.', 14) + 1);
        if (sub == "0.weight") { SSRB_EXPECT((int64_t)Hh * D); SSRB_TRY(upload(lm, host, n, (char*)lm->hw1 + (size_t)k * Hh * D * lm->esz, wdt)); }
        else if (sub == "0.bias") { SSRB_EXPECT(Hh); SSRB_TRY(upload(lm, host, n, lm->hb1 + k * Hh, 0)); }
        else if (sub == "2.weight") { SSRB_EXPECT((int64_t)V * Hh); SSRB_TRY(upload(lm, host, n, (char*)lm->hw2 + (size_t)k * V * Hh * lm->esz, wdt)); }
        else if (sub == "2.bias") { SSRB_EXPECT(V); SSRB_TRY(upload(lm, host, n, lm->hb2 + k * V, 0)); }
        else return 0;
    } else {
        return 0;   // keys that carry no inference state (e.g. accuracy_metrics.*) are ignored
    }
    lm->loaded[name] = true;
    lm->fold_dirty = true;
    lm->mega_dirty = true;
    return 0;
}

// (re)build the gamma-folded copies of the LayerNorm-consuming matrices once all tensors are loaded
static int fold_weights(ssrb_lm* lm, cudaStream_t s) {
    if (!lm->fold_ok || !lm->fold_dirty) return 0;
    const int D = lm->D, F = lm->F;
    for (auto& w : lm->layers) {
        SSRB_TRY(launch_fold_ln(w.wqkv, 3 * D, D, w.ln1w, w.ln1b, w.bqkv, w.wqkv_f, w.cqkv, w.bqkv_f, s));
        SSRB_TRY(launch_fold_ln(w.w1, F, D, w.ln2w, w.ln2b, w.b1, w.w1_f, w.c1, w.b1_f, s));
    }
    SSRB_TRY(launch_fold_ln(lm->hw1, lm->K * lm->Hh, D, lm->lnfw, lm->lnfb, lm->hb1, lm->hw1_f, lm->hc1, lm->hb1_f, s));
    SSRB_CUDA(cudaStreamSynchronize(s));
    lm->fold_dirty = false;
    return 0;
}

// (re)build the streaming-order copies of every matrix the persistent small-batch kernel reads (lm_mega.cu), and its per-layer
// records.  1.65 GB for the 830M model, allocated the first time a batch of <= 16 rows is opened.
static int mega_prepare(ssrb_lm* lm, cudaStream_t s) {
    if (!lm->mega_dirty && !lm->mega_w.empty()) return 0;
    const int D = lm->D, F = lm->F, L = lm->L, K = lm->K, V = lm->V, Hh = lm->Hh;
    const size_t e = 2;
    if (lm->mega_w.empty()) {
        lm->mega_w.assign((size_t)4 * L + 2, nullptr);
        for (int n = 0; n < L; n++) {
            SSRB_TRY(dev_alloc(&lm->mega_w[4 * n + 0], (size_t)3 * D * D * e)); SSRB_TRY(dev_alloc(&lm->mega_w[4 * n + 1], (size_t)D * D * e));
            SSRB_TRY(dev_alloc(&lm->mega_w[4 * n + 2], (size_t)F * D * e)); SSRB_TRY(dev_alloc(&lm->mega_w[4 * n + 3], (size_t)D * F * e));
        }
        SSRB_TRY(dev_alloc(&lm->mega_w[4 * L], (size_t)K * Hh * D * e)); SSRB_TRY(dev_alloc(&lm->mega_w[4 * L + 1], (size_t)K * V * Hh * e));
        SSRB_TRY(dev_alloc(&lm->mega_layers, (size_t)L * mega_layer_bytes()));
    }
    std::vector<char> recs((size_t)L * mega_layer_bytes());
    for (int n = 0; n < L; n++) {
        const LayerW& w = lm->layers[n];
        SSRB_TRY(mega_pack(w.wqkv, lm->mega_w[4 * n + 0], 3 * D, D, MEGA_QKV, s));
        SSRB_TRY(mega_pack(w.wo, lm->mega_w[4 * n + 1], D, D, MEGA_OUT, s));
        SSRB_TRY(mega_pack(w.w1, lm->mega_w[4 * n + 2], F, D, MEGA_FFN1, s));
        SSRB_TRY(mega_pack(w.w2, lm->mega_w[4 * n + 3], D, F, MEGA_FFN2, s));
        MegaLayerHost h{};
        h.D = D; h.F = F;
        h.wqkv = lm->mega_w[4 * n + 0]; h.wo = lm->mega_w[4 * n + 1]; h.w1 = lm->mega_w[4 * n + 2]; h.w2 = lm->mega_w[4 * n + 3];
        h.bqkv = w.bqkv; h.bo = w.bo; h.b1 = w.b1; h.b2 = w.b2; h.ln1g = w.ln1w; h.ln1b = w.ln1b; h.ln2g = w.ln2w; h.ln2b = w.ln2b;
        h.kc = (char*)lm->kcache + (size_t)n * lm->kv_layer_elems * lm->esz;
        h.vc = (char*)lm->vcache + (size_t)n * lm->kv_layer_elems * lm->esz;
        SSRB_TRY(mega_fill_layer(recs.data() + (size_t)n * mega_layer_bytes(), h));
    }
    SSRB_TRY(mega_pack(lm->hw1, lm->mega_w[4 * L], K * Hh, D, MEGA_H1, s));
    SSRB_TRY(mega_pack(lm->hw2, lm->mega_w[4 * L + 1], K * V, Hh, MEGA_H2, s));
    SSRB_CUDA(cudaMemcpyAsync(lm->mega_layers, recs.data(), recs.size(), cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaStreamSynchronize(s));
    lm->mega_dirty = false;
    return 0;
}

int ssrb_lm_check_loaded(ssrb_lm* lm) {
    std::vector<std::string> need = {"pe_table", "text_embedding.word_embeddings.weight", "text_positional_embedding.alpha",
                                     "audio_positional_embedding.alpha", "decoder.norm.weight", "decoder.norm.bias"};
    for (int k = 0; k < lm->K; k++) {
        need.push_back("audio_embedding." + std::to_string(k) + ".word_embeddings.weight");
        for (const char* s : {"0.weight", "0.bias", "2.weight", "2.bias"}) need.push_back("predict_layer." + std::to_string(k) + "." + s);
    }
    for (int n = 0; n < lm->L; n++)
        for (const char* s : {"self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight",
                              "self_attn.out_proj.bias", "linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias",
                              "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias"})
            need.push_back("decoder.layers." + std::to_string(n) + "." + s);
    std::string missing;
    for (auto& k : need) if (!lm->loaded.count(k)) missing += k + " ";
    if (!missing.empty()) { set_error("missing tensors: " + missing); return 1; }
    return 0;
}

// ---- GEMM dispatch ---------------------------------------------------------------------------------
static int gemm(ssrb_lm* lm, GemmArgs g, cudaStream_t s) {
    ProfScope ps(lm, PC_GEMM, s);
    g.ab_dtype = lm->wdt;
    if (lm->wdt == SSRB_DTYPE_BF16 && lm->cfg.gemm_impl != 1 && gemm_tc_supported(g))
        return gemm_tc(g, lm->tc_ws, lm->tc_ws_bytes, s);
    SSRB_CHECK(!g.ln_part && !g.part_out && !g.C2, "folded LayerNorm needs the tcgen05 GEMM");
    return gemm_simt(g, s);
}

// one decoder layer over M packed positions.  prefill: rows/slots/attn over packed rows; decode: M = R.
static int run_layer(ssrb_lm* lm, int n, int M, bool prefill, int n_rows, int max_len, cudaStream_t s) {
    const int D = lm->D, F = lm->F, H = lm->H;
    const LayerW& w = lm->layers[n];
    const size_t e = lm->esz;
    void* kc = (char*)lm->kcache + (size_t)n * lm->kv_layer_elems * e;
    void* vc = (char*)lm->vcache + (size_t)n * lm->kv_layer_elems * e;
    { ProfScope ps(lm, PC_SMALL, s); SSRB_TRY(launch_layernorm(lm->x, nullptr, M, D, w.ln1w, w.ln1b, lm->hn, lm->wdt, s)); }
    GemmArgs g;
    g.A = lm->hn; g.lda = D; g.W = w.wqkv; g.ldw = D; g.bias = w.bqkv; g.C = lm->qkv; g.ldc = 3 * D;
    g.M = M; g.N = 3 * D; g.K = D; g.c_dtype = SSRB_DTYPE_F32;
    SSRB_TRY(gemm(lm, g, s));
    if (prefill) {
        SSRB_TRY(launch_kv_append(lm->qkv, M, D, H, lm->d_rows, lm->d_slots, nullptr, kc, vc, lm->wdt, lm->cfg.max_seq, s));
        SSRB_TRY(launch_attn_prefill(lm->qkv, D, H, kc, vc, lm->wdt, lm->cfg.max_seq, n_rows, lm->d_row_ids,
                                     lm->d_row_start, lm->d_row_len, max_len, lm->ao, lm->wdt, s));
    } else {
        ProfScope ps(lm, PC_ATTN, s);   // the decode attention kernel also appends this step's K/V row in place
        SSRB_TRY(launch_attn_decode(lm->qkv, M, D, H, kc, vc, lm->wdt, lm->cfg.max_seq, lm->d_seq_len, lm->d_state,
                                    lm->rpu, lm->attn_ws, lm->tickets, lm->ao, lm->wdt, attn_prefetch_enabled() ? 1 : 0, s));
    }
    g = GemmArgs();
    g.A = lm->ao; g.lda = D; g.W = w.wo; g.ldw = D; g.bias = w.bo; g.residual = lm->x; g.ldr = D; g.C = lm->x; g.ldc = D;
    g.M = M; g.N = D; g.K = D; g.c_dtype = SSRB_DTYPE_F32;
    SSRB_TRY(gemm(lm, g, s));
    { ProfScope ps(lm, PC_SMALL, s); SSRB_TRY(launch_layernorm(lm->x, nullptr, M, D, w.ln2w, w.ln2b, lm->hn, lm->wdt, s)); }
    g = GemmArgs();
    g.A = lm->hn; g.lda = D; g.W = w.w1; g.ldw = D; g.bias = w.b1; g.C = lm->hid; g.ldc = F;
    g.M = M; g.N = F; g.K = D; g.act = ACT_RELU; g.c_dtype = lm->wdt;
    SSRB_TRY(gemm(lm, g, s));
    g = GemmArgs();
    g.A = lm->hid; g.lda = F; g.W = w.w2; g.ldw = F; g.bias = w.b2; g.residual = lm->x; g.ldr = D; g.C = lm->x; g.ldc = D;
    g.M = M; g.N = D; g.K = F; g.c_dtype = SSRB_DTYPE_F32;
    SSRB_TRY(gemm(lm, g, s));
    return 0;
}

// one decoder layer of a decode iteration with both LayerNorms folded into the GEMMs that consume them (gemm_tc.cu):
// hn holds bf16(x), ln_part the row statistics of x; the two residual GEMMs refresh both while they write x.
static int fold_qkv(ssrb_lm* lm, int n, int M, cudaStream_t s) {
    const int D = lm->D;
    const LayerW& w = lm->layers[n];
    GemmArgs g;
    g.A = lm->hn; g.lda = D; g.W = w.wqkv_f; g.ldw = D; g.bias = w.bqkv_f; g.C = lm->qkv; g.ldc = 3 * D;
    g.M = M; g.N = 3 * D; g.K = D; g.c_dtype = SSRB_DTYPE_F32;
    g.ln_part = lm->ln_part; g.part_ld = lm->cfg.max_rows; g.ln_blocks = D / 128; g.ln_colsum = w.cqkv;
    return gemm(lm, g, s);
}

static int fold_attn(ssrb_lm* lm, int n, int M, cudaStream_t s) {
    const size_t e = lm->esz;
    void* kc = (char*)lm->kcache + (size_t)n * lm->kv_layer_elems * e;
    void* vc = (char*)lm->vcache + (size_t)n * lm->kv_layer_elems * e;
    ProfScope ps(lm, PC_ATTN, s);
    return launch_attn_decode(lm->qkv, M, lm->D, lm->H, kc, vc, lm->wdt, lm->cfg.max_seq, lm->d_seq_len, lm->d_state,
                              lm->rpu, lm->attn_ws, lm->tickets, lm->ao, lm->wdt, attn_prefetch_enabled() ? 1 : 0, s);
}

static int run_layer_fold(ssrb_lm* lm, int n, int M, cudaStream_t s) {
    const int D = lm->D, F = lm->F;
    const LayerW& w = lm->layers[n];
    SSRB_TRY(fold_qkv(lm, n, M, s));
    SSRB_TRY(fold_attn(lm, n, M, s));
    GemmArgs g;
    g.A = lm->ao; g.lda = D; g.W = w.wo; g.ldw = D; g.bias = w.bo; g.residual = lm->x; g.ldr = D; g.C = lm->x; g.ldc = D;
    g.M = M; g.N = D; g.K = D; g.c_dtype = SSRB_DTYPE_F32;
    g.C2 = lm->hn; g.ldc2 = D; g.part_out = lm->ln_part; g.part_ld = lm->cfg.max_rows;
    SSRB_TRY(gemm(lm, g, s));
    g = GemmArgs();
    g.A = lm->hn; g.lda = D; g.W = w.w1_f; g.ldw = D; g.bias = w.b1_f; g.C = lm->hid; g.ldc = F;
    g.M = M; g.N = F; g.K = D; g.act = ACT_RELU; g.c_dtype = lm->wdt;
    g.ln_part = lm->ln_part; g.part_ld = lm->cfg.max_rows; g.ln_blocks = D / 128; g.ln_colsum = w.c1;
    SSRB_TRY(gemm(lm, g, s));
    g = GemmArgs();
    g.A = lm->hid; g.lda = F; g.W = w.w2; g.ldw = F; g.bias = w.b2; g.residual = lm->x; g.ldr = D; g.C = lm->x; g.ldc = D;
    g.M = M; g.N = D; g.K = F; g.c_dtype = SSRB_DTYPE_F32;
    g.C2 = lm->hn; g.ldc2 = D; g.part_out = lm->ln_part; g.part_ld = lm->cfg.max_rows;
    SSRB_TRY(gemm(lm, g, s));
    return 0;
}

// the same layer with its four GEMM launches replaced by one persistent launch (gemm_layer.cu, SSRB_LAYER_KERNEL=1): layer n's
// attention, then out-proj -> FFN1 -> FFN2 -> layer n+1's QKV projection (layer 0's projection is launched by the caller)
static int run_layer_persistent(ssrb_lm* lm, int n, int M, cudaStream_t s) {
    const LayerW& w = lm->layers[n];
    SSRB_TRY(fold_attn(lm, n, M, s));
    ProfScope ps(lm, PC_GEMM, s);
    LayerChainArgs a;
    a.M = M; a.D = lm->D; a.F = lm->F;
    a.ao = lm->ao; a.x = lm->x; a.hn = lm->hn; a.hid = lm->hid; a.qkv = lm->qkv;
    a.ln_part = lm->ln_part; a.part_ld = lm->cfg.max_rows;
    a.wo = w.wo; a.bo = w.bo; a.w1f = w.w1_f; a.b1f = w.b1_f; a.c1 = w.c1; a.w2 = w.w2; a.b2 = w.b2;
    if (n + 1 < lm->L) {
        const LayerW& nx = lm->layers[n + 1];
        a.wqkv_next = nx.wqkv_f; a.bqkv_next = nx.bqkv_f; a.cqkv_next = nx.cqkv;
    }
    a.gbar = lm->gbar;
    return gemm_layer(a, s);
}

// final LN (gathered rows) + 4 prediction heads -> logits [M, K, V] fp32
static int run_heads(ssrb_lm* lm, const int* gather_idx, int M, void* hl, void* hhbuf, float* logits, cudaStream_t s) {
    const int D = lm->D, K = lm->K, V = lm->V, Hh = lm->Hh;
    SSRB_TRY(launch_layernorm(lm->x, gather_idx, M, D, lm->lnfw, lm->lnfb, hl, lm->wdt, s));
    GemmArgs g;
    g.A = hl; g.lda = D; g.W = lm->hw1; g.ldw = D; g.bias = lm->hb1; g.C = hhbuf; g.ldc = K * Hh;
    g.M = M; g.N = K * Hh; g.K = D; g.act = ACT_GELU; g.c_dtype = lm->wdt;
    SSRB_TRY(gemm(lm, g, s));
    g = GemmArgs();
    g.A = hhbuf; g.lda = K * Hh; g.a_gs = Hh; g.W = lm->hw2; g.ldw = Hh; g.w_gs = (int64_t)V * Hh;
    g.bias = lm->hb2; g.bias_gs = V; g.C = logits; g.ldc = (int64_t)K * V; g.c_gs = V;
    g.M = M; g.N = V; g.K = Hh; g.groups = K; g.c_dtype = SSRB_DTYPE_F32;
    SSRB_TRY(gemm(lm, g, s));
    return 0;
}

// a decode iteration of a small batch (R <= 16): embedding, ONE persistent kernel for the 16 layers and the heads, sampler
static int enqueue_step_mega(ssrb_lm* lm, cudaStream_t s) {
    { ProfScope ps(lm, PC_SMALL, s);
      SSRB_TRY(launch_embed_step(lm->d_next_tok, lm->d_state, lm->R, lm->rpu, lm->K, lm->D, lm->audio_emb, lm->V, lm->pe,
                                 lm->alpha_a, lm->x, s, lm->mega_bar)); }
    { ProfScope ps(lm, PC_GEMM, s);
      MegaArgs a{};
      a.R = lm->R; a.D = lm->D; a.H = lm->H; a.F = lm->F; a.L = lm->L; a.NCB = lm->K; a.V = lm->V; a.Hh = lm->Hh;
      a.Smax = lm->cfg.max_seq; a.rpu = lm->rpu; a.max_pieces = attn_decode_tma_max_nsplit(lm->cfg.max_seq);
      a.x = lm->x; a.qkv = lm->qkv; a.ao = lm->ao; a.hid = lm->hid; a.hh = lm->hh; a.logits = lm->logits;
      a.seq_len = lm->d_seq_len; a.st = lm->d_state; a.attn_ws = lm->attn_ws; a.tickets = lm->tickets; a.bar = lm->mega_bar;
      a.layers_dev = lm->mega_layers;
      a.h1_w = lm->mega_w[(size_t)4 * lm->L]; a.h2_w = lm->mega_w[(size_t)4 * lm->L + 1]; a.h1_b = lm->hb1; a.h2_b = lm->hb2;
      a.lnf_g = lm->lnfw; a.lnf_b = lm->lnfb;
      SSRB_TRY(launch_mega(a, s)); }
    ProfScope ps(lm, PC_SAMPLE, s);
    return launch_sample(lm->logits, lm->d_state, lm->d_seq_len, lm->d_next_tok, lm->d_gen_tok, lm->noise, lm->d_iter, lm->sp, s);
}

static int enqueue_step(ssrb_lm* lm, cudaStream_t s) {
    if (lm->mega) return enqueue_step_mega(lm, s);
    if (lm->fold) {
        { ProfScope ps(lm, PC_SMALL, s);
          SSRB_TRY(launch_embed_step_fold(lm->d_next_tok, lm->d_state, lm->R, lm->rpu, lm->K, lm->D, lm->audio_emb, lm->V,
                                          lm->pe, lm->alpha_a, lm->x, lm->hn, lm->ln_part, lm->cfg.max_rows, s)); }
        if (lm->layer_kernel) {
            SSRB_TRY(fold_qkv(lm, 0, lm->R, s));
            for (int n = 0; n < lm->L; n++) SSRB_TRY(run_layer_persistent(lm, n, lm->R, s));
        } else {
            for (int n = 0; n < lm->L; n++) SSRB_TRY(run_layer_fold(lm, n, lm->R, s));
        }
        // final LayerNorm folded into the first head layer
        const int D = lm->D, K = lm->K, V = lm->V, Hh = lm->Hh;
        GemmArgs g;
        g.A = lm->hn; g.lda = D; g.W = lm->hw1_f; g.ldw = D; g.bias = lm->hb1_f; g.C = lm->hh; g.ldc = K * Hh;
        g.M = lm->R; g.N = K * Hh; g.K = D; g.act = ACT_GELU; g.c_dtype = lm->wdt;
        g.ln_part = lm->ln_part; g.part_ld = lm->cfg.max_rows; g.ln_blocks = D / 128; g.ln_colsum = lm->hc1;
        SSRB_TRY(gemm(lm, g, s));
        g = GemmArgs();
        g.A = lm->hh; g.lda = K * Hh; g.a_gs = Hh; g.W = lm->hw2; g.ldw = Hh; g.w_gs = (int64_t)V * Hh;
        g.bias = lm->hb2; g.bias_gs = V; g.C = lm->logits; g.ldc = (int64_t)K * V; g.c_gs = V;
        g.M = lm->R; g.N = V; g.K = Hh; g.groups = K; g.c_dtype = SSRB_DTYPE_F32;
        SSRB_TRY(gemm(lm, g, s));
    } else {
    { ProfScope ps(lm, PC_SMALL, s);
      SSRB_TRY(launch_embed_step(lm->d_next_tok, lm->d_state, lm->R, lm->rpu, lm->K, lm->D, lm->audio_emb, lm->V, lm->pe,
                                 lm->alpha_a, lm->x, s)); }
    for (int n = 0; n < lm->L; n++) SSRB_TRY(run_layer(lm, n, lm->R, false, 0, 0, s));
    SSRB_TRY(run_heads(lm, nullptr, lm->R, lm->hlast, lm->hh, lm->logits, s));
    }
    ProfScope ps(lm, PC_SAMPLE, s);
    SSRB_TRY(launch_sample(lm->logits, lm->d_state, lm->d_seq_len, lm->d_next_tok, lm->d_gen_tok, lm->noise, lm->d_iter,
                           lm->sp, s));
    return 0;
}

static int heads_on_rows(ssrb_lm* lm, int r0, int M, cudaStream_t s);

// ---- prefill ----------------------------------------------------------------------------------------
struct RowPlan { int r, u, lx, plen, len; };

static int prefill_chunk(ssrb_lm* lm, const std::vector<RowPlan>& rows, const ssrb_lm_batch* b, cudaStream_t s,
                         const std::vector<int>* tf_audio /* teacher forcing: K*Ty tokens, no <mts> */) {
    const int K = lm->K;
    std::vector<PosDesc> desc; std::vector<int> prow, pslot, rid, rstart, rlen, last;
    int M = 0, max_len = 0;
    for (const RowPlan& rp : rows) {
        rid.push_back(rp.r); rstart.push_back(M); rlen.push_back(rp.len);
        if (rp.len > max_len) max_len = rp.len;
        for (int i = 0; i < rp.len; i++) {
            PosDesc pd{};
            if (i < rp.lx) { pd.text_tok = b->text[(size_t)rp.r * b->text_stride + i]; pd.pe_idx = i; }
            else {
                const int j = i - rp.lx;
                pd.text_tok = -1; pd.pe_idx = j;
                int t[4];
                for (int k = 0; k < K; k++) {
                    if (tf_audio) t[k] = (*tf_audio)[(size_t)k * rp.plen + j];
                    else if (j < rp.plen) t[k] = b->prompt[((size_t)rp.u * K + k) * b->prompt_stride + j];
                    else t[k] = lm->cfg.mts;   // first <mts> of span 0 (ssr.py:654-660)
                }
                pd.a0 = t[0]; pd.a1 = t[1]; pd.a2 = t[2]; pd.a3 = t[3];
                SSRB_CHECK(pd.pe_idx < lm->n_pos, "audio position exceeds the PE table");
            }
            desc.push_back(pd); prow.push_back(rp.r); pslot.push_back(i);
        }
        M += rp.len;
        last.push_back(M - 1);
    }
    SSRB_CHECK(M <= lm->Mmax, "prefill chunk exceeds max_prefill_tokens");
    SSRB_CUDA(cudaMemcpyAsync(lm->d_desc, desc.data(), M * sizeof(PosDesc), cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaMemcpyAsync(lm->d_rows, prow.data(), M * 4, cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaMemcpyAsync(lm->d_slots, pslot.data(), M * 4, cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaMemcpyAsync(lm->d_row_ids, rid.data(), rid.size() * 4, cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaMemcpyAsync(lm->d_row_start, rstart.data(), rid.size() * 4, cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaMemcpyAsync(lm->d_row_len, rlen.data(), rid.size() * 4, cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaMemcpyAsync(lm->d_last_idx, last.data(), last.size() * 4, cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaStreamSynchronize(s));   // host vectors go out of scope
    SSRB_TRY(launch_embed_prefill(lm->d_desc, M, lm->D, lm->text_emb, lm->audio_emb, lm->V, lm->pe, lm->alpha_t,
                                  lm->alpha_a, lm->x, s));
    for (int n = 0; n < lm->L; n++) SSRB_TRY(run_layer(lm, n, M, true, (int)rows.size(), max_len, s));
    if (!tf_audio) {
        // final LN of each row's last position -> hlast[row] (rows of a chunk are consecutive)
        SSRB_TRY(launch_layernorm(lm->x, lm->d_last_idx, (int)rows.size(), lm->D, lm->lnfw, lm->lnfb,
                                  (char*)lm->hlast + (size_t)rows[0].r * lm->D * lm->esz, lm->wdt, s));
    }
    return 0;
}

int ssrb_lm_begin(ssrb_lm* lm, const ssrb_lm_batch* b, const ssrb_sampling* sp, const float* noise_dev, void* stream) {
    SSRB_CHECK(lm && b && sp, "null argument");
    SSRB_CUDA(cudaSetDevice(lm->device));
    SSRB_TRY(ssrb_lm_check_loaded(lm));
    cudaStream_t s = (cudaStream_t)stream;
    const int rpu = sp->aug_text ? 2 : 1;
    const int U = b->n_utt, R = U * rpu, K = lm->K;
    SSRB_CHECK(U > 0 && R <= lm->cfg.max_rows, "batch exceeds max_rows");
    SSRB_CHECK(sp->cfg_coef >= 1.0f, "cfg_coef must be >= 1.0");           // ssr.py:552
    SSRB_CHECK(sp->n_silence <= SSRB_MAX_SILENCE, "too many silence tokens");
    lm->n_utt = U; lm->rpu = rpu; lm->R = R; lm->noise = noise_dev;
    SSRB_TRY(fold_weights(lm, s));
    lm->fold = lm->fold_ok && R <= 128;
    lm->mega = lm->mega_ok && R <= 16 && mega_supported(R, lm->D, lm->H, lm->F, lm->K, lm->V, lm->Hh);
    if (lm->mega) SSRB_TRY(mega_prepare(lm, s));
    lm->layer_kernel = lm->fold && layer_kernel_enabled() && gemm_layer_supported(R, lm->D, lm->F);
    if (lm->layer_kernel) SSRB_CUDA(cudaMemsetAsync(lm->gbar, 0, 64 * 4, s));   // a launch that died mid-barrier must not poison the next batch
    if (lm->graph) { cudaGraphExecDestroy(lm->graph); lm->graph = nullptr; }
    SampleParams& p = lm->sp;
    p.K = K; p.V = lm->V; p.rpu = rpu; p.empty_token = lm->cfg.empty_token; p.eog = lm->cfg.eog; p.eos = lm->cfg.eos;
    p.sos = lm->cfg.sos; p.mts = lm->cfg.mts; p.max_n_spans = lm->cfg.max_n_spans; p.top_k = sp->top_k; p.top_p = sp->top_p;
    p.temperature = sp->temperature; p.stop_repetition = sp->stop_repetition; p.n_silence = sp->n_silence;
    for (int i = 0; i < sp->n_silence; i++) p.silence[i] = sp->silence_tokens[i];
    p.cfg_coef = sp->cfg_coef; p.cfg_stride = sp->cfg_stride; p.seed = sp->seed; p.max_steps = lm->cfg.max_steps; p.n_utt = U;
    // state + plans
    std::vector<UttState> st(U); std::vector<int> seq(R); std::vector<RowPlan> plan;
    for (int u = 0; u < U; u++) {
        const int lx = b->text_len[u], pl = b->prompt_len[u];
        SSRB_CHECK(lx > 0 && lx <= b->text_stride && pl >= 0 && pl <= b->prompt_stride, "bad text/prompt length");
        SSRB_CHECK(b->n_spans[u] >= 1 && b->n_spans[u] <= lm->cfg.max_n_spans, "n_spans out of range");
        SSRB_CHECK(lx + pl + 1 + lm->cfg.max_steps <= lm->cfg.max_seq, "max_seq too small for text + prompt + max_steps");
        SSRB_CHECK(pl + 1 + lm->cfg.max_steps <= lm->n_pos && lx <= lm->n_pos, "PE table too small");
        UttState z{}; z.cfg_tag = 1; z.prev_token = -1; z.n_spans = b->n_spans[u]; z.x_len = lx; z.y_len = pl; z.rng_id = u;
        st[u] = z;
        for (int j = 0; j < rpu; j++) { seq[u * rpu + j] = lx + pl; plan.push_back({u * rpu + j, u, lx, pl, lx + pl + 1}); }
    }
    SSRB_CUDA(cudaMemcpyAsync(lm->d_state, st.data(), U * sizeof(UttState), cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaMemcpyAsync(lm->d_seq_len, seq.data(), R * 4, cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaMemsetAsync(lm->d_iter, 0, 4, s));
    SSRB_CUDA(cudaStreamSynchronize(s));
    // chunked prefill
    std::vector<RowPlan> chunk; int tok = 0;
    for (size_t i = 0; i < plan.size(); i++) {
        SSRB_CHECK(plan[i].len <= lm->cfg.max_prefill_tokens, "one prompt exceeds max_prefill_tokens");
        if (tok + plan[i].len > lm->cfg.max_prefill_tokens) { SSRB_TRY(prefill_chunk(lm, chunk, b, s, nullptr)); chunk.clear(); tok = 0; }
        chunk.push_back(plan[i]); tok += plan[i].len;
    }
    if (!chunk.empty()) SSRB_TRY(prefill_chunk(lm, chunk, b, s, nullptr));
    // heads on every row's last position, then iteration 1's sample
    SSRB_TRY(heads_on_rows(lm, 0, R, s));
    SSRB_TRY(launch_sample(lm->logits, lm->d_state, lm->d_seq_len, lm->d_next_tok, lm->d_gen_tok, lm->noise, lm->d_iter,
                           lm->sp, s));
    return 0;
}

// heads on rows [r0, r0 + M) of hlast -> logits rows [r0, r0 + M)
static int heads_on_rows(ssrb_lm* lm, int r0, int M, cudaStream_t s) {
    const int D = lm->D, V = lm->V, Hh = lm->Hh, K = lm->K;
    const size_t e = lm->esz;
    GemmArgs g;
    g.A = (char*)lm->hlast + (size_t)r0 * D * e; g.lda = D; g.W = lm->hw1; g.ldw = D; g.bias = lm->hb1;
    g.C = (char*)lm->hh + (size_t)r0 * K * Hh * e; g.ldc = K * Hh;
    g.M = M; g.N = K * Hh; g.K = D; g.act = ACT_GELU; g.c_dtype = lm->wdt;
    SSRB_TRY(gemm(lm, g, s));
    g = GemmArgs();
    g.A = (char*)lm->hh + (size_t)r0 * K * Hh * e; g.lda = K * Hh; g.a_gs = Hh; g.W = lm->hw2; g.ldw = Hh; g.w_gs = (int64_t)V * Hh;
    g.bias = lm->hb2; g.bias_gs = V; g.C = lm->logits + (size_t)r0 * K * V; g.ldc = (int64_t)K * V; g.c_gs = V;
    g.M = M; g.N = V; g.K = Hh; g.groups = K; g.c_dtype = SSRB_DTYPE_F32;
    SSRB_TRY(gemm(lm, g, s));
    return 0;
}

int ssrb_lm_admit(ssrb_lm* lm, int utt, const int32_t* text, int text_len, const int32_t* prompt, int prompt_len,
                  int n_spans, int rng_stream, void* stream) {
    SSRB_CHECK(lm && lm->R > 0 && text && (prompt || prompt_len == 0), "no open batch / null argument");
    SSRB_CHECK(utt >= 0 && utt < lm->n_utt, "bad utterance slot");
    SSRB_CHECK(lm->noise == nullptr, "a batch with injected sampling noise cannot admit utterances");
    SSRB_CUDA(cudaSetDevice(lm->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int rpu = lm->rpu, K = lm->K, r0 = utt * rpu;
    SSRB_CHECK(text_len > 0 && prompt_len >= 0, "bad text/prompt length");
    SSRB_CHECK(n_spans >= 1 && n_spans <= lm->cfg.max_n_spans, "n_spans out of range");
    SSRB_CHECK(text_len + prompt_len + 1 + lm->cfg.max_steps <= lm->cfg.max_seq, "max_seq too small for text + prompt + max_steps");
    SSRB_CHECK(prompt_len + 1 + lm->cfg.max_steps <= lm->n_pos && text_len <= lm->n_pos, "PE table too small");
    SSRB_CHECK(text_len + prompt_len + 1 <= lm->cfg.max_prefill_tokens, "one prompt exceeds max_prefill_tokens");
    // the slot must have finished: its rows are about to be overwritten
    UttState cur;
    SSRB_CUDA(cudaMemcpyAsync(&cur, lm->d_state + utt, sizeof(UttState), cudaMemcpyDeviceToHost, s));
    SSRB_CUDA(cudaStreamSynchronize(s));
    SSRB_CHECK(cur.done, "slot still decoding");
    // host view addressed by absolute row / utterance index, as prefill_chunk expects
    const int ps = prompt_len > 0 ? prompt_len : 1;
    std::vector<int32_t> text_full((size_t)lm->R * text_len, 0), prompt_full((size_t)lm->n_utt * K * ps, 0);
    for (int j = 0; j < rpu; j++)
        std::copy(text + (size_t)j * text_len, text + (size_t)(j + 1) * text_len, text_full.begin() + (size_t)(r0 + j) * text_len);
    for (int k = 0; k < K; k++)
        std::copy(prompt + (size_t)k * prompt_len, prompt + (size_t)(k + 1) * prompt_len, prompt_full.begin() + ((size_t)utt * K + k) * ps);
    ssrb_lm_batch b{};
    b.n_utt = lm->n_utt; b.text = text_full.data(); b.text_stride = text_len; b.prompt = prompt_full.data(); b.prompt_stride = ps;
    UttState z{}; z.cfg_tag = 1; z.prev_token = -1; z.n_spans = n_spans; z.x_len = text_len; z.y_len = prompt_len; z.rng_id = rng_stream;
    std::vector<int> seq(rpu, text_len + prompt_len);
    SSRB_CUDA(cudaMemcpyAsync(lm->d_state + utt, &z, sizeof(UttState), cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaMemcpyAsync(lm->d_seq_len + r0, seq.data(), rpu * 4, cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaStreamSynchronize(s));
    std::vector<RowPlan> chunk; int tok = 0;
    for (int j = 0; j < rpu; j++) {
        const int len = text_len + prompt_len + 1;
        if (tok + len > lm->cfg.max_prefill_tokens) { SSRB_TRY(prefill_chunk(lm, chunk, &b, s, nullptr)); chunk.clear(); tok = 0; }
        chunk.push_back({r0 + j, utt, text_len, prompt_len, len}); tok += len;
    }
    if (!chunk.empty()) SSRB_TRY(prefill_chunk(lm, chunk, &b, s, nullptr));
    SSRB_TRY(heads_on_rows(lm, r0, rpu, s));
    SSRB_TRY(launch_sample(lm->logits, lm->d_state, lm->d_seq_len, lm->d_next_tok, lm->d_gen_tok, nullptr, lm->d_iter, lm->sp, s, utt));
    return 0;
}

int ssrb_lm_poll_flags(ssrb_lm* lm, void* stream, int32_t* done_flags, int* n_iter) {
    SSRB_CHECK(lm && lm->n_utt > 0 && done_flags, "no active batch");
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<UttState> st(lm->n_utt);
    int it = 0;
    SSRB_CUDA(cudaMemcpyAsync(st.data(), lm->d_state, lm->n_utt * sizeof(UttState), cudaMemcpyDeviceToHost, s));
    SSRB_CUDA(cudaMemcpyAsync(&it, lm->d_iter, 4, cudaMemcpyDeviceToHost, s));
    SSRB_CUDA(cudaStreamSynchronize(s));
    for (int u = 0; u < lm->n_utt; u++) done_flags[u] = st[u].done ? 1 : 0;
    if (n_iter) *n_iter = it;
    return 0;
}

int ssrb_lm_decode(ssrb_lm* lm, int n_steps, void* stream) {
    SSRB_CHECK(lm && lm->R > 0, "ssrb_lm_begin has not been called");
    SSRB_CUDA(cudaSetDevice(lm->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (!lm->use_graph) {
        for (int i = 0; i < n_steps; i++) SSRB_TRY(enqueue_step(lm, s));
        return 0;
    }
    if (!lm->graph) {
        cudaStream_t cs;
        SSRB_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t gr = nullptr;
        SSRB_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        const unsigned long long before = g_launch_count;
        const int rc = enqueue_step(lm, cs);
        lm->graph_kernels = g_launch_count - before;
        g_launch_count = before;                 // captured, not launched
        cudaError_t ce = cudaStreamEndCapture(cs, &gr);
        cudaStreamDestroy(cs);
        if (rc) { if (gr) cudaGraphDestroy(gr); return rc; }
        SSRB_CUDA(ce);
        SSRB_CUDA(cudaGraphInstantiate(&lm->graph, gr, 0));
        cudaGraphDestroy(gr);
    }
    for (int i = 0; i < n_steps; i++) SSRB_CUDA(cudaGraphLaunch(lm->graph, s));
    g_launch_count += lm->graph_kernels * (unsigned long long)n_steps;
    return 0;
}

int ssrb_lm_poll(ssrb_lm* lm, void* stream, int* n_done, int* n_iter) {
    SSRB_CHECK(lm && lm->n_utt > 0, "no active batch");
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<UttState> st(lm->n_utt);
    int it = 0;
    SSRB_CUDA(cudaMemcpyAsync(st.data(), lm->d_state, lm->n_utt * sizeof(UttState), cudaMemcpyDeviceToHost, s));
    SSRB_CUDA(cudaMemcpyAsync(&it, lm->d_iter, 4, cudaMemcpyDeviceToHost, s));
    SSRB_CUDA(cudaStreamSynchronize(s));
    int d = 0;
    for (auto& x : st) d += x.done ? 1 : 0;
    if (n_done) *n_done = d;
    if (n_iter) *n_iter = it;
    return 0;
}

int ssrb_lm_read_tokens(ssrb_lm* lm, void* stream, int utt, int32_t* out, int cap, int* n_tokens, int32_t* span_len) {
    SSRB_CHECK(lm && utt >= 0 && utt < lm->n_utt, "bad utterance index");
    cudaStream_t s = (cudaStream_t)stream;
    UttState st;
    SSRB_CUDA(cudaMemcpyAsync(&st, lm->d_state + utt, sizeof(UttState), cudaMemcpyDeviceToHost, s));
    SSRB_CUDA(cudaStreamSynchronize(s));
    SSRB_CHECK(st.n_tok <= cap, "token buffer too small");
    SSRB_CUDA(cudaMemcpyAsync(out, lm->d_gen_tok + (size_t)utt * lm->cfg.max_steps * lm->K, (size_t)st.n_tok * lm->K * 4,
                              cudaMemcpyDeviceToHost, s));
    SSRB_CUDA(cudaStreamSynchronize(s));
    if (n_tokens) *n_tokens = st.n_tok;
    if (span_len) for (int i = 0; i < SSRB_MAX_SPANS; i++) span_len[i] = st.span_len[i];
    return 0;
}

int ssrb_lm_decode_path(ssrb_lm* lm) {
    if (!lm || lm->R <= 0) return -1;
    if (lm->mega) return 3;
    if (lm->fold) return lm->layer_kernel ? 2 : 1;
    return 0;
}

int ssrb_lm_read_logits(ssrb_lm* lm, void* stream, float* host_out) {
    SSRB_CHECK(lm && lm->R > 0 && host_out, "no active batch");
    cudaStream_t s = (cudaStream_t)stream;
    SSRB_CUDA(cudaMemcpyAsync(host_out, lm->logits, (size_t)lm->R * lm->K * lm->V * 4, cudaMemcpyDeviceToHost, s));
    SSRB_CUDA(cudaStreamSynchronize(s));
    return 0;
}

// grow-only scratch for the two entry points below (a cudaMalloc / cudaFree pair per buffer and call costs more than the forward
// itself: measured 50 ms per 600-position utterance against 4.3 ms of kernels)
static int tf_reserve(ssrb_lm* lm, size_t bytes) {
    if (bytes <= lm->tf_cap) return 0;
    if (lm->tf_buf) { SSRB_CUDA(cudaDeviceSynchronize()); cudaFree(lm->tf_buf); lm->tf_buf = nullptr; lm->tf_cap = 0; }
    const size_t cap = bytes + bytes / 2;
    SSRB_TRY(dev_alloc((void**)&lm->tf_buf, cap));
    lm->tf_cap = cap;
    return 0;
}
struct TfScratch { void* hl; void* hhb; float* lg; int* idx; int* aud; unsigned char* flags; unsigned char* hit; float* nll; double* out; };
static int tf_carve(ssrb_lm* lm, int Ty, TfScratch* t) {
    const size_t K = lm->K, V = lm->V, Hh = lm->Hh, D = lm->D, n = (size_t)Ty;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t sz[9] = {up(n * D * lm->esz), up(n * K * Hh * lm->esz), up(n * K * V * 4), up(n * 4), up(K * n * 4), up(K * n), up(K * n),
                          up(K * n * 4), up(K * 4 * 8)};
    size_t tot = 0;
    for (size_t b : sz) tot += b;
    SSRB_TRY(tf_reserve(lm, tot));
    char* p = lm->tf_buf;
    t->hl = p; p += sz[0]; t->hhb = p; p += sz[1]; t->lg = (float*)p; p += sz[2]; t->idx = (int*)p; p += sz[3];
    t->aud = (int*)p; p += sz[4]; t->flags = (unsigned char*)p; p += sz[5]; t->hit = (unsigned char*)p; p += sz[6];
    t->nll = (float*)p; p += sz[7]; t->out = (double*)p;
    return 0;
}

// teacher forcing through the prefill path: logits [Ty, K, V] fp32 for every audio position, left on the device in t->lg
static int teacher_forced_device(ssrb_lm* lm, const int32_t* text, int Lx, const int32_t* audio, int Ty, cudaStream_t s, TfScratch* t) {
    SSRB_CHECK(Lx + Ty <= lm->cfg.max_prefill_tokens && Lx + Ty <= lm->cfg.max_seq, "sequence too long for teacher forcing");
    SSRB_CHECK(Ty <= lm->n_pos && Lx <= lm->n_pos, "PE table too small");
    SSRB_TRY(tf_carve(lm, Ty, t));
    ssrb_lm_batch b{};
    b.n_utt = 1; b.text = text; b.text_stride = Lx;
    std::vector<int> aud(audio, audio + (size_t)lm->K * Ty);
    std::vector<RowPlan> rows = {{0, 0, Lx, Ty, Lx + Ty}};
    lm->rpu = 1;
    SSRB_TRY(prefill_chunk(lm, rows, &b, s, &aud));
    // heads over all audio positions
    std::vector<int> hidx(Ty);
    for (int i = 0; i < Ty; i++) hidx[i] = Lx + i;
    SSRB_CUDA(cudaMemcpyAsync(t->idx, hidx.data(), Ty * 4, cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaStreamSynchronize(s));                   // hidx is a stack-lifetime host buffer
    return run_heads(lm, t->idx, Ty, t->hl, t->hhb, t->lg, s);
}

int ssrb_lm_teacher_forced(ssrb_lm* lm, const int32_t* text, int Lx, const int32_t* audio, int Ty, float* host_logits,
                           void* stream) {
    SSRB_CHECK(lm && text && audio && host_logits, "null argument");
    SSRB_CUDA(cudaSetDevice(lm->device));
    SSRB_TRY(ssrb_lm_check_loaded(lm));
    cudaStream_t s = (cudaStream_t)stream;
    TfScratch t{};
    SSRB_TRY(teacher_forced_device(lm, text, Lx, audio, Ty, s, &t));
    SSRB_CUDA(cudaMemcpyAsync(host_logits, t.lg, (size_t)Ty * lm->K * lm->V * 4, cudaMemcpyDeviceToHost, s));
    SSRB_CUDA(cudaStreamSynchronize(s));
    return 0;
}

// SSR_Speech.forward of ONE utterance at its own length (models/ssr.py:280-379; padded key positions are masked in the reference,
// so a batch is the sum over its utterances): forward over [text ; audio], heads on every audio position, then the masked
// per-codebook cross entropy / top-10 accuracy on the device.  out [K][4] = {sum of nll, positions in the loss, top-10 hits,
// token count}; the wrapper combines utterances and codebooks as ssr.py:352-372 does.
int ssrb_lm_forward_loss(ssrb_lm* lm, const int32_t* text, int Lx, const int32_t* audio, int Ty, const uint8_t* flags, double* out,
                         void* stream) {
    SSRB_CHECK(lm && text && audio && flags && out, "null argument");
    SSRB_CHECK(Ty >= 2, "forward_loss: at least two audio positions");
    SSRB_CUDA(cudaSetDevice(lm->device));
    SSRB_TRY(ssrb_lm_check_loaded(lm));
    cudaStream_t s = (cudaStream_t)stream;
    const int K = lm->K, V = lm->V, n = Ty - 1;
    for (size_t i = 0; i < (size_t)K * Ty; i++) SSRB_CHECK(audio[i] >= 0 && audio[i] < V, "forward_loss: audio token outside the embedding table");
    TfScratch t{};
    SSRB_TRY(teacher_forced_device(lm, text, Lx, audio, Ty, s, &t));
    SSRB_CUDA(cudaMemcpyAsync(t.aud, audio, (size_t)K * Ty * 4, cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaMemcpyAsync(t.flags, flags, (size_t)K * n, cudaMemcpyHostToDevice, s));
    SSRB_TRY(launch_masked_ce(t.lg, t.aud, t.flags, Ty, K, V, t.nll, t.hit, t.out, s));
    SSRB_CUDA(cudaMemcpyAsync(out, t.out, (size_t)K * 4 * 8, cudaMemcpyDeviceToHost, s));
    SSRB_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int ssrb_lm_step_bytes(ssrb_lm* lm, void* stream, double* weight_bytes, double* kv_bytes) {
    SSRB_CHECK(lm && lm->n_utt > 0, "no active batch");
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<UttState> st(lm->n_utt); std::vector<int> seq(lm->R);
    SSRB_CUDA(cudaMemcpyAsync(st.data(), lm->d_state, lm->n_utt * sizeof(UttState), cudaMemcpyDeviceToHost, s));
    SSRB_CUDA(cudaMemcpyAsync(seq.data(), lm->d_seq_len, lm->R * 4, cudaMemcpyDeviceToHost, s));
    SSRB_CUDA(cudaStreamSynchronize(s));
    const double D = lm->D, F = lm->F, K = lm->K, V = lm->V, Hh = lm->Hh, L = lm->L, e = (double)lm->esz;
    // W_step = per-layer matrices + biases + LN + final LN + heads, streamed once per step
    const double per_layer = (3 * D * D + D * D + 2 * D * F) * e + (3 * D + D + F + D + 4 * D) * 4.0;
    const double heads = (K * Hh * D + K * V * Hh) * e + (K * Hh + K * V) * 4.0;
    if (weight_bytes) *weight_bytes = L * per_layer + 2 * D * 4.0 + heads;
    double kv = 0;
    for (int r = 0; r < lm->R; r++)
        if (!st[r / lm->rpu].done) kv += L * 2.0 * D * e * ((double)seq[r] + 1.0 /*read S+1*/ + 1.0 /*write 1*/);
    if (kv_bytes) *kv_bytes = kv;
    return 0;
}

int ssrb_lm_profile_steps(ssrb_lm* lm, int n_steps, void* stream, double* ms_by_class, double* total_ms) {
    SSRB_CHECK(lm && lm->R > 0 && n_steps > 0, "no active batch");
    SSRB_CUDA(cudaSetDevice(lm->device));
    cudaStream_t s = (cudaStream_t)stream;
    cudaEvent_t t0, t1;
    SSRB_CUDA(cudaEventCreate(&t0)); SSRB_CUDA(cudaEventCreate(&t1));
    lm->prof = true; lm->prof_ev.clear();
    SSRB_CUDA(cudaEventRecord(t0, s));
    int rc = 0;
    for (int i = 0; i < n_steps && !rc; i++) rc = enqueue_step(lm, s);
    cudaEventRecord(t1, s);
    lm->prof = false;
    cudaError_t ce = cudaStreamSynchronize(s);
    double acc[PC_COUNT] = {0, 0, 0, 0};
    for (auto& t : lm->prof_ev) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, std::get<1>(t), std::get<2>(t));
        acc[std::get<0>(t)] += ms;
        cudaEventDestroy(std::get<1>(t)); cudaEventDestroy(std::get<2>(t));
    }
    lm->prof_ev.clear();
    float tot = 0.f;
    cudaEventElapsedTime(&tot, t0, t1);
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    if (rc) return rc;
    SSRB_CUDA(ce);
    if (ms_by_class) for (int i = 0; i < PC_COUNT; i++) ms_by_class[i] = acc[i] / n_steps;
    if (total_ms) *total_ms = tot / n_steps;
    return 0;
}

int ssrb_debug_timeline(unsigned long long* dev_buf, unsigned int* dev_idx, unsigned int cap) {
    TsBuf t{dev_buf, dev_idx, cap};
    SSRB_CHECK(!ts_arm_gemm_tc(t) && !ts_arm_attn_tma(t) && !ts_arm_lm_kernels(t) && !ts_arm_gemm_layer(t) && !ts_arm_gemm_flat2(t), "cudaMemcpyToSymbol failed");
    return 0;
}

int ssrb_debug_mega_trace(unsigned long long* dev_buf, int cap_per_cta, int* n_ctas) {
    int G = 0;
    SSRB_TRY(mega_grid(&G));
    if (n_ctas) *n_ctas = G;
    return mega_trace_arm(dev_buf, dev_buf ? cap_per_cta : 0);
}

int ssrb_op_gemm(const void* A, const void* W, const float* bias, const float* residual, float* C, int M, int N, int K,
                 int dtype, int act, int impl, void* stream) {
    GemmArgs g;
    g.A = A; g.lda = K; g.W = W; g.ldw = K; g.bias = bias; g.residual = residual; g.ldr = N; g.C = C; g.ldc = N;
    g.M = M; g.N = N; g.K = K; g.act = act; g.ab_dtype = dtype; g.c_dtype = SSRB_DTYPE_F32;
    if (impl == 2) {
        SSRB_CHECK(dtype == SSRB_DTYPE_BF16 && gemm_tc_supported(g), "tcgen05 GEMM does not support this problem");
        static void* ws = nullptr; static size_t wsb = 0;
        if (!ws) { wsb = gemm_tc_workspace_bytes(256, 16384); SSRB_CUDA(cudaMalloc(&ws, wsb)); SSRB_CUDA(cudaMemset(ws, 0, wsb)); }
        return gemm_tc(g, ws, wsb, (cudaStream_t)stream);
    }
    return gemm_simt(g, (cudaStream_t)stream);
}

int ssrb_op_gemm_ln(const void* A1, const void* W1, const float* bias1, const float* residual, float* X_out, int M, int D,
                    int K1, const void* W2, const float* gamma, const float* beta, const float* bias2, float* C_out, int N,
                    int act, void* stream) {
    SSRB_CHECK(A1 && W1 && X_out && W2 && gamma && beta && C_out, "null argument");
    SSRB_CHECK(D % 128 == 0 && D <= 2048 && M >= 1 && M <= 128, "op_gemm_ln: d_model must be a multiple of 128 <= 2048, M <= 128");
    cudaStream_t s = (cudaStream_t)stream;
    void *xb = nullptr, *w2f = nullptr; float *cs = nullptr, *bf = nullptr; float2* part = nullptr;
    SSRB_CUDA(cudaMalloc(&xb, (size_t)M * D * 2)); SSRB_CUDA(cudaMalloc(&w2f, (size_t)N * D * 2));
    SSRB_CUDA(cudaMalloc((void**)&cs, (size_t)N * 4)); SSRB_CUDA(cudaMalloc((void**)&bf, (size_t)N * 4));
    SSRB_CUDA(cudaMalloc((void**)&part, (size_t)M * (D / 128) * sizeof(float2)));
    int rc = launch_fold_ln(W2, N, D, gamma, beta, bias2, w2f, cs, bf, s);
    GemmArgs g;
    g.A = A1; g.lda = K1; g.W = W1; g.ldw = K1; g.bias = bias1; g.residual = residual; g.ldr = D; g.C = X_out; g.ldc = D;
    g.M = M; g.N = D; g.K = K1; g.ab_dtype = SSRB_DTYPE_BF16; g.c_dtype = SSRB_DTYPE_F32;
    g.C2 = xb; g.ldc2 = D; g.part_out = part; g.part_ld = M;
    if (!rc) rc = gemm_tc_supported(g) ? gemm_tc(g, nullptr, 0, s) : 1;
    g = GemmArgs();
    g.A = xb; g.lda = D; g.W = w2f; g.ldw = D; g.bias = bf; g.C = C_out; g.ldc = N;
    g.M = M; g.N = N; g.K = D; g.act = act; g.ab_dtype = SSRB_DTYPE_BF16; g.c_dtype = SSRB_DTYPE_F32;
    g.ln_part = part; g.part_ld = M; g.ln_blocks = D / 128; g.ln_colsum = cs;
    if (!rc) rc = gemm_tc_supported(g) ? gemm_tc(g, nullptr, 0, s) : 1;
    cudaError_t ce = cudaStreamSynchronize(s);
    cudaFree(xb); cudaFree(w2f); cudaFree(cs); cudaFree(bf); cudaFree(part);
    if (rc) return rc;
    SSRB_CUDA(ce);
    return 0;
}

// One layer's GEMM chain between two attention kernels, on caller-provided buffers: impl 0 = the four per-GEMM launches of
// run_layer_fold, impl 1 = the persistent layer kernel (gemm_layer.cu).  Same folded weights, same buffers, same outputs.
int ssrb_op_layer_chain(const void* ao, float* x_inout, const void* wo, const float* bo, const void* w1, const float* b1,
                        const float* gamma2, const float* beta2, const void* w2, const float* b2, const void* wqkv,
                        const float* bqkv, const float* gamma1n, const float* beta1n, void* hid_out, float* qkv_out, int M,
                        int D, int F, int impl, void* stream) {
    SSRB_CHECK(ao && x_inout && wo && bo && w1 && b1 && gamma2 && beta2 && w2 && b2 && hid_out, "null argument");
    SSRB_CHECK(!wqkv || (bqkv && gamma1n && beta1n && qkv_out), "the QKV phase needs bias, LayerNorm parameters and an output");
    SSRB_CHECK(M >= 1 && M <= 128 && D % 128 == 0 && D <= 2048 && F % 128 == 0, "op_layer_chain: M <= 128, d_model % 128 == 0 <= 2048");
    cudaStream_t s = (cudaStream_t)stream;
    void *hn = nullptr, *w1f = nullptr, *wqf = nullptr; float *c1 = nullptr, *b1f = nullptr, *cq = nullptr, *bqf = nullptr;
    float2* part = nullptr; unsigned int* gbar = nullptr;
    SSRB_CUDA(cudaMalloc(&hn, (size_t)M * D * 2)); SSRB_CUDA(cudaMalloc(&w1f, (size_t)F * D * 2));
    SSRB_CUDA(cudaMalloc((void**)&c1, (size_t)F * 4)); SSRB_CUDA(cudaMalloc((void**)&b1f, (size_t)F * 4));
    SSRB_CUDA(cudaMalloc((void**)&part, (size_t)M * (D / 128) * sizeof(float2)));
    SSRB_CUDA(cudaMalloc((void**)&gbar, 64 * 4)); SSRB_CUDA(cudaMemsetAsync(gbar, 0, 64 * 4, s));
    int rc = launch_fold_ln(w1, F, D, gamma2, beta2, b1, w1f, c1, b1f, s);
    if (wqkv) {
        SSRB_CUDA(cudaMalloc(&wqf, (size_t)3 * D * D * 2));
        SSRB_CUDA(cudaMalloc((void**)&cq, (size_t)3 * D * 4)); SSRB_CUDA(cudaMalloc((void**)&bqf, (size_t)3 * D * 4));
        if (!rc) rc = launch_fold_ln(wqkv, 3 * D, D, gamma1n, beta1n, bqkv, wqf, cq, bqf, s);
    }
    if (!rc && impl == 1) {
        if (!gemm_layer_supported(M, D, F)) { set_error("op_layer_chain: the persistent layer kernel does not support this shape / device"); rc = 1; }
        LayerChainArgs a;
        a.M = M; a.D = D; a.F = F; a.ao = ao; a.x = x_inout; a.hn = hn; a.hid = hid_out; a.qkv = qkv_out;
        a.ln_part = part; a.part_ld = M;
        a.wo = wo; a.bo = bo; a.w1f = w1f; a.b1f = b1f; a.c1 = c1; a.w2 = w2; a.b2 = b2;
        if (wqkv) { a.wqkv_next = wqf; a.bqkv_next = bqf; a.cqkv_next = cq; }
        a.gbar = gbar;
        if (!rc) rc = gemm_layer(a, s);
    } else if (!rc) {
        GemmArgs g;
        g.A = ao; g.lda = D; g.W = wo; g.ldw = D; g.bias = bo; g.residual = x_inout; g.ldr = D; g.C = x_inout; g.ldc = D;
        g.M = M; g.N = D; g.K = D; g.ab_dtype = SSRB_DTYPE_BF16; g.c_dtype = SSRB_DTYPE_F32;
        g.C2 = hn; g.ldc2 = D; g.part_out = part; g.part_ld = M;
        rc = gemm_tc_supported(g) ? gemm_tc(g, nullptr, 0, s) : 1;
        g = GemmArgs();
        g.A = hn; g.lda = D; g.W = w1f; g.ldw = D; g.bias = b1f; g.C = hid_out; g.ldc = F;
        g.M = M; g.N = F; g.K = D; g.act = ACT_RELU; g.ab_dtype = SSRB_DTYPE_BF16; g.c_dtype = SSRB_DTYPE_BF16;
        g.ln_part = part; g.part_ld = M; g.ln_blocks = D / 128; g.ln_colsum = c1;
        if (!rc) rc = gemm_tc_supported(g) ? gemm_tc(g, nullptr, 0, s) : 1;
        g = GemmArgs();
        g.A = hid_out; g.lda = F; g.W = w2; g.ldw = F; g.bias = b2; g.residual = x_inout; g.ldr = D; g.C = x_inout; g.ldc = D;
        g.M = M; g.N = D; g.K = F; g.ab_dtype = SSRB_DTYPE_BF16; g.c_dtype = SSRB_DTYPE_F32;
        g.C2 = hn; g.ldc2 = D; g.part_out = part; g.part_ld = M;
        if (!rc) rc = gemm_tc_supported(g) ? gemm_tc(g, nullptr, 0, s) : 1;
        if (wqkv) {
            g = GemmArgs();
            g.A = hn; g.lda = D; g.W = wqf; g.ldw = D; g.bias = bqf; g.C = qkv_out; g.ldc = 3 * D;
            g.M = M; g.N = 3 * D; g.K = D; g.ab_dtype = SSRB_DTYPE_BF16; g.c_dtype = SSRB_DTYPE_F32;
            g.ln_part = part; g.part_ld = M; g.ln_blocks = D / 128; g.ln_colsum = cq;
            if (!rc) rc = gemm_tc_supported(g) ? gemm_tc(g, nullptr, 0, s) : 1;
        }
    }
    cudaError_t ce = cudaStreamSynchronize(s);
    cudaFree(hn); cudaFree(w1f); cudaFree(wqf); cudaFree(c1); cudaFree(b1f); cudaFree(cq); cudaFree(bqf); cudaFree(part); cudaFree(gbar);
    if (rc) return rc;
    SSRB_CUDA(ce);
    return 0;
}

// Decode attention exactly as the bf16 decode chain runs it (attn_decode_tma_kernel incl. the fused in-place KV append), on
// caller-provided buffers.  Test hook for tests/test_gpu_attn_ops.py.
int ssrb_op_attn_decode(const float* qkv, void* kcache, void* vcache, const int32_t* seq_len_host, const int32_t* done_host,
                        void* out, int R, int D, int H, int Smax, void* stream) {
    SSRB_CHECK(qkv && kcache && vcache && seq_len_host && out, "null argument");
    SSRB_CHECK(R >= 1 && H >= 1 && D == H * 128 && Smax >= 1, "op_attn_decode: head_dim must be 128");
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<UttState> st(R);
    for (int r = 0; r < R; r++) {
        SSRB_CHECK(seq_len_host[r] >= 0 && seq_len_host[r] < Smax, "op_attn_decode: seq_len must leave room for the appended row");
        UttState z{}; z.done = done_host ? done_host[r] : 0; st[r] = z;
    }
    UttState* d_st = nullptr; int *d_seq = nullptr, *d_tk = nullptr; float* ws = nullptr;
    SSRB_CUDA(cudaMalloc((void**)&d_st, R * sizeof(UttState))); SSRB_CUDA(cudaMalloc((void**)&d_seq, R * 4));
    SSRB_CUDA(cudaMalloc((void**)&d_tk, (size_t)R * H * 4)); SSRB_CUDA(cudaMalloc((void**)&ws, attn_decode_ws_floats(R, H, Smax) * 4));
    SSRB_CUDA(cudaMemcpyAsync(d_st, st.data(), R * sizeof(UttState), cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaMemcpyAsync(d_seq, seq_len_host, R * 4, cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaMemsetAsync(d_tk, 0, (size_t)R * H * 4, s));
    int rc = launch_attn_decode(qkv, R, D, H, kcache, vcache, SSRB_DTYPE_BF16, Smax, d_seq, d_st, 1, ws, d_tk, out,
                                SSRB_DTYPE_BF16, 0, s);
    cudaError_t ce = cudaStreamSynchronize(s);
    cudaFree(d_st); cudaFree(d_seq); cudaFree(d_tk); cudaFree(ws);
    if (rc) return rc;
    SSRB_CUDA(ce);
    return 0;
}

// Prefill attention exactly as the bf16 prefill runs it: K/V of the packed positions are scattered into the cache
// (kv_append_kernel), then attn_prefill_mma_kernel attends causally within each row.  Row i of the call uses cache row i.
int ssrb_op_attn_prefill(const float* qkv, void* kcache, void* vcache, const int32_t* row_len_host, int n_rows, void* out,
                         int D, int H, int Smax, void* stream) {
    SSRB_CHECK(qkv && kcache && vcache && row_len_host && out && n_rows >= 1, "null argument");
    SSRB_CHECK(H >= 1 && D == H * 128, "op_attn_prefill: head_dim must be 128");
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<int> rows, slots, rid, rstart, rlen;
    int M = 0, max_len = 0;
    for (int i = 0; i < n_rows; i++) {
        const int len = row_len_host[i];
        SSRB_CHECK(len >= 1 && len <= Smax, "op_attn_prefill: bad row length");
        rid.push_back(i); rstart.push_back(M); rlen.push_back(len);
        for (int j = 0; j < len; j++) { rows.push_back(i); slots.push_back(j); }
        M += len; if (len > max_len) max_len = len;
    }
    int *d_rows = nullptr, *d_slots = nullptr, *d_rid = nullptr, *d_rstart = nullptr, *d_rlen = nullptr;
    SSRB_CUDA(cudaMalloc((void**)&d_rows, M * 4)); SSRB_CUDA(cudaMalloc((void**)&d_slots, M * 4));
    SSRB_CUDA(cudaMalloc((void**)&d_rid, n_rows * 4)); SSRB_CUDA(cudaMalloc((void**)&d_rstart, n_rows * 4));
    SSRB_CUDA(cudaMalloc((void**)&d_rlen, n_rows * 4));
    SSRB_CUDA(cudaMemcpyAsync(d_rows, rows.data(), M * 4, cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaMemcpyAsync(d_slots, slots.data(), M * 4, cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaMemcpyAsync(d_rid, rid.data(), n_rows * 4, cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaMemcpyAsync(d_rstart, rstart.data(), n_rows * 4, cudaMemcpyHostToDevice, s));
    SSRB_CUDA(cudaMemcpyAsync(d_rlen, rlen.data(), n_rows * 4, cudaMemcpyHostToDevice, s));
    int rc = launch_kv_append(qkv, M, D, H, d_rows, d_slots, nullptr, kcache, vcache, SSRB_DTYPE_BF16, Smax, s);
    if (!rc) rc = launch_attn_prefill(qkv, D, H, kcache, vcache, SSRB_DTYPE_BF16, Smax, n_rows, d_rid, d_rstart, d_rlen, max_len,
                                      out, SSRB_DTYPE_BF16, s);
    cudaError_t ce = cudaStreamSynchronize(s);
    cudaFree(d_rows); cudaFree(d_slots); cudaFree(d_rid); cudaFree(d_rstart); cudaFree(d_rlen);
    if (rc) return rc;
    SSRB_CUDA(ce);
    return 0;
}
