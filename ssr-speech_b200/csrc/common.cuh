// common.cuh — shared helpers for libssr_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

namespace ssrb {

typedef __nv_bfloat16 bf16;

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const std::string& msg);
extern unsigned long long g_launch_count;   // kernels launched by this library (bench gpu_launches)

#define SSRB_CUDA(expr)                                                                           \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ::ssrb::set_error(std::string(#expr) + " -> " + cudaGetErrorString(_e) + " at " +     \
                              __FILE__ + ":" + std::to_string(__LINE__));                         \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

#define SSRB_CHECK(cond, msg)                                                                     \
    do {                                                                                          \
        if (!(cond)) {                                                                            \
            ::ssrb::set_error(std::string(msg) + " (" #cond ") at " + __FILE__ + ":" +            \
                              std::to_string(__LINE__));                                          \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

#define SSRB_TRY(expr)                                                                            \
    do {                                                                                          \
        int _r = (expr);                                                                          \
        if (_r) return _r;                                                                        \
    } while (0)

// every kernel launch goes through this so the launch count is honest
#define SSRB_LAUNCH(kernel, grid, block, smem, stream, ...)                                       \
    do {                                                                                          \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                               \
        ::ssrb::g_launch_count++;                                                                 \
        SSRB_CUDA(cudaGetLastError());                                                            \
    } while (0)

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------
// Kernels of the decode chain are launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel may start
// while its predecessor is still running, so EVERY kernel calls pdl_wait() before touching anything a predecessor
// writes (only immutable weights may be prefetched earlier) and pdl_launch_dependents() as early as possible.
// Both are no-ops when the kernel was launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();     // SSRB_NO_PDL=1 disables the launch attribute (lm_engine.cu)

template <typename... KArgs, typename... Args>
int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_y, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (cluster_y > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = (unsigned)cluster_y; attr[na].val.clusterDim.z = 1;
        na++;
    }
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        na++;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    SSRB_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
    g_launch_count++;
    return 0;
}

// ---- poor man's timeline (no nsys in this image): when armed (ssrb_debug_timeline), thread 0 of every CTA of the decode-chain
// kernels records of 10 u64 {kernel id, linear CTA id, entry, after griddepcontrol.wait, exit, aux0..aux4}; tools/timeline.py reconstructs the
// overlap between kernels of the PDL chain from the dump.
struct TsBuf { unsigned long long* buf; unsigned int* idx; unsigned int cap; };
static __device__ TsBuf g_ts = {nullptr, nullptr, 0};      // one copy per translation unit (no -rdc); armed by ts_arm_tu()
static inline int ts_arm_tu(const TsBuf& t) { return cudaMemcpyToSymbol(g_ts, &t, sizeof(t)) == cudaSuccess ? 0 : 1; }
int ts_arm_gemm_tc(const TsBuf& t); int ts_arm_attn_tma(const TsBuf& t); int ts_arm_lm_kernels(const TsBuf& t);
int ts_arm_gemm_layer(const TsBuf& t); int ts_arm_gemm_flat2(const TsBuf& t);
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int ts_begin(int kernel_id) {
    if (g_ts.buf == nullptr || threadIdx.x != 0 || threadIdx.y != 0) return -1;
    const unsigned int slot = atomicAdd(g_ts.idx, 1u);
    if (slot >= g_ts.cap) return -1;
    g_ts.buf[slot * 10 + 0] = (unsigned long long)kernel_id;
    g_ts.buf[slot * 10 + 1] = (unsigned long long)(blockIdx.x + gridDim.x * (blockIdx.y + (unsigned long long)gridDim.y * blockIdx.z));
    g_ts.buf[slot * 10 + 2] = globaltimer_ns();
    for (int i = 3; i < 10; i++) g_ts.buf[slot * 10 + i] = 0;
    return (int)slot;
}
__device__ __forceinline__ void ts_dep(int slot) { if (slot >= 0) g_ts.buf[slot * 10 + 3] = globaltimer_ns(); }
__device__ __forceinline__ void ts_aux(int slot, int i = 0) { if (slot >= 0) g_ts.buf[slot * 10 + 5 + i] = globaltimer_ns(); }
__device__ __forceinline__ void ts_end(int slot) { if (slot >= 0) g_ts.buf[slot * 10 + 4] = globaltimer_ns(); }
enum TsKernel { TSK_EMBED = 1, TSK_LN = 2, TSK_GEMM = 3, TSK_ATTN = 4, TSK_SAMPLE = 5 };

// launch of a kernel that belongs to a PDL chain (the kernel itself calls pdl_launch_dependents()/pdl_wait())
#define SSRB_LAUNCH_PDL(kernel, grid, block, smem, stream, ...)                                   \
    SSRB_TRY(::ssrb::launch_pdl(kernel, dim3(grid), dim3(block), (smem), (stream), 1, __VA_ARGS__))

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 8 consecutive elements -> fp32 registers (16 B for bf16, 32 B for fp32); pointer must be 16B aligned
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
    uint2 raw = *reinterpret_cast<const uint2*>(p);
    v[0] = __uint_as_float(raw.x << 16); v[1] = __uint_as_float(raw.x & 0xffff0000u);
    v[2] = __uint_as_float(raw.y << 16); v[3] = __uint_as_float(raw.y & 0xffff0000u);
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }
// ELU for results that are stored as bf16: exp(x) - 1 through MUFU.EX2 (absolute error ~2.4e-7, i.e. <= 2.4e-4 relative for
// x <= -1e-3) and the series x + x^2/2 above that (error x^3/6 <= 1.7e-10): both far inside half a bf16 ulp (2e-3 relative),
// at 4 instructions instead of expm1f's ~25 — the tensor-core conv epilogues are instruction-bound on exactly this.
__device__ __forceinline__ float elu1_bf16(float x) { return x > 0.f ? x : (x > -1e-3f ? fmaf(0.5f * x, x, x) : __expf(x) - 1.f); }

// ---- GEMM interface (gemm_simt.cu / gemm_tc.cu) -----------------------------------------------------
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

struct GemmArgs {
    const void* A = nullptr;  int64_t lda = 0;          // [M, K] activations, dtype a_dtype
    const void* W = nullptr;  int64_t ldw = 0;          // [N, K] weights, dtype w_dtype (same as A)
    const float* bias = nullptr;                        // [N] or null
    const float* residual = nullptr; int64_t ldr = 0;   // fp32 [M, N] or null (added after activation)
    void* C = nullptr;        int64_t ldc = 0;          // [M, N], dtype c_dtype
    int M = 0, N = 0, K = 0;
    int act = ACT_NONE;
    int ab_dtype = 0;                                   // SSRB_DTYPE_* of A and W
    int c_dtype = 0;                                    // SSRB_DTYPE_* of C
    // grouped GEMM (blockIdx.z): element strides between groups
    int groups = 1;
    int64_t a_gs = 0, w_gs = 0, bias_gs = 0, c_gs = 0;
    // ---- LayerNorm folded into the decode GEMMs (tcgen05 swap-AB path only, M <= 128, groups == 1) -----------------
    // consumer: A holds bf16(x) (NOT normalised), W holds bf16(gamma_k * W_nk), bias holds b_n + sum_k beta_k W_nk and
    //   C = act(rstd_r * (A.W^T - mean_r * ln_colsum_n) + bias) (+ residual); the row statistics are combined in a fixed
    //   order from ln_blocks per-128-column partials {mean, M2} written by the kernel that produced x.
    const float2* ln_part = nullptr; int ln_blocks = 0; const float* ln_colsum = nullptr; float ln_eps = 1e-5f;
    int part_ld = 0;                                    // rows per block of the partial buffer, laid out [block][part_ld]
    // producer: besides C (fp32, the residual stream) also write C2 = bf16(C) and the {mean, M2} partial of every
    //   (row, 128-column block) into part_out[block * part_ld + row]
    void* C2 = nullptr; int64_t ldc2 = 0; float2* part_out = nullptr;
};

int gemm_simt(const GemmArgs& g, cudaStream_t stream);
// tcgen05 path: bf16 A/W only.  `workspace` (device, >= gemm_tc_workspace_bytes) is needed for split-K.
int gemm_tc(const GemmArgs& g, void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t gemm_tc_workspace_bytes(int max_rows_decode, int max_n);
bool gemm_tc_supported(const GemmArgs& g);

// ---- prefill GEMM on CTA pairs (gemm_flat2.cu; the default since round 2, SSRB_FLAT_2CTA=0 opts out): tcgen05.mma.cta_group::2, 256 x 256 tiles
bool gemm_flat2_enabled();
bool gemm_flat2_supported(const GemmArgs& g);
int gemm_flat2(const GemmArgs& g, cudaStream_t stream);

// ---- persistent per-layer decode GEMM chain (gemm_layer.cu; experimental, SSRB_LAYER_KERNEL=1) ---------------------------
// One launch runs the GEMMs between two attention kernels of a decode iteration with the LayerNorms folded exactly as the
// per-GEMM chain does (run_layer_fold in lm_engine.cu):
//   phase 0  x += ao . Wo^T + bo                      (writes fp32 x, bf16(x) -> hn, row statistics -> ln_part)
//   phase 1  hid = relu(LN2(x) . W1^T + b1)           (bf16)
//   phase 2  x += hid . W2^T + b2                     (writes x, hn, ln_part)
//   phase 3  qkv = LN1'(x) . Wqkv'^T + bqkv'          (fp32; the NEXT layer's projection; skipped when wqkv_next == null)
struct LayerChainArgs {
    int M = 0, D = 0, F = 0;                            // rows (<= 128), d_model, ffn width
    const void* ao = nullptr;                           // bf16 [M, D] attention output
    float* x = nullptr;                                 // fp32 [M, D] residual stream (in / out)
    void* hn = nullptr;                                 // bf16 [M, D]
    void* hid = nullptr;                                // bf16 [M, F]
    float* qkv = nullptr;                               // fp32 [M, 3D]
    float2* ln_part = nullptr; int part_ld = 0;         // [D / 128][part_ld] {mean, M2}
    float ln_eps = 1e-5f;
    const void *wo = nullptr, *w1f = nullptr, *w2 = nullptr, *wqkv_next = nullptr;     // bf16, LayerNorm gamma folded (w1f, wqkv_next)
    const float *bo = nullptr, *b1f = nullptr, *c1 = nullptr, *b2 = nullptr, *bqkv_next = nullptr, *cqkv_next = nullptr;
    unsigned int* gbar = nullptr;                       // 64 zero-initialised words: grid barrier {count @0, generation @32}
};
bool gemm_layer_supported(int M, int D, int F);         // shape + co-residency of the persistent grid on this device
int gemm_layer(const LayerChainArgs& a, cudaStream_t stream);

}  // namespace ssrb
