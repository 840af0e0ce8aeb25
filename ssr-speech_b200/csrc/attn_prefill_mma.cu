// attn_prefill_mma.cu — bf16 causal prefill attention on the tensor cores (flash-attention tiling, mma.sync m16n8k16).
//
// Replaces F.scaled_dot_product_attention (models/modules/activation.py:634) for the first dec_forward of
// SSR_Speech.inference (models/ssr.py:673-684): every position of the packed prompt [text ; audio] attends causally
// (the reference's mask is exactly triu(ones(S,S),1): models/ssr.py:227-257).
//
// grid (ceil(max_len/64), H, n_rows), 128 threads: each warp owns 16 query rows; K/V tiles of 64 keys are staged from the
// in-place KV cache with cp.async (double buffered, 272-byte row pitch: conflict-free fragment loads); S = QK^T and O += PV
// on tensor cores with fp32 accumulators, online softmax in registers (exp2 with the 1/sqrt(128)*log2e scale folded into
// Q), P is re-used from the S accumulators as the A operand of PV.  This is prefill only (one shot per batch, ~1.6 TFLOP
// for the benchmark batch); the legacy mma.sync path is ample here — the tcgen05 pipelines carry the GEMMs.
#include "lm_kernels.cuh"

namespace ssrb {

namespace {

constexpr int PQ = 64, PK = 64, PITCH = 136;                 // bf16 elements per smem row (128 + 8 pad)
constexpr int TILE_ELEMS = PK * PITCH;
constexpr int PM_SMEM = 2 * 2 * TILE_ELEMS * 2;              // K,V x double buffer, bytes

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

__global__ void __launch_bounds__(128) attn_prefill_mma_kernel(const float* __restrict__ qkv, int D, int H,
                                                               const bf16* __restrict__ kc, const bf16* __restrict__ vc, int Smax,
                                                               const int* __restrict__ row_ids, const int* __restrict__ row_start,
                                                               const int* __restrict__ row_len, bf16* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    pdl_launch_dependents();
    pdl_wait();
    bf16* sm = reinterpret_cast<bf16*>(smem_raw);             // [buf][K|V][64][PITCH]
    const int qt = blockIdx.x, h = blockIdx.y, ri = blockIdx.z;
    const int len = row_len[ri];
    if (qt * PQ >= len) return;
    const int r = row_ids[ri], base = row_start[ri];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int q0 = qt * PQ + warp * 16;                        // first query row of this warp
    const bf16* kb = kc + ((int64_t)r * H + h) * Smax * 128;
    const bf16* vb = vc + ((int64_t)r * H + h) * Smax * 128;
    const int kend = min(len, (qt + 1) * PQ);
    const int ntiles = (kend + PK - 1) / PK;
    const uint32_t sm_base = (uint32_t)__cvta_generic_to_shared(sm);

    auto load_tile = [&](int kt, int buf) {
        const int k0 = kt * PK;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int c = tid + i * 128;                        // 1024 16-byte chunks per tensor
            const int key = c >> 4, ch = c & 15;
            const int kk = min(k0 + key, kend - 1);             // clamp: rows past the end are masked anyway
            const uint32_t d = sm_base + (uint32_t)(((buf * 2 + 0) * TILE_ELEMS + key * PITCH + ch * 8) * 2);
            cp_async16(d, kb + (int64_t)kk * 128 + ch * 8);
            cp_async16(d + TILE_ELEMS * 2, vb + (int64_t)kk * 128 + ch * 8);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    load_tile(0, 0);

    // Q fragments (A operand), scaled by 1/sqrt(128) * log2(e)
    const float qs = 0.08838834764831845f * 1.4426950408889634f;
    uint32_t qf[8][4];
    {
        const int ra = min(q0 + g, len - 1), rb = min(q0 + g + 8, len - 1);
        const float* pa = qkv + (int64_t)(base + ra) * 3 * D + h * 128;
        const float* pb = qkv + (int64_t)(base + rb) * 3 * D + h * 128;
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
            const float2 a0 = *reinterpret_cast<const float2*>(pa + ks * 16 + 2 * t);
            const float2 a1 = *reinterpret_cast<const float2*>(pb + ks * 16 + 2 * t);
            const float2 a2 = *reinterpret_cast<const float2*>(pa + ks * 16 + 2 * t + 8);
            const float2 a3 = *reinterpret_cast<const float2*>(pb + ks * 16 + 2 * t + 8);
            qf[ks][0] = pack_bf16(a0.x * qs, a0.y * qs); qf[ks][1] = pack_bf16(a1.x * qs, a1.y * qs);
            qf[ks][2] = pack_bf16(a2.x * qs, a2.y * qs); qf[ks][3] = pack_bf16(a3.x * qs, a3.y * qs);
        }
    }
    float o[16][4];
#pragma unroll
    for (int i = 0; i < 16; i++) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;   // rows g and g+8
    const int qa = q0 + g, qb = q0 + g + 8;

    for (int kt = 0; kt < ntiles; kt++) {
        const int buf = kt & 1;
        if (kt + 1 < ntiles) { load_tile(kt + 1, buf ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const bf16* Ks = sm + (buf * 2 + 0) * TILE_ELEMS;
        const uint32_t Vs_addr = sm_base + (uint32_t)((buf * 2 + 1) * TILE_ELEMS * 2);
        // ---- S = Q K^T (16 x 64 per warp) ----
        float s[8][4];
#pragma unroll
        for (int nb = 0; nb < 8; nb++) {
            s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
            const bf16* krow = Ks + (nb * 8 + g) * PITCH + 2 * t;
#pragma unroll
            for (int ks = 0; ks < 8; ks++) {
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(krow + ks * 16);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(krow + ks * 16 + 8);
                mma16816(s[nb], qf[ks], b0, b1);
            }
        }
        // ---- causal / length mask (only the last tile can be partial) ----
        const int k0 = kt * PK;
        if (k0 + PK - 1 > q0 || k0 + PK > kend) {
#pragma unroll
            for (int nb = 0; nb < 8; nb++) {
                const int key = k0 + nb * 8 + 2 * t;
                if (key > qa || key >= kend) s[nb][0] = -INFINITY;
                if (key + 1 > qa || key + 1 >= kend) s[nb][1] = -INFINITY;
                if (key > qb || key >= kend) s[nb][2] = -INFINITY;
                if (key + 1 > qb || key + 1 >= kend) s[nb][3] = -INFINITY;
            }
        }
        // ---- online softmax ----
        float mx_a = m_a, mx_b = m_b;
#pragma unroll
        for (int nb = 0; nb < 8; nb++) {
            mx_a = fmaxf(mx_a, fmaxf(s[nb][0], s[nb][1]));
            mx_b = fmaxf(mx_b, fmaxf(s[nb][2], s[nb][3]));
        }
        mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1)); mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
        mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1)); mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
        const float ua = (mx_a == -INFINITY) ? 0.f : mx_a, ub = (mx_b == -INFINITY) ? 0.f : mx_b;   // rows past `len`
        const float ca = exp2f(m_a - ua), cb = exp2f(m_b - ub);
        m_a = mx_a; m_b = mx_b;
        l_a *= ca; l_b *= cb;
#pragma unroll
        for (int i = 0; i < 16; i++) { o[i][0] *= ca; o[i][1] *= ca; o[i][2] *= cb; o[i][3] *= cb; }
        uint32_t pf[4][4];                                      // P as A fragments: k-step kk covers keys kk*16..+15
#pragma unroll
        for (int nb = 0; nb < 8; nb++) {
            const float p0 = exp2f(s[nb][0] - ua), p1 = exp2f(s[nb][1] - ua), p2 = exp2f(s[nb][2] - ub), p3 = exp2f(s[nb][3] - ub);
            l_a += p0 + p1; l_b += p2 + p3;
            pf[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16(p0, p1);
            pf[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16(p2, p3);
        }
        // ---- O += P V ----
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
#pragma unroll
            for (int n2 = 0; n2 < 16; n2 += 2) {
                uint32_t bv[4];
                const uint32_t addr = Vs_addr + (uint32_t)(((kk * 16 + (lane & 15)) * PITCH + (n2 + (lane >> 4)) * 8) * 2);
                ldmatrix_x4_trans(bv, addr);
                mma16816(o[n2], pf[kk], bv[0], bv[1]);
                mma16816(o[n2 + 1], pf[kk], bv[2], bv[3]);
            }
        }
        __syncthreads();
    }
    // row sums live spread over the quad
    l_a += __shfl_xor_sync(0xffffffffu, l_a, 1); l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);
    l_b += __shfl_xor_sync(0xffffffffu, l_b, 1); l_b += __shfl_xor_sync(0xffffffffu, l_b, 2);
    const float ia = 1.f / l_a, ib = 1.f / l_b;
    if (qa < len) {
        bf16* op = out + (int64_t)(base + qa) * D + h * 128 + 2 * t;
#pragma unroll
        for (int i = 0; i < 16; i++) *reinterpret_cast<uint32_t*>(op + i * 8) = pack_bf16(o[i][0] * ia, o[i][1] * ia);
    }
    if (qb < len) {
        bf16* op = out + (int64_t)(base + qb) * D + h * 128 + 2 * t;
#pragma unroll
        for (int i = 0; i < 16; i++) *reinterpret_cast<uint32_t*>(op + i * 8) = pack_bf16(o[i][2] * ib, o[i][3] * ib);
    }
}

}  // namespace

int launch_attn_prefill_mma(const float* qkv, int D, int H, const void* kcache, const void* vcache, int Smax, int n_rows,
                            const int* row_ids, const int* row_start, const int* row_len, int max_len, void* out, cudaStream_t s) {
    if (n_rows <= 0) return 0;
    static bool done = false;
    if (!done) {
        SSRB_CUDA(cudaFuncSetAttribute(attn_prefill_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PM_SMEM));
        done = true;
    }
    dim3 grid(cdiv(max_len, PQ), H, n_rows);
    return launch_pdl(attn_prefill_mma_kernel, grid, dim3(128), (size_t)PM_SMEM, s, 1, qkv, D, H, (const bf16*)kcache, (const bf16*)vcache,
                      Smax, row_ids, row_start, row_len, (bf16*)out);
}

}  // namespace ssrb
