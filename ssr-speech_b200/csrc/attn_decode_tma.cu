// attn_decode_tma.cu — bf16 decode attention with bulk-async (TMA engine) staging of the K/V stream.
//
// One query per (row, head) against the in-place cache — replaces F.scaled_dot_product_attention at
// models/modules/activation.py:634 for tgt_len == 1, plus the KV "append" the reference performs by
// re-materialising the cache (activation.py:626-631, ssr.py:685-686).
//
// grid (H, R, ceil(Smax/128)); each CTA owns 128 consecutive keys of one (row, head):
//   * thread 0 issues four cp.async.bulk copies (K and V, two 64-key sub-tiles, <= 16 KB each) straight into
//     shared memory, completion tracked by one mbarrier per sub-tile — the bytes in flight do not depend on
//     registers or occupancy (64 KB per CTA, 3 CTAs per SM = 192 KB of loads in flight per SM);
//   * 4 warps consume the tiles from shared memory (16 lanes x 16 B per key row, conflict-free), fp32 online
//     softmax, flash-decoding merge of the splits by the last-arriving CTA;
//   * the CTA whose range ends at the current position takes this step's K/V row from the QKV GEMM output,
//     rounds it to bf16, stores it into the cache in place and scores it from registers (so the bulk copy never
//     reads bytes written in the same kernel).
// HBM roofline: algorithmic bytes per launch = R*H*(S+1)*2*128*2 B.
#include "lm_kernels.cuh"

namespace ssrb {

namespace {

constexpr int AT_CHUNK = 128, AT_SUB = 64, AT_ROWB = 256;            // keys per CTA / per sub-tile, bytes per key row
constexpr int AT_SMEM = 2 * AT_CHUNK * AT_ROWB + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void lds8_bf16(uint32_t addr, float (&v)[8]) {
    uint32_t w0, w1, w2, w3;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(addr));
    const uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

__global__ void __launch_bounds__(128, 3) attn_decode_tma_kernel(const float* __restrict__ qkv, int D, int H, bf16* kc, bf16* vc,
                                                                 int Smax, const int* __restrict__ seq_len,
                                                                 const UttState* __restrict__ st, int rpu, float* ws,
                                                                 int* __restrict__ tickets, bf16* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int h = blockIdx.x, r = blockIdx.y, z = blockIdx.z, nz = gridDim.z;
    if (st[r / rpu].done) return;
    const int n_keys = seq_len[r] + 1;
    const int nsplit = (n_keys + AT_CHUNK - 1) / AT_CHUNK;
    if (z >= nsplit) return;
    const int s0 = z * AT_CHUNK, s1 = min(n_keys, s0 + AT_CHUNK);
    const bool has_new = (s1 == n_keys);
    const int n_old = (has_new ? s1 - 1 : s1) - s0;                      // keys streamed from the cache
    const uint32_t ks = smem_u32(smem), vs = ks + AT_CHUNK * AT_ROWB, bar0 = vs + AT_CHUNK * AT_ROWB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, half = lane >> 4, dl = (lane & 15) * 8;
    const bf16* kbase = kc + ((int64_t)r * H + h) * Smax * 128;
    const bf16* vbase = vc + ((int64_t)r * H + h) * Smax * 128;
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int sub = 0; sub < 2; sub++) {
            const int nk = min(AT_SUB, n_old - sub * AT_SUB);
            if (nk > 0) {
                const uint32_t bytes = (uint32_t)nk * AT_ROWB;
                mbar_expect_tx(bar0 + 8 * sub, 2 * bytes);
                bulk_g2s(ks + sub * AT_SUB * AT_ROWB, kbase + (int64_t)(s0 + sub * AT_SUB) * 128, bytes, bar0 + 8 * sub);
                bulk_g2s(vs + sub * AT_SUB * AT_ROWB, vbase + (int64_t)(s0 + sub * AT_SUB) * 128, bytes, bar0 + 8 * sub);
            }
        }
    }
    const float scale = 0.08838834764831845f;   // 1/sqrt(128)
    float q[8];
    load8(qkv + (int64_t)r * 3 * D + h * 128 + dl, q);
#pragma unroll
    for (int i = 0; i < 8; i++) q[i] *= scale;
    float mrun = -INFINITY, lrun = 0.f, o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (has_new && warp == 0) {
        // this step's K (half 0) / V (half 1) row: round to bf16 exactly as later steps will read it back
        float nv[8];
        load8(qkv + (int64_t)r * 3 * D + (1 + half) * D + h * 128 + dl, nv);
        bf16* dst = (half ? vc : kc) + (((int64_t)r * H + h) * Smax + (n_keys - 1)) * 128 + dl;
        store8(dst, nv);
#pragma unroll
        for (int i = 0; i < 8; i++) nv[i] = __bfloat162float(__float2bfloat16_rn(nv[i]));
        float p = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) p = fmaf(q[i], nv[i], p);
        p += __shfl_xor_sync(0xffffffffu, p, 1);
        p += __shfl_xor_sync(0xffffffffu, p, 2);
        p += __shfl_xor_sync(0xffffffffu, p, 4);
        p += __shfl_xor_sync(0xffffffffu, p, 8);
        float vv[8];
#pragma unroll
        for (int i = 0; i < 8; i++) vv[i] = __shfl_sync(0xffffffffu, nv[i], (lane & 15) + 16);
        if (half == 0) {
            mrun = p; lrun = 1.f;
#pragma unroll
            for (int i = 0; i < 8; i++) o[i] = vv[i];
        }
    }
    __syncthreads();                                   // mbarrier init visible to the waiting threads
#pragma unroll 1
    for (int sub = 0; sub < 2; sub++) {
        const int nk = min(AT_SUB, n_old - sub * AT_SUB);
        if (nk <= 0) break;                            // CTA-uniform
        mbar_wait(bar0 + 8 * sub, 0);
        const uint32_t kt = ks + sub * AT_SUB * AT_ROWB + dl * 2, vt = vs + sub * AT_SUB * AT_ROWB + dl * 2;
#pragma unroll
        for (int it = 0; it < 2; it++) {
            float sc[4];
            float mnew = mrun;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int kl = warp * 16 + it * 8 + j * 2 + half;
                float kk[8];
                lds8_bf16(kt + kl * AT_ROWB, kk);
                float p = 0.f;
#pragma unroll
                for (int i = 0; i < 8; i++) p = fmaf(q[i], kk[i], p);
                p += __shfl_xor_sync(0xffffffffu, p, 1);
                p += __shfl_xor_sync(0xffffffffu, p, 2);
                p += __shfl_xor_sync(0xffffffffu, p, 4);
                p += __shfl_xor_sync(0xffffffffu, p, 8);
                sc[j] = kl < nk ? p : -INFINITY;
                mnew = fmaxf(mnew, sc[j]);
            }
            if (mnew > -INFINITY) {
                const float corr = __expf(mrun - mnew);
                lrun *= corr;
#pragma unroll
                for (int i = 0; i < 8; i++) o[i] *= corr;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int kl = warp * 16 + it * 8 + j * 2 + half;
                    if (kl < nk) {
                        float vv[8];
                        lds8_bf16(vt + kl * AT_ROWB, vv);
                        const float p = __expf(sc[j] - mnew);
                        lrun += p;
#pragma unroll
                        for (int i = 0; i < 8; i++) o[i] = fmaf(p, vv[i], o[i]);
                    }
                }
                mrun = mnew;
            }
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
    const int d = tid;
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
    bf16* op = out + (int64_t)r * D + h * 128 + d;
    if (nsplit == 1) { *op = __float2bfloat16_rn(O / L); return; }
    float* wsp = ws + ((int64_t)(r * H + h) * nz + z) * 130;
    wsp[2 + d] = O;
    if (d == 0) { wsp[0] = M; wsp[1] = L; }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
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
    *op = __float2bfloat16_rn(O2 / L2);
}

}  // namespace

int attn_decode_tma_nsplit(int Smax) { return cdiv(Smax, AT_CHUNK); }

int launch_attn_decode_tma(const float* qkv, int R, int D, int H, void* kcache, void* vcache, int Smax, const int* seq_len,
                           const UttState* st, int rpu, float* ws, int* tickets, void* out, cudaStream_t s) {
    static bool attr_done = false;
    if (!attr_done) {
        SSRB_CUDA(cudaFuncSetAttribute(attn_decode_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
        attr_done = true;
    }
    dim3 grid(H, R, attn_decode_tma_nsplit(Smax));
    SSRB_LAUNCH(attn_decode_tma_kernel, grid, 128, AT_SMEM, s, qkv, D, H, (bf16*)kcache, (bf16*)vcache, Smax, seq_len, st, rpu,
                ws, tickets, (bf16*)out);
    return 0;
}

}  // namespace ssrb
