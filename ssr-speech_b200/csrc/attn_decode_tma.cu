// attn_decode_tma.cu — bf16 decode attention with bulk-async (TMA engine) staging of the K/V stream.
//
// One query per (row, head) against the in-place cache — replaces F.scaled_dot_product_attention at
// models/modules/activation.py:634 for tgt_len == 1, plus the KV "append" the reference performs by
// re-materialising the cache (activation.py:626-631, ssr.py:685-686).
//
// grid (H, R, nsplit).  A CTA streams a contiguous range of one (row, head)'s keys — the whole row when the batch
// alone fills the machine (R*H >= 2 CTAs per SM), a slice of it otherwise (flash-decoding split, merged by the
// last-arriving CTA):
//   * warp 4 is the producer: one lane issues cp.async.bulk copies of 64-key K and V sub-tiles (16 KB each) into a
//     3-stage shared-memory ring, each stage guarded by a full (tx-count) and an empty mbarrier; the bytes in flight
//     (96 KB per CTA, 2 CTAs per SM) do not depend on registers or occupancy, and the ring keeps streaming while the
//     consumers compute;
//   * warps 0-3 consume: 16 lanes x 16 B per key row (conflict-free), fp32 online softmax, 8 partial states merged
//     through shared memory at the end;
//   * the CTA whose range ends at the current position takes this step's K/V row from the QKV GEMM output, rounds it
//     to bf16, stores it into the cache in place and scores it from registers (the bulk copies never read bytes
//     written by this kernel);
//   * launched as a programmatic dependent (PDL): CTA scheduling and barrier set-up overlap the tail of the QKV GEMM.
// HBM roofline: algorithmic bytes per launch = R*H*(S+1)*2*128*2 B.
#include "lm_kernels.cuh"

namespace ssrb {

namespace {

constexpr int AT_SUB = 64, AT_ROWB = 256, AT_STAGES = 3;            // keys per sub-tile, bytes per key row
constexpr int AT_STAGE_BYTES = 2 * AT_SUB * AT_ROWB;                // K + V
constexpr int AT_SMEM = AT_STAGES * AT_STAGE_BYTES + 64;

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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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

__global__ void __launch_bounds__(160, 2) attn_decode_tma_kernel(const float* __restrict__ qkv, int D, int H, bf16* kc, bf16* vc,
                                                                 int Smax, const int* __restrict__ seq_len,
                                                                 const UttState* __restrict__ st, int rpu, float* ws,
                                                                 int* __restrict__ tickets, bf16* __restrict__ out, int prefetch) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ float sm_m[8], sm_l[8], sm_o[8][128];
    __shared__ int sm_last;
    pdl_launch_dependents();
    const int ts = ts_begin(TSK_ATTN);
    // prefetch == 0: wait for the predecessor before reading anything.  prefetch == 1 (decode chain): the row state (done,
    // seq_len) was written by the previous iteration's sampler and the cached K/V rows [0, n_old) by earlier iterations; the
    // embed kernel that opens every iteration releases its dependents only AFTER its own griddepcontrol.wait, so no kernel
    // of this iteration — this one included — can start before the previous iteration has completed and flushed.  Only q
    // and this step's K/V row come from the QKV GEMM right before: the consumers wait for it, the producer warp starts
    // streaming the cache at once, so the K/V stream overlaps the tail of the GEMM.
    if (!prefetch) { pdl_wait(); ts_dep(ts); }
    const int h = blockIdx.x, r = blockIdx.y, z = blockIdx.z, nz = gridDim.z;
    if (st[r / rpu].done) { if (prefetch) pdl_wait(); return; }     // every CTA waits: completion stays transitive along the chain
    const int n_keys = seq_len[r] + 1;
    const int n_old = n_keys - 1;                                         // keys already in the cache
    const int tiles_total = (n_old + AT_SUB - 1) / AT_SUB;               // sub-tiles of cached keys
    // balanced contiguous split of the cached sub-tiles over the nz CTAs of this (row, head); the last non-empty CTA
    // also owns the new key.  CTAs with no work exit (nsplit_eff counts the ones that take a ticket).
    const int per = (tiles_total + nz - 1) / nz;
    const int nsplit_eff = per > 0 ? (tiles_total + per - 1) / per : 1;   // >= 1 (a row with no cached key: 1 CTA)
    if (z >= nsplit_eff) { if (prefetch) pdl_wait(); return; }
    const int t0 = z * per, t1 = min(tiles_total, t0 + per);
    const int my_tiles = t1 - t0;
    const bool has_new = (z == nsplit_eff - 1);
    const uint32_t ring = smem_u32(smem), bar0 = ring + AT_STAGES * AT_STAGE_BYTES;   // full[s] = bar0+8s, empty[s] = bar0+24+8s
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, half = lane >> 4, dl = (lane & 15) * 8;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < AT_STAGES; s++) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar0 + 24 + 8 * s, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 4) {
        // ===== producer =====
        if (lane == 0) {
            const bf16* kbase = kc + ((int64_t)r * H + h) * Smax * 128;
            const bf16* vbase = vc + ((int64_t)r * H + h) * Smax * 128;
            for (int i = 0; i < my_tiles; i++) {
                const int s = i % AT_STAGES;
                const uint32_t ph = (i / AT_STAGES) & 1;
                mbar_wait(bar0 + 24 + 8 * s, ph ^ 1);
                const int k0 = (t0 + i) * AT_SUB;
                const uint32_t bytes = (uint32_t)min(AT_SUB, n_old - k0) * AT_ROWB;
                mbar_expect_tx(bar0 + 8 * s, 2 * bytes);
                bulk_g2s(ring + s * AT_STAGE_BYTES, kbase + (int64_t)k0 * 128, bytes, bar0 + 8 * s);
                bulk_g2s(ring + s * AT_STAGE_BYTES + AT_SUB * AT_ROWB, vbase + (int64_t)k0 * 128, bytes, bar0 + 8 * s);
            }
        }
        return;
    }

    // ===== consumers (warps 0-3) =====
    if (prefetch) { pdl_wait(); ts_dep(ts); }
    const float scale = 0.08838834764831845f;          // 1/sqrt(128)
    float q[8];
    load8(qkv + (int64_t)r * 3 * D + h * 128 + dl, q);
#pragma unroll
    for (int i = 0; i < 8; i++) q[i] *= scale;
    float mrun = -INFINITY, lrun = 0.f, o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (has_new && warp == 0) {
        // this step's K (half 0) / V (half 1) row: round to bf16 exactly as later steps will read it back
        float nv[8];
        load8(qkv + (int64_t)r * 3 * D + (1 + half) * D + h * 128 + dl, nv);
        bf16* dst = (half ? vc : kc) + (((int64_t)r * H + h) * Smax + n_old) * 128 + dl;
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
#pragma unroll 1
    for (int i = 0; i < my_tiles; i++) {
        const int s = i % AT_STAGES;
        const uint32_t ph = (i / AT_STAGES) & 1;
        const int nk = min(AT_SUB, n_old - (t0 + i) * AT_SUB);
        mbar_wait(bar0 + 8 * s, ph);
        const uint32_t kt = ring + s * AT_STAGE_BYTES + dl * 2, vt = kt + AT_SUB * AT_ROWB;
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
                for (int e = 0; e < 8; e++) p = fmaf(q[e], kk[e], p);
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
                for (int e = 0; e < 8; e++) o[e] *= corr;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int kl = warp * 16 + it * 8 + j * 2 + half;
                    if (kl < nk) {
                        float vv[8];
                        lds8_bf16(vt + kl * AT_ROWB, vv);
                        const float p = __expf(sc[j] - mnew);
                        lrun += p;
#pragma unroll
                        for (int e = 0; e < 8; e++) o[e] = fmaf(p, vv[e], o[e]);
                    }
                }
                mrun = mnew;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + 24 + 8 * s);     // this warp is done with the stage
    }
    // merge the 8 (warp, half) partial states of this CTA (named barrier: the producer warp has left)
    const int slot = warp * 2 + half;
    if ((lane & 15) == 0) { sm_m[slot] = mrun; sm_l[slot] = lrun; }
#pragma unroll
    for (int i = 0; i < 8; i++) sm_o[slot][dl + i] = o[i];
    asm volatile("bar.sync 1, 128;" ::: "memory");
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
    if (nsplit_eff == 1) { *op = __float2bfloat16_rn(O / L); ts_end(ts); return; }
    float* wsp = ws + ((int64_t)(r * H + h) * nz + z) * 130;
    wsp[2 + d] = O;
    if (d == 0) { wsp[0] = M; wsp[1] = L; }
    __threadfence();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (tid == 0) {
        const int t = atomicAdd(&tickets[r * H + h], 1);
        sm_last = (t == nsplit_eff - 1);
        if (sm_last) tickets[r * H + h] = 0;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (!sm_last) return;
    __threadfence();
    const float* wb = ws + (int64_t)(r * H + h) * nz * 130;
    float M2 = -INFINITY;
    for (int i = 0; i < nsplit_eff; i++) M2 = fmaxf(M2, __ldcg(wb + i * 130));
    float L2 = 0.f, O2 = 0.f;
    for (int i = 0; i < nsplit_eff; i++) {
        const float w = __expf(__ldcg(wb + i * 130) - M2);
        L2 += __ldcg(wb + i * 130 + 1) * w;
        O2 += __ldcg(wb + i * 130 + 2 + d) * w;
    }
    *op = __float2bfloat16_rn(O2 / L2);
}

}  // namespace

// number of CTAs per (row, head): 1 when the batch alone gives >= 2 CTAs per SM, else enough slices to get there
int attn_decode_tma_nsplit(int R, int H, int Smax) {
    const int max_split = cdiv(Smax, 2 * AT_SUB);
    int ns = cdiv(2 * 148, R * H);
    if (ns < 1) ns = 1;
    if (ns > max_split) ns = max_split;
    return ns;
}
int attn_decode_tma_max_nsplit(int Smax) { return cdiv(Smax, 2 * AT_SUB); }

int launch_attn_decode_tma(const float* qkv, int R, int D, int H, void* kcache, void* vcache, int Smax, const int* seq_len,
                           const UttState* st, int rpu, float* ws, int* tickets, void* out, int prefetch, cudaStream_t s) {
    static bool attr_done = false;
    if (!attr_done) {
        SSRB_CUDA(cudaFuncSetAttribute(attn_decode_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
        attr_done = true;
    }
    dim3 grid(H, R, attn_decode_tma_nsplit(R, H, Smax));
    return launch_pdl(attn_decode_tma_kernel, grid, dim3(160), AT_SMEM, s, 1, qkv, D, H, (bf16*)kcache, (bf16*)vcache, Smax, seq_len,
                      st, rpu, ws, tickets, (bf16*)out, prefetch);
}

int ts_arm_attn_tma(const TsBuf& t) { return ts_arm_tu(t); }

}  // namespace ssrb
