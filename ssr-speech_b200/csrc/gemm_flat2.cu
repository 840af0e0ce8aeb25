// gemm_flat2.cu — prefill GEMM on CTA pairs (tcgen05.mma.cta_group::2).  Default path for M > 128 since round 2: verified on a B200
// (tests/test_gpu_flat2.py: 10 shapes vs torch fp32 and vs the 1-CTA kernel, whole roll-outs) and measured: prefill of the bench
// batch 92.3 -> 78.8 ms (profiles/r02a_summary.md).  SSRB_FLAT_2CTA=0 selects the 1-CTA gemm_flat_kernel of gemm_tc.cu.
// sm_100a only.
//
//   C[M,N] = act(A[M,K] . W[N,K]^T + bias) (+ residual)         A, W bf16 K-major; fp32 accumulate; M > 128 (prefill)
//
// Why: ncu on the 1-CTA kernel shows the tensor pipe 43 % active (profiles/r01e_summary.md) — a 128 x 256 x 64 k-block costs
// 48 KB of operands per SM, 26 TB/s of L2 -> SM traffic at full rate.  A CTA pair computes one 256 x 256 tile: each CTA stages
// its own 128 activation rows and HALF of the weight tile (128 of the 256 output columns), the pair's tensor cores read both
// halves (UMMA M = 256 across two SMs), so the same math needs 32 KB per SM and k-block.  Accumulators: 128 lanes x 256 fp32
// columns in each CTA's TMEM, double-buffered (512 columns), drained by each CTA's own epilogue warps.
//
// Protocol (the one CUTLASS's sm100 2-SM collectives use; PTX forms checked against the vendored headers
// cute/arch/copy_sm100_tma.hpp, cute/arch/mma_sm100_umma.hpp, cutlass/arch/barrier.h, cute/arch/tmem_allocator_sm100.hpp):
//   * both CTAs issue their TMA loads with .cta_group::2 and the LEADER's (even rank) full barrier as completion target
//     (barrier address with the peer bit cleared); the leader's producer arms it with the bytes of BOTH CTAs;
//   * only the leader's MMA thread waits on it and issues tcgen05.mma.cta_group::2; tcgen05.commit ... multicast::cluster
//     with mask 0b11 releases the ring stage in both CTAs and publishes the accumulator to both epilogues;
//   * the epilogue warps of both CTAs arrive on the leader's accumulator-empty barrier (8 arrivals);
//   * tcgen05.alloc / dealloc with .cta_group::2 by the same warp of both CTAs, cluster barrier before teardown.
//
// warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (leader), warps 2-5 = epilogue.
#include <algorithm>

#include "tc_ptx.cuh"
#include "../../include/ssr_b200.h"

namespace ssrb {

namespace {

constexpr int F2_N = 256;                               // output columns per tile (both CTAs together)
constexpr int F2_HALF = 128;                            // weight rows each CTA stages per k-block
constexpr int F2_STAGES = 6;
constexpr int F2_STAGE_BYTES = P_BYTES + F2_HALF * BK * 2;          // 16 KB activations + 16 KB weights
constexpr size_t F2_SMEM = (size_t)F2_STAGES * F2_STAGE_BYTES + 1024 + 256;
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;         // cute::Sm100MmaPeerBitMask: shared::cluster address of the even CTA

struct Flat2Prm {
    int M, N, nkb, m_tiles, n_tiles;
    const float* bias; const float* residual; long long ldr;
    void* C; long long ldc; int c_dtype, act;
};

__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((unsigned long long)map), "r"(leader_bar & PEER_BIT_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma2_commit(uint32_t bar, uint16_t cta_mask) {      // arrives at the same barrier offset in every CTA of the mask
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void arrive_on_leader(uint32_t local_bar) {                // mbarrier of rank 0 at the same offset
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_bar), "r"(0u));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");     // what it guards lives in TMEM (tcgen05 fences)
}
__device__ __forceinline__ void wait_guarded(uint32_t bar, uint32_t parity) {         // bounded spin: trap instead of hanging the box
    uint32_t ok;
    unsigned long long t0 = 0;
    for (;;) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 2000000000ull) __trap();
    }
}
__device__ __forceinline__ void tmem_ld32_2(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
        "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// drains one 128 x 256 fp32 accumulator (the calling thread's TMEM lane = output row m): bias, activation, residual, 16-byte
// stores; same arithmetic and order as gemm_flat_kernel's epilogue
__device__ __forceinline__ void flat2_store_tile(const Flat2Prm& prm, uint32_t taddr, int m, bool mok, int n_base) {
#pragma unroll 1
    for (int c0 = 0; c0 < F2_N; c0 += 32) {
        const int n0 = n_base + c0;
        if (n0 >= prm.N) break;                            // warp-uniform
        float v[32];
        tmem_ld32_2(taddr + c0, v);
        if (!mok) continue;
        if (n0 + 32 <= prm.N) {
            if (prm.bias) {
#pragma unroll
                for (int q = 0; q < 32; q += 4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(prm.bias + n0 + q));
                    v[q] += b4.x; v[q + 1] += b4.y; v[q + 2] += b4.z; v[q + 3] += b4.w;
                }
            }
#pragma unroll
            for (int q = 0; q < 32; q++) v[q] = apply_act(v[q], prm.act);
            if (prm.residual) {
                const float* rp = prm.residual + (long long)m * prm.ldr + n0;
#pragma unroll
                for (int q = 0; q < 32; q += 4) {
                    const float4 r4 = *reinterpret_cast<const float4*>(rp + q);
                    v[q] += r4.x; v[q + 1] += r4.y; v[q + 2] += r4.z; v[q + 3] += r4.w;
                }
            }
            if (prm.c_dtype == SSRB_DTYPE_F32) {
                float* cp = reinterpret_cast<float*>(prm.C) + (long long)m * prm.ldc + n0;
#pragma unroll
                for (int q = 0; q < 32; q += 4) *reinterpret_cast<float4*>(cp + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
            } else {
                bf16* cp = reinterpret_cast<bf16*>(prm.C) + (long long)m * prm.ldc + n0;
#pragma unroll
                for (int q = 0; q < 32; q += 8) {
                    float w8[8];
#pragma unroll
                    for (int z = 0; z < 8; z++) w8[z] = v[q + z];
                    store8(cp + q, w8);
                }
            }
        } else {
            for (int q = 0; q < 32 && n0 + q < prm.N; q++) {
                const int n = n0 + q;
                float x = apply_act(v[q] + (prm.bias ? prm.bias[n] : 0.f), prm.act);
                if (prm.residual) x += prm.residual[(long long)m * prm.ldr + n];
                const long long o = (long long)m * prm.ldc + n;
                if (prm.c_dtype == SSRB_DTYPE_F32) reinterpret_cast<float*>(prm.C)[o] = x;
                else reinterpret_cast<bf16*>(prm.C)[o] = __float2bfloat16_rn(x);
            }
        }
    }
}

__global__ void __launch_bounds__(192, 1) gemm_flat2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                            const Flat2Prm prm) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = base + F2_STAGES * F2_STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (F2_STAGES + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * F2_STAGES + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * F2_STAGES + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * F2_STAGES + 4);

    pdl_launch_dependents();
    const int ts = ts_begin(TSK_GEMM);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();               // 0 = leader (issues the MMAs), 1 = peer
    const int pair = (int)(blockIdx.x >> 1), n_pairs = (int)(gridDim.x >> 1);
    if (threadIdx.x == 0) {
        for (int s = 0; s < F2_STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; b++) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 8); }     // 4 epilogue warps x 2 CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    cluster_sync_all();                                    // the peer's barriers exist before anything is signalled across the pair
    const int total = prm.m_tiles * prm.n_tiles, nkb = prm.nkb;

    if (warp == 0) {
        if (lane == 0) {                                   // ===== TMA producer (both CTAs: own activation rows, own half of the weights) =====
            pdl_wait();
            ts_dep(ts);
            int it = 0;
            for (int t = pair; t < total; t += n_pairs) {
                const int mt = t / prm.n_tiles, nt = t - mt * prm.n_tiles;
                const int a_row = mt * 2 * P_ROWS + (int)rank * P_ROWS, w_row = nt * F2_N + (int)rank * F2_HALF;
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % F2_STAGES;
                    wait_guarded(empty_bar(s), (uint32_t)(((it / F2_STAGES) & 1) ^ 1));
                    if (rank == 0) mbar_expect_tx(full_bar(s), 2 * F2_STAGE_BYTES);       // both CTAs' bytes land on the leader's barrier
                    const uint32_t sp = base + s * F2_STAGE_BYTES;
                    tma2_load_2d(sp, &mapA, full_bar(s), kb * BK, a_row);
                    tma2_load_2d(sp + P_BYTES, &mapB, full_bar(s), kb * BK, w_row);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {                      // ===== MMA issuer: the leader drives both SMs' tensor cores =====
            // instruction descriptor: c=f32 (1<<4), a=b=bf16 (1<<7, 1<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24; M = 256 over the pair
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(F2_N >> 3) << 17) | ((uint32_t)((2 * P_ROWS) >> 4) << 24);
            int it = 0, j = 0;
            for (int t = pair; t < total; t += n_pairs, j++) {
                const int buf = j & 1;
                wait_guarded(tempty_bar(buf), (uint32_t)(((j >> 1) & 1) ^ 1));            // both epilogues have drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + (uint32_t)(buf * F2_N);
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % F2_STAGES;
                    wait_guarded(full_bar(s), (uint32_t)((it / F2_STAGES) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sp = base + s * F2_STAGE_BYTES;
                    const uint64_t da = make_desc(sp), db = make_desc(sp + P_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; k++) umma2_bf16(acc, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    umma2_commit(empty_bar(s), (uint16_t)3);                               // the stage is free in both CTAs
                }
                umma2_commit(tfull_bar(buf), (uint16_t)3);                                 // the accumulator is complete in both CTAs
            }
        }
    } else {
        pdl_wait();                                        // residual / output buffers belong to the kernel chain
        const int lg = warp & 3, nl = lg * 32 + lane;
        int j = 0;
        for (int t = pair; t < total; t += n_pairs, j++) {
            const int mt = t / prm.n_tiles, nt = t - mt * prm.n_tiles;
            const int buf = j & 1;
            wait_guarded(tfull_bar(buf), (uint32_t)((j >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * F2_N);
            const int m = mt * 2 * P_ROWS + (int)rank * P_ROWS + nl;
            flat2_store_tile(prm, taddr, m, m < prm.M, nt * F2_N);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) arrive_on_leader(tempty_bar(buf));
        }
    }
    __syncwarp();
    __syncthreads();
    cluster_sync_all();                                    // the leader's MMAs read the peer's shared memory and TMEM until the end
    ts_end(ts);
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace

bool gemm_flat2_enabled() {
    static const bool on = [] { const char* e = getenv("SSRB_FLAT_2CTA"); return !(e && e[0] == '0'); }();     // default on since round 2 (SSRB_FLAT_2CTA=0: the 1-CTA gemm_flat_kernel)
    return on;
}

bool gemm_flat2_supported(const GemmArgs& g) {
    if (g.ab_dtype != SSRB_DTYPE_BF16 || g.groups != 1 || g.M <= 128) return false;
    if (g.K % BK != 0 || g.K < BK || g.lda % 8 != 0 || g.ldw % 8 != 0) return false;
    if (g.ln_part || g.C2 || g.part_out) return false;
    return g.N % 4 == 0 && g.ldc % 8 == 0 && (!g.residual || g.ldr % 4 == 0) && ((uintptr_t)g.C & 15) == 0 &&
           ((uintptr_t)g.bias & 15) == 0 && ((uintptr_t)g.residual & 15) == 0;
}

int gemm_flat2(const GemmArgs& g, cudaStream_t s) {
    SSRB_CHECK(gemm_flat2_supported(g), "gemm_flat2: unsupported problem");
    static int n_sm = 0;
    if (n_sm == 0) { int dev = 0; SSRB_CUDA(cudaGetDevice(&dev)); SSRB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev)); }
    static bool attr_done = false;
    if (!attr_done) {
        SSRB_CUDA(cudaFuncSetAttribute(gemm_flat2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F2_SMEM));
        attr_done = true;
    }
    CUtensorMap mapA, mapB;
    SSRB_TRY(tc_make_map(&mapA, g.A, g.M, g.K, g.lda, P_ROWS));
    SSRB_TRY(tc_make_map(&mapB, g.W, g.N, g.K, g.ldw, F2_HALF));
    Flat2Prm prm{};
    prm.M = g.M; prm.N = g.N; prm.nkb = g.K / BK;
    prm.m_tiles = cdiv(g.M, 2 * P_ROWS); prm.n_tiles = cdiv(g.N, F2_N);
    prm.bias = g.bias; prm.residual = g.residual; prm.ldr = g.ldr; prm.C = g.C; prm.ldc = g.ldc; prm.c_dtype = g.c_dtype; prm.act = g.act;
    const int pairs = std::min(prm.m_tiles * prm.n_tiles, n_sm / 2);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = F2_SMEM; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int na = 0;
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    na++;
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        na++;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    SSRB_CUDA(cudaLaunchKernelEx(&cfg, gemm_flat2_kernel, mapA, mapB, prm));
    g_launch_count++;
    return 0;
}

int ts_arm_gemm_flat2(const TsBuf& t) { return ts_arm_tu(t); }

}  // namespace ssrb
