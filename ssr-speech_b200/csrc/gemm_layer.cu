// gemm_layer.cu — EXPERIMENTAL persistent per-layer GEMM chain for the decode iteration (off by default:
// SSRB_LAYER_KERNEL=1).  sm_100a only.  Verified and A/B'd on a B200 in round 2 (profiles/r02a_summary.md): correct, but slower
// than the per-GEMM chain (its grid barriers cost what the kernel boundaries cost: iteration 1.79 -> 1.95 ms); the
// per-GEMM chain of gemm_tc.cu stays the product path.
//
// Why: inside a decode iteration the GEMM phase of a layer is latency-bound — 38 us for a 15.4 us weight stream
// (profiles/r01e_summary.md): each of the four launches pays activation tile, accumulate, park, cluster barrier, DSMEM
// reduce, exit + dependent release, and FFN2 cannot even prefetch because its CTAs are not resident before FFN1 exits.
// Here ONE launch of 16 clusters x 8 CTAs (one CTA per SM, all co-resident) runs
//     phase 0  out-proj (+ residual, bf16 copy, row statistics)      8-way split-K  (transformer.py:321-343, activation.py:637)
//     phase 1  FFN1 (folded LayerNorm, ReLU)                         4-way split-K  (transformer.py:386-388)
//     phase 2  FFN2 (+ residual, bf16 copy, row statistics)          8-way split-K
//     phase 3  the NEXT layer's QKV projection (folded LayerNorm)    4-way split-K  (activation.py:83-89)
// with a grid barrier between phases.  The TMA producer thread keeps ONE weight ring running across the phases: the weight
// tiles of phase p+1 are requested while phase p is still reducing / waiting at the barrier, only the (tiny) activation
// tiles wait for it.  Tiles, split-K slices, DSMEM reduce-scatter and every epilogue expression are those of
// gemm_dec_kernel (same summation order -> the same bits).
//
// Synchronisation inside a cluster uses mbarriers with remote arrives (only the epilogue warps take part, so the producer
// and the MMA thread never stall on it):  parked (every peer's fp32 partial tile is in its shared memory) and consumed
// (every peer has finished reading mine).  Grid barrier: one arrival per CTA on a global counter, generation word polled
// by the producer thread and one epilogue thread.
//
// warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue.
#include <algorithm>

#include "tc_ptx.cuh"
#include "../../include/ssr_b200.h"

namespace ssrb {

namespace {

constexpr int LK_CLUSTER = 8;
constexpr int LK_PHASES = 4;
constexpr int LK_MAX_CLUSTERS = 16;
enum { LK_RES = 0, LK_FFN1 = 1, LK_QKV = 2 };

struct alignas(64) LayerMaps { CUtensorMap p[LK_PHASES], q[LK_PHASES]; };

struct LayerPrm {
    int M, n_phases, n_clusters;
    int n_tiles[LK_PHASES], kbps[LK_PHASES];        // 128-row weight tiles of the phase; k-blocks per split-K slice
    const float* bias[LK_PHASES];
    const float* colsum[LK_PHASES];
    float* x; long long ld_x;
    bf16* hn; long long ld_hn;
    bf16* hid; long long ld_hid;
    float* qkv; long long ld_qkv;
    float2* ln_part; int part_ld, ln_blocks; float ln_eps;
    unsigned int* gbar;
};

template <int QROWS> struct LayerCfg {
    static constexpr int Q_BYTES = QROWS * BK * 2;
    static constexpr int STAGE_BYTES = P_BYTES + Q_BYTES;
    static constexpr int PARK_BYTES = QROWS * 128 * 4;
    static constexpr int BAR_BYTES = 512;
    static constexpr int BUDGET = 227 * 1024 - 1024 /*align slack*/ - BAR_BYTES - PARK_BYTES - 2048 /*static*/;
    static constexpr int STAGES = BUDGET / STAGE_BYTES > 8 ? 8 : BUDGET / STAGE_BYTES;
    static constexpr int TMEM_COLS = 2 * QROWS < 32 ? 32 : 2 * QROWS;      // two accumulators
    static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
    static constexpr size_t SMEM = (size_t)RING_BYTES + PARK_BYTES + 1024 + BAR_BYTES;
    static_assert(STAGES >= 3, "ring too shallow");
};

__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Every spin in this kernel is bounded: a protocol bug would otherwise hang the GPU until the box is reclaimed.  After
// LK_WATCHDOG_NS of waiting the thread traps; the launch fails with an error instead of never finishing.
#ifndef LK_WATCHDOG_NS
#define LK_WATCHDOG_NS 2000000000ull
#endif
__device__ __forceinline__ void spin_guard(unsigned long long& t0) {
    const unsigned long long now = globaltimer_ns();
    if (t0 == 0) t0 = now;
    else if (now - t0 > LK_WATCHDOG_NS) __trap();
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {       // non-blocking: has that phase completed?
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void lk_wait(uint32_t bar, uint32_t parity) {       // mbar_wait of tc_ptx.cuh with the watchdog
    uint32_t ok;
    unsigned long long t0 = 0;
    for (;;) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        spin_guard(t0);
    }
}
// Cluster-level signalling between the epilogue warps of peer CTAs.  Default: fence-fence synchronisation — a release fence
// restricted to this CTA's shared memory (MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC.S in SASS, no GPU-scope membar on the per-tile
// critical path), relaxed remote arrive, relaxed wait, acquire fence restricted to shared::cluster.  LK_STRONG_SYNC=1 uses
// release/acquire.cluster mbarrier operations instead (MEMBAR.ALL.GPU + CCTL.IVALL each) for A/B runs on hardware.
#ifndef LK_STRONG_SYNC
#define LK_STRONG_SYNC 0
#endif
__device__ __forceinline__ uint32_t map_to_peer(uint32_t local_addr, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
    return ra;
}
// publishes this thread's earlier st.shared to the cluster; followed by a CTA barrier and peer_arrive() by a few lanes
__device__ __forceinline__ void smem_release_cluster() {
#if LK_STRONG_SYNC
    asm volatile("fence.acq_rel.cluster;" ::: "memory");
#else
    asm volatile("fence.release.sync_restrict::shared::cta.cluster;" ::: "memory");
#endif
}
__device__ __forceinline__ void peer_arrive(uint32_t local_bar, uint32_t rank) {
    const uint32_t ra = map_to_peer(local_bar, rank);
#if LK_STRONG_SYNC
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
#else
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
#endif
}
// the same without publishing anything: "I have finished READING your tile" (the loads are complete by data dependence)
__device__ __forceinline__ void peer_arrive_relaxed(uint32_t local_bar, uint32_t rank) {
    const uint32_t ra = map_to_peer(local_bar, rank);
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// wait, then acquire the peers' shared-memory writes (before ld.shared::cluster)
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    unsigned long long t0 = 0;
    for (;;) {
#if LK_STRONG_SYNC
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
#else
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
#endif
        if (ok) break;
        spin_guard(t0);
    }
#if !LK_STRONG_SYNC
    asm volatile("fence.acquire.sync_restrict::shared::cluster.cluster;" ::: "memory");
#endif
}
// wait only (write-after-read protection of the park buffer: nothing to acquire)
__device__ __forceinline__ void mbar_wait_relaxed_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    unsigned long long t0 = 0;
    for (;;) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        spin_guard(t0);
    }
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }      // the 4 epilogue warps

// ---- grid barrier: word 0 = arrival counter, word 32 = generation (separate 128-byte lines) --------------------------
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// one thread per CTA, after a CTA-level barrier behind the CTA's writers (acq_rel: release of their stores, cumulative)
__device__ __forceinline__ void grid_arrive(unsigned int* gbar, unsigned int n_cta) {
    unsigned int old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(gbar) : "memory");
    if (old == n_cta - 1) {                                  // last arrival: reset the counter, then publish the generation
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(gbar), "r"(0u) : "memory");
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gbar + 32) : "memory");
    }
}
__device__ __forceinline__ void grid_wait(const unsigned int* gbar, unsigned int target) {
    unsigned long long t0 = 0;
    while ((int)(ld_acquire_gpu(gbar + 32) - target) < 0) {
        __nanosleep(32);
        spin_guard(t0);
    }
}

// split-K geometry of a phase inside the 8-CTA cluster: S slices per tile, 8 / S tiles side by side
struct PhaseGeo { int S, gbase, s, n_act, tpr, off; };
__device__ __forceinline__ PhaseGeo phase_geo(const LayerPrm& prm, int p, int cluster, int rank) {
    PhaseGeo g;
    g.S = (p & 1) ? 4 : 8;
    const int grp = rank / g.S;
    g.s = rank % g.S;
    g.gbase = grp * g.S;
    g.tpr = prm.n_clusters * (LK_CLUSTER / g.S);             // tiles per round over the whole grid
    g.off = grp * prm.n_clusters + cluster;                  // this group's tile in round 0
    const int nt = prm.n_tiles[p];
    g.n_act = g.off < nt ? (nt - g.off + g.tpr - 1) / g.tpr : 0;
    return g;
}

struct EpiCtx {
    uint32_t tmem_base, park, tfull0, tempty0, parked8, parked4, cons8, cons4;
    uint32_t pend_bar, pend_par;       // consumed-barrier of the previous tile (peers may still be reading the park buffer)
    int j, n8, n4;                     // tiles drained so far (all / per split mode): barrier parities
    int cluster, rank, warp, lane;
    unsigned int gen0;
};

template <int QROWS, int ROLE>
__device__ __forceinline__ void epi_phase(const LayerPrm& prm, const int p, EpiCtx& e, float* s_mu, float* s_rs) {
    constexpr bool LN = ROLE != LK_RES;
    constexpr int S = LN ? 4 : 8;
    const PhaseGeo geo = phase_geo(prm, p, e.cluster, e.rank);
    const int lg = e.warp & 3, lane = e.lane, nl = lg * 32 + lane, ew = e.warp - 2;
    const uint32_t parked = LN ? e.parked4 : e.parked8, cons = LN ? e.cons4 : e.cons8;
    int& nmode = LN ? e.n4 : e.n8;

    if (p > 0) {                                             // the previous phase's outputs of EVERY CTA must be visible
        if (threadIdx.x == 64) grid_wait(prm.gbar, e.gen0 + (unsigned int)p);
        epi_bar();
    }
    if (LN) {
        if (nl < prm.M) {                                    // Chan's combination of the per-block {mean, M2}, fixed order
            const float2* pp = prm.ln_part + nl;
            float2 pb[LN_MAX_BLOCKS];
#pragma unroll
            for (int b = 0; b < LN_MAX_BLOCKS; b++)
                pb[b] = b < prm.ln_blocks ? __ldcg(pp + (long long)b * prm.part_ld) : make_float2(0.f, 0.f);
            float ms = 0.f;
#pragma unroll
            for (int b = 0; b < LN_MAX_BLOCKS; b++) ms += pb[b].x;
            const float mean = ms / (float)prm.ln_blocks;
            float m2 = 0.f;
#pragma unroll
            for (int b = 0; b < LN_MAX_BLOCKS; b++)
                if (b < prm.ln_blocks) { const float d = pb[b].x - mean; m2 += pb[b].y + (float)LN_BLOCK * d * d; }
            s_mu[nl] = mean;
            s_rs[nl] = rsqrtf(m2 / (float)(LN_BLOCK * prm.ln_blocks) + prm.ln_eps);
        }
        epi_bar();
    }
    uint32_t rbase[S];
#pragma unroll
    for (int s = 0; s < S; s++) {
        rbase[s] = map_to_peer(e.park, (uint32_t)(geo.gbase + s)) + (uint32_t)(16 * lane);
    }

#pragma unroll 1
    for (int a = 0; a < geo.n_act; a++) {
        const int tile = a * geo.tpr + geo.off;
        const int buf = e.j & 1;
        lk_wait(e.tfull0 + 8u * buf, (uint32_t)((e.j >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (e.pend_bar) {                                    // peers have finished reading the previous tile parked here
            mbar_wait_relaxed_cluster(e.pend_bar, e.pend_par);
            e.pend_bar = 0;
        }
        // ---- TMEM -> fp32 partial tile [r][128 n] parked in this CTA's shared memory ----
        const uint32_t taddr = e.tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * QROWS);
#pragma unroll 1
        for (int c0 = 0; c0 < QROWS; c0 += 16) {
            if (c0 >= prm.M) break;
            float v[16];
            tmem_ld16(taddr + c0, v);
#pragma unroll
            for (int q = 0; q < 16; q++)
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(e.park + (uint32_t)(((c0 + q) * 128 + nl) * 4)), "f"(v[q]) : "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_local(e.tempty0 + 8u * buf);      // 4 arrivals free the accumulator for the MMA thread
        smem_release_cluster();
        epi_bar();
        if (e.warp == 2) {
            smem_release_cluster();                          // cumulative over the other warps' stores (ordered by the barrier)
            if (lane < S) peer_arrive(parked, (uint32_t)(geo.gbase + lane));
        }
        mbar_wait_acquire_cluster(parked, (uint32_t)(nmode & 1));    // every peer's partial tile is parked and visible

        // ---- reduce-scatter through distributed shared memory: rows s, s+S, ...; 512 B of one peer per warp request ----
        const int n4 = tile * P_ROWS + 4 * lane;
        const float4 bv = *reinterpret_cast<const float4*>(prm.bias[p] + n4);
        float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (LN) c4 = *reinterpret_cast<const float4*>(prm.colsum[p] + n4);
        bool arrived = false;
#pragma unroll 1
        for (int r0 = geo.s + ew * S; r0 < prm.M; r0 += 8 * S) {
            const int r1 = r0 + 4 * S;
            const bool ok1 = r1 < prm.M;                     // warp-uniform
            float4 res0 = make_float4(0.f, 0.f, 0.f, 0.f), res1 = res0, a0 = res0, a1 = res0;
            if (!LN) {
                res0 = __ldcg(reinterpret_cast<const float4*>(prm.x + (long long)r0 * prm.ld_x + n4));
                if (ok1) res1 = __ldcg(reinterpret_cast<const float4*>(prm.x + (long long)r1 * prm.ld_x + n4));
            }
#pragma unroll
            for (int sh = 0; sh < S; sh += 4) {              // 8 vector loads in flight per lane
                float4 p0[4], p1[4];
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    p1[s] = make_float4(0.f, 0.f, 0.f, 0.f);
                    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(p0[s].x), "=f"(p0[s].y), "=f"(p0[s].z), "=f"(p0[s].w) : "r"(rbase[sh + s] + (uint32_t)(r0 * 512)) : "memory");
                    if (ok1)
                        asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];"
                                     : "=f"(p1[s].x), "=f"(p1[s].y), "=f"(p1[s].z), "=f"(p1[s].w) : "r"(rbase[sh + s] + (uint32_t)(r1 * 512)) : "memory");
                }
#pragma unroll
                for (int s = 0; s < 4; s++) {                // fixed order: deterministic
                    a0.x += p0[s].x; a0.y += p0[s].y; a0.z += p0[s].z; a0.w += p0[s].w;
                    a1.x += p1[s].x; a1.y += p1[s].y; a1.z += p1[s].z; a1.w += p1[s].w;
                }
            }
            if (r0 + 8 * S >= prm.M) {                       // warp-uniform: the sums consumed this warp's last remote loads
                asm volatile("// consumed %0 %1" ::"f"(a0.x + a1.x), "f"(a0.w + a1.w) : "memory");
                __syncwarp();
                if (lane < S) peer_arrive_relaxed(cons, (uint32_t)(geo.gbase + lane));
                arrived = true;
            }
#pragma unroll
            for (int q = 0; q < 2; q++) {
                if (q == 1 && !ok1) break;
                const int r = q ? r1 : r0;
                float4 a = q ? a1 : a0;
                const float4 rv = q ? res1 : res0;
                if (LN) {
                    const float mu = s_mu[r], rs = s_rs[r];
                    a.x = rs * (a.x - mu * c4.x); a.y = rs * (a.y - mu * c4.y); a.z = rs * (a.z - mu * c4.z); a.w = rs * (a.w - mu * c4.w);
                }
                float4 x = make_float4(a.x + bv.x, a.y + bv.y, a.z + bv.z, a.w + bv.w);
                if (ROLE == LK_FFN1) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                if (!LN) { x.x += rv.x; x.y += rv.y; x.z += rv.z; x.w += rv.w; }
                if (ROLE != LK_QKV) {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
                    uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
                    if (ROLE == LK_FFN1) *reinterpret_cast<uint2*>(prm.hid + (long long)r * prm.ld_hid + n4) = pk;
                    else *reinterpret_cast<uint2*>(prm.hn + (long long)r * prm.ld_hn + n4) = pk;
                }
                if (ROLE == LK_RES) *reinterpret_cast<float4*>(prm.x + (long long)r * prm.ld_x + n4) = x;
                if (ROLE == LK_QKV) *reinterpret_cast<float4*>(prm.qkv + (long long)r * prm.ld_qkv + n4) = x;
                if (ROLE == LK_RES) {                        // {mean, M2} of this 128-column block of row r, shifted one-pass
                    const float x0 = __shfl_sync(0xffffffffu, x.x, 0);
                    const float dx = x.x - x0, dy = x.y - x0, dz = x.z - x0, dw = x.w - x0;
                    float s1 = (dx + dy) + (dz + dw), s2 = (dx * dx + dy * dy) + (dz * dz + dw * dw);
#pragma unroll
                    for (int o2 = 16; o2 > 0; o2 >>= 1) {
                        s1 += __shfl_xor_sync(0xffffffffu, s1, o2);
                        s2 += __shfl_xor_sync(0xffffffffu, s2, o2);
                    }
                    if (lane == 0)
                        prm.ln_part[(long long)tile * prm.part_ld + r] =
                            make_float2(x0 + s1 * (1.f / (float)LN_BLOCK), fmaxf(s2 - s1 * s1 * (1.f / (float)LN_BLOCK), 0.f));
                }
            }
        }
        if (!arrived) {                                      // a warp without rows still reports "done reading"
            __syncwarp();
            if (lane < S) peer_arrive_relaxed(cons, (uint32_t)(geo.gbase + lane));
        }
        e.pend_bar = cons; e.pend_par = (uint32_t)(nmode & 1);
        nmode++; e.j++;
    }
    if (p + 1 < prm.n_phases) {                              // publish this CTA's outputs, then arrive at the grid barrier
        epi_bar();                                           // every epilogue thread's stores are ordered before thread 64 ...
        if (threadIdx.x == 64) {                             // ... whose release is cumulative (the grid.sync() pattern; the
            asm volatile("fence.proxy.async;" ::: "memory"); // acq_rel atomic of grid_arrive carries the GPU-scope fence).  The
            grid_arrive(prm.gbar, gridDim.x);                // next phase reads hn / hid through TMA: async-proxy fence first
        }
    }
}

template <int QROWS>
__global__ void __launch_bounds__(192, 1) gemm_layer_kernel(const __grid_constant__ LayerMaps maps, const LayerPrm prm) {
    using Cfg = LayerCfg<QROWS>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t park = base + Cfg::RING_BYTES;
    const uint32_t bar_base = park + Cfg::PARK_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
    const uint32_t tfull0 = bar_base + 8u * (2 * Cfg::STAGES);          // [2]
    const uint32_t tempty0 = tfull0 + 16u;                               // [2]
    const uint32_t parked8 = tempty0 + 16u, parked4 = parked8 + 8u, cons8 = parked4 + 8u, cons4 = cons8 + 8u;
    const uint32_t tmem_slot = cons4 + 8u;

    pdl_launch_dependents();
    __shared__ float s_mu[P_ROWS], s_rs[P_ROWS];
    __shared__ volatile unsigned int s_gen0[2];              // {generation of the grid barrier at kernel entry, valid}
    if (threadIdx.x == 0) s_gen0[1] = 0u;                    // (ordered before every reader by the __syncthreads below)
    const int ts = ts_begin(TSK_GEMM);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int cluster = blockIdx.x / LK_CLUSTER;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; b++) { mbar_init(tfull0 + 8u * b, 1); mbar_init(tempty0 + 8u * b, 4); }
        mbar_init(parked8, 8); mbar_init(parked4, 4);
        mbar_init(cons8, 4 * 8); mbar_init(cons4, 4 * 4);               // one arrival per epilogue warp of every peer
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    cluster_sync_all();                                      // peers arrive on these barriers remotely: all initialised first

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: one ring across all phases; weights run ahead of the grid barrier, activations wait =====
            uint32_t it = 0;
            unsigned int gen0 = 0;
            int carried = 0;                                 // weight stages of this phase already requested during the previous one
            for (int p = 0; p < prm.n_phases; p++) {
                const PhaseGeo geo = phase_geo(prm, p, cluster, rank);
                const int kbps = prm.kbps[p];
                const int total = geo.n_act * kbps;
                const int npre = min(total, Cfg::STAGES);
                const CUtensorMap* mp = &maps.p[p];
                const CUtensorMap* mq = &maps.q[p];
                const int kb0 = geo.s * kbps;
                for (int i = carried; i < npre; i++) {       // immutable weights: requested before the dependency is met
                    const uint32_t g = it + (uint32_t)i;
                    const int slot = (int)(g % Cfg::STAGES);
                    lk_wait(empty_bar(slot), ((g / Cfg::STAGES) & 1) ^ 1);
                    mbar_expect_tx(full_bar(slot), Cfg::STAGE_BYTES);
                    const int tile = (i / kbps) * geo.tpr + geo.off;
                    tma_load_2d(base + slot * Cfg::STAGE_BYTES, mp, full_bar(slot), (kb0 + i % kbps) * BK, tile * P_ROWS);
                }
                carried = 0;
                if (total < Cfg::STAGES && p + 1 < prm.n_phases) {
                    // a phase shorter than the ring (out-proj: 4 k-blocks per CTA): the free stages take the NEXT phase's first
                    // weight tiles before this phase's dependency is even met — at kernel entry that is HBM traffic issued
                    // under the tail of the attention kernel.  Non-blocking: a stage that is still in use ends the run.
                    const PhaseGeo gn = phase_geo(prm, p + 1, cluster, rank);
                    const int kn = prm.kbps[p + 1];
                    const int extra = min(gn.n_act * kn, Cfg::STAGES - total);
                    for (int i = 0; i < extra; i++) {
                        const uint32_t g = it + (uint32_t)(total + i);
                        const int slot = (int)(g % Cfg::STAGES);
                        if (!mbar_test(empty_bar(slot), ((g / Cfg::STAGES) & 1) ^ 1)) break;
                        mbar_expect_tx(full_bar(slot), Cfg::STAGE_BYTES);
                        const int tile = (i / kn) * gn.tpr + gn.off;
                        tma_load_2d(base + slot * Cfg::STAGE_BYTES, &maps.p[p + 1], full_bar(slot), (gn.s * kn + i % kn) * BK, tile * P_ROWS);
                        carried = i + 1;
                    }
                }
                if (p == 0) {
                    pdl_wait();                              // attention output / residual stream of the preceding kernels
                    ts_dep(ts);
                } else {
                    if (p == 1) {
                        // the base generation is read ONCE per CTA, by the thread that later arrives for this CTA (thread 64):
                        // a second read here could come after barrier 0 has already completed (a CTA without work in phase 0
                        // arrives immediately) and would then wait for a generation that needs this very thread to progress
                        unsigned long long t0 = 0;
                        while (s_gen0[1] == 0u) spin_guard(t0);
                        gen0 = s_gen0[0];
                    }
                    grid_wait(prm.gbar, gen0 + (unsigned int)p);
                    ts_aux(ts, p - 1);
                }
                asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy writes of other CTAs -> TMA reads below
                for (int i = 0; i < npre; i++) {
                    const int slot = (int)((it + (uint32_t)i) % Cfg::STAGES);
                    tma_load_2d(base + slot * Cfg::STAGE_BYTES + P_BYTES, mq, full_bar(slot), (kb0 + i % kbps) * BK, 0);
                }
                for (int i = npre; i < total; i++) {
                    const uint32_t g = it + (uint32_t)i;
                    const int slot = (int)(g % Cfg::STAGES);
                    lk_wait(empty_bar(slot), ((g / Cfg::STAGES) & 1) ^ 1);
                    mbar_expect_tx(full_bar(slot), Cfg::STAGE_BYTES);
                    const int tile = (i / kbps) * geo.tpr + geo.off;
                    const uint32_t sp = base + slot * Cfg::STAGE_BYTES;
                    tma_load_2d(sp, mp, full_bar(slot), (kb0 + i % kbps) * BK, tile * P_ROWS);
                    tma_load_2d(sp + P_BYTES, mq, full_bar(slot), (kb0 + i % kbps) * BK, 0);
                }
                it += (uint32_t)total;
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer: two TMEM accumulators, tile j+1 accumulates while the epilogue parks tile j =====
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(QROWS >> 3) << 17) | ((uint32_t)(P_ROWS >> 4) << 24);
            uint32_t it = 0;
            int j = 0;
            for (int p = 0; p < prm.n_phases; p++) {
                const PhaseGeo geo = phase_geo(prm, p, cluster, rank);
                const int kbps = prm.kbps[p];
                for (int a = 0; a < geo.n_act; a++, j++) {
                    const int buf = j & 1;
                    lk_wait(tempty0 + 8u * buf, (uint32_t)(((j >> 1) & 1) ^ 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t acc = tmem_base + (uint32_t)(buf * QROWS);
                    for (int kb = 0; kb < kbps; kb++, it++) {
                        const int slot = (int)(it % Cfg::STAGES);
                        lk_wait(full_bar(slot), (it / Cfg::STAGES) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t sp = base + slot * Cfg::STAGE_BYTES;
                        const uint64_t da = make_desc(sp), db = make_desc(sp + P_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 16; k++) umma_bf16(acc, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_commit(empty_bar(slot));
                    }
                    umma_commit(tfull0 + 8u * buf);
                }
            }
        }
    } else {
        // ===== epilogue warps =====
        pdl_wait();
        EpiCtx e;
        e.tmem_base = tmem_base; e.park = park; e.tfull0 = tfull0; e.tempty0 = tempty0;
        e.parked8 = parked8; e.parked4 = parked4; e.cons8 = cons8; e.cons4 = cons4;
        e.pend_bar = 0; e.pend_par = 0; e.j = 0; e.n8 = 0; e.n4 = 0;
        e.cluster = cluster; e.rank = rank; e.warp = warp; e.lane = lane;
        e.gen0 = 0u;
        if (threadIdx.x == 64) {                             // after griddepcontrol.wait: every earlier launch has completed
            e.gen0 = ld_acquire_gpu(prm.gbar + 32);
            s_gen0[0] = e.gen0;
            __threadfence_block();
            s_gen0[1] = 1u;
        }
        for (int p = 0; p < prm.n_phases; p++) {
            switch (p) {
                case 1: epi_phase<QROWS, LK_FFN1>(prm, p, e, s_mu, s_rs); break;
                case 3: epi_phase<QROWS, LK_QKV>(prm, p, e, s_mu, s_rs); break;
                default: epi_phase<QROWS, LK_RES>(prm, p, e, s_mu, s_rs); break;
            }
        }
        if (e.pend_bar) mbar_wait_relaxed_cluster(e.pend_bar, e.pend_par);   // no exit while a peer still reads this CTA's tile
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncwarp();
    __syncthreads();
    ts_end(ts);
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    }
}

int lk_qrows(int M) { return M <= 16 ? 16 : (M <= 32 ? 32 : (M <= 64 ? 64 : 128)); }

template <int QROWS>
int lk_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int n_clusters, cudaStream_t s) {
    using Cfg = LayerCfg<QROWS>;
    static bool attr_done = false;
    if (!attr_done) {
        SSRB_CUDA(cudaFuncSetAttribute(gemm_layer_kernel<QROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_done = true;
    }
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3(n_clusters * LK_CLUSTER); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = s;
    int na = 0;
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = LK_CLUSTER; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    na++;
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        na++;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    return 0;
}

// clusters of 8 that can be co-resident at one CTA per SM (every CTA of the grid spins at the grid barrier, so the whole grid
// must fit at once); cached per instantiation.  B200: 8 GPCs of 16-20 SMs -> 2 clusters each -> 16.
template <int QROWS>
int lk_max_clusters(int* out) {
    static int cached = -1;
    if (cached < 0) {
        cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[2];
        SSRB_TRY(lk_config<QROWS>(cfg, attr, LK_MAX_CLUSTERS, nullptr));
        cfg.numAttrs = 1;                                    // the query takes the cluster shape only
        int n = 0;
        SSRB_CUDA(cudaOccupancyMaxActiveClusters(&n, gemm_layer_kernel<QROWS>, &cfg));
        cached = n;
    }
    *out = cached;
    return 0;
}

template <int QROWS>
int lk_launch(const LayerMaps& maps, LayerPrm& prm, cudaStream_t s) {
    int nmax = 0;
    SSRB_TRY(lk_max_clusters<QROWS>(&nmax));
    SSRB_CHECK(nmax >= 8, "gemm_layer: fewer than 8 co-resident clusters of 8 CTAs on this device");
    prm.n_clusters = std::min(nmax, LK_MAX_CLUSTERS);
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[2];
    SSRB_TRY(lk_config<QROWS>(cfg, attr, prm.n_clusters, s));
    SSRB_CUDA(cudaLaunchKernelEx(&cfg, gemm_layer_kernel<QROWS>, maps, prm));
    g_launch_count++;
    return 0;
}

}  // namespace

bool gemm_layer_supported(int M, int D, int F) {
    if (M < 1 || M > 128) return false;
    if (D % 512 != 0 || F % 512 != 0 || D > LN_BLOCK * LN_MAX_BLOCKS) return false;   // k-blocks divisible by the 8-way split
    int n = 0, rc = 0;
    switch (lk_qrows(M)) {
        case 16: rc = lk_max_clusters<16>(&n); break;
        case 32: rc = lk_max_clusters<32>(&n); break;
        case 64: rc = lk_max_clusters<64>(&n); break;
        default: rc = lk_max_clusters<128>(&n); break;
    }
    return rc == 0 && n >= 8;
}

int gemm_layer(const LayerChainArgs& a, cudaStream_t s) {
    SSRB_CHECK(a.M >= 1 && a.M <= 128 && a.D % 512 == 0 && a.F % 512 == 0 && a.D <= LN_BLOCK * LN_MAX_BLOCKS, "gemm_layer: unsupported shape");
    SSRB_CHECK(a.ao && a.x && a.hn && a.hid && a.ln_part && a.wo && a.w1f && a.w2 && a.bo && a.b1f && a.c1 && a.b2 && a.gbar, "gemm_layer: null argument");
    SSRB_CHECK(a.part_ld >= a.M, "gemm_layer: part_ld must cover the M rows");
    const bool qkv = a.wqkv_next != nullptr;
    SSRB_CHECK(!qkv || (a.qkv && a.bqkv_next && a.cqkv_next), "gemm_layer: the QKV phase needs its bias, column sums and output");
    const int D = a.D, F = a.F, q = lk_qrows(a.M);
    LayerMaps maps;
    memset(&maps, 0, sizeof(maps));
    LayerPrm prm{};
    prm.M = a.M; prm.n_phases = qkv ? 4 : 3;
    const int nkbD = D / BK, nkbF = F / BK;
    // phase 0: out-proj
    SSRB_TRY(tc_make_map(&maps.p[0], a.wo, D, D, D, P_ROWS)); SSRB_TRY(tc_make_map(&maps.q[0], a.ao, a.M, D, D, q));
    prm.n_tiles[0] = D / P_ROWS; prm.kbps[0] = nkbD / 8; prm.bias[0] = a.bo; prm.colsum[0] = nullptr;
    // phase 1: FFN1
    SSRB_TRY(tc_make_map(&maps.p[1], a.w1f, F, D, D, P_ROWS)); SSRB_TRY(tc_make_map(&maps.q[1], a.hn, a.M, D, D, q));
    prm.n_tiles[1] = F / P_ROWS; prm.kbps[1] = nkbD / 4; prm.bias[1] = a.b1f; prm.colsum[1] = a.c1;
    // phase 2: FFN2
    SSRB_TRY(tc_make_map(&maps.p[2], a.w2, D, F, F, P_ROWS)); SSRB_TRY(tc_make_map(&maps.q[2], a.hid, a.M, F, F, q));
    prm.n_tiles[2] = D / P_ROWS; prm.kbps[2] = nkbF / 8; prm.bias[2] = a.b2; prm.colsum[2] = nullptr;
    // phase 3: the next layer's QKV projection
    if (qkv) {
        SSRB_TRY(tc_make_map(&maps.p[3], a.wqkv_next, 3 * D, D, D, P_ROWS)); SSRB_TRY(tc_make_map(&maps.q[3], a.hn, a.M, D, D, q));
        prm.n_tiles[3] = 3 * D / P_ROWS; prm.kbps[3] = nkbD / 4; prm.bias[3] = a.bqkv_next; prm.colsum[3] = a.cqkv_next;
    }
    prm.x = a.x; prm.ld_x = D;
    prm.hn = reinterpret_cast<bf16*>(a.hn); prm.ld_hn = D;
    prm.hid = reinterpret_cast<bf16*>(a.hid); prm.ld_hid = F;
    prm.qkv = a.qkv; prm.ld_qkv = 3 * D;
    prm.ln_part = a.ln_part; prm.part_ld = a.part_ld; prm.ln_blocks = D / LN_BLOCK; prm.ln_eps = a.ln_eps;
    prm.gbar = a.gbar;
    switch (q) {
        case 16: return lk_launch<16>(maps, prm, s);
        case 32: return lk_launch<32>(maps, prm, s);
        case 64: return lk_launch<64>(maps, prm, s);
        default: return lk_launch<128>(maps, prm, s);
    }
}

int ts_arm_gemm_layer(const TsBuf& t) { return ts_arm_tu(t); }

}  // namespace ssrb
