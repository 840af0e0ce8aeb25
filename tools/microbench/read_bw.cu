// read_bw.cu — what a pure READ stream achieves on this B200 (MEASURED_PEAKS.json's hbm_gbs is a copy: read + write).
// Two kernels over a buffer much larger than L2: (a) 16-byte LDG with 8 independent loads in flight per thread,
// (b) the decode-attention staging pattern: one lane issues 16 KB cp.async.bulk copies into a 3-stage x 32 KB ring,
// 2 CTAs per SM, consumers only touch one word per tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o read_bw read_bw.cu && ./read_bw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__global__ void __launch_bounds__(256) ldg_kernel(const uint4* __restrict__ p, size_t n, unsigned int* sink) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned int acc = 0;
    for (; i + 7 * stride < n; i += 8 * stride) {
        uint4 v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = __ldcs(p + i + j * stride);
#pragma unroll
        for (int j = 0; j < 8; j++) acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int STAGES = 3, STAGE_BYTES = 32768;
__global__ void __launch_bounds__(160, 2) bulk_kernel(const char* __restrict__ p, size_t bytes_per_cta, unsigned int* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t ring = smem_u32(smem), bar0 = ring + STAGES * STAGE_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8 * s), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 24 + 8 * s), "r"(4));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const char* base = p + (size_t)blockIdx.x * bytes_per_cta;
    const int tiles = (int)(bytes_per_cta / STAGE_BYTES);
    auto wait = [](uint32_t bar, uint32_t parity) {
        uint32_t ok;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        } while (!ok);
    };
    if (warp == 4) {
        if (lane == 0)
            for (int i = 0; i < tiles; i++) {
                const int s = i % STAGES; const uint32_t ph = (i / STAGES) & 1;
                wait(bar0 + 24 + 8 * s, ph ^ 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * s), "r"(STAGE_BYTES) : "memory");
                for (int h = 0; h < 2; h++)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(ring + s * STAGE_BYTES + h * 16384), "l"(base + (size_t)i * STAGE_BYTES + h * 16384), "r"(16384), "r"(bar0 + 8 * s) : "memory");
            }
        return;
    }
    unsigned int acc = 0;
    for (int i = 0; i < tiles; i++) {
        const int s = i % STAGES; const uint32_t ph = (i / STAGES) & 1;
        wait(bar0 + 8 * s, ph);
        acc ^= *reinterpret_cast<const unsigned int*>(smem + s * STAGE_BYTES + tid * 16);
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 24 + 8 * s) : "memory");
    }
    if (acc == 0x12345678u) *sink = acc;
}

int main() {
    const size_t bytes = (size_t)8 << 30;
    char* buf; unsigned int* sink;
    cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int ctas_per_sm : {4, 8, 16}) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            ldg_kernel<<<148 * ctas_per_sm, 256>>>((const uint4*)buf, bytes / 16, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("ldg128 x8 in flight, %2d CTAs/SM : %.1f GB/s\n", ctas_per_sm, bytes / best / 1e6);
    }
    const int smem = STAGES * STAGE_BYTES + 64;
    cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int waves : {1, 4, 16}) {
        const int ctas = 296 * waves;
        const size_t per = (bytes / ctas) / STAGE_BYTES * STAGE_BYTES;
        float best = 1e9f;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            bulk_kernel<<<ctas, 160, smem>>>(buf, per, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("bulk ring 3x32KB, 2 CTAs/SM, %2d waves (%.0f KB per CTA): %.1f GB/s\n", waves, per / 1024.0, (double)per * ctas / best / 1e6);
    }
    // the attention geometry: 1024 CTAs x 440 KB
    {
        const int ctas = 1024; const size_t per = 13 * STAGE_BYTES + 16384 * 2;
        float best = 1e9f;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            bulk_kernel<<<ctas, 160, smem>>>(buf, per / STAGE_BYTES * STAGE_BYTES, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("bulk ring, 1024 CTAs x %zu KB (attention geometry, %0.1f us): %.1f GB/s\n", per / STAGE_BYTES * STAGE_BYTES / 1024, best * 1e3,
               (double)(per / STAGE_BYTES * STAGE_BYTES) * ctas / best / 1e6);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
