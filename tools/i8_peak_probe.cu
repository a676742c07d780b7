// Dense INT8 tensor-core ceiling of this GPU for the instruction the Gram kernel uses:
// tcgen05.mma.cta_group::2.kind::i8, M = N = 256, K = 32, operands resident in shared memory (SWIZZLE_128B
// K-major tiles, no loads in the loop), accumulators in TMEM.  One CTA pair per SM pair, the leader's elected
// thread issues back-to-back MMAs exactly like warp 1 of k_gram_i8<2> (8 per 128-byte K stage, one commit per
// stage).  Prints burst (one short launch) and sustained (back-to-back launches for ~4 s, under the power cap)
// TOP/s; the sustained figure is the roofline denominator of a kernel timed inside a long step.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/i8_peak_probe tools/i8_peak_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kTileBytes = 128 * 128;   // 128 rows x 128 bytes of K
constexpr int kStages = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t instr_desc(uint32_t mn) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((mn >> 3) << 17) | ((mn >> 4) << 24);
}

__global__ void __launch_bounds__(128, 1) k_probe(int iters) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t bars = base + 3 * kTileBytes;
    const uint32_t tmem_slot = bars + 8 * kStages;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    // operand bytes: pseudo-random, like the digit planes (the switching activity of the data sets the power the
    // MMAs draw, and the sustained figure is a power-capped one)
    for (int i = threadIdx.x; i < 3 * kTileBytes / 4; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u + 12345u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + 4 * i), "r"(h) : "memory");
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
    if (threadIdx.x == 32 && rank == 0) {
        const uint64_t da = umma_smem_desc(base), dar = umma_smem_desc(base + kTileBytes), db = umma_smem_desc(base + 2 * kTileBytes);
        const uint32_t idesc = instr_desc(256);
        uint32_t phase = 0;
        int stage = 0;
        for (int it = 0; it < iters; ++it) {
            if (it >= kStages) mbar_wait(bars + 8 * stage, phase ^ 1u);   // at most kStages stages of MMAs in flight
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint32_t acc = (it | kk) ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem_base), "l"(da + 2u * kk), "l"(db + 2u * kk), "r"(idesc), "r"(acc) : "memory");
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem_base + 256u), "l"(dar + 2u * kk), "l"(db + 2u * kk), "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bars + 8 * stage) : "memory");
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        // drain: the last kStages commits
        for (int s = 0; s < kStages && s < iters; ++s) {
            mbar_wait(bars + 8 * stage, phase ^ 1u);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

static float launch(int pairs, int iters, int reps) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 3 * kTileBytes + 2048;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a));
    for (int r = 0; r < reps; ++r) CK(cudaLaunchKernelEx(&cfg, k_probe, iters));
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int pairs = sms / 2;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * kTileBytes + 2048));
    const int iters = 200000;                                   // stages per pair and launch
    const double ops_per_launch = (double)pairs * iters * 8.0 * 2.0 * 256.0 * 256.0 * 32.0;
    launch(pairs, 1000, 1);                                     // warm-up
    double best = 0;
    for (int r = 0; r < 5; ++r) {
        const float ms = launch(pairs, iters, 1);
        const double tops = ops_per_launch / (ms * 1e-3) / 1e12;
        if (tops > best) best = tops;
    }
    // sustained: back-to-back launches for about 4 s
    const float ms1 = launch(pairs, iters, 1);
    const int reps = (int)(4000.0f / ms1) + 1;
    const float ms = launch(pairs, iters, reps);
    const double sustained = ops_per_launch * reps / (ms * 1e-3) / 1e12;
    printf("{\"probe\": \"tcgen05.mma.cta_group::2.kind::i8 M=N=256 K=32, SMEM-resident operands\", \"sms\": %d, "
           "\"pairs\": %d, \"burst_tops\": %.1f, \"burst_ms_per_launch\": %.3f, \"sustained_tops\": %.1f, "
           "\"sustained_seconds\": %.2f}\n",
           sms, pairs, best, ops_per_launch / best / 1e9, sustained, ms * 1e-3);
    return 0;
}
