// Fidelity Gram matrix on the 5th-generation tensor cores: split-integer ("Ozaki scheme") FP64 emulation.
// Replaces the O(N^2) Python pair loop `np.abs(np.matmul(x.conj(), y))**2` of
// FidelityKernelStatevector.evaluate_kernel_sv (src/squlearn/kernel/lowlevel_kernel/
// fidelity_kernel_statevector.py:303-310, 358-383) for large problems; semantics (symmetric mirror, diagonal
// policy) as in gram_kernels.cu.
//
// tcgen05 has no f64 kind and B200's FP64 DMMA peaks at ~37 TFLOP/s, while the INT8 kind runs at
// ~4.5 POP/s.  Every amplitude component v of a state row is therefore written in fixed point relative to
// the row's power-of-two scale 2^e,
//        v = 2^e * sum_{s=0}^{S-1} d_s * 2^(-6 - 8 s),      d_0 in [-65, 65], d_s in [-128, 127]  (S = 6)
// (exact integer digit decomposition of round(v * 2^(46 - e)): 47 significant bits per row), and the complex
// inner products become sums of EXACT int8 x int8 -> int32 tensor-core products, grouped by weight class
// c = i + j:
//        <x|y> = 2^(ex + ey - 12) * sum_c 2^(-8 c) * sum_{i + j = c} D_i(x)^T D_j(y)
// Classes c <= 5 (21 slice pairs) plus the pair (3, 3) are kept; what is dropped is below 2^-48 of the row
// scales (measured |K - K_fp64| ~ 1e-13 .. 1e-11, see DESIGN.md for the bound).
//
// Complex arithmetic without data rearrangement in the GEMM: a state row is the real vector
// (re_0, im_0, re_1, im_1, ...), K' = 2 * dim.  With  nat(v) = (vr_k, vi_k)  and  rot(v) = (-vi_k, vr_k):
//        Re <x|y> = nat(x) . nat(y),     Im <x|y> = rot(x) . nat(y)
// so ONE B tile (digit plane of nat(y)) feeds two accumulators through two A tiles (digit planes of nat(x) and
// rot(x)): the column operand only needs its 6 nat planes, which is what the multi-GPU exchange sends.
//
// SECOND SCHEME (B200SQ_GRAM_CRT, "Ozaki scheme II"): instead of 6 x 6 digit pairs, the integers
// q = round(v * 2^(B - e)) are reduced modulo s pairwise coprime numbers p_m <= 256 (256 and odd numbers); one exact
// int8 x int8 -> int32 product per modulus gives the inner product mod p_m, and the Chinese remainder theorem recovers
// the exact integer inner product (|.| < P / 2, P = prod p_m) in the epilogue kernel with a double-double accumulation
// of k_m / p_m.  The row scale 2^e comes from the row's 2-NORM, not its maximum: by Cauchy-Schwarz
// |q_x . q_y| <= |q_x| |q_y| <= R^2 with R = sqrt(P / 2), so e is the smallest integer with 2^(B - e) |x| <= R and no
// factor K' is wasted on a worst case that cannot happen.  s = 11, B = 43 for K' = 2^17: 11 tensor-core products
// instead of 22 for a worst-case bound of 2 sqrt(K') 2^-B = 8e-11 on K (unit-norm rows; the error is the input
// quantisation only: no product terms are dropped).  Planes: [2 s][rows][2 dim] int8, residues of nat(v) then of
// rot(v); W holds int32 residue sums per (tile, modulus).
//
// JOB TABLE.  One launch covers several Gram blocks ("jobs") that share the local row operand: job j multiplies
// rows [x0, x0 + nx) of the local planes with rows [y0, y0 + ny) of its own column planes (the local planes for
// the diagonal block, a received block otherwise) and writes its own K pointer.  A job may carry a READY FLAG:
// its work units start when *flag >= value, so blocks that are still being copied in by the copy engines
// (peer pushes over NVLink) are multiplied as they land, inside the same persistent kernel.
//
// Kernel k_gram_i8 (one CTA per SM, persistent over work units = (tile, K-chunk), job-major order):
//   warp 0    TMA producer: 3 x 16 KiB boxes (Anat_i, Arot_i, Bnat_j; 128 rows x 128 B, SWIZZLE_128B) per stage
//   warp 1    MMA issuer:   8 x tcgen05.mma.kind::i8 (M = N = 128, K = 32) per stage into TMEM (s32)
//   warp 2    TMEM allocator (512 columns: 2 accumulator stages x (Re 128 + Im 128))
//   warps 4-7 epilogue: tcgen05.ld -> int32 -> double * 2^(-8 c) -> red.add.f64 into the per-tile FP64
//             accumulator W (int32 accumulation is exact for <= 7 pairs x 16 KiB: 7 * 2^14 * 2^14 < 2^31)
// k_slice_i8 produces the digit planes, k_gram_i8_final turns W into K = |.|^2 with the symmetric
// mirror / diagonal policy of the FP64 kernel.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>

#include <algorithm>

#include "kernels.h"

namespace b200sq {

namespace {

constexpr int kSlices = 6;
constexpr int kTopBits = 6 + 8 * (kSlices - 1);  // v * 2^(kTopBits - e) is the integer that gets sliced
constexpr int kTile = 128;
constexpr int kStageK = 128;                     // bytes of K per pipeline stage (one swizzle atom)
constexpr int kTileBytes = kTile * kStageK;      // 16 KiB
constexpr int kStageBytes = 3 * kTileBytes;      // A, Bnat, Brot
constexpr int kStages = 4;
constexpr int kChunkK = 16384;                   // bytes of K per int32 accumulation chunk
constexpr int kMaxClasses = 16, kMaxPairs = 8;
constexpr int kCrtMaxBits = 46;                   // upper limit of B
constexpr double kCrtMaxLog2R = 46.9;             // upper limit of the scaled row norm: |q| <= 2^46.9 fits the 6 balanced
                                                  // base-256 exchange digits (top digit <= 120) and is exact in a double
constexpr int kCrtMaxModuli = 16;
// pairwise coprime moduli (256; 255 = 3 5 17, 253 = 11 23, 247 = 13 19, 217 = 7 31, the rest prime): symmetric
// residues fit int8 (256: -128 stands for +-128), and 127^2 * 2^17 < 2^31 keeps a whole 128 KiB K chunk exact in int32
// (for p = 256 a wrapped int32 sum is still right modulo 256)
#define B200SQ_CRT_MODULI 256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193
__constant__ int c_crt_moduli[kCrtMaxModuli] = {B200SQ_CRT_MODULI};
__constant__ double c_crt_inv[kCrtMaxModuli] = {1.0 / 256, 1.0 / 255, 1.0 / 253, 1.0 / 251, 1.0 / 247, 1.0 / 241, 1.0 / 239, 1.0 / 233,
                                                1.0 / 229, 1.0 / 227, 1.0 / 223, 1.0 / 217, 1.0 / 211, 1.0 / 199, 1.0 / 197, 1.0 / 193};
static const int h_crt_moduli[kCrtMaxModuli] = {B200SQ_CRT_MODULI};
constexpr int kThreads = 256;

struct I8Schedule {
    int n_classes;
    int cls_shift[kMaxClasses];   // weight of the class = 2^(-8 * cls_shift)
    int n_pairs[kMaxClasses];
    uint8_t pi[kMaxClasses][kMaxPairs], pj[kMaxClasses][kMaxPairs];
};

constexpr int kMaxJobs = 12;

struct I8Job {
    int64_t row0_a;           // first row of the job's row operand inside the local planes
    int64_t row0_b, rows_b;   // first row / rows per digit plane of the column operand
    int64_t ldk;
    double* K;                // [nx][ldk]
    double* K_mirror;         // nullptr, or [ny][ldm]: the transpose of the block (rectangular jobs)
    int64_t ldm;
    const int* ex;            // row exponents of the local planes, of the column planes
    const int* ey;
    const uint32_t* ready;    // nullptr, or: the column planes are complete when *ready >= ready_value
    const double* align_yx;   // nullptr, or row / column labels: align_out[0] += w sum K_ij yx_i yy_j,
    const double* align_yy;   //                                   align_out[1] += w sum K_ij^2
    double* align_out;
    double align_weight;
    uint32_t ready_value;
    uint32_t flags;           // B200SQ_GRAM_* of this block
    int32_t nx, ny;
    int32_t tiles_n, tile0, n_tiles, symmetric;
    int32_t map_b;            // which tensor map holds the column planes
    int32_t pad_;
};

struct I8Args {
    int64_t rows_a;           // rows per digit plane of the local (row) operand
    int kbytes;               // K' = 2 * dim bytes per row
    int n_chunks, chunk_bytes;
    int n_jobs, n_tiles;      // n_tiles: over all jobs
    int mn;                   // edge of an output tile: 128 (one CTA) or 256 (CTA pair)
    int n_planes;             // nat planes per operand (rot planes follow): 6 digit slices, or s residue planes
    int crt;                  // 1: CRT scheme (classes = moduli, W = int32 residue sums [n_tiles][s][2][mn][mn])
    int crt_c[kCrtMaxModuli];    // (P / p_m)^-1 mod p_m
    double crt_scale;            // P * 2^(-2 B)
    double* W;                // [n_tiles][2][mn cols][mn rows]
    I8Schedule sched;
    I8Job jobs[kMaxJobs];
};

struct I8Maps {
    CUtensorMap b[kMaxJobs];  // column-operand planes (deduplicated by base pointer)
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Both CTAs of a pair run this; the transaction bytes land on the LEADER's barrier (peer bit cleared).
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar, uint32_t cta) {  // arrive on CTA `cta` of the cluster
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(cta)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// D[tmem] (+)= A[smem] * B[smem], int8 x int8 -> int32, K = 32; CG = 1: M = N = 128 on one SM,
// CG = 2: M = N = 256 on a CTA pair (each CTA holds 128 rows of A and 128 of the 256 rows of B).
template <int CG>
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
    if (CG == 1)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
}
// Arrive on `bar` (same offset in every CTA of the pair) when all MMAs issued so far have completed.
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                     : "memory");
    else
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
            ::"r"(bar), "h"((uint16_t)3)
            : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void red_add_f64(double* p, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// K-major operand tile, 128-byte rows, SWIZZLE_128B (8-row atoms of 1024 bytes): LBO unused, SBO = 1024.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::i8: D = s32 (bits 4-5 = 2), A = B = signed int8 (bits 7-9 = 1, 10-12 = 1), both K-major,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28.
constexpr uint32_t instr_desc(uint32_t mn) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((mn >> 3) << 17) | ((mn >> 4) << 24);
}

__device__ __forceinline__ void decode_tile(const I8Job& A, int t, int& bm, int& bn) {
    if (A.symmetric) {
        bm = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while ((bm + 1) * (bm + 2) / 2 <= t) ++bm;
        while (bm * (bm + 1) / 2 > t) --bm;
        bn = t - bm * (bm + 1) / 2;
    } else {
        bm = t / A.tiles_n;
        bn = t - bm * A.tiles_n;
    }
}

// ---------------------------------------------------------------------------------------------

// Work unit u -> (job, tile inside the job, K chunk).  Units are job-major (a job's column planes may still be
// in flight: earlier jobs run first), chunk-major inside a job.
// The digit scheme runs all its weight classes inside a unit (c0 = 0, c1 = n_classes); the CRT scheme makes every
// (tile, K chunk, modulus) a unit of its own (upc = n_classes units per tile and chunk, modulus-major so that the
// workers of a wave read the same residue planes): a whole 128 KiB of K accumulates in TMEM before one drain.
__device__ __forceinline__ int decode_unit(const I8Args& A, int u, int& tile, int& chunk, int& c0, int& c1) {
    const int upc = A.crt ? A.sched.n_classes : 1;
    const int per_tile = A.n_chunks * upc;
    int j = 0;
    while (j + 1 < A.n_jobs && u >= (A.jobs[j].tile0 + A.jobs[j].n_tiles) * per_tile) ++j;
    int ul = u - A.jobs[j].tile0 * per_tile;
    const int nt = A.jobs[j].n_tiles;
    const int cls = ul / (A.n_chunks * nt);
    ul -= cls * A.n_chunks * nt;
    chunk = ul / nt;
    tile = ul - chunk * nt;
    c0 = A.crt ? cls : 0;
    c1 = A.crt ? cls + 1 : A.sched.n_classes;
    return j;
}

// Blocks that never arrive must not hang the GPU: a producer gives up after kReadyTimeoutNs, counts the event here
// (b200sq_gram_ready_timeouts) and multiplies whatever the buffer holds; the host side treats a non-zero count
// as an error.
__device__ unsigned int g_ready_timeouts = 0;
constexpr unsigned long long kReadyTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int CG>
__global__ void __launch_bounds__(kThreads, 1)
k_gram_i8(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ I8Maps mapsB,
          const __grid_constant__ I8Args A) {
    constexpr int MN = kTile * CG;            // edge of the output tile of a CTA (pair)
    constexpr int ACC = CG == 1 ? 2 : 1;      // accumulator stages: 2 * MN columns each, 512 columns in total
    constexpr uint32_t kIdesc = instr_desc(MN);
    extern __shared__ unsigned char smem_dyn[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t bars = smem_base + kStages * kStageBytes;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages + 2 + s); };
    const uint32_t tmem_slot = bars + 8u * (2 * kStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), CG);   // producer arrivals of every CTA of the pair (leader's barrier is used)
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 4 * CG);  // one arrival per epilogue warp of the pair
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 2) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    const int n_units = A.n_tiles * A.n_chunks * (A.crt ? A.sched.n_classes : 1);
    const int worker = blockIdx.x / CG, n_workers = gridDim.x / CG;
    const I8Schedule& S = A.sched;

    if (warp == 0) {
        // ===== TMA producer (every CTA loads its own 128 rows of Anat, Arot and B) =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t ready_jobs = 0;   // jobs whose ready flag has been seen
            for (int u = worker; u < n_units; u += n_workers) {
                int tile, chunk, c0, c1;
                const int ji = decode_unit(A, u, tile, chunk, c0, c1);
                const I8Job& J = A.jobs[ji];
                if (J.ready && !(ready_jobs >> ji & 1u)) {
                    // column planes pushed by a peer's copy engine: wait for its flag, then order the TMA
                    // (async proxy) reads behind the acquire
                    const unsigned long long t_start = global_timer_ns();
                    while (ld_acquire_sys(J.ready) < J.ready_value) {
                        __nanosleep(200);
                        if (global_timer_ns() - t_start > kReadyTimeoutNs) {
                            atomicAdd(&g_ready_timeouts, 1u);
                            break;
                        }
                    }
                    asm volatile("fence.proxy.async;" ::: "memory");
                    ready_jobs |= 1u << ji;
                }
                const CUtensorMap* mapB = &mapsB.b[J.map_b];
                int bm, bn;
                decode_tile(J, tile, bm, bn);
                const int k0 = chunk * A.chunk_bytes;
                const int k1 = min(A.kbytes, k0 + A.chunk_bytes);
                for (int ci = c0; ci < c1; ++ci)
                    for (int p = 0; p < S.n_pairs[ci]; ++p) {
                        const int row_an = (int)(S.pi[ci][p] * A.rows_a + J.row0_a) + bm * MN + (int)cta_rank * kTile;
                        const int row_ar = (int)((A.n_planes + S.pi[ci][p]) * A.rows_a + J.row0_a) + bm * MN + (int)cta_rank * kTile;
                        const int row_b = (int)(S.pj[ci][p] * J.rows_b + J.row0_b) + bn * MN + (int)cta_rank * kTile;
                        for (int k = k0; k < k1; k += kStageK) {
                            mbar_wait(empty_bar(stage), phase ^ 1u);
                            const uint32_t dst = smem_base + stage * kStageBytes;
                            if (CG == 1) {
                                mbar_expect_tx(full_bar(stage), kStageBytes);
                                tma_load_2d(dst, &mapA, full_bar(stage), k, row_an);
                                tma_load_2d(dst + kTileBytes, &mapA, full_bar(stage), k, row_ar);
                                tma_load_2d(dst + 2 * kTileBytes, mapB, full_bar(stage), k, row_b);
                            } else {
                                tma_load_2d_2sm(dst, &mapA, full_bar(stage), k, row_an);
                                tma_load_2d_2sm(dst + kTileBytes, &mapA, full_bar(stage), k, row_ar);
                                tma_load_2d_2sm(dst + 2 * kTileBytes, mapB, full_bar(stage), k, row_b);
                                if (leader) mbar_expect_tx(full_bar(stage), 2 * kStageBytes);
                                else mbar_arrive_cta(full_bar(stage), 0);
                            }
                            if (++stage == kStages) { stage = 0; phase ^= 1u; }
                        }
                    }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (lane == 0 && leader) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t acc_it = 0;
            for (int u = worker; u < n_units; u += n_workers) {
                int tile, chunk, c0, c1;
                decode_unit(A, u, tile, chunk, c0, c1);
                const int k0 = chunk * A.chunk_bytes;
                const int k1 = min(A.kbytes, k0 + A.chunk_bytes);
                for (int ci = c0; ci < c1; ++ci, ++acc_it) {
                    const uint32_t acc = acc_it % ACC, acc_phase = (acc_it / ACC) & 1u;
                    mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                    tc_fence_after();
                    const uint32_t d_re = tmem_base + acc * (2u * MN), d_im = d_re + MN;
                    uint32_t accumulate = 0;
                    for (int p = 0; p < S.n_pairs[ci]; ++p)
                        for (int k = k0; k < k1; k += kStageK) {
                            mbar_wait(full_bar(stage), phase);
                            tc_fence_after();
                            const uint32_t sa = smem_base + stage * kStageBytes;
                            const uint64_t dan = umma_smem_desc(sa);
                            const uint64_t dar = umma_smem_desc(sa + kTileBytes);
                            const uint64_t db = umma_smem_desc(sa + 2 * kTileBytes);
#pragma unroll
                            for (int kk = 0; kk < kStageK / 32; ++kk) {
                                // +32 bytes along K inside the swizzle atom = +2 in the (addr >> 4) field
                                umma_i8<CG>(d_re, dan + 2u * kk, db + 2u * kk, kIdesc, accumulate | (uint32_t)kk);
                                umma_i8<CG>(d_im, dar + 2u * kk, db + 2u * kk, kIdesc, accumulate | (uint32_t)kk);
                            }
                            accumulate = 1;
                            umma_commit<CG>(empty_bar(stage));  // frees the stage when these MMAs have read it
                            if (++stage == kStages) { stage = 0; phase ^= 1u; }
                        }
                    umma_commit<CG>(tfull_bar(acc));  // accumulator of this class complete
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> FP64 accumulators in global memory (every CTA drains its 128 rows) =====
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int row = (int)cta_rank * kTile + q * 32 + lane;
        uint32_t acc_it = 0;
        for (int u = worker; u < n_units; u += n_workers) {
            int tile, chunk, c0, c1;
            const int ji = decode_unit(A, u, tile, chunk, c0, c1);
            double* wt = A.W + (size_t)(A.jobs[ji].tile0 + tile) * (2 * MN * MN) + row;
            int* wt32 = reinterpret_cast<int*>(A.W) + (size_t)(A.jobs[ji].tile0 + tile) * S.n_classes * (2 * MN * MN) + row;
            for (int ci = c0; ci < c1; ++ci, ++acc_it) {
                const uint32_t acc = acc_it % ACC, acc_phase = (acc_it / ACC) & 1u;
                const double w = __hiloint2double((1023 - 8 * S.cls_shift[ci]) << 20, 0);  // 2^(-8 c)
                const int pm = c_crt_moduli[ci];
                const double inv_pm = c_crt_inv[ci];
                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (2u * MN);
#pragma unroll 1
                for (int cb = 0; cb < 2 * MN / 32; ++cb) {  // blocks of 32 columns: Re [0, MN), Im [MN, 2 MN)
                    uint32_t v[32];
                    tmem_ld32(t0 + cb * 32, v);
                    tmem_ld_wait();
                    if (A.crt && A.n_chunks == 1) {
                        // the unit covers the whole K (the usual case: 128 KiB of K stay exact in int32): the inner
                        // products modulo p_m, as symmetric residues, are STORED as int8: W8[tile][modulus][part][row][col]
                        uint32_t pk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const int iv = (int)v[c];
                            const int r = iv - pm * __double2int_rn((double)iv * inv_pm);   // (p = 256: +128 and -128 are the same byte)
                            pk[c >> 2] |= ((uint32_t)r & 255u) << (8 * (c & 3));
                        }
                        const int col0 = cb * 32;   // Re: [0, MN), Im: [MN, 2 MN)
                        int8_t* w8 = reinterpret_cast<int8_t*>(A.W) +
                                     ((size_t)(A.jobs[ji].tile0 + tile) * S.n_classes + ci) * (2 * MN * MN) +
                                     (size_t)(col0 / MN) * (MN * MN) + (size_t)row * MN + (col0 % MN);
                        reinterpret_cast<uint4*>(w8)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        reinterpret_cast<uint4*>(w8)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    } else if (A.crt) {
                        // K cut into chunks: the residues of the chunks are summed (int32 W[tile][modulus][part][col][row])
                        int* wp = wt32 + (size_t)ci * (2 * MN * MN) + (size_t)(cb * 32) * MN;
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const int iv = (int)v[c];
                            int r = iv - pm * __double2int_rn((double)iv * inv_pm);
                            if (r == 128) r = -128;   // (p = 256 only)
                            if (r != 0) atomicAdd(wp + (size_t)c * MN, r);
                        }
                    } else {
                        double* wp = wt + (size_t)(cb * 32) * MN;  // [part][col][row]
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const int iv = (int)v[c];
                            if (iv != 0) red_add_f64(wp + (size_t)c * MN, (double)iv * w);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG == 1) mbar_arrive(tempty_bar(acc));
                    else mbar_arrive_cta(tempty_bar(acc), 0);
                }
            }
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if (CG == 1)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// Digit planes.  planes: [2 * kSlices][n_rows][2 * dim] int8; plane s = digits of nat = (re, im), plane
// kSlices + s = digits of rot = (-im, re).  One CTA per state row.  `row_stats` (optional): high words of the row
// maxima max(|re|, |im|) already reduced by the producer of the states (the last gate pass, see TileArgs::row_stats),
// which saves this kernel's first sweep over the state (the high word has the exponent: all the scale needs).

__device__ __forceinline__ uint32_t pack_digit(long long& q, uint32_t word, int byte) {
    const long long d = ((q + 128) & 255) - 128;
    q = (q - d) >> 8;
    return word | (((uint32_t)d & 255u) << (8 * byte));
}

__global__ void __launch_bounds__(256) k_slice_i8(const double2* __restrict__ psi, int64_t plane_rows, int64_t row0,
                                                  int64_t dim, int8_t* __restrict__ planes, int* __restrict__ expo,
                                                  const RowStats* __restrict__ row_stats) {
    const int64_t row = row0 + blockIdx.x;   // row inside the planes; state blockIdx.x of psi
    const double2* src = psi + (int64_t)blockIdx.x * dim;
    __shared__ double red[8];
    __shared__ int s_e;
    double m = 0.0;
    const uint32_t max_hi = row_stats ? row_stats[blockIdx.x].max_hi : 0u;   // 0: no statistics for this row
    if (max_hi != 0u) {
        m = __hiloint2double((int)max_hi, 0);
        if (m == 0.0) m = 4.9e-324;   // (cannot happen for a normalised state)
    } else {
        for (int64_t i = threadIdx.x; i < dim; i += blockDim.x) {
            const double2 v = src[i];
            m = fmax(m, fmax(fabs(v.x), fabs(v.y)));
        }
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (max_hi == 0u)
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
        int e = 0;
        if (m > 0.0) frexp(m, &e);  // m = f * 2^e, f in [0.5, 1): every |v| < 2^e
        s_e = e;
        expo[row] = e;
    }
    __syncthreads();
    const double scale = ldexp(1.0, kTopBits - s_e);
    const size_t plane_stride = (size_t)plane_rows * 2 * dim;
    int8_t* dst_row = planes + (size_t)row * 2 * dim;
    // 8 complex amplitudes per thread and iteration -> 16 bytes per plane
    for (int64_t i0 = (int64_t)threadIdx.x * 8; i0 < dim; i0 += (int64_t)blockDim.x * 8) {
        long long qn[16], qr[16];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const double2 v = __ldcs(src + i0 + c);
            const long long a = __double2ll_rn(v.x * scale), b = __double2ll_rn(v.y * scale);
            qn[2 * c] = a;
            qn[2 * c + 1] = b;
            qr[2 * c] = -b;
            qr[2 * c + 1] = a;
        }
#pragma unroll
        for (int s = kSlices - 1; s >= 0; --s) {
            uint32_t wn[4] = {0, 0, 0, 0}, wr[4] = {0, 0, 0, 0};
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                wn[b >> 2] = pack_digit(qn[b], wn[b >> 2], b & 3);
                wr[b >> 2] = pack_digit(qr[b], wr[b >> 2], b & 3);
            }
            *reinterpret_cast<uint4*>(dst_row + (size_t)s * plane_stride + 2 * i0) = make_uint4(wn[0], wn[1], wn[2], wn[3]);
            *reinterpret_cast<uint4*>(dst_row + (size_t)(kSlices + s) * plane_stride + 2 * i0) =
                make_uint4(wr[0], wr[1], wr[2], wr[3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// W -> K.  One CTA per tile; 32 x 32 sub-blocks go through shared memory so that both the direct and the
// mirrored store are coalesced.  Jobs with `align_y` also reduce sum K_ij y_i y_j and sum K_ij^2 here (the
// kernel-target-alignment sums of src/squlearn/kernel/loss/target_alignment.py:75-90), so that a training step
// never re-reads K.

__global__ void __launch_bounds__(256) k_gram_i8_final(const double* __restrict__ W, const __grid_constant__ I8Args A) {
    __shared__ double sm[32][33];
    __shared__ double s_red[2][8];
    int ji = 0;
    while (ji + 1 < A.n_jobs && (int)blockIdx.x >= A.jobs[ji].tile0 + A.jobs[ji].n_tiles) ++ji;
    const I8Job& J = A.jobs[ji];
    int bm, bn;
    decode_tile(J, (int)blockIdx.x - J.tile0, bm, bn);
    const int MN = A.mn, nsb = MN / 32;
    const double* wt = W + (size_t)blockIdx.x * (2 * (size_t)MN * MN);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const uint32_t flags = J.flags;
    const bool sym = flags & B200SQ_GRAM_SYMMETRIC;
    const bool diag_tile = sym && bm == bn;
    const int64_t nx = J.nx, ny = J.ny, ldk = J.ldk;
    double* __restrict__ K = J.K;
    double* __restrict__ Km = sym ? nullptr : J.K_mirror;
    const int* __restrict__ ex = J.ex + J.row0_a;
    const int* __restrict__ ey = J.ey + J.row0_b;
    const double* __restrict__ ayx = J.align_yx;
    const double* __restrict__ ayy = J.align_yy;
    double a_ky = 0.0, a_kk = 0.0;
    for (int sb = 0; sb < nsb * nsb; ++sb) {
        const int sr = sb / nsb, sc = sb - sr * nsb;
        if ((int64_t)bm * MN + sr * 32 >= nx || (int64_t)bn * MN + sc * 32 >= ny) continue;  // CTA-uniform
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            const int col = sc * 32 + ty + 8 * qd, row = sr * 32 + tx;
            const int64_t i = (int64_t)bm * MN + row, j = (int64_t)bn * MN + col;
            double k = 0.0;
            if (i < nx && j < ny) {
                // diagonal tiles of a symmetric Gram: take both triangles from the lower one, so that K is
                // bit-exactly symmetric whatever order the atomic FP64 accumulation happened in
                const int r2 = (diag_tile && col > row) ? col : row, c2 = (diag_tile && col > row) ? row : col;
                const double re = wt[(size_t)c2 * MN + r2], im = wt[(size_t)(MN + c2) * MN + r2];
                k = ldexp(re * re + im * im, 2 * (ex[i] + ey[j] - 12));
                if (sym && i == j) {
                    if (flags & B200SQ_GRAM_DIAG_ONE) k = 1.0;
                    else if (flags & B200SQ_GRAM_DIAG_ZERO) k = 0.0;
                }
                if (ayx) {   // off-diagonal tiles of a symmetric block stand for their mirror image too
                    const double wgt = sym ? (bm != bn ? 2.0 : 1.0) : J.align_weight;
                    a_ky = fma(wgt * k, ayx[i] * ayy[j], a_ky);
                    a_kk = fma(wgt * k, k, a_kk);
                }
            }
            sm[ty + 8 * qd][tx] = k;  // [col][row]
        }
        __syncthreads();
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            {   // direct: K[i][j], consecutive threads -> consecutive j
                const int row = sr * 32 + ty + 8 * qd, col = sc * 32 + tx;
                const int64_t i = (int64_t)bm * MN + row, j = (int64_t)bn * MN + col;
                if (i < nx && j < ny) K[i * ldk + j] = sm[tx][ty + 8 * qd];
            }
            if (sym && bm != bn) {  // mirror: K[j][i], consecutive threads -> consecutive i
                const int col = sc * 32 + ty + 8 * qd, row = sr * 32 + tx;
                const int64_t i = (int64_t)bm * MN + row, j = (int64_t)bn * MN + col;
                if (i < nx && j < ny) K[j * ldk + i] = sm[ty + 8 * qd][tx];
            } else if (Km) {        // transpose of a rectangular block for the owner of its columns
                const int col = sc * 32 + ty + 8 * qd, row = sr * 32 + tx;
                const int64_t i = (int64_t)bm * MN + row, j = (int64_t)bn * MN + col;
                if (i < nx && j < ny) Km[j * J.ldm + i] = sm[ty + 8 * qd][tx];
            }
        }
        __syncthreads();
    }
    if (Km) __threadfence_system();   // peer stores are performed before the kernel (and the flag behind it) completes
    if (ayx) {
        for (int o = 16; o > 0; o >>= 1) {
            a_ky += __shfl_xor_sync(0xffffffffu, a_ky, o);
            a_kk += __shfl_xor_sync(0xffffffffu, a_kk, o);
        }
        if (tx == 0) { s_red[0][ty] = a_ky; s_red[1][ty] = a_kk; }
        __syncthreads();
        if (threadIdx.x < 2) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += s_red[threadIdx.x][w];
            atomicAdd(J.align_out + threadIdx.x, t);
        }
    }
}

// Symmetric residue of an exact integer q held in a double, modulo p, WITHOUT conversion instructions (FRND /
// F2I on 64-bit types run at a quarter of the FP64 rate): with M = 1.5 * 2^52, fma(q, 1/p, M) - M is q / p rounded
// to the nearest integer k (odd p: q / p is never within 1 / (2 p) of a half-integer and the product's error is
// < 2^-8; p = 256: the product is exact), and fma(-p, k, q + M) = M + r exactly, whose low 32 bits are r in two's
// complement.  qm = q + M.
constexpr double kRoundMagic = 6755399441055744.0;   // 1.5 * 2^52
__device__ __forceinline__ int crt_residue(double q, double qm, double pm, double inv) {
    const double k = fma(q, inv, kRoundMagic) - kRoundMagic;
    return __double2loint(fma(-pm, k, qm));
}
// round to nearest (ties to even) without FRND: |x| < 2^51
__device__ __forceinline__ double magic_rint(double x) { return (x + kRoundMagic) - kRoundMagic; }

// ---------------------------------------------------------------------------------------------
// CRT scheme: residue planes.  planes: [2 s][n_rows][2 * dim] int8; plane m = nat(q) mod p_m (symmetric residues),
// plane s + m = rot(q) mod p_m with rot = (-im, re), q = round(v * 2^(B - e)), e = the smallest integer with
// 2^(B - e) |row|_2 <= R (CrtPlan).  The residue of an exact integer q held in a double is q - p * rint(q / p): rint
// is exact because q / p is never within 1 / (2 p) of a half-integer for odd p while the rounding error of q * (1 / p)
// is below 2^-8; for p = 256 the product is exact and a tie leaves r = +-128, both the byte 0x80.

struct CrtPlan {
    int n_mod;       // moduli in use
    int bits;        // B
    double e_bias;   // B - log2(R_eff): e = ceil(log2 |row|_2 + e_bias)
};

// `digits` (optional): also write the 6 nat DIGIT planes of the rows ([6][digit_rows][2 dim], what the multi-GPU
// exchange sends) from the same read of the states: the balanced base-256 digits of the SAME integers q, so a
// receiver's residues (k_digits_to_crt) are bit-identical to the sender's and the exchange costs no precision.
__global__ void __launch_bounds__(256) k_slice_crt(const double2* __restrict__ psi, int64_t plane_rows, int64_t row0,
                                                   int64_t dim, const CrtPlan plan, int8_t* __restrict__ planes,
                                                   int* __restrict__ expo, const RowStats* __restrict__ row_stats,
                                                   int8_t* __restrict__ digits, int64_t digit_rows, int64_t digit_row0) {
    const int64_t row = row0 + blockIdx.x;
    const double2* src = psi + (int64_t)blockIdx.x * dim;
    const int n_mod = plan.n_mod;
    __shared__ double red[8];
    __shared__ int s_e;
    double n2 = row_stats ? row_stats[blockIdx.x].norm2 : 0.0;   // 0: no statistics for this row
    const bool have = n2 > 0.0;
    if (!have) {
        double a0 = 0.0, a1 = 0.0;
        for (int64_t i = threadIdx.x; i < dim; i += blockDim.x) {
            const double2 v = src[i];
            a0 = fma(v.x, v.x, a0);
            a1 = fma(v.y, v.y, a1);
        }
        n2 = a0 + a1;
        for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = n2;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (!have)
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) n2 += red[w];
        // (1e-9: the summation error of n2 and of log2 can never push a row over R)
        int e = 0;
        if (n2 > 0.0) e = (int)ceil(0.5 * log2(n2) + plan.e_bias + 1e-9);
        s_e = e;
        expo[row] = e;
    }
    __syncthreads();
    const double scale = ldexp(1.0, plan.bits - s_e);
    const size_t plane_stride = (size_t)plane_rows * 2 * dim;
    int8_t* dst_row = planes + (size_t)row * 2 * dim;
    const size_t digit_stride = (size_t)digit_rows * 2 * dim;
    int8_t* dig_row = digits ? digits + (size_t)(digit_row0 + blockIdx.x) * 2 * dim : nullptr;
    for (int64_t i0 = (int64_t)threadIdx.x * 8; i0 < dim; i0 += (int64_t)blockDim.x * 8) {
        double q[16];   // (re, im) of 8 amplitudes as exact integers
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const double2 v = __ldcs(src + i0 + c);
            q[2 * c] = magic_rint(v.x * scale);
            q[2 * c + 1] = magic_rint(v.y * scale);
        }
        if (digits) {
            long long qd[16];
#pragma unroll
            for (int b = 0; b < 16; ++b) qd[b] = (long long)q[b];
#pragma unroll
            for (int sl = kSlices - 1; sl >= 0; --sl) {
                uint32_t wd[4] = {0, 0, 0, 0};
#pragma unroll
                for (int b = 0; b < 16; ++b) wd[b >> 2] = pack_digit(qd[b], wd[b >> 2], b & 3);
                *reinterpret_cast<uint4*>(dig_row + (size_t)sl * digit_stride + 2 * i0) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
            }
        }
        double qm[16];
#pragma unroll
        for (int b = 0; b < 16; ++b) qm[b] = q[b] + kRoundMagic;
        for (int mi = 0; mi < n_mod; ++mi) {
            const double pm = (double)c_crt_moduli[mi], inv = c_crt_inv[mi];
            uint32_t wn[4] = {0, 0, 0, 0}, wr[4] = {0, 0, 0, 0};
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int rr = crt_residue(q[2 * c], qm[2 * c], pm, inv);
                const int ri = crt_residue(q[2 * c + 1], qm[2 * c + 1], pm, inv);
                const int b = 2 * c;   // byte positions b (re), b + 1 (im) of the 16-byte group
                wn[b >> 2] |= ((uint32_t)rr & 255u) << (8 * (b & 3));
                wn[(b + 1) >> 2] |= ((uint32_t)ri & 255u) << (8 * ((b + 1) & 3));
                wr[b >> 2] |= ((uint32_t)(-ri) & 255u) << (8 * (b & 3));          // rot = (-im, re)
                wr[(b + 1) >> 2] |= ((uint32_t)rr & 255u) << (8 * ((b + 1) & 3));
            }
            *reinterpret_cast<uint4*>(dst_row + (size_t)mi * plane_stride + 2 * i0) = make_uint4(wn[0], wn[1], wn[2], wn[3]);
            *reinterpret_cast<uint4*>(dst_row + (size_t)(n_mod + mi) * plane_stride + 2 * i0) =
                make_uint4(wr[0], wr[1], wr[2], wr[3]);
        }
    }
}

// Digit planes -> residue planes, for column blocks that arrive as digit planes (the multi-GPU exchange sends the
// 6 nat digit planes: 6 instead of s bytes per component; the receiver derives the residues).  The digits are those
// of the CRT integer q = round(v 2^(B - e)) itself (k_slice_crt), so the residues are the sender's.  nat planes only.
// The kernel runs NEXT TO the persistent Gram kernel, which keeps ~30 K registers and 198 KB of shared memory per SM:
// at most 64 registers per thread let two CTAs fit beside it, and every thread fetches kConvCols independent words
// of each digit plane (24 loads in flight) before it starts to compute, because with so few resident warps the
// kernel is bound by the number of bytes in flight, not by arithmetic (measured on 8 GPUs: the conversions, not the
// NVLink pushes, were the critical path of the exchange).
constexpr int kConvCols = 4;
__global__ void __launch_bounds__(256, 4) k_digits_to_crt(const int8_t* __restrict__ digits, int64_t rows_d, int64_t row0_d,
                                                          int64_t kbytes, int n_mod, int8_t* __restrict__ out,
                                                          int64_t rows_o, int64_t row0_o) {
    const size_t stride_d = (size_t)rows_d * kbytes, stride_o = (size_t)rows_o * kbytes;
    const int8_t* src = digits + (size_t)(row0_d + blockIdx.y) * kbytes;
    int8_t* dst = out + (size_t)(row0_o + blockIdx.y) * kbytes;
    const int64_t kb0 = ((int64_t)blockIdx.x * kConvCols * blockDim.x + threadIdx.x) * 4;   // word u: kb0 + u * 4 * blockDim.x
    uint32_t wd[kConvCols][kSlices];
#pragma unroll
    for (int u = 0; u < kConvCols; ++u) {
        const int64_t k0 = kb0 + (int64_t)u * 4 * blockDim.x;
#pragma unroll
        for (int sl = 0; sl < kSlices; ++sl)
            wd[u][sl] = k0 < kbytes ? __ldcs(reinterpret_cast<const uint32_t*>(src + (size_t)sl * stride_d + k0)) : 0u;
    }
#pragma unroll
    for (int u = 0; u < kConvCols; ++u) {
        const int64_t k0 = kb0 + (int64_t)u * 4 * blockDim.x;
        if (k0 >= kbytes) break;
        // the integer from two 24-bit halves (int32 arithmetic; exact in a double)
        double q[4], qm[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int hi = 0, lo = 0;
#pragma unroll
            for (int sl = 0; sl < kSlices; ++sl) {   // most significant digit first
                const int d = (int)(int8_t)((wd[u][sl] >> (8 * b)) & 255u);
                if (sl < 3) hi = hi * 256 + d;
                else lo = lo * 256 + d;
            }
            q[b] = fma((double)hi, 16777216.0, (double)lo);
            qm[b] = q[b] + kRoundMagic;
        }
        for (int mi = 0; mi < n_mod; ++mi) {
            const double pm = (double)c_crt_moduli[mi], inv = c_crt_inv[mi];
            uint32_t wn = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int r = crt_residue(q[b], qm[b], pm, inv);
                wn |= ((uint32_t)r & 255u) << (8 * b);
            }
            *reinterpret_cast<uint32_t*>(dst + (size_t)mi * stride_o + k0) = wn;
        }
    }
}

// CRT reconstruction: residues r_m of an integer x with |x| < P / 2  ->  x / P = frac(sum_m k_m / p_m) mapped to
// (-1/2, 1/2], k_m = r_m c_m mod p_m, c_m = (P / p_m)^-1 mod p_m.  Every k_m / p_m is split into a double-double
// (hi + lo) and the sum is accumulated error-free, so the result keeps full double precision although it is
// ~2^-13 of the terms.  The (hi, lo) pairs come from a shared-memory table indexed by (modulus, residue + 128),
// built once per CTA: one table look-up and one error-free addition per modulus and entry.
constexpr int kCrtTabStride = 256;   // residues -128 .. 127 per modulus

__device__ __forceinline__ double2 crt_term(int r, int pm, int c) {
    int k = (r % pm) * c % pm;
    if (k < 0) k += pm;
    const double pd = (double)pm, inv = 1.0 / pd, kd = (double)k;
    const double hi = kd * inv;
    return double2{hi, fma(-hi, pd, kd) * inv};
}

__device__ __forceinline__ double crt_fraction(const int* __restrict__ w, size_t mod_stride, int n_mod,
                                               const int* __restrict__ crt_c, const double2* __restrict__ tab) {
    double sh = 0.0, sl = 0.0;
#pragma unroll 4
    for (int mi = 0; mi < n_mod; ++mi) {
        const int r = w[(size_t)mi * mod_stride];
        // a single K chunk stores symmetric residues (-128 <= r <= 127): table; sums over chunks: reduce first
        const double2 t2 = ((unsigned)(r + 128) < 256u) ? tab[mi * kCrtTabStride + r + 128]
                                                        : crt_term(r, c_crt_moduli[mi], crt_c[mi]);
        const double t = sh + t2.x;
        const double bb = t - sh;
        sl += ((sh - (t - bb)) + (t2.x - bb)) + t2.y;
        sh = t;
    }
    double f = (sh - rint(sh)) + sl;
    if (f > 0.5) f -= 1.0;
    if (f < -0.5) f += 1.0;
    return f;
}

// Final kernel of the usual case (one K chunk per unit): W8[tile][modulus][part][row][col] holds int8 residues.  A
// thread reconstructs FOUR consecutive columns of one row: one 32-bit load per (modulus, part), all 2 s loads issued
// before the first use (the int32 form below keeps one load in flight per thread and is latency-bound).  The integer
// arithmetic is exact, so a diagonal tile is symmetric as computed.
__global__ void __launch_bounds__(256, 3) k_gram_crt_final8(const int8_t* __restrict__ W8, const __grid_constant__ I8Args A) {
    __shared__ double sm[32][33];   // [row][col]
    __shared__ double s_red[2][8];
    extern __shared__ double2 s_tab[];   // [n_mod][kCrtTabStride]
    const int n_mod = A.sched.n_classes;
    for (int t = threadIdx.x; t < n_mod * kCrtTabStride; t += blockDim.x) {
        const int mi = t / kCrtTabStride, r = t - mi * kCrtTabStride - 128;
        s_tab[t] = crt_term(r, c_crt_moduli[mi], A.crt_c[mi]);
    }
    __syncthreads();
    int ji = 0;
    while (ji + 1 < A.n_jobs && (int)blockIdx.x >= A.jobs[ji].tile0 + A.jobs[ji].n_tiles) ++ji;
    const I8Job& J = A.jobs[ji];
    int bm, bn;
    decode_tile(J, (int)blockIdx.x - J.tile0, bm, bn);
    const int MN = A.mn, nsb = MN / 32;
    const size_t part = (size_t)MN * MN, mod_stride = 2 * part;
    const int8_t* wt = W8 + (size_t)blockIdx.x * n_mod * mod_stride;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int lrow = threadIdx.x >> 3, lcol = (threadIdx.x & 7) * 4;   // compute phase: 4 columns of one row
    const uint32_t flags = J.flags;
    const bool sym = flags & B200SQ_GRAM_SYMMETRIC;
    const int64_t nx = J.nx, ny = J.ny, ldk = J.ldk;
    double* __restrict__ K = J.K;
    double* __restrict__ Km = sym ? nullptr : J.K_mirror;
    const int* __restrict__ ex = J.ex + J.row0_a;
    const int* __restrict__ ey = J.ey + J.row0_b;
    const double* __restrict__ ayx = J.align_yx;
    const double* __restrict__ ayy = J.align_yy;
    double a_ky = 0.0, a_kk = 0.0;
    for (int sb = blockIdx.y; sb < nsb * nsb; sb += gridDim.y) {   // a tile is shared by gridDim.y CTAs
        const int sr = sb / nsb, sc = sb - sr * nsb;
        if ((int64_t)bm * MN + sr * 32 >= nx || (int64_t)bn * MN + sc * 32 >= ny) continue;  // CTA-uniform
        {
            const int row = sr * 32 + lrow, col = sc * 32 + lcol;
            const int64_t i = (int64_t)bm * MN + row;
            const uint32_t* w0 = reinterpret_cast<const uint32_t*>(wt + (size_t)row * MN + col);
            uint32_t wre[kCrtMaxModuli], wim[kCrtMaxModuli];
#pragma unroll
            for (int mi = 0; mi < kCrtMaxModuli; ++mi)
                if (mi < n_mod) {
                    wre[mi] = __ldcs(w0 + (size_t)mi * (mod_stride / 4));
                    wim[mi] = __ldcs(w0 + (size_t)mi * (mod_stride / 4) + part / 4);
                }
            const int e_i = i < nx ? ex[i] : 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int64_t j = (int64_t)bn * MN + col + c;
                double k = 0.0;
                if (i < nx && j < ny) {
                    double rh = 0.0, rl = 0.0, ih = 0.0, il = 0.0;
#pragma unroll
                    for (int mi = 0; mi < kCrtMaxModuli; ++mi)
                        if (mi < n_mod) {
                            const double2 tr = s_tab[mi * kCrtTabStride + (int)(int8_t)(wre[mi] >> (8 * c)) + 128];
                            const double2 ti = s_tab[mi * kCrtTabStride + (int)(int8_t)(wim[mi] >> (8 * c)) + 128];
                            double t = rh + tr.x, bb = t - rh;
                            rl += ((rh - (t - bb)) + (tr.x - bb)) + tr.y;
                            rh = t;
                            t = ih + ti.x;
                            bb = t - ih;
                            il += ((ih - (t - bb)) + (ti.x - bb)) + ti.y;
                            ih = t;
                        }
                    double fr = (rh - rint(rh)) + rl, fi = (ih - rint(ih)) + il;
                    if (fr > 0.5) fr -= 1.0;
                    if (fr < -0.5) fr += 1.0;
                    if (fi > 0.5) fi -= 1.0;
                    if (fi < -0.5) fi += 1.0;
                    const double re = fr * A.crt_scale, im = fi * A.crt_scale;
                    k = ldexp(re * re + im * im, 2 * (e_i + ey[j]));
                    if (sym && bm == bn && i == j) {
                        if (flags & B200SQ_GRAM_DIAG_ONE) k = 1.0;
                        else if (flags & B200SQ_GRAM_DIAG_ZERO) k = 0.0;
                    }
                    if (ayx) {
                        const double wgt = sym ? (bm != bn ? 2.0 : 1.0) : J.align_weight;
                        a_ky = fma(wgt * k, ayx[i] * ayy[j], a_ky);
                        a_kk = fma(wgt * k, k, a_kk);
                    }
                }
                sm[lrow][lcol + c] = k;
            }
        }
        __syncthreads();
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            {   // direct: K[i][j], consecutive threads -> consecutive j
                const int row = sr * 32 + ty + 8 * qd, col = sc * 32 + tx;
                const int64_t i = (int64_t)bm * MN + row, j = (int64_t)bn * MN + col;
                if (i < nx && j < ny) K[i * ldk + j] = sm[ty + 8 * qd][tx];
            }
            if (sym && bm != bn) {  // mirror: K[j][i], consecutive threads -> consecutive i
                const int col = sc * 32 + ty + 8 * qd, row = sr * 32 + tx;
                const int64_t i = (int64_t)bm * MN + row, j = (int64_t)bn * MN + col;
                if (i < nx && j < ny) K[j * ldk + i] = sm[tx][ty + 8 * qd];
            } else if (Km) {        // transpose of a rectangular block for the owner of its columns
                const int col = sc * 32 + ty + 8 * qd, row = sr * 32 + tx;
                const int64_t i = (int64_t)bm * MN + row, j = (int64_t)bn * MN + col;
                if (i < nx && j < ny) Km[j * J.ldm + i] = sm[tx][ty + 8 * qd];
            }
        }
        __syncthreads();
    }
    if (Km) __threadfence_system();
    if (ayx) {
        for (int o = 16; o > 0; o >>= 1) {
            a_ky += __shfl_xor_sync(0xffffffffu, a_ky, o);
            a_kk += __shfl_xor_sync(0xffffffffu, a_kk, o);
        }
        if (tx == 0) { s_red[0][ty] = a_ky; s_red[1][ty] = a_kk; }
        __syncthreads();
        if (threadIdx.x < 2) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += s_red[threadIdx.x][w];
            atomicAdd(J.align_out + threadIdx.x, t);
        }
    }
}

// Final kernel when K is cut into chunks: W[tile][modulus][part][col][row] holds int32 sums of chunk residues.
__global__ void __launch_bounds__(256) k_gram_crt_final(const int* __restrict__ W, const __grid_constant__ I8Args A) {
    __shared__ double sm[32][33];
    __shared__ double s_red[2][8];
    __shared__ int s_c[kCrtMaxModuli];
    extern __shared__ double2 s_tab[];   // [n_mod][kCrtTabStride]
    if (threadIdx.x < kCrtMaxModuli) s_c[threadIdx.x] = A.crt_c[threadIdx.x];
    for (int t = threadIdx.x; t < A.sched.n_classes * kCrtTabStride; t += blockDim.x) {
        const int mi = t / kCrtTabStride, r = t - mi * kCrtTabStride - 128;
        s_tab[t] = crt_term(r, c_crt_moduli[mi], A.crt_c[mi]);
    }
    __syncthreads();
    int ji = 0;
    while (ji + 1 < A.n_jobs && (int)blockIdx.x >= A.jobs[ji].tile0 + A.jobs[ji].n_tiles) ++ji;
    const I8Job& J = A.jobs[ji];
    int bm, bn;
    decode_tile(J, (int)blockIdx.x - J.tile0, bm, bn);
    const int MN = A.mn, nsb = MN / 32, n_mod = A.sched.n_classes;
    const size_t part = (size_t)MN * MN, mod_stride = 2 * part;
    const int* wt = W + (size_t)blockIdx.x * n_mod * mod_stride;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const uint32_t flags = J.flags;
    const bool sym = flags & B200SQ_GRAM_SYMMETRIC;
    const bool diag_tile = sym && bm == bn;
    const int64_t nx = J.nx, ny = J.ny, ldk = J.ldk;
    double* __restrict__ K = J.K;
    double* __restrict__ Km = sym ? nullptr : J.K_mirror;
    const int* __restrict__ ex = J.ex + J.row0_a;
    const int* __restrict__ ey = J.ey + J.row0_b;
    const double* __restrict__ ayx = J.align_yx;
    const double* __restrict__ ayy = J.align_yy;
    double a_ky = 0.0, a_kk = 0.0;
    for (int sb = blockIdx.y; sb < nsb * nsb; sb += gridDim.y) {   // a tile is shared by gridDim.y CTAs
        const int sr = sb / nsb, sc = sb - sr * nsb;
        if ((int64_t)bm * MN + sr * 32 >= nx || (int64_t)bn * MN + sc * 32 >= ny) continue;  // CTA-uniform
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            const int col = sc * 32 + ty + 8 * qd, row = sr * 32 + tx;
            const int64_t i = (int64_t)bm * MN + row, j = (int64_t)bn * MN + col;
            double k = 0.0;
            if (i < nx && j < ny) {
                const int r2 = (diag_tile && col > row) ? col : row, c2 = (diag_tile && col > row) ? row : col;
                const int* w0 = wt + (size_t)c2 * MN + r2;
                const double re = crt_fraction(w0, mod_stride, n_mod, s_c, s_tab) * A.crt_scale;
                const double im = crt_fraction(w0 + part, mod_stride, n_mod, s_c, s_tab) * A.crt_scale;
                k = ldexp(re * re + im * im, 2 * (ex[i] + ey[j]));
                if (sym && i == j) {
                    if (flags & B200SQ_GRAM_DIAG_ONE) k = 1.0;
                    else if (flags & B200SQ_GRAM_DIAG_ZERO) k = 0.0;
                }
                if (ayx) {
                    const double wgt = sym ? (bm != bn ? 2.0 : 1.0) : J.align_weight;
                    a_ky = fma(wgt * k, ayx[i] * ayy[j], a_ky);
                    a_kk = fma(wgt * k, k, a_kk);
                }
            }
            sm[ty + 8 * qd][tx] = k;  // [col][row]
        }
        __syncthreads();
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            {
                const int row = sr * 32 + ty + 8 * qd, col = sc * 32 + tx;
                const int64_t i = (int64_t)bm * MN + row, j = (int64_t)bn * MN + col;
                if (i < nx && j < ny) K[i * ldk + j] = sm[tx][ty + 8 * qd];
            }
            if (sym && bm != bn) {
                const int col = sc * 32 + ty + 8 * qd, row = sr * 32 + tx;
                const int64_t i = (int64_t)bm * MN + row, j = (int64_t)bn * MN + col;
                if (i < nx && j < ny) K[j * ldk + i] = sm[ty + 8 * qd][tx];
            } else if (Km) {
                const int col = sc * 32 + ty + 8 * qd, row = sr * 32 + tx;
                const int64_t i = (int64_t)bm * MN + row, j = (int64_t)bn * MN + col;
                if (i < nx && j < ny) Km[j * J.ldm + i] = sm[ty + 8 * qd][tx];
            }
        }
        __syncthreads();
    }
    if (Km) __threadfence_system();
    if (ayx) {
        for (int o = 16; o > 0; o >>= 1) {
            a_ky += __shfl_xor_sync(0xffffffffu, a_ky, o);
            a_kk += __shfl_xor_sync(0xffffffffu, a_kk, o);
        }
        if (tx == 0) { s_red[0][ty] = a_ky; s_red[1][ty] = a_kk; }
        __syncthreads();
        if (threadIdx.x < 2) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += s_red[threadIdx.x][w];
            atomicAdd(J.align_out + threadIdx.x, t);
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_planes(CUtensorMap* map, void* base, uint64_t kbytes, uint64_t rows_total) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) return (int)cudaErrorNotSupported;
        fn = (EncodeTiledFn)p;
    }
    const cuuint64_t dims[2] = {kbytes, rows_total};
    const cuuint64_t strides[1] = {kbytes};
    const cuuint32_t box[2] = {(cuuint32_t)kStageK, (cuuint32_t)kTile};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

}  // namespace

struct WaitArgs {
    int32_t which[32];
    int n;
    uint32_t value;
};

__global__ void k_wait_flags(const uint32_t* __restrict__ flags, const WaitArgs w) {
    if ((int)threadIdx.x < w.n) {
        const uint32_t* f = flags + w.which[threadIdx.x];
        const unsigned long long t_start = global_timer_ns();
        while (ld_acquire_sys(f) < w.value) {
            __nanosleep(500);
            if (global_timer_ns() - t_start > kReadyTimeoutNs) {
                atomicAdd(&g_ready_timeouts, 1u);
                break;
            }
        }
    }
}

int launch_wait_flags(const uint32_t* flags, const int32_t* which, int n, uint32_t value, cudaStream_t stream) {
    if (n <= 0) return 0;
    if (n > 32) return (int)cudaErrorInvalidValue;
    WaitArgs w;
    memset(&w, 0, sizeof(w));
    for (int k = 0; k < n; ++k) w.which[k] = which[k];
    w.n = n;
    w.value = value;
    k_wait_flags<<<1, 32, 0, stream>>>(flags, w);
    count_launch();
    return (int)cudaGetLastError();
}

unsigned int gram_ready_timeouts() {
    unsigned int v = 0;
    if (cudaMemcpyFromSymbol(&v, g_ready_timeouts, sizeof(v)) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return v;
}

bool gram_i8_supported(int64_t dim) { return dim >= 64 && dim % 8 == 0 && 2 * dim <= (1ll << 30); }

int launch_slice_i8(const double2* psi, int64_t n, int64_t dim, int8_t* planes, int64_t plane_rows, int64_t row0,
                    int* expo, cudaStream_t stream, const RowStats* row_stats) {
    if (n <= 0) return 0;
    k_slice_i8<<<(unsigned)n, 256, 0, stream>>>(psi, plane_rows, row0, dim, planes, expo, row_stats);
    count_launch();
    return (int)cudaGetLastError();
}

// Moduli and bits of the CRT scheme for K' = 2 * dim int8 per row.  With s moduli, P = prod p_m and R = sqrt(P / 2),
// a row scaled to 2^(B - e) |x|_2 <= R_eff = R - (rounding slack) has |q_x . q_y| <= |q_x|_2 |q_y|_2 < P / 2
// (Cauchy-Schwarz), so the symmetric CRT reconstruction is exact; B = floor(log2 R_eff), at most kCrtMaxBits.  A
// unit-norm row gets e = 0, components are off by at most 2^-B / 2, an inner product by sqrt(K') 2^-B and a Gram
// entry by 2 sqrt(K') 2^-B: s is the smallest count whose B keeps that below the tolerance (1e-10, B200SQ_CRT_TOL).
static double crt_tolerance() {
    static const double tol = [] {
        const char* t = getenv("B200SQ_CRT_TOL");
        const double v = t ? atof(t) : 0.0;
        return v > 0.0 ? v : 1e-10;
    }();
    return tol;
}

static CrtPlan crt_plan(int64_t dim) {
    CrtPlan pl;
    pl.n_mod = 0;
    pl.bits = 0;
    pl.e_bias = 0.0;
    const double kp = (double)(2 * dim);
    const int need = (int)ceil(log2(2.0 * sqrt(kp) / crt_tolerance()));
    long double P = 1.0L;
    for (int s = 1; s <= kCrtMaxModuli; ++s) {
        P *= (long double)h_crt_moduli[s - 1];
        // every component of q is off by at most 1/2 (rounding): |q|_2 <= 2^(B - e) |x|_2 + sqrt(K') / 2
        long double r_eff = sqrtl(P / 2.0L) - 0.5L * sqrtl((long double)kp) - 1.0L;
        if (r_eff < 2.0L) continue;
        r_eff = std::min(r_eff, powl(2.0L, (long double)kCrtMaxLog2R));
        const int bits = std::min((int)floorl(log2l(r_eff)), kCrtMaxBits);
        if (bits >= need || (s == kCrtMaxModuli && bits >= 30)) {
            pl.n_mod = s;
            pl.bits = bits;
            pl.e_bias = (double)((long double)bits - log2l(r_eff));
            return pl;
        }
    }
    return pl;
}

int gram_crt_moduli(int64_t dim) { return crt_plan(dim).n_mod; }
int gram_crt_bits(int64_t dim) { return crt_plan(dim).bits; }

bool gram_crt_supported(int64_t dim) { return gram_i8_supported(dim) && gram_crt_moduli(dim) > 0; }

int launch_slice_crt(const double2* psi, int64_t n, int64_t dim, int8_t* planes, int64_t plane_rows, int64_t row0,
                     int* expo, cudaStream_t stream, const RowStats* row_stats, int8_t* digits, int64_t digit_rows,
                     int64_t digit_row0) {
    if (n <= 0) return 0;
    const CrtPlan pl = crt_plan(dim);
    if (pl.n_mod <= 0) return (int)cudaErrorInvalidValue;
    k_slice_crt<<<(unsigned)n, 256, 0, stream>>>(psi, plane_rows, row0, dim, pl, planes, expo, row_stats, digits,
                                                 digit_rows, digit_row0);
    count_launch();
    return (int)cudaGetLastError();
}

int launch_digits_to_crt(const int8_t* digits, int64_t rows_d, int64_t row0_d, int64_t n, int64_t dim, int8_t* out,
                         int64_t rows_o, int64_t row0_o, cudaStream_t stream) {
    if (n <= 0) return 0;
    const CrtPlan pl = crt_plan(dim);
    if (pl.n_mod <= 0) return (int)cudaErrorInvalidValue;
    const int64_t kbytes = 2 * dim;
    const dim3 grid((unsigned)((kbytes / 4 + 256 * kConvCols - 1) / (256 * kConvCols)), (unsigned)n);
    k_digits_to_crt<<<grid, 256, 0, stream>>>(digits, rows_d, row0_d, kbytes, pl.n_mod, out, rows_o, row0_o);
    count_launch();
    return (int)cudaGetLastError();
}

int launch_gram_i8(const double2* x, int64_t nx, const double2* y, int64_t ny, int64_t dim, double* k,
                   int64_t ldk, uint32_t flags, cudaStream_t stream, const RowStats* row_stats_x,
                   const RowStats* row_stats_y) {
    if (nx <= 0 || ny <= 0) return 0;
    const bool sym = flags & B200SQ_GRAM_SYMMETRIC;
    const bool crt = (flags & B200SQ_GRAM_CRT) && gram_crt_supported(dim);
    if (!crt) flags &= ~B200SQ_GRAM_CRT;
    const int n_planes = crt ? gram_crt_moduli(dim) : kSlices;
    auto slice = [&](const double2* psi, int64_t n, int8_t* planes, int* expo, const RowStats* rm) {
        return crt ? launch_slice_crt(psi, n, dim, planes, n, 0, expo, stream, rm)
                   : launch_slice_i8(psi, n, dim, planes, n, 0, expo, stream, rm);
    };
    const size_t plane_bytes = (size_t)(2 * dim);  // per row
    int8_t *px = nullptr, *py = nullptr;
    int *ex = nullptr, *ey = nullptr;
    cudaError_t e;
    auto cleanup = [&]() {
        if (px) cudaFreeAsync(px, stream);
        if (py && py != px) cudaFreeAsync(py, stream);
        if (ex) cudaFreeAsync(ex, stream);
        if (ey && ey != ex) cudaFreeAsync(ey, stream);
    };
    if ((e = cudaMallocAsync(&py, 2 * n_planes * (size_t)ny * plane_bytes, stream)) != cudaSuccess) return (int)e;
    if ((e = cudaMallocAsync(&ey, ny * sizeof(int), stream)) != cudaSuccess) { cleanup(); return (int)e; }
    slice(y, ny, py, ey, row_stats_y);
    if (sym) {
        px = py;
        ex = ey;
    } else {
        if ((e = cudaMallocAsync(&px, 2 * n_planes * (size_t)nx * plane_bytes, stream)) != cudaSuccess) { cleanup(); return (int)e; }
        if ((e = cudaMallocAsync(&ex, nx * sizeof(int), stream)) != cudaSuccess) { cleanup(); return (int)e; }
        slice(x, nx, px, ex, row_stats_x);
    }
    const int rc = launch_gram_planes(px, ex, nx, 0, nx, py, ey, ny, 0, ny, dim, k, ldk, flags, stream);
    cleanup();
    return rc;
}

// Gram of rows [x0, x0 + nx) of the planes px against rows [y0, y0 + ny) of py (see k_slice_i8 for the layout;
// rows_x / rows_y = rows per plane).  Symmetric flag: px == py and x0 == y0.
int launch_gram_planes(const int8_t* px, const int* ex, int64_t rows_x, int64_t x0, int64_t nx, const int8_t* py,
                       const int* ey, int64_t rows_y, int64_t y0, int64_t ny, int64_t dim, double* k, int64_t ldk,
                       uint32_t flags, cudaStream_t stream) {
    b200sq_gram_job job;
    memset(&job, 0, sizeof(job));
    job.x0 = x0;
    job.nx = nx;
    job.planes_y_dev = py;
    job.expo_y_dev = ey;
    job.plane_rows_y = rows_y;
    job.y0 = y0;
    job.ny = ny;
    job.K_dev = k;
    job.ldk = ldk;
    job.flags = flags;
    return launch_gram_jobs(px, ex, rows_x, dim, 1, &job, stream);
}

// Several Gram blocks that share the row operand px in one persistent launch (see the JOB TABLE note at the top).
// Column planes only need their kSlices nat planes resident (a received block carries no rot planes).
int launch_gram_jobs(const int8_t* px, const int* ex, int64_t rows_x, int64_t dim, int n_jobs,
                     const b200sq_gram_job* jobs, cudaStream_t stream) {
    if (n_jobs <= 0) return 0;
    if (n_jobs > kMaxJobs) {   // more blocks than one launch holds: split
        for (int j0 = 0; j0 < n_jobs; j0 += kMaxJobs) {
            const int rc = launch_gram_jobs(px, ex, rows_x, dim, std::min(kMaxJobs, n_jobs - j0), jobs + j0, stream);
            if (rc) return rc;
        }
        return 0;
    }
    const int64_t kbytes = 2 * dim;
    double* W = nullptr;
    cudaError_t e;
    auto cleanup = [&]() {
        if (W) cudaFreeAsync(W, stream);
    };
    I8Args a;
    memset(&a, 0, sizeof(a));
    a.rows_a = rows_x;
    a.kbytes = (int)kbytes;
    // one scheme per launch: digit slices (default) or CRT residues (B200SQ_GRAM_CRT on the jobs)
    const bool crt = jobs[0].flags & B200SQ_GRAM_CRT;
    for (int j = 1; j < n_jobs; ++j)
        if (((jobs[j].flags & B200SQ_GRAM_CRT) != 0) != crt) return (int)cudaErrorInvalidValue;
    const int n_mod = crt ? gram_crt_moduli(dim) : 0;
    if (crt && n_mod <= 0) return (int)cudaErrorInvalidValue;
    a.crt = crt ? 1 : 0;
    a.n_planes = crt ? n_mod : kSlices;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static const int force_cg = getenv("B200SQ_I8_CG") ? atoi(getenv("B200SQ_I8_CG")) : 0;
    static const int force_chunk = getenv("B200SQ_I8_CHUNK") ? atoi(getenv("B200SQ_I8_CHUNK")) : 0;
    auto count_tiles = [&](int mn) {
        int64_t t = 0;
        for (int j = 0; j < n_jobs; ++j) {
            if (jobs[j].nx <= 0 || jobs[j].ny <= 0) continue;
            const int64_t tm = (jobs[j].nx + mn - 1) / mn, tn = (jobs[j].ny + mn - 1) / mn;
            t += (jobs[j].flags & B200SQ_GRAM_SYMMETRIC) ? tm * (tm + 1) / 2 : tm * tn;
        }
        return t;
    };
    if (count_tiles(kTile) == 0) return 0;
    // K is cut into chunks of at most 16 KiB (the int32 accumulators stay exact).  CTA pairs (256 x 256 tiles,
    // cta_group::2) halve the L2 -> SMEM operand traffic per MMA but need enough units to fill the machine;
    // otherwise single CTAs with 128 x 128 tiles (TMEM double buffered) run.  The chunk is halved (down to 4 KiB)
    // while the last wave of the persistent workers would be less than ~85 % full.
    // (CRT: one plane pair per accumulator, so 128 KiB of K stay exact: 127^2 * 2^17 < 2^31)
    const int chunk_cap = crt ? (1 << 17) : kChunkK;
    int chunk_max = (int)(kbytes < chunk_cap ? ((kbytes + kStageK - 1) / kStageK) * kStageK : chunk_cap);
    auto n_chunks_of = [&](int cb) { return (kbytes + cb - 1) / cb; };
    const int cg = force_cg ? force_cg
                            : (count_tiles(256) * (crt ? n_mod : 1) * n_chunks_of(std::min(chunk_max, 4096)) >= sms / 2 ? 2 : 1);
    const int workers_max = sms / cg;
    a.chunk_bytes = chunk_max;
    {
        const int64_t tiles = count_tiles(kTile * cg) * (crt ? n_mod : 1);
        auto efficiency = [&](int cb) {
            const int64_t units = tiles * n_chunks_of(cb);
            const int64_t waves = (units + workers_max - 1) / workers_max;
            return (double)units / (double)(waves * workers_max);
        };
        while (a.chunk_bytes > 4096 && a.chunk_bytes / 2 >= kStageK && efficiency(a.chunk_bytes) < 0.85 &&
               efficiency(a.chunk_bytes / 2) > efficiency(a.chunk_bytes) + 0.02)
            a.chunk_bytes /= 2;
        if (force_chunk >= kStageK && force_chunk <= chunk_cap && force_chunk % kStageK == 0) a.chunk_bytes = force_chunk;
    }
    a.n_chunks = (int)n_chunks_of(a.chunk_bytes);
    a.mn = kTile * cg;
    if (crt) {
        // one class per modulus: plane m x plane m; (P / p_m)^-1 mod p_m and P * 2^(-2 B) for the epilogue
        I8Schedule& s = a.sched;
        s.n_classes = n_mod;
        long double P = 1.0L;
        for (int m = 0; m < n_mod; ++m) {
            s.cls_shift[m] = 0;
            s.n_pairs[m] = 1;
            s.pi[m][0] = s.pj[m][0] = (uint8_t)m;
            P *= (long double)h_crt_moduli[m];
            int prod = 1;
            for (int j = 0; j < n_mod; ++j)
                if (j != m) prod = (prod * (h_crt_moduli[j] % h_crt_moduli[m])) % h_crt_moduli[m];
            int inv = 1;
            while ((prod * inv) % h_crt_moduli[m] != 1) ++inv;
            a.crt_c[m] = inv;
        }
        a.crt_scale = (double)ldexpl(P, -2 * gram_crt_bits(dim));
    } else {
        // classes from the smallest weight to the largest: c = 5 .. 0, plus the correlated pair (3, 3) of c = 6
        I8Schedule& s = a.sched;
        int nc = 0;
        s.cls_shift[nc] = 6;
        s.n_pairs[nc] = 1;
        s.pi[nc][0] = 3;
        s.pj[nc][0] = 3;
        ++nc;
        for (int c = kSlices - 1; c >= 0; --c, ++nc) {
            s.cls_shift[nc] = c;
            s.n_pairs[nc] = 0;
            for (int i = 0; i <= c; ++i) {
                s.pi[nc][s.n_pairs[nc]] = (uint8_t)i;
                s.pj[nc][s.n_pairs[nc]] = (uint8_t)(c - i);
                ++s.n_pairs[nc];
            }
        }
        s.n_classes = nc;
    }
    CUtensorMap mapA;
    I8Maps maps;
    memset(&maps, 0, sizeof(maps));
    int rc = encode_planes(&mapA, (void*)px, (uint64_t)kbytes, (uint64_t)(2 * a.n_planes) * rows_x);
    if (rc) return rc;
    const void* map_base[kMaxJobs];
    int n_maps = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const b200sq_gram_job& in = jobs[j];
        if (in.nx <= 0 || in.ny <= 0) continue;
        I8Job& J = a.jobs[a.n_jobs];
        const bool sym = in.flags & B200SQ_GRAM_SYMMETRIC;
        J.row0_a = in.x0;
        J.row0_b = in.y0;
        J.rows_b = in.plane_rows_y;
        J.ldk = in.ldk;
        J.K = in.K_dev;
        J.K_mirror = in.K_mirror_dev;
        J.ldm = in.ldk_mirror;
        J.ex = ex;
        J.ey = in.expo_y_dev;
        J.ready = in.ready_flag_dev;
        J.ready_value = in.ready_value;
        J.align_yx = in.align_yx_dev;
        J.align_yy = in.align_yy_dev ? in.align_yy_dev : in.align_yx_dev;
        J.align_out = in.align_out_dev;
        J.align_weight = in.align_weight != 0.0 ? in.align_weight : 1.0;
        J.flags = in.flags;
        J.nx = (int32_t)in.nx;
        J.ny = (int32_t)in.ny;
        const int tiles_m = (int)((in.nx + a.mn - 1) / a.mn);
        J.tiles_n = (int)((in.ny + a.mn - 1) / a.mn);
        J.symmetric = sym ? 1 : 0;
        J.n_tiles = sym ? tiles_m * (tiles_m + 1) / 2 : tiles_m * J.tiles_n;
        J.tile0 = a.n_tiles;
        a.n_tiles += J.n_tiles;
        int m = 0;
        while (m < n_maps && map_base[m] != in.planes_y_dev) ++m;
        if (m == n_maps) {
            map_base[n_maps++] = in.planes_y_dev;
            // only the nat planes of a column operand are read
            rc = encode_planes(&maps.b[m], (void*)in.planes_y_dev, (uint64_t)kbytes, (uint64_t)a.n_planes * in.plane_rows_y);
            if (rc) return rc;
        }
        J.map_b = m;
        ++a.n_jobs;
    }
    const size_t wbytes = crt ? (size_t)a.n_tiles * n_mod * 2 * a.mn * a.mn * (a.n_chunks == 1 ? sizeof(int8_t) : sizeof(int))
                              : (size_t)a.n_tiles * 2 * a.mn * a.mn * sizeof(double);
    if ((e = cudaMallocAsync(&W, wbytes, stream)) != cudaSuccess) { cleanup(); return (int)e; }
    a.W = W;
    if (!(crt && a.n_chunks == 1)) cudaMemsetAsync(W, 0, wbytes, stream);   // (whole-K CRT units store, they do not add)
    const int smem = kStages * kStageBytes + 1024 + 256;
    const int n_units = a.n_tiles * a.n_chunks * (crt ? n_mod : 1);
    {
        auto kern = cg == 2 ? k_gram_i8<2> : k_gram_i8<1>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) {
            cleanup();
            return (int)e;
        }
        const int workers = n_units < workers_max ? n_units : workers_max;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)(workers * cg));
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = cg;
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        if ((e = cudaLaunchKernelEx(&cfg, kern, mapA, maps, a)) != cudaSuccess) {
            cleanup();
            return (int)e;
        }
        count_launch();
    }
    if (crt) {
        const int tab_bytes = n_mod * kCrtTabStride * (int)sizeof(double2);
        if ((e = cudaFuncSetAttribute(k_gram_crt_final, cudaFuncAttributeMaxDynamicSharedMemorySize, tab_bytes)) != cudaSuccess) {
            cleanup();
            return (int)e;
        }
        const int split = a.mn == 256 ? 4 : 2;   // CTAs per tile: enough of them to fill the machine
        if (a.n_chunks == 1) {
            if ((e = cudaFuncSetAttribute(k_gram_crt_final8, cudaFuncAttributeMaxDynamicSharedMemorySize, tab_bytes)) != cudaSuccess) {
                cleanup();
                return (int)e;
            }
            k_gram_crt_final8<<<dim3((unsigned)a.n_tiles, (unsigned)split), 256, tab_bytes, stream>>>(
                reinterpret_cast<const int8_t*>(W), a);
        } else {
            k_gram_crt_final<<<dim3((unsigned)a.n_tiles, (unsigned)split), 256, tab_bytes, stream>>>(
                reinterpret_cast<const int*>(W), a);
        }
    }
    else k_gram_i8_final<<<a.n_tiles, 256, 0, stream>>>(W, a);
    count_launch();
    e = cudaGetLastError();
    cleanup();
    return (int)e;
}

}  // namespace b200sq
