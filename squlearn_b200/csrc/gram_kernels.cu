// Fidelity Gram matrix  K[i][j] = | <x_i | y_j> |^2  (K4 of SURVEY.md §2.2).
//
// The contraction is a complex GEMM  C = conj(X) * Y^T  over the amplitude index; both operands
// are K-major (one state per row), which is the layout FP64 tensor-core fragments want.  tcgen05
// has no f64 kind, so the tensor path for FP64 on sm_100a is DMMA (mma.sync.m8n8k4.f64; measured
// 37.1 TFLOP/s on B200, above cuBLAS DGEMM's 35.5).  Complex arithmetic is folded into real DMMAs
// without rearranging data: a lane that loads one complex element (re, im) of X and of Y with a
// single 16-byte LDS feeds
//   4M mode:  Re += Xre*Yre, Re += Xim*Yim, Im += Xre*Yim, Im += Xim*(-Yre)       (4 DMMA)
//   3M mode:  T1 += Xre*Yre, T2 += Xim*Yim, T3 += (Xre+Xim)*(Yim-Yre)             (3 DMMA)
//             Re = T1 + T2,  Im = T3 + T1 - T2   (Gauss / Karatsuba: 25 % fewer tensor flops)
// i.e. the interleaved (re, im) storage IS the K' = 2K real operand.  The modulus square and the
// symmetric mirror / diagonal policy are fused into the epilogue, so K is written exactly once
// and C never exists in memory.
//
// Tiling: 8 warps per CTA, BK = 16 complex per stage, 4-stage cp.async pipeline.
//   4M: CTA tile 128 x 64, warp tile 32 x 32 (128 accumulator registers per thread)
//   3M: CTA tile  64 x 64, warp tile 16 x 32 ( 96 accumulator registers per thread)
// Shared-memory traffic is ~1/8 (4M) / ~1/4 (3M) of the DMMA issue time and L2->SM traffic is
// <= 3 TB/s of the measured ~11.8 TB/s, so the kernel is FP64-pipe bound by construction.
//
// Work decomposition (split-K for the ragged last wave): the (lower-triangle) tiles are cut into
// UNITS.  The first floor(tiles / G) * G tiles (G = CTAs resident on the GPU) are whole-tile units;
// each of the remaining R tiles is split along K into S slices, S chosen to minimise the tail
// time ceil(R * S / G) / S, so that the last wave also fills the machine.  Slices park their partial sums in a workspace; the slice that
// arrives last adds them in slice order (deterministic) and runs the epilogue.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "kernels.h"

namespace b200sq {

namespace {

constexpr int BK = 16, STAGES = 4, NT = 256;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem_src),
                 "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// 16-byte slot of (row, chunk) inside a [rows][16] tile; odd rows have chunk bit 2 flipped so
// the two rows a quarter-warp touches land in different halves of the 128-byte bank window.
__device__ __forceinline__ int slot(int row, int chunk) { return row * BK + (chunk ^ ((row & 1) << 2)); }

struct GramArgs {
    const double2* X;
    const double2* Y;
    int64_t nx, ny, dim;
    double* K;
    int64_t ldk;
    uint32_t flags;
    int tiles_m, tiles_n;
    int n_tiles;         // tiles that are computed (lower triangle when symmetric)
    int full_units;      // whole-tile units
    int slices;          // K slices of each remaining tile (>= 1)
    double2* partial;    // [n_tiles - full_units][slices][BM*BN] (re, im) partial sums
    unsigned* counters;  // [n_tiles - full_units] arrival counters (zero-initialised)
};

template <int BM, int BN>
__device__ __forceinline__ int tiles_in_row(const GramArgs& A, int bi) {
    if (!(A.flags & B200SQ_GRAM_SYMMETRIC)) return A.tiles_n;
    // column tiles bj with bj * BN <= bi * BM + BM - 1
    int c = (bi * BM + BM - 1) / BN + 1;
    return c < A.tiles_n ? c : A.tiles_n;
}

// MODE 0: 4M, MODE 1: 3M.  WM x WN = warp tile; warps are laid out (BM / WM) x (BN / WN) = 8.
template <int BM, int BN, int WM, int WN, int MODE>
__global__ void __launch_bounds__(NT, 1) k_gram(const __grid_constant__ GramArgs A) {
    constexpr int MI = WM / 8, NI = WN / 8;
    constexpr int WARPS_M = BM / WM;
    constexpr int A_ELEMS = BM * BK, B_ELEMS = BN * BK, STAGE_ELEMS = A_ELEMS + B_ELEMS;
    constexpr int NACC = MODE == 0 ? 2 : 3;
    static_assert((BM / WM) * (BN / WN) == NT / 32, "eight warps");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* sm = reinterpret_cast<double2*>(smem_raw);
    __shared__ int s_tile[2];
    __shared__ unsigned s_last;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int lr = lane >> 2, lq = lane & 3;

    // ---- decode the unit
    const int unit = blockIdx.x;
    int tile, slice = 0, nslices = 1;
    if (unit < A.full_units) {
        tile = unit;
    } else {
        const int r = (unit - A.full_units) / A.slices;
        slice = (unit - A.full_units) - r * A.slices;
        nslices = A.slices;
        tile = A.full_units + r;
    }
    if (tid == 0) {
        int bi = 0, t = tile;
        for (; bi < A.tiles_m; ++bi) {
            const int c = tiles_in_row<BM, BN>(A, bi);
            if (t < c) break;
            t -= c;
        }
        s_tile[0] = bi;
        s_tile[1] = t;
    }
    __syncthreads();
    const int bi = s_tile[0], bj = s_tile[1];
    const int64_t row0 = (int64_t)bi * BM, col0 = (int64_t)bj * BN;
    const int KT_all = (int)((A.dim + BK - 1) / BK);
    const int kt_begin = (int)((int64_t)KT_all * slice / nslices);
    const int kt_end = (int)((int64_t)KT_all * (slice + 1) / nslices);
    const int KT = kt_end - kt_begin;

    const double2* __restrict__ X = A.X;
    const double2* __restrict__ Y = A.Y;
    auto load_stage = [&](int st, int kt) {
        double2* sa = sm + st * STAGE_ELEMS;
        double2* sb = sa + A_ELEMS;
        const int64_t k0 = (int64_t)(kt_begin + kt) * BK;
#pragma unroll
        for (int j = 0; j < A_ELEMS / NT; ++j) {
            const int c = tid + j * NT, r = c >> 4, ch = c & 15;
            const bool ok = (row0 + r < A.nx) && (k0 + ch < A.dim);
            const double2* src = ok ? X + (row0 + r) * A.dim + k0 + ch : X;
            cp_async16(sa + slot(r, ch), src, ok ? 16 : 0);
        }
#pragma unroll
        for (int j = 0; j < B_ELEMS / NT; ++j) {
            const int c = tid + j * NT, r = c >> 4, ch = c & 15;
            const bool ok = (col0 + r < A.ny) && (k0 + ch < A.dim);
            const double2* src = ok ? Y + (col0 + r) * A.dim + k0 + ch : Y;
            cp_async16(sb + slot(r, ch), src, ok ? 16 : 0);
        }
    };

    double acc[NACC][MI][NI][2];
#pragma unroll
    for (int s = 0; s < NACC; ++s)
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) acc[s][i][j][0] = acc[s][i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < KT) load_stage(nk % STAGES, nk);
            cp_async_commit();
        }
        const double2* sa = sm + (kt % STAGES) * STAGE_ELEMS;
        const double2* sb = sa + A_ELEMS;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double2 a[MI], b[NI];
            double a2[MI], b2[NI];
            const int ch = kk * 4 + lq;
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                a[i] = sa[slot(wm * WM + i * 8 + lr, ch)];
                if (MODE == 1) a2[i] = a[i].x + a[i].y;
            }
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                b[j] = sb[slot(wn * WN + j * 8 + lr, ch)];
                b2[j] = MODE == 1 ? b[j].y - b[j].x : -b[j].x;
            }
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) dmma(acc[0][i][j][0], acc[0][i][j][1], a[i].x, b[j].x);
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) dmma(acc[1][i][j][0], acc[1][i][j][1], a[i].x, b[j].y);
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) dmma(acc[0][i][j][0], acc[0][i][j][1], a[i].y, b[j].y);
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) dmma(acc[1][i][j][0], acc[1][i][j][1], a[i].y, b2[j]);
            } else {
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) dmma(acc[0][i][j][0], acc[0][i][j][1], a[i].x, b[j].x);
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) dmma(acc[1][i][j][0], acc[1][i][j][1], a[i].y, b[j].y);
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) dmma(acc[NACC - 1][i][j][0], acc[NACC - 1][i][j][1], a2[i], b2[j]);
            }
        }
    }
    cp_async_wait<0>();

    // ---- (re, im) of this unit's share of the sum
    double re[MI][NI][2], im[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (MODE == 0) {
                    re[i][j][e] = acc[0][i][j][e];
                    im[i][j][e] = acc[1][i][j][e];
                } else {
                    re[i][j][e] = acc[0][i][j][e] + acc[1][i][j][e];
                    im[i][j][e] = acc[NACC - 1][i][j][e] + (acc[0][i][j][e] - acc[1][i][j][e]);
                }
            }

    // ---- split tiles: park partial sums, last arrival reduces them in slice order
    if (nslices > 1) {
        const int r = tile - A.full_units;
        double2* mine = A.partial + ((size_t)r * nslices + slice) * (BM * BN);
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    mine[((i * NI + j) * 2 + e) * NT + tid] = double2{re[i][j][e], im[i][j][e]};
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(A.counters + r, 1u) == (unsigned)(nslices - 1));
        __syncthreads();
        if (!s_last) return;
        __threadfence();
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) re[i][j][e] = im[i][j][e] = 0.0;
        for (int s = 0; s < nslices; ++s) {
            const double2* p = A.partial + ((size_t)r * nslices + s) * (BM * BN);
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double2 v = __ldcg(p + ((i * NI + j) * 2 + e) * NT + tid);
                        re[i][j][e] += v.x;
                        im[i][j][e] += v.y;
                    }
        }
    }

    // ---- epilogue: |.|^2, diagonal policy, mirror
    const bool sym = A.flags & B200SQ_GRAM_SYMMETRIC;
    const bool diag_one = A.flags & B200SQ_GRAM_DIAG_ONE, diag_zero = A.flags & B200SQ_GRAM_DIAG_ZERO;
    double* __restrict__ K = A.K;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int64_t r = row0 + wm * WM + i * 8 + lr;
        if (r >= A.nx) continue;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t c = col0 + wn * WN + j * 8 + 2 * lq + e;
                if (c >= A.ny) continue;
                double v = re[i][j][e] * re[i][j][e] + im[i][j][e] * im[i][j][e];
                if (sym) {
                    if (c > r) continue;  // filled by the mirror of (c, r)
                    if (c == r) v = diag_one ? 1.0 : (diag_zero ? 0.0 : v);
                    K[c * A.ldk + r] = v;
                }
                K[r * A.ldk + c] = v;
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// 3M kernel, warp-specialised (the default path).  Eight consumer warps issue DMMAs; a ninth
// PRODUCER warp streams the operand tiles L2 -> SMEM with 16-byte cp.async (LDGSTS) into the same
// XOR-swizzled layout as above and lets the copy unit signal completion on an mbarrier
// (cp.async.mbarrier.arrive.noinc).  Stages are handed over through full/empty mbarriers, so the
// consumer warps never meet at a CTA-wide barrier inside the k-loop: they drift apart and the
// FP64 pipe always finds a ready warp.  (The classic kernel above pays ~1300 cycles of barrier
// drain + load issue per k-tile: 15 % at 128x64 4M tiles, 30 % at 64x64 3M tiles.)
// Measured alternative, kept out: one cp.async.bulk per 256-byte operand row (128 bulk copies
// per stage) costs ~60 cycles per copy in the copy engine and ran 1.7x SLOWER than the classic
// kernel.
constexpr int WS_BM = 64, WS_BN = 64, WS_STAGES = 5;
constexpr int WS_STAGE_ELEMS = (WS_BM + WS_BN) * BK;
constexpr int WS_SMEM = WS_STAGES * WS_STAGE_ELEMS * (int)sizeof(double2) + 2 * WS_STAGES * 8;

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(addr), "r"(parity) : "memory");
}
// WM x WN = consumer warp tile; (64 / WM) * (64 / WN) consumer warps + one producer warp group
// (registers are allotted per group of 4 warps, so the producer occupies a whole group).
template <int WM, int WN>
__global__ void __launch_bounds__((64 / WM) * (64 / WN) * 32 + 128, 1) k_gram3m_ws(const __grid_constant__ GramArgs A) {
    constexpr int BM = WS_BM, BN = WS_BN;
    constexpr int NCW = (BM / WM) * (BN / WN);      // consumer warps
    constexpr int CT = NCW * 32;                     // consumer threads
    constexpr int REG_CONS = NCW == 8 ? 208 : 112, REG_PROD = NCW == 8 ? 40 : 24;
    constexpr int MI = WM / 8, NI = WN / 8, WARPS_M = BM / WM;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* sm = reinterpret_cast<double2*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + WS_STAGES * WS_STAGE_ELEMS);
    uint64_t* empty = full + WS_STAGES;
    __shared__ int s_tile[2];
    __shared__ unsigned s_last;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int unit = blockIdx.x;
    int tile, slice = 0, nslices = 1;
    if (unit < A.full_units) {
        tile = unit;
    } else {
        const int r = (unit - A.full_units) / A.slices;
        slice = (unit - A.full_units) - r * A.slices;
        nslices = A.slices;
        tile = A.full_units + r;
    }
    if (tid == 0) {
        int bi = 0, t = tile;
        for (; bi < A.tiles_m; ++bi) {
            const int c = tiles_in_row<BM, BN>(A, bi);
            if (t < c) break;
            t -= c;
        }
        s_tile[0] = bi;
        s_tile[1] = t;
        for (int s = 0; s < WS_STAGES; ++s) {
            mbar_init(full + s, 32);   // one asynchronous arrival per producer lane
            mbar_init(empty + s, NCW);  // one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int bi = s_tile[0], bj = s_tile[1];
    const int64_t row0 = (int64_t)bi * BM, col0 = (int64_t)bj * BN;
    const int KT_all = (int)((A.dim + BK - 1) / BK);
    const int kt_begin = (int)((int64_t)KT_all * slice / nslices);
    const int kt_end = (int)((int64_t)KT_all * (slice + 1) / nslices);
    const int KT = kt_end - kt_begin;

    if (warp >= NCW) {
        // ================= producer warp group =================
        // hand most of this group's registers to the consumers (setmaxnreg works per warp group)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_PROD));
        if (warp != NCW) return;
        // lane -> (row parity, 16-byte chunk); chunk c = lane + 32 j  <=>  row 2j + (lane >> 4)
        const int pr = lane >> 4, pc = lane & 15;
        const bool full_tile = (row0 + BM <= A.nx) && (col0 + BN <= A.ny) && (A.dim % BK == 0);
        const double2* pa = A.X + (row0 + pr) * A.dim + (int64_t)kt_begin * BK + pc;
        const double2* pb = A.Y + (col0 + pr) * A.dim + (int64_t)kt_begin * BK + pc;
        const int64_t step = 2 * A.dim;
        int s = 0;
        unsigned eph = 1;  // parity of the empty-barrier phase to wait for (first ring pass: none)
        for (int kt = 0; kt < KT; ++kt) {
            if (kt >= WS_STAGES) mbar_wait(empty + s, eph);
            double2* sa = sm + s * WS_STAGE_ELEMS;
            double2* sb = sa + BM * BK;
            if (full_tile) {
                const double2* qa = pa;
                const double2* qb = pb;
#pragma unroll 8
                for (int j = 0; j < BM / 2; ++j, qa += step) cp_async16(sa + slot(2 * j + pr, pc), qa, 16);
#pragma unroll 8
                for (int j = 0; j < BN / 2; ++j, qb += step) cp_async16(sb + slot(2 * j + pr, pc), qb, 16);
            } else {
                const int64_t k0 = (int64_t)(kt_begin + kt) * BK;
                for (int j = 0; j < BM / 2; ++j) {
                    const int r = 2 * j + pr;
                    const bool ok = (row0 + r < A.nx) && (k0 + pc < A.dim);
                    cp_async16(sa + slot(r, pc), ok ? A.X + (row0 + r) * A.dim + k0 + pc : A.X, ok ? 16 : 0);
                }
                for (int j = 0; j < BN / 2; ++j) {
                    const int r = 2 * j + pr;
                    const bool ok = (col0 + r < A.ny) && (k0 + pc < A.dim);
                    cp_async16(sb + slot(r, pc), ok ? A.Y + (col0 + r) * A.dim + k0 + pc : A.Y, ok ? 16 : 0);
                }
            }
            // the copy unit arrives on full[s] once all of this lane's copies have landed
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(
                             (unsigned)__cvta_generic_to_shared(full + s))
                         : "memory");
            pa += BK;
            pb += BK;
            if (++s == WS_STAGES) {
                s = 0;
                eph ^= 1u;
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        return;
    }

    // ================= consumers =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_CONS));
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int lr = lane >> 2, lq = lane & 3;
    double acc[3][MI][NI][2];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) acc[s][i][j][0] = acc[s][i][j][1] = 0.0;

    // Software pipeline over (k-tile, k-step): the fragments AND the 3M sums (Xr+Xi, Yi-Yr) of the
    // next step are produced while the current step feeds the tensor pipe, also across k-tile
    // boundaries, so neither LDS nor DADD latency nor the mbarrier wait sits in front of a DMMA.
    const int a_row = wm * WM + lr, b_row = wn * WN + lr;
    double2 a[MI], b[NI], an[MI], bn[NI];
    double a2[MI], b2[NI];
    if (KT > 0) {
        mbar_wait(full + 0, 0);
        const double2* sa = sm;
        const double2* sb = sa + BM * BK;
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            a[i] = sa[slot(a_row + i * 8, lq)];
            a2[i] = a[i].x + a[i].y;
        }
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            b[j] = sb[slot(b_row + j * 8, lq)];
            b2[j] = b[j].y - b[j].x;
        }
    }
    int s = 0;
    unsigned fph = 0;  // parity of the full-barrier phase of stage s
    for (int kt = 0; kt < KT; ++kt) {
        const double2* sa = sm + s * WS_STAGE_ELEMS;
        const double2* sb = sa + BM * BK;
        int sn = s + 1;
        unsigned nph = fph;
        if (sn == WS_STAGES) {
            sn = 0;
            nph ^= 1u;
        }
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            bool have_next = true;
            if (kk + 1 < BK / 4) {
                const int ch = (kk + 1) * 4 + lq;
#pragma unroll
                for (int i = 0; i < MI; ++i) an[i] = sa[slot(a_row + i * 8, ch)];
#pragma unroll
                for (int j = 0; j < NI; ++j) bn[j] = sb[slot(b_row + j * 8, ch)];
            } else if (kt + 1 < KT) {
                mbar_wait(full + sn, nph);
                const double2* na = sm + sn * WS_STAGE_ELEMS;
                const double2* nb = na + BM * BK;
#pragma unroll
                for (int i = 0; i < MI; ++i) an[i] = na[slot(a_row + i * 8, lq)];
#pragma unroll
                for (int j = 0; j < NI; ++j) bn[j] = nb[slot(b_row + j * 8, lq)];
            } else {
                have_next = false;
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma(acc[0][i][j][0], acc[0][i][j][1], a[i].x, b[j].x);
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma(acc[1][i][j][0], acc[1][i][j][1], a[i].y, b[j].y);
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma(acc[2][i][j][0], acc[2][i][j][1], a2[i], b2[j]);
            if (kk + 1 == BK / 4) {
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + s);
            }
            if (have_next) {
#pragma unroll
                for (int i = 0; i < MI; ++i) {
                    a[i] = an[i];
                    a2[i] = an[i].x + an[i].y;
                }
#pragma unroll
                for (int j = 0; j < NI; ++j) {
                    b[j] = bn[j];
                    b2[j] = bn[j].y - bn[j].x;
                }
            }
        }
        s = sn;
        fph = nph;
    }

    double re[MI][NI][2], im[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                re[i][j][e] = acc[0][i][j][e] + acc[1][i][j][e];
                im[i][j][e] = acc[2][i][j][e] + (acc[0][i][j][e] - acc[1][i][j][e]);
            }

    if (nslices > 1) {
        const int r = tile - A.full_units;
        double2* mine = A.partial + ((size_t)r * nslices + slice) * (BM * BN);
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    mine[((i * NI + j) * 2 + e) * CT + tid] = double2{re[i][j][e], im[i][j][e]};
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory");
        if (tid == 0) s_last = (atomicAdd(A.counters + r, 1u) == (unsigned)(nslices - 1));
        asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory");
        if (!s_last) return;
        __threadfence();
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) re[i][j][e] = im[i][j][e] = 0.0;
        for (int s = 0; s < nslices; ++s) {
            const double2* p = A.partial + ((size_t)r * nslices + s) * (BM * BN);
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double2 v = __ldcg(p + ((i * NI + j) * 2 + e) * CT + tid);
                        re[i][j][e] += v.x;
                        im[i][j][e] += v.y;
                    }
        }
    }

    const bool sym = A.flags & B200SQ_GRAM_SYMMETRIC;
    const bool diag_one = A.flags & B200SQ_GRAM_DIAG_ONE, diag_zero = A.flags & B200SQ_GRAM_DIAG_ZERO;
    double* __restrict__ K = A.K;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int64_t r = row0 + wm * WM + i * 8 + lr;
        if (r >= A.nx) continue;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t c = col0 + wn * WN + j * 8 + 2 * lq + e;
                if (c >= A.ny) continue;
                double v = re[i][j][e] * re[i][j][e] + im[i][j][e] * im[i][j][e];
                if (sym) {
                    if (c > r) continue;
                    if (c == r) v = diag_one ? 1.0 : (diag_zero ? 0.0 : v);
                    K[c * A.ldk + r] = v;
                }
                K[r * A.ldk + c] = v;
            }
        }
    }
}

// KERNEL: 0 = classic template, 1 = warp-specialised 3M with 8 consumer warps, 2 = with 16
template <int BM, int BN, int WM, int WN, int MODE, int KERNEL>
int launch_variant(const double2* x, int64_t nx, const double2* y, int64_t ny, int64_t dim, double* k,
                   int64_t ldk, uint32_t flags, cudaStream_t stream) {
    constexpr int SMEM = KERNEL != 0 ? WS_SMEM : STAGES * (BM + BN) * BK * (int)sizeof(double2);
    constexpr int THREADS = KERNEL == 1 ? 8 * 32 + 128 : (KERNEL == 2 ? 16 * 32 + 128 : NT);
    cudaError_t e;
    if constexpr (KERNEL == 1) e = cudaFuncSetAttribute(k_gram3m_ws<16, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    else if constexpr (KERNEL == 2) e = cudaFuncSetAttribute(k_gram3m_ws<16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    else e = cudaFuncSetAttribute(k_gram<BM, BN, WM, WN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

    GramArgs A;
    A.X = x; A.Y = y; A.nx = nx; A.ny = ny; A.dim = dim; A.K = k; A.ldk = ldk; A.flags = flags;
    A.tiles_m = (int)((nx + BM - 1) / BM);
    A.tiles_n = (int)((ny + BN - 1) / BN);
    int64_t n_tiles = 0;
    for (int bi = 0; bi < A.tiles_m; ++bi) {
        int c = A.tiles_n;
        if (flags & B200SQ_GRAM_SYMMETRIC) {
            int cc = (bi * BM + BM - 1) / BN + 1;
            c = cc < c ? cc : c;
        }
        n_tiles += c;
    }
    if (n_tiles > 0x3fffffff) return (int)cudaErrorInvalidConfiguration;
    A.n_tiles = (int)n_tiles;
    const int G = sms;  // one CTA per SM (launch bounds)
    const int KT = (int)((dim + BK - 1) / BK);
    // Tail policy: the last rem = n_tiles mod G tiles are cut into `slices` K-slices each.  With s
    // slices the tail takes ceil(rem * s / G) / s tile-times; pick the s (<= 16, and leaving every
    // slice at least 16 k-tiles) that minimises it.
    int rem = A.n_tiles % G;
    int slices = 1;
    if (rem) {
        const int s_max = KT / 16 < 16 ? KT / 16 : 16;
        double best = 1.0;
        for (int sl = 2; sl <= s_max; ++sl) {
            // + 1 % of a tile-time per slice: partial-sum traffic, pipeline fill, launch
            const double t = (double)((rem * sl + G - 1) / G) / sl + 0.01 * sl;
            if (t < best - 1e-9) {
                best = t;
                slices = sl;
            }
        }
    }
    if (const char* env = getenv("B200SQ_GRAM_SLICES")) {  // tuning override
        const int forced = atoi(env);
        if (forced >= 1 && rem) slices = forced;
    }
    if (slices == 1) rem = 0;
    A.full_units = A.n_tiles - rem;
    A.slices = slices;
    A.partial = nullptr;
    A.counters = nullptr;
    const int units = A.full_units + rem * slices;
    void* ws = nullptr;
    if (rem) {
        static bool pool_configured[64] = {};
        if (dev >= 0 && dev < 64 && !pool_configured[dev]) {
            // keep freed workspace in the stream-ordered pool instead of returning it to the driver
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            pool_configured[dev] = true;
        }
        const size_t part_bytes = (size_t)rem * slices * BM * BN * sizeof(double2);
        const size_t cnt_bytes = ((size_t)rem * sizeof(unsigned) + 255) & ~(size_t)255;
        e = cudaMallocAsync(&ws, part_bytes + cnt_bytes, stream);
        if (e != cudaSuccess) return (int)e;
        A.counters = (unsigned*)ws;
        A.partial = (double2*)((char*)ws + cnt_bytes);
        cudaMemsetAsync(ws, 0, cnt_bytes, stream);
    }
    if constexpr (KERNEL == 1) k_gram3m_ws<16, 32><<<units, THREADS, SMEM, stream>>>(A);
    else if constexpr (KERNEL == 2) k_gram3m_ws<16, 16><<<units, THREADS, SMEM, stream>>>(A);
    else k_gram<BM, BN, WM, WN, MODE><<<units, THREADS, SMEM, stream>>>(A);
    count_launch();
    e = cudaGetLastError();
    if (ws) cudaFreeAsync(ws, stream);
    return (int)e;
}

}  // namespace

int launch_gram(const double2* x, int64_t nx, const double2* y, int64_t ny, int64_t dim, double* k,
                int64_t ldk, uint32_t flags, cudaStream_t stream) {
    if (nx <= 0 || ny <= 0) return 0;
    if (flags & B200SQ_GRAM_3M) {
        if (!(flags & B200SQ_GRAM_CLASSIC)) {
            static const int ws_warps = getenv("B200SQ_GRAM_WARPS") ? atoi(getenv("B200SQ_GRAM_WARPS")) : 8;
            if (ws_warps == 16)
                return launch_variant<64, 64, 16, 16, 1, 2>(x, nx, y, ny, dim, k, ldk, flags, stream);
            return launch_variant<64, 64, 16, 32, 1, 1>(x, nx, y, ny, dim, k, ldk, flags, stream);
        }
        return launch_variant<64, 64, 16, 32, 1, 0>(x, nx, y, ny, dim, k, ldk, flags, stream);
    }
    return launch_variant<128, 64, 32, 32, 0, 0>(x, nx, y, ny, dim, k, ldk, flags, stream);
}

}  // namespace b200sq
