// Fidelity Gram matrix  K[i][j] = | <x_i | y_j> |^2  (K4 of SURVEY.md §2.2).
//
// The contraction is a complex GEMM  C = conj(X) * Y^T  over the amplitude index; both operands
// are K-major (one state per row), which is the layout FP64 tensor-core fragments want.  tcgen05
// has no f64 kind, so the tensor path for FP64 on sm_100a is DMMA (mma.sync.m8n8k4.f64).  Complex
// arithmetic is folded into real DMMAs without any data rearrangement: a lane that loads one
// complex element (re, im) of X and of Y with a single 16-byte LDS feeds four DMMAs
//      Re += Xre*Yre,  Re += Xim*Yim,  Im += Xre*Yim,  Im += Xim*(-Yre)
// i.e. the interleaved (re, im) storage IS the K' = 2K real operand.  The modulus square and the
// symmetric mirror / diagonal policy are fused into the epilogue, so K is written exactly once
// and C never exists in memory.
//
// Tiling: CTA tile 128 x 64 outputs, 8 warps (4 x 2) of 32 x 32, BK = 16 complex per stage,
// 4-stage cp.async pipeline (48 KiB / stage).  Per k-step a warp issues 8 LDS.128 for 64 DMMAs,
// so shared-memory bandwidth sits at ~1/8 of the DMMA issue time: the kernel is FP64-pipe bound
// by construction.  Symmetric Grams only run the tiles that touch the lower triangle.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace b200sq {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, STAGES = 4, NT = 256;
constexpr int A_ELEMS = BM * BK, B_ELEMS = BN * BK, STAGE_ELEMS = A_ELEMS + B_ELEMS;
constexpr int GRAM_SMEM = STAGES * STAGE_ELEMS * (int)sizeof(double2);

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
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

__global__ void __launch_bounds__(NT, 1)
k_gram(const double2* __restrict__ X, int64_t nx, const double2* __restrict__ Y, int64_t ny,
       int64_t dim, double* __restrict__ K, int64_t ldk, uint32_t flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* sm = reinterpret_cast<double2*>(smem_raw);

    const int bi = blockIdx.y, bj = blockIdx.x;
    const int64_t row0 = (int64_t)bi * BM, col0 = (int64_t)bj * BN;
    const bool sym = flags & B200SQ_GRAM_SYMMETRIC;
    if (sym && col0 > row0 + BM - 1) return;  // tile lies strictly above the diagonal

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp & 3, wn = warp >> 2;
    const int lr = lane >> 2, lq = lane & 3;
    const int KT = (int)((dim + BK - 1) / BK);

    auto load_stage = [&](int st, int kt) {
        double2* sa = sm + st * STAGE_ELEMS;
        double2* sb = sa + A_ELEMS;
        const int64_t k0 = (int64_t)kt * BK;
#pragma unroll
        for (int j = 0; j < A_ELEMS / NT; ++j) {
            const int c = tid + j * NT, r = c >> 4, ch = c & 15;
            const bool ok = (row0 + r < nx) && (k0 + ch < dim);
            const double2* src = ok ? X + (row0 + r) * dim + k0 + ch : X;
            cp_async16(sa + slot(r, ch), src, ok ? 16 : 0);
        }
#pragma unroll
        for (int j = 0; j < B_ELEMS / NT; ++j) {
            const int c = tid + j * NT, r = c >> 4, ch = c & 15;
            const bool ok = (col0 + r < ny) && (k0 + ch < dim);
            const double2* src = ok ? Y + (col0 + r) * dim + k0 + ch : Y;
            cp_async16(sb + slot(r, ch), src, ok ? 16 : 0);
        }
    };

    double cre[4][4][2], cim[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) cre[i][j][0] = cre[i][j][1] = cim[i][j][0] = cim[i][j][1] = 0.0;

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
            double2 a[4], b[4];
            double nb[4];
            const int ch = kk * 4 + lq;
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sa[slot(wm * 32 + i * 8 + lr, ch)];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                b[j] = sb[slot(wn * 32 + j * 8 + lr, ch)];
                nb[j] = -b[j].x;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(cre[i][j][0], cre[i][j][1], a[i].x, b[j].x);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(cim[i][j][0], cim[i][j][1], a[i].x, b[j].y);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(cre[i][j][0], cre[i][j][1], a[i].y, b[j].y);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(cim[i][j][0], cim[i][j][1], a[i].y, nb[j]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: |.|^2, diagonal policy, mirror
    const bool diag_one = flags & B200SQ_GRAM_DIAG_ONE, diag_zero = flags & B200SQ_GRAM_DIAG_ZERO;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = row0 + wm * 32 + i * 8 + lr;
        if (r >= nx) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t c = col0 + wn * 32 + j * 8 + 2 * lq + e;
                if (c >= ny) continue;
                double v = cre[i][j][e] * cre[i][j][e] + cim[i][j][e] * cim[i][j][e];
                if (sym) {
                    if (c > r) continue;  // filled by the mirror of (c, r)
                    if (c == r) v = diag_one ? 1.0 : (diag_zero ? 0.0 : v);
                    K[c * ldk + r] = v;
                }
                K[r * ldk + c] = v;
            }
        }
    }
}

}  // namespace

int launch_gram(const double2* x, int64_t nx, const double2* y, int64_t ny, int64_t dim, double* k,
                int64_t ldk, uint32_t flags, cudaStream_t stream) {
    if (nx <= 0 || ny <= 0) return 0;
    cudaError_t e = cudaFuncSetAttribute(k_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM);
    if (e != cudaSuccess) return (int)e;
    const int64_t tm = (nx + BM - 1) / BM, tn = (ny + BN - 1) / BN;
    if (tm > 65535) return (int)cudaErrorInvalidConfiguration;
    dim3 grid((unsigned)tn, (unsigned)tm);
    k_gram<<<grid, NT, GRAM_SMEM, stream>>>(x, nx, y, ny, dim, k, ldk, flags);
    count_launch();
    return (int)cudaGetLastError();
}

}  // namespace b200sq
