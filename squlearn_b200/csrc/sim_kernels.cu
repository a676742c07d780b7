// Gate application (tile kernel), Pauli expectation values and the Gaussian outer kernel.
//
// Tile kernel (K2/K3 of SURVEY.md §2.2): one TEAM of threads per (variant, sample, outer tile)
// work item.  The team builds the per-sample gate matrices in shared memory, stages 2^T
// amplitudes (HBM -> SMEM, 16-byte coalesced, streaming), runs the pass's rounds on them with the
// register micro-op interpreter of uop_exec.cuh, and streams the tile back — or, when the tile
// is the whole state, reduces it straight to Pauli expectation values.  HBM traffic per pass is
// exactly one read and one write of the state batch (2 * 16 * N * 2^n bytes; the first pass only
// writes).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "kernels.h"

namespace b200sq {

namespace {

// The generic pass kernel: rounds run through the step interpreter of tile_exec.cuh.  jit.cu compiles
// the same body (tile_kernel.cuh) at run time with the interpreter replaced by straight-line code.
__global__ void __launch_bounds__(kCtaThreads, B200SQ_TILE_MINBLOCKS) k_tile_pass(const __grid_constant__ TileArgs A) {
    tile_pass_body(A, InterpretedRounds{});
}

// Expectation values of states resident in HBM: one CTA per state, one warp per Pauli term.
__global__ void __launch_bounds__(kCtaThreads) k_expvals(const double2* __restrict__ psi, int n,
                                                         int n_terms,
                                                         const uint32_t* __restrict__ xmask,
                                                         const uint32_t* __restrict__ zmask,
                                                         double* __restrict__ out,
                                                         int64_t out_stride) {
    const int64_t s = blockIdx.x;
    const double2* st = psi + (s << n);
    const uint32_t dim = 1u << n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int t = blockIdx.y * nwarps + warp; t < n_terms; t += gridDim.y * nwarps) {
        const uint32_t xm = xmask[t], zm = zmask[t];
        const int ny = __popc(xm & zm);
        double acc = 0.0;
        for (uint32_t l = lane; l < dim; l += 32) acc += pauli_term(st[l ^ xm], st[l], l, zm, ny);
        acc = warp_sum(acc);
        if (lane == 0) out[s * out_stride + t] = acc;
    }
}

__global__ void k_rbf(const double* __restrict__ fx, int64_t nx, const double* __restrict__ fy,
                      int64_t ny, int nf, double gamma, double* __restrict__ k, int64_t ldk) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ny) return;
    for (int64_t i = blockIdx.y; i < nx; i += gridDim.y) {
        double d2 = 0.0;
        for (int f = 0; f < nf; ++f) {
            const double d = fx[i * nf + f] - fy[j * nf + f];
            d2 = fma(d, d, d2);
        }
        k[i * ldk + j] = exp(-gamma * d2);
    }
}

}  // namespace

// Team geometry and dynamic shared memory of a pass.
//   team_elems: double2 elements per team = tile buffer(s) + matrices/tables + scratch + per-j tables
void tile_pass_geometry(const DPass& P, TileArgs* a) {
    const int T = P.T;
    // one thread per two register groups of eight amplitudes (the interpreter decodes a step once for
    // both), at least a warp, at most the CTA
    int groups = T >= 3 ? (1 << (T - 3)) : 1;
    static const int gpt = getenv("B200SQ_TILE_G") ? atoi(getenv("B200SQ_TILE_G")) : 2;  // groups per thread
    int ts = groups / gpt < 32 ? 32 : (groups / gpt > kCtaThreads ? kCtaThreads : groups / gpt);
    int log2ts = 0;
    while ((1 << log2ts) < ts) ++log2ts;
    const bool plain = P.has_src && P.n_fresh == 0;
    const int J = std::max(1, (1 << T) >> log2ts);
    int elems = (plain ? 2 : 1) * (1 << T) + P.mat_elems + kScratchElems + J;
    elems = (elems + 7) & ~7;  // keep every team 128-byte aligned
    int teams = kCtaThreads / ts;
    while (teams > 1 && (size_t)teams * elems * sizeof(double2) + P.blob_bytes > (size_t)kMaxSmem / 2) teams >>= 1;
    a->team_size = ts;
    a->log2ts = log2ts;
    a->teams_per_cta = teams;
    a->team_elems = elems;
}

int launch_tile_pass(const TileArgs& args, cudaStream_t stream, const void* jit_kernel) {
    const int smem = args.blob_bytes + args.teams_per_cta * args.team_elems * (int)sizeof(double2);
    if (smem > kMaxSmem) return (int)cudaErrorInvalidConfiguration;
    const void* kern = jit_kernel ? jit_kernel : (const void*)k_tile_pass;
    {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        if (e != cudaSuccess) return (int)e;
    }
    const int threads = args.team_size * args.teams_per_cta;
    int64_t ctas = (args.total_items + args.teams_per_cta - 1) / args.teams_per_cta;
    if (ctas <= 0) return 0;
    // persistent grid: as many CTAs as fit on the GPU at once; teams take contiguous item chunks
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    const int64_t resident = (int64_t)sms * per_sm;
    if (ctas > resident) ctas = resident;
    void* params[] = {(void*)&args};
    cudaError_t e = cudaLaunchKernel(kern, dim3((unsigned)ctas), dim3((unsigned)threads), params, (size_t)smem, stream);
    count_launch();
    return (int)(e != cudaSuccess ? e : cudaGetLastError());
}

int launch_expvals(const double2* psi, int64_t n_states, int n_qubits, int n_terms,
                   const uint32_t* xmask_dev, const uint32_t* zmask_dev, double* out,
                   int64_t out_stride, cudaStream_t stream) {
    if (n_states <= 0 || n_terms <= 0) return 0;
    const int nwarps = kCtaThreads / 32;
    int gy = (n_terms + nwarps - 1) / nwarps;
    // enough CTAs to fill the machine when there are few states
    int64_t want = (148 * 8 + n_states - 1) / n_states;
    if (gy > want) gy = (int)(want < 1 ? 1 : want);
    int64_t done = 0;
    while (done < n_states) {  // gridDim.x limit is 2^31-1; chunk defensively
        int64_t chunk = n_states - done;
        if (chunk > (1 << 30)) chunk = 1 << 30;
        k_expvals<<<dim3((unsigned)chunk, (unsigned)gy), kCtaThreads, 0, stream>>>(
            psi + (done << n_qubits), n_qubits, n_terms, xmask_dev, zmask_dev,
            out + done * out_stride, out_stride);
        count_launch();
        done += chunk;
    }
    return (int)cudaGetLastError();
}

int launch_rbf(const double* fx, int64_t nx, const double* fy, int64_t ny, int nf, double gamma,
               double* k, int64_t ldk, cudaStream_t stream) {
    if (nx <= 0 || ny <= 0) return 0;
    dim3 grid((unsigned)((ny + 127) / 128), (unsigned)(nx > 65535 ? 65535 : nx));
    k_rbf<<<grid, 128, 0, stream>>>(fx, nx, fy, ny, nf, gamma, k, ldk);
    count_launch();
    return (int)cudaGetLastError();
}

}  // namespace b200sq
