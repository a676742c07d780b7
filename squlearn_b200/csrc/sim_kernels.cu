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

#include "kernels.h"
#include "uop_exec.cuh"

namespace b200sq {

namespace {

#ifndef B200SQ_TILE_MINBLOCKS
#define B200SQ_TILE_MINBLOCKS 3
#endif
constexpr int kCtaThreads = 256;
constexpr int kMaxSmem = 227 * 1024;

__device__ __forceinline__ void team_sync(int team_size, int team) {
    if (team_size == 32) {
        __syncwarp();
    } else if (team_size == (int)blockDim.x) {
        __syncthreads();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(team_size) : "memory");
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kCtaThreads, B200SQ_TILE_MINBLOCKS) k_tile_pass(const __grid_constant__ TileArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* smem = reinterpret_cast<double2*>(smem_raw);

    const int TS = A.team_size;
    const int team = threadIdx.x / TS;
    const int tid = threadIdx.x - team * TS;
    const int64_t item = (int64_t)blockIdx.x * A.teams_per_cta + team;
    if (item >= A.total_items) return;  // the whole team leaves together

    const DPass& P = A.pass;
    const int T = P.T;
    const uint32_t tile_elems = 1u << T;
    const uint32_t n_outer = 1u << (A.n - T);
    const uint32_t outer = (uint32_t)(item % n_outer);
    const int64_t rest = item / n_outer;
    const int64_t sample = A.sample0 + rest % A.n_samples;
    const int variant = A.variant0 + (int)(rest / A.n_samples);

    double2* tile = smem + (size_t)team * A.team_elems;
    double2* mats = tile + tile_elems;

    const double* xrow = A.x ? A.x + sample * A.d : nullptr;
    const int vgate = A.var_gate ? A.var_gate[variant] : -1;
    const double vshift = A.var_gate ? A.var_shift[variant] : 0.0;
    double2* state = A.psi ? A.psi + (((int64_t)variant * A.n_total + sample - A.psi_off) << A.n) : nullptr;
    const uint32_t obase = scatter_runs(outer, P.n_oruns, P.omask, P.oshift);

    // ---- first pass: a tile whose outer qubits are not all |0> is identically zero
    if (P.first && outer != 0) {
        bool zero = false;  // outer qubits are never folded, so any set outer bit means a zero tile
        for (int j = 0; j < A.n - T; ++j) zero |= (outer >> j) & 1u;
        if (zero) {
            if (A.store)
                for (uint32_t l = tid; l < tile_elems; l += TS)
                    __stcs(state + (obase | scatter_runs(l, P.n_lruns, P.lmask, P.lshift)), double2{0.0, 0.0});
            return;
        }
    }

    // ---- per-sample gate matrices
    for (int ui = tid; ui < P.n_uops; ui += TS) {
        const DUop u = A.uops[ui];
        const b200sq_gate g = A.gates[u.gate];
        double theta = gate_angle(g, xrow, A.table, A.n_total, sample);
        if ((int)u.gate == vgate) theta += vshift;
        build_matrix(u, g, theta, mats + u.mat);
    }

    // ---- stage the tile
    bool zero_tile = false;
    if (P.first) {
        // product-state start: |psi> = (x)_q v_q with v_q = (folded 1-qubit gates of q)|0>
        double2* qv = mats + P.mat_elems;  // [32][2]
        double2* lo = qv + 64;             // [64] partial products over the low tile bits
        double2* hi = lo + 64;             // [<=128] partial products over the high tile bits
        for (int q = tid; q < A.n; q += TS)
            fold_qubit_vector(q, A.inits, A.n_init, A.gates, xrow, A.table, A.n_total, sample, vgate,
                              vshift, qv + 2 * q);
        team_sync(TS, team);
        const int nlo = T < 6 ? T : 6, nhi = T - nlo;
        const double2 oscal = product_amp(outer, A.n - T, P.extq + T, qv);
        zero_tile = (oscal.x == 0.0 && oscal.y == 0.0);
        for (int t = tid; t < (1 << nlo); t += TS) lo[t] = product_amp(t, nlo, P.extq, qv);
        for (int t = tid; t < (1 << nhi); t += TS) hi[t] = cmul(oscal, product_amp(t, nhi, P.extq + nlo, qv));
        team_sync(TS, team);
        for (uint32_t l = tid; l < tile_elems; l += TS)
            tile[tile_sw(l)] = cmul(lo[l & ((1u << nlo) - 1u)], hi[l >> nlo]);
    } else {
        // batches of eight independent 16-byte loads per thread keep the memory pipe full
        for (uint32_t l0 = tid; l0 < tile_elems; l0 += 8 * TS) {
            double2 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t l = l0 + j * TS;
                if (l < tile_elems)
                    v[j] = __ldcs(state + (obase | scatter_runs(l, P.n_lruns, P.lmask, P.lshift)));
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t l = l0 + j * TS;
                if (l < tile_elems) tile[tile_sw(l)] = v[j];
            }
        }
    }
    team_sync(TS, team);

    // ---- rounds
    if (!zero_tile) {
        const uint32_t ext_base = outer << T;
        for (int r = 0; r < P.n_rounds; ++r) {
            exec_round_thread(tile, A.rounds[r], A.uops, mats, ext_base, T, tid, TS);
            team_sync(TS, team);
        }
    }

    // ---- epilogue
    if (A.store) {
        for (uint32_t l = tid; l < tile_elems; l += TS) {
            const uint32_t gidx = obase | scatter_runs(l, P.n_lruns, P.lmask, P.lshift);
            __stcs(state + gidx, tile[tile_sw(l)]);
        }
    }
    if (A.n_terms > 0) {
        // T == n here: the tile is the state and local index == amplitude index.
        const int nwarps = TS >> 5, warp = tid >> 5, lane = tid & 31;
        double* out = A.out + ((int64_t)variant * A.n_total + sample) * A.n_terms;
        for (int t = warp; t < A.n_terms; t += nwarps) {
            const uint32_t xm = A.xmask[t], zm = A.zmask[t];
            const int ny = __popc(xm & zm);
            double acc = 0.0;
            for (uint32_t l = lane; l < tile_elems; l += 32)
                acc += pauli_term(tile[tile_sw(l ^ xm)], tile[tile_sw(l)], l, zm, ny);
            acc = warp_sum(acc);
            if (lane == 0) out[t] = acc;
        }
    }
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

int tile_pass_smem_bytes(int T, int mat_elems, int* team_size, int* teams_per_cta,
                         int* team_elems) {
    int groups = T >= 3 ? (1 << (T - 3)) : 1;
    int ts = groups < 32 ? 32 : (groups > kCtaThreads ? kCtaThreads : groups);
    int elems = (1 << T) + mat_elems + 64 + 64 + 128;  // tile + matrices + product-start scratch
    elems = (elems + 7) & ~7;  // keep every team 128-byte aligned
    int teams = kCtaThreads / ts;
    while (teams > 1 && (size_t)teams * elems * sizeof(double2) > (size_t)kMaxSmem / 2) teams >>= 1;
    *team_size = ts;
    *teams_per_cta = teams;
    *team_elems = elems;
    return teams * elems * (int)sizeof(double2);
}

int launch_tile_pass(const TileArgs& args, cudaStream_t stream) {
    const int smem = args.teams_per_cta * args.team_elems * (int)sizeof(double2);
    if (smem > kMaxSmem) return (int)cudaErrorInvalidConfiguration;
    {
        cudaError_t e = cudaFuncSetAttribute(k_tile_pass, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kMaxSmem);
        if (e != cudaSuccess) return (int)e;
    }
    const int64_t ctas = (args.total_items + args.teams_per_cta - 1) / args.teams_per_cta;
    if (ctas <= 0) return 0;
    if (ctas > 0x7fffffffLL) return (int)cudaErrorInvalidConfiguration;
    k_tile_pass<<<(unsigned)ctas, args.team_size * args.teams_per_cta, smem, stream>>>(args);
    count_launch();
    return (int)cudaGetLastError();
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
