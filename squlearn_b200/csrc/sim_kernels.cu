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

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src));
}

// Persistent teams: a team walks over its work items (variant, sample, outer tile) with the grid
// stride and keeps TWO tile buffers — while tile i is being processed, tile i+1 streams in with
// 16-byte cp.async (no register staging), so HBM reads overlap the FP64 work and ~64 KiB per team
// are always in flight.
__global__ void __launch_bounds__(kCtaThreads, B200SQ_TILE_MINBLOCKS) k_tile_pass(const __grid_constant__ TileArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* smem = reinterpret_cast<double2*>(smem_raw);

    const int TS = A.team_size;
    const int team = threadIdx.x / TS;
    const int tid = threadIdx.x - team * TS;
    // contiguous chunk of work items per team (balanced to within one item; consecutive items are
    // the tiles of one state, so expensive and zero tiles of the first pass are spread evenly)
    const int64_t n_teams = (int64_t)gridDim.x * A.teams_per_cta;
    const int64_t my_team = (int64_t)blockIdx.x * A.teams_per_cta + team;
    const int64_t chunk = A.total_items / n_teams, extra = A.total_items % n_teams;
    int64_t item = my_team * chunk + (my_team < extra ? my_team : extra);
    const int64_t item_end = item + chunk + (my_team < extra ? 1 : 0);
    const int64_t stride = 1;
    if (item >= item_end) return;  // the whole team leaves together

    const DPass& P = A.pass;
    const int T = P.T;
    const uint32_t tile_elems = 1u << T;
    const uint32_t n_outer = 1u << (A.n - T);

    double2* buf0 = smem + (size_t)team * A.team_elems;
    double2* buf1 = P.first ? buf0 : buf0 + tile_elems;   // the first pass loads nothing
    double2* mats = buf1 + tile_elems;

    const int outer_bits = A.n - T;
    auto state_of = [&](int64_t it, uint32_t& outer, int64_t& sample, int& variant) {
        outer = (uint32_t)it & (n_outer - 1u);
        const int64_t rest = it >> outer_bits;
        int64_t v, smp;
        if (A.n_variants == 1) {          // the common case: no division at all
            v = 0;
            smp = rest;
        } else if (rest < 0x7fffffffLL) {  // 32-bit division
            const uint32_t r32 = (uint32_t)rest, ns = (uint32_t)A.n_samples;
            v = r32 / ns;
            smp = r32 - (uint32_t)v * ns;
        } else {
            v = rest / A.n_samples;
            smp = rest - v * A.n_samples;
        }
        sample = A.sample0 + smp;
        variant = A.variant0 + (int)v;
        return A.psi ? A.psi + (((int64_t)variant * A.n_total + sample - A.psi_off) << A.n) : nullptr;
    };
    // Global / shared addressing of tile element l = tid + j * TS:  the low log2(TS) bits of l are
    // the thread's, so their share of the scattered global index (g_lo) and of the swizzle is fixed
    // per thread; the high bits' share (ghi[j]) is the same for the whole team and sits in SMEM.
    const int J = (int)(tile_elems / (uint32_t)TS) > 0 ? (int)(tile_elems / (uint32_t)TS) : 1;
    uint32_t* ghi = reinterpret_cast<uint32_t*>(mats + P.mat_elems + 256);
    for (int j = tid; j < J; j += TS) ghi[j] = scatter_runs((uint32_t)j * (uint32_t)TS, P.n_lruns, P.lmask, P.lshift);
    const uint32_t g_lo = scatter_runs((uint32_t)tid, P.n_lruns, P.lmask, P.lshift);
    const bool wide = TS >= 64;  // then (l >> 3) & 7 does not depend on j
    const uint32_t sw_tid = tile_sw((uint32_t)tid);
    auto smem_idx = [&](int j) { return wide ? sw_tid + (uint32_t)j * TS : tile_sw((uint32_t)tid + (uint32_t)j * TS); };
    const bool in_range = (uint32_t)tid < tile_elems;  // tiles smaller than a warp (n < 5)
    team_sync(TS, team);
    auto prefetch = [&](int64_t it, double2* dst) {
        uint32_t o;
        int64_t smp;
        int var;
        const double2* st = state_of(it, o, smp, var);
        const double2* src = st + (scatter_runs(o, P.n_oruns, P.omask, P.oshift) | g_lo);
        if (in_range)
            for (int j = 0; j < J; ++j) cp_async16(dst + smem_idx(j), src + ghi[j]);
    };

    if (!P.first) prefetch(item, buf0);
    asm volatile("cp.async.commit_group;" ::: "memory");

    for (int iter = 0; item < item_end; item += stride, ++iter) {
        double2* tile = (iter & 1) ? buf1 : buf0;
        if (!P.first && item + stride < item_end) prefetch(item + stride, (iter & 1) ? buf0 : buf1);
        asm volatile("cp.async.commit_group;" ::: "memory");

        uint32_t outer;
        int64_t sample;
        int variant;
        double2* state = state_of(item, outer, sample, variant);
        const double* xrow = A.x ? A.x + sample * A.d : nullptr;
        const int vgate = A.var_gate ? A.var_gate[variant] : -1;
        const double vshift = A.var_gate ? A.var_shift[variant] : 0.0;
        const uint32_t obase = scatter_runs(outer, P.n_oruns, P.omask, P.oshift);

        // ---- first pass: a tile whose outer qubits are not all |0> is identically zero
        // (outer qubits are never folded into the product start)
        if (P.first && outer != 0) {
            if (A.store && in_range)
                for (int j = 0; j < J; ++j) __stcs(state + (obase | g_lo | ghi[j]), double2{0.0, 0.0});
            continue;
        }

        // ---- per-sample gate matrices (the previous item's rounds are behind a team barrier)
        for (int ui = tid; ui < P.n_uops; ui += TS) {
            const DUop u = A.uops[ui];
            const b200sq_gate g = A.gates[u.gate];
            double theta = gate_angle(g, xrow, A.table, A.n_total, sample);
            if ((int)u.gate == vgate) theta += vshift;
            build_matrix(u, g, theta, mats + u.mat);
        }

        // ---- stage the tile
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
            for (int t = tid; t < (1 << nlo); t += TS) lo[t] = product_amp(t, nlo, P.extq, qv);
            for (int t = tid; t < (1 << nhi); t += TS) hi[t] = product_amp(t, nhi, P.extq + nlo, qv);
            team_sync(TS, team);
            for (uint32_t l = tid; l < tile_elems; l += TS)
                tile[tile_sw(l)] = cmul(lo[l & ((1u << nlo) - 1u)], hi[l >> nlo]);
        } else {
            asm volatile("cp.async.wait_group 1;" ::: "memory");  // this item's tile has landed
        }
        team_sync(TS, team);

        // ---- rounds
        const uint32_t ext_base = outer << T;
        for (int r = 0; r < P.n_rounds; ++r) {
            exec_round_thread(tile, A.rounds[r], A.uops, mats, ext_base, T, tid, TS);
            team_sync(TS, team);
        }

        // ---- epilogue
        if (A.store && in_range) {
            double2* dst = state + (obase | g_lo);
            for (int j = 0; j < J; ++j) __stcs(dst + ghi[j], tile[smem_idx(j)]);
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
        // the tile buffer, the matrices and the scratch are reused by the next items
        team_sync(TS, team);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
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

int tile_pass_smem_bytes(int T, int mat_elems, int first, int* team_size, int* teams_per_cta,
                         int* team_elems) {
    int groups = T >= 3 ? (1 << (T - 3)) : 1;
    int ts = groups < 32 ? 32 : (groups > kCtaThreads ? kCtaThreads : groups);
    // one tile (first pass: nothing is loaded) or two (double buffering) + matrices + product-start scratch
    int elems = (first ? 1 : 2) * (1 << T) + mat_elems + 64 + 64 + 128 + 16;  // + ghi table
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
    const int threads = args.team_size * args.teams_per_cta;
    int64_t ctas = (args.total_items + args.teams_per_cta - 1) / args.teams_per_cta;
    if (ctas <= 0) return 0;
    // persistent grid: as many CTAs as fit on the GPU at once; teams stride over the work items
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tile_pass, threads, smem) != cudaSuccess ||
        per_sm < 1)
        per_sm = 1;
    const int64_t resident = (int64_t)sms * per_sm;
    if (ctas > resident) ctas = resident;
    k_tile_pass<<<(unsigned)ctas, threads, smem, stream>>>(args);
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
