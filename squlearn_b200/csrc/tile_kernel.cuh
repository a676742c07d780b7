// Body of the pass kernel (see sim_kernels.cu for the overview), shared by the ahead-of-time compiled
// generic kernel k_tile_pass and the run-time specialised kernels of jit.cu.  Device code only; compiles
// under nvcc and NVRTC.
#pragma once

#include "tile_exec.cuh"

namespace b200sq {

// Per-state statistics reduced by the store epilogue of the last pass (same layout as b200sq_row_stats).
struct RowStats {
    double norm2;
    uint32_t max_hi;
    uint32_t reserved;
};

struct TileArgs {
    const unsigned char* blob;   // device copy of the pass blob (plan.hpp: build_pass_blob)
    int32_t blob_bytes;
    const b200sq_gate* gates;    // device gate table (angle coefficients)
    // state (v, i) of the input / output layout lives at base + ((v * vstride + i - off) << layout bits)
    const double2* src;
    double2* dst;
    int64_t src_vstride, src_off, dst_vstride, dst_off;
    const double* x;           // [n_total][d]
    const double* table;       // [slots][n_total]
    int64_t n_total;           // samples in the whole call (row stride of table, variant stride of out)
    int64_t sample0;           // first sample of this launch
    int64_t n_samples;         // samples in this launch
    int32_t n, d;
    int32_t n_variants;        // variants in this launch
    int32_t variant0;          // first variant of this launch
    const int32_t* var_gate;   // device, indexed by absolute variant; nullptr = unshifted
    const double* var_shift;
    const int32_t* var_gate2;  // optional second shift per variant (second-order derivatives); nullptr = none
    const double* var_shift2;
    int32_t store;             // write the tile to dst
    int32_t n_terms;           // > 0: reduce the tile to expectation values (requires T == n)
    const uint32_t* xmask;     // device
    const uint32_t* zmask;
    double* out;               // [n_variants_total][n_total][n_terms]
    RowStats* row_stats;       // nullptr, or per output state (indexed like dst; zeroed by the launcher): max_hi = running
                               // maximum of the high words of max(|re|, |im|) over the stored amplitudes (atomicMax),
                               // norm2 = sum of |amplitude|^2 (atomicAdd) — the row scales the INT8 Gram's slicing
                               // needs (digit form: maximum; CRT form: 2-norm), without a sweep of its own
    int32_t team_size, log2ts, teams_per_cta, team_elems;
    int64_t total_items;
};


#if defined(__CUDACC__)

#ifndef B200SQ_TILE_MINBLOCKS
#define B200SQ_TILE_MINBLOCKS 2
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

// Persistent teams: a team walks over a contiguous chunk of work items (variant, sample, outer tile).
// The pass blob (rounds, steps, matrix / table records) is copied to shared memory once per CTA, so the
// interpreter reads it with broadcast LDS instead of indexed constant loads.
//
// Plain passes (input layout == tile bits all active) keep TWO tile buffers: while tile i is being
// processed, tile i+1 streams in with 16-byte cp.async, so HBM reads overlap the FP64 work.
// Expanding passes (fresh tile qubits) gather the few source amplitudes of the tile through L1/L2
// and multiply them with the product tables of the fresh qubits: no full-state read at all.
struct InterpretedRounds {  // every round through the step interpreter
    __device__ __forceinline__ void operator()(double2* tile, const PassView& pv, const double2* mats,
                                               uint32_t ext_base, int T, int tid, int TS, int team) const {
        for (int r = 0; r < pv.P->n_rounds; ++r) {
            exec_round_thread(tile, &pv.rounds[r], pv.steps, mats, ext_base, T, tid, TS);
            team_sync(TS, team);
        }
    }
};

template <class Rounds>
__device__ __forceinline__ void tile_pass_body(const TileArgs& A, Rounds run_rounds) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    {   // pass blob -> SMEM
        const uint4* g = reinterpret_cast<const uint4*>(A.blob);
        uint4* sdst = reinterpret_cast<uint4*>(smem_raw);
        for (int i = threadIdx.x; i < A.blob_bytes / 16; i += blockDim.x) sdst[i] = g[i];
    }
    __syncthreads();
    const PassView pv = view_pass(smem_raw);
    const DPass& P = *pv.P;
    double2* smem = reinterpret_cast<double2*>(smem_raw + A.blob_bytes);

    const int TS = A.team_size;
    const int team = threadIdx.x >> A.log2ts;
    const int tid = threadIdx.x - (team << A.log2ts);
    // contiguous chunk of work items per team (balanced to within one item; consecutive items are
    // the tiles of one state)
    const int64_t n_teams = (int64_t)gridDim.x * A.teams_per_cta;
    const int64_t my_team = (int64_t)blockIdx.x * A.teams_per_cta + team;
    const int64_t chunk = A.total_items / n_teams, extra = A.total_items % n_teams;
    int64_t item = my_team * chunk + (my_team < extra ? my_team : extra);
    const int64_t item_end = item + chunk + (my_team < extra ? 1 : 0);
    if (item >= item_end) return;  // the whole team leaves together (after the CTA-wide barrier above)

    const int T = P.T;
    const uint32_t tile_elems = 1u << T;
    const bool plain = P.has_src && P.n_fresh == 0;
    TeamMem tm;
    double2* buf0 = smem + (size_t)team * A.team_elems;
    double2* buf1 = plain ? buf0 + tile_elems : buf0;
    tm.mats = buf1 + tile_elems;
    tm.qv = tm.mats + P.mat_elems;
    tm.plo = tm.qv + 32;
    tm.phi = tm.plo + 64;
    tm.jt = reinterpret_cast<uint32_t*>(tm.phi + 128);
    tm.J = (int)(tile_elems >> A.log2ts) > 0 ? (int)(tile_elems >> A.log2ts) : 1;
    tm.tile = buf0;
    ThreadMap map;
    setup_thread(pv, tm, tid, TS, A.log2ts, map);
    const int J = tm.J;
    const bool wide = TS >= 64;  // then (l >> 3) & 7 does not depend on j
    const uint32_t sw_tid = tile_sw((uint32_t)tid);
    auto smem_idx = [&](int j) { return wide ? sw_tid + ((uint32_t)j << A.log2ts) : tile_sw((uint32_t)tid + ((uint32_t)j << A.log2ts)); };
    const bool in_range = (uint32_t)tid < tile_elems;  // tiles smaller than a team (n < 5)
    team_sync(TS, team);

    const uint32_t omask = (1u << P.n_obits) - 1u;
    auto decode = [&](int64_t it, uint32_t& o, int64_t& sample, int& variant) {
        o = (uint32_t)it & omask;
        const int64_t rest = it >> P.n_obits;
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
    };
    auto prefetch = [&](int64_t it, double2* dstbuf) {
        uint32_t o, eb, oin, oout;
        int64_t smp;
        int var;
        decode(it, o, smp, var);
        item_outer(P, o, eb, oin, oout);
        const double2* st = A.src + (((int64_t)var * A.src_vstride + smp - A.src_off) << P.in_bits);
        const double2* sp = st + (oin | map.in_lo);
        if (in_range)
            for (int j = 0; j < J; ++j) cp_async16(dstbuf + smem_idx(j), sp + tm.jt[j]);
    };

    if (plain) prefetch(item, buf0);
    asm volatile("cp.async.commit_group;" ::: "memory");

    int64_t prev_sample = -1;
    int prev_variant = -1;
    for (int iter = 0; item < item_end; ++item, ++iter) {
        tm.tile = (plain && (iter & 1)) ? buf1 : buf0;
        if (plain && item + 1 < item_end) prefetch(item + 1, (iter & 1) ? buf0 : buf1);
        asm volatile("cp.async.commit_group;" ::: "memory");

        uint32_t o, obase_in, obase_out;
        int variant;
        ItemCtx ic;
        decode(item, o, ic.sample, variant);
        const bool zero = item_outer(P, o, ic.ext_base, obase_in, obase_out);
        double2* dstate = A.dst ? A.dst + (((int64_t)variant * A.dst_vstride + ic.sample - A.dst_off) << P.out_bits) : nullptr;

        // ---- a tile with a set bit of an outer qubit that is still |0> is identically zero
        if (zero) {
            if (A.store && in_range) {
                double2* dp = dstate + (obase_out | map.out_lo);
                for (int j = 0; j < J; ++j) __stcs(dp + tm.jt[J + j], double2{0.0, 0.0});
            }
            continue;
        }
        ic.gates = A.gates;
        ic.xrow = A.x ? A.x + ic.sample * A.d : nullptr;
        ic.table = A.table;
        ic.n_total = A.n_total;
        ic.vgate = A.var_gate ? A.var_gate[variant] : -1;
        ic.vshift = A.var_gate ? A.var_shift[variant] : 0.0;
        ic.vgate2 = A.var_gate2 ? A.var_gate2[variant] : -1;
        ic.vshift2 = A.var_gate2 ? A.var_shift2[variant] : 0.0;

        // ---- phase A: per-sample gate matrices + start vectors of the fresh qubits.  Consecutive items
        // of a team are tiles of the same state, so this runs once per state, not once per tile.
        // (the previous item's rounds / epilogue are behind a team barrier)
        const bool new_sample = ic.sample != prev_sample || variant != prev_variant;
        prev_sample = ic.sample;
        prev_variant = variant;
        if (new_sample) {
            build_mats_thread(pv, tm, ic, tid, TS);
            if (P.n_tabs > 0 || P.n_fresh > 0) team_sync(TS, team);
        }
        // ---- phase B: product tables, phase tables
        if (P.n_tabs > 0 || P.n_fresh > 0) build_tables_thread(pv, tm, ic, tid, TS, new_sample);
        // ---- phase C: stage the tile
        if (plain) {
            asm volatile("cp.async.wait_group 1;" ::: "memory");  // this item's tile has landed
        } else {
            team_sync(TS, team);
            const double2* sstate = P.has_src
                ? A.src + (((int64_t)variant * A.src_vstride + ic.sample - A.src_off) << P.in_bits) : nullptr;
            stage_thread(pv, tm, sstate, obase_in, map, tid, TS);
        }
        team_sync(TS, team);

        // ---- rounds (one team barrier after each)
        run_rounds(tm.tile, pv, tm.mats, ic.ext_base, T, tid, TS, team);

        // ---- epilogue
        if (A.store && in_range) {
            double2* dp = dstate + (obase_out | map.out_lo);
            if (A.row_stats) {
                uint32_t hi = 0;
                double n2 = 0.0;
#pragma unroll 8
                for (int j = 0; j < J; ++j) {
                    const double2 v = tm.tile[smem_idx(j)];
                    __stcs(dp + tm.jt[J + j], v);
                    const uint32_t hx = (uint32_t)__double2hiint(v.x) & 0x7fffffffu;
                    const uint32_t hy = (uint32_t)__double2hiint(v.y) & 0x7fffffffu;
                    hi = max(hi, max(hx, hy));
                    n2 = fma(v.x, v.x, fma(v.y, v.y, n2));
                }
                const unsigned am = __activemask();
                hi = __reduce_max_sync(am, hi);
                RowStats* rs = A.row_stats + ((int64_t)variant * A.dst_vstride + ic.sample - A.dst_off);
                if (am == 0xffffffffu) {
                    for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
                    if ((tid & 31) == 0) {
                        atomicMax(&rs->max_hi, hi);
                        atomicAdd(&rs->norm2, n2);
                    }
                } else {   // (teams are whole warps: not reached)
                    atomicMax(&rs->max_hi, hi);
                    atomicAdd(&rs->norm2, n2);
                }
            } else {
#pragma unroll 8
                for (int j = 0; j < J; ++j) __stcs(dp + tm.jt[J + j], tm.tile[smem_idx(j)]);
            }
        }
        if (A.n_terms > 0) {
            // T == n here: the tile is the state and local index == amplitude index.
            const int nwarps = TS >> 5, warp = tid >> 5, lane = tid & 31;
            double* out = A.out + ((int64_t)variant * A.n_total + ic.sample) * A.n_terms;
            for (int t = warp; t < A.n_terms; t += nwarps) {
                const uint32_t xm = A.xmask[t], zm = A.zmask[t];
                const int ny = __popc(xm & zm);
                double acc = 0.0;
                for (uint32_t l = lane; l < tile_elems; l += 32)
                    acc += pauli_term(tm.tile[tile_sw(l ^ xm)], tm.tile[tile_sw(l)], l, zm, ny);
                acc = warp_sum(acc);
                if (lane == 0) out[t] = acc;
            }
        }
        // the tile buffer, the matrices and the scratch are reused by the next items
        team_sync(TS, team);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
}


#endif  // __CUDACC__

}  // namespace b200sq
