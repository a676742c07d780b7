// TEST INFRASTRUCTURE: CPU emulation of the tile kernel's control flow around the SAME planner
// (csrc/plan.hpp) and the SAME per-thread micro-op interpreter (csrc/uop_exec.cuh) that the CUDA
// kernel runs.  It lets the CPU-only test suite check plans and the interpreter against the
// oracle without a GPU.  It is built into tests/hostemu/libb200sq_hostemu.so, is never loaded by
// the squlearn_b200 package, and is not a fallback for anything.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../squlearn_b200/csrc/plan.hpp"
#include "../../squlearn_b200/csrc/uop_exec.cuh"

using namespace b200sq;

static std::string g_err;

extern "C" {

__attribute__((visibility("default"))) const char* hostemu_last_error() { return g_err.c_str(); }

// psi_out: [n_variants][N][2^n] complex128.  Returns the number of passes, or -1 on error.
__attribute__((visibility("default"))) int hostemu_simulate(
    int n, int n_gates, const b200sq_gate* gates, int tile_bits, int64_t N, int d, const double* x,
    const double* table, int n_variants, const int32_t* var_gate, const double* var_shift,
    double* psi_out, int team_size) {
    try {
        std::vector<b200sq_gate> g(gates, gates + n_gates);
        Plan pl = make_plan(n, g, tile_bits);
        double2* psi = reinterpret_cast<double2*>(psi_out);
        for (const Pass& P : pl.passes) {
            const int T = P.T;
            const uint32_t tile_elems = 1u << T;
            const uint32_t n_outer = 1u << (n - T);
            std::vector<double2> tile(tile_elems), mats(P.mat_elems + 16);
            const int TS = team_size;
            for (int v = 0; v < n_variants; ++v)
                for (int64_t s = 0; s < N; ++s)
                    for (uint32_t outer = 0; outer < n_outer; ++outer) {
                        const double* xrow = x ? x + s * d : nullptr;
                        const int vgate = var_gate ? var_gate[v] : -1;
                        const double vshift = var_gate ? var_shift[v] : 0.0;
                        std::vector<DUop> duops;
                        std::vector<DRound> drounds;
                        for (int ui = P.u0; ui < P.u1; ++ui) duops.push_back(pack_uop(pl.uops[ui], P.u0));
                        for (int r = P.r0; r < P.r1; ++r) drounds.push_back(pack_round(pl.rounds[r], P.u0));
                        for (int ui = 0; ui < (int)duops.size(); ++ui) {
                            const DUop& u = duops[ui];
                            const b200sq_gate& gg = pl.gates[u.gate];
                            double theta = gate_angle(gg, xrow, table, N, s);
                            if (u.gate == vgate) theta += vshift;
                            build_matrix(u, gg, theta, mats.data() + u.mat);
                        }
                        const uint32_t obase = scatter_runs(outer, P.n_oruns, P.omask, P.oshift);
                        double2* state = psi + (((int64_t)v * N + s) << n);
                        bool zero_tile = false;
                        if (P.first) {
                            std::vector<DInit> inits;
                            for (int32_t gi : pl.init_gates)
                                inits.push_back(DInit{(uint16_t)gi, (uint8_t)pl.gates[gi].q[0], 0});
                            std::vector<double2> qv(64);
                            for (int q = 0; q < n; ++q)
                                fold_qubit_vector(q, inits.data(), (int)inits.size(), pl.gates.data(), xrow, table,
                                                  N, s, vgate, vshift, qv.data() + 2 * q);
                            const int nlo = T < 6 ? T : 6, nhi = T - nlo;
                            const double2 oscal = product_amp(outer, n - T, P.extq + T, qv.data());
                            zero_tile = (oscal.x == 0.0 && oscal.y == 0.0);
                            for (uint32_t l = 0; l < tile_elems; ++l) {
                                const double2 lo = product_amp(l & ((1u << nlo) - 1u), nlo, P.extq, qv.data());
                                const double2 hi = cmul(oscal, product_amp(l >> nlo, nhi, P.extq + nlo, qv.data()));
                                tile[tile_sw(l)] = cmul(lo, hi);
                            }
                        } else {
                            for (uint32_t l = 0; l < tile_elems; ++l)
                                tile[tile_sw(l)] = state[obase | scatter_runs(l, P.n_lruns, P.lmask, P.lshift)];
                        }
                        if (!zero_tile) {
                            const uint32_t ext_base = outer << T;
                            for (size_t r = 0; r < drounds.size(); ++r)
                                for (int tid = 0; tid < TS; ++tid)  // rounds are barrier-separated
                                    exec_round_thread(tile.data(), drounds[r], duops.data(), mats.data(),
                                                      ext_base, T, tid, TS);
                        }
                        for (uint32_t l = 0; l < tile_elems; ++l)
                            state[obase | scatter_runs(l, P.n_lruns, P.lmask, P.lshift)] = tile[tile_sw(l)];
                    }
        }
        return (int)pl.passes.size();
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// out[s][t] = <psi_s| P_t |psi_s> with the kernel's pauli_term helper.
__attribute__((visibility("default"))) void hostemu_expvals(const double* psi_in, int64_t N, int n,
                                                            int n_terms, const uint32_t* xm,
                                                            const uint32_t* zm, double* out) {
    const double2* psi = reinterpret_cast<const double2*>(psi_in);
    const uint32_t dim = 1u << n;
    for (int64_t s = 0; s < N; ++s)
        for (int t = 0; t < n_terms; ++t) {
            const int ny = __builtin_popcount(xm[t] & zm[t]);
            double acc = 0.0;
            const double2* st = psi + (s << n);
            for (uint32_t l = 0; l < dim; ++l) acc += pauli_term(st[l ^ xm[t]], st[l], l, zm[t], ny);
            out[s * n_terms + t] = acc;
        }
}

}  // extern "C"
