// TEST INFRASTRUCTURE: CPU emulation of the tile kernel's control flow around the SAME planner
// (csrc/plan.hpp) and the SAME per-thread phase functions and step interpreter (csrc/tile_exec.cuh) that the CUDA
// kernel runs.  It lets the CPU-only test suite check plans and the interpreter against the
// oracle without a GPU.  It is built into tests/hostemu/libb200sq_hostemu.so, is never loaded by
// the squlearn_b200 package, and is not a fallback for anything.
#include <cstdint>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../squlearn_b200/csrc/plan.hpp"
#include "../../squlearn_b200/csrc/tile_exec.cuh"

using namespace b200sq;

static std::string g_err;

extern "C" {

__attribute__((visibility("default"))) const char* hostemu_last_error() { return g_err.c_str(); }

// psi_out: [n_variants][N][2^n] complex128.  Returns the number of passes, or -1 on error.
// Every pass runs exactly the phases of the CUDA kernel (build matrices | build tables | stage |
// rounds | store), thread by thread, through the shared per-thread functions of tile_exec.cuh; the
// compact inter-pass layouts live in two scratch buffers like the library's device workspace.
__attribute__((visibility("default"))) int hostemu_simulate(
    int n, int n_gates, const b200sq_gate* gates, int tile_bits, int64_t N, int d, const double* x,
    const double* table, int n_variants, const int32_t* var_gate, const double* var_shift,
    double* psi_out, int team_size, const int32_t* var_gate2, const double* var_shift2) {
    try {
        std::vector<b200sq_gate> g(gates, gates + n_gates);
        const char* lift = getenv("B200SQ_ROT_LIFT");
        Plan pl = make_plan(n, g, tile_bits, 4, true, !(lift && lift[0] == '0'));
        double2* psi = reinterpret_cast<double2*>(psi_out);
        const int TS = team_size;
        int log2ts = 0;
        while ((1 << log2ts) < TS) ++log2ts;
        if ((1 << log2ts) != TS) throw std::invalid_argument("team_size must be a power of two");
        const int64_t n_states = (int64_t)n_variants * N;
        std::vector<double2> ws[2];
        const double2* src = nullptr;
        int src_bits = 0, flip = 0;
        for (size_t ip = 0; ip < pl.passes.size(); ++ip) {
            std::vector<uint8_t> blob = build_pass_blob(pl, (int)ip);
            const PassView pv = view_pass(blob.data());
            const DPass& P = *pv.P;
            const int T = P.T;
            const uint32_t tile_elems = 1u << T;
            double2* dst;
            if (P.out_bits == n) {
                dst = psi;
            } else if (src && P.out_bits == src_bits && P.n_fresh == 0) {
                dst = const_cast<double2*>(src);  // same layout: in place
            } else {
                ws[flip].assign((size_t)n_states << P.out_bits, double2{0.0, 0.0});
                dst = ws[flip].data();
                flip ^= 1;
            }
            if (P.has_src && P.in_bits != src_bits) throw std::logic_error("layout chain broken");
            std::vector<double2> tile(tile_elems), mats(P.mat_elems + 16), scratch(kScratchElems);
            const int J = std::max<int>(1, (int)(tile_elems >> log2ts));
            std::vector<uint32_t> jt(4 * J);
            TeamMem tm{tile.data(), mats.data(), scratch.data(), scratch.data() + 32, scratch.data() + 96,
                       jt.data(), J};
            std::vector<ThreadMap> maps(TS);
            for (int tid = 0; tid < TS; ++tid) setup_thread(pv, tm, tid, TS, log2ts, maps[tid]);
            int64_t prev_state = -1;
            for (int v = 0; v < n_variants; ++v)
                for (int64_t s = 0; s < N; ++s)
                    for (uint32_t o = 0; o < (1u << P.n_obits); ++o) {
                        ItemCtx ic;
                        ic.gates = pl.gates.data();
                        ic.xrow = x ? x + s * d : nullptr;
                        ic.table = table;
                        ic.n_total = N;
                        ic.sample = s;
                        ic.vgate = var_gate ? var_gate[v] : -1;
                        ic.vshift = var_gate ? var_shift[v] : 0.0;
                        ic.vgate2 = var_gate2 ? var_gate2[v] : -1;
                        ic.vshift2 = var_gate2 ? var_shift2[v] : 0.0;
                        uint32_t obase_in, obase_out;
                        const bool zero = item_outer(P, o, ic.ext_base, obase_in, obase_out);
                        const int64_t st = (int64_t)v * N + s;
                        double2* dstate = dst + (st << P.out_bits);
                        const double2* sstate = P.has_src ? src + (st << P.in_bits) : nullptr;
                        if (zero) {
                            std::fill(tile.begin(), tile.end(), double2{0.0, 0.0});
                        } else {
                            const bool new_sample = st != prev_state;  // like the kernel: matrices once per state
                            prev_state = st;
                            if (new_sample)
                                for (int tid = 0; tid < TS; ++tid) build_mats_thread(pv, tm, ic, tid, TS);
                            for (int tid = 0; tid < TS; ++tid) build_tables_thread(pv, tm, ic, tid, TS, new_sample);
                            for (int tid = 0; tid < TS; ++tid) stage_thread(pv, tm, sstate, obase_in, maps[tid], tid, TS);
                            for (int r = 0; r < P.n_rounds; ++r)
                                for (int tid = 0; tid < TS; ++tid)  // rounds are barrier-separated
                                    exec_round_thread(tile.data(), &pv.rounds[r], pv.steps, mats.data(),
                                                      ic.ext_base, T, tid, TS);
                        }
                        for (int tid = 0; tid < TS; ++tid) store_thread(pv, tm, dstate, obase_out, maps[tid], tid, TS);
                    }
            src = dst;
            src_bits = P.out_bits;
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
