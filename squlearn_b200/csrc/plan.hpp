// Gate-list -> pass / round / step plan for the tile kernel (host-only C++, no CUDA).
//
// A PASS sweeps every state once: each thread team stages a TILE of 2^T amplitudes (the
// amplitudes that differ only in the T "tile qubits") in shared memory, applies every gate the
// tile can absorb, and writes the tile back.  Inside a pass, gates are grouped into ROUNDS: in a
// round every thread holds the 2^nb (nb <= 3) amplitudes spanned by the round's register bits
// and applies the round's STEPS to them in registers; rounds are separated by one team barrier.
// A gate can be absorbed by a pass when all qubits it acts on NON-diagonally are tile qubits
// (controls and diagonal gates only read index bits, so they may sit anywhere).
//
// ACTIVE QUBITS.  A qubit no gate has acted on non-diagonally is still |0>, so the state is
// psi_active (x) |0..0>.  Between passes only psi_active is stored: the "compact layout" of an
// active set A keeps amplitude bits in ascending qubit order over A (2^|A| amplitudes per
// state).  A pass reads layout a_in and writes layout a_out = a_in | tile qubits (the last pass
// writes the full 2^n layout).  Tile qubits that are not active yet are FRESH: the 1-qubit gates
// that hit them before anything else are folded into 2-vectors v_q, and the tile is staged as
//     tile[l] = src[l restricted to a_in] * prod_{q fresh} v_q[bit_q(l)]
// (two small product tables in shared memory, one or two complex multiplies per amplitude).  For
// the first pass (a_in empty) this is the product-state start; for ChebyshevPQC(16,2) it means the
// first pass touches 2^11 amplitudes per state and the second pass never READS the full state.
//
// STEPS.  Register work is coarse so that the per-step decode cost is amortised:
//   L_RX / L_REAL / L_GEN  one 2x2 gate of one family on each register slot of a slot mask
//   DTAB                   multiply by a precomputed phase table: a run of commuting diagonal
//                          gates (rz, p, cz, crz, rzz, ...) on one group of tile bits collapses to
//                          a single complex multiply per amplitude
//   DIAG / U1CTRL / U2     single diagonal gate, controlled 2x2, 4x4
//
// Index spaces:  "local" = position inside the tile (0..T-1, tile qubits in ascending order);
// "ext" = local bits followed by the remaining ("outer") qubits in ascending order, so
// ext = local | outer_index << T.  All masks / bit positions in steps are ext positions.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "plan_dev.h"

namespace b200sq {

struct Step {        // host-side step (the device gets the packed DStep below)
    int32_t kind;    // StepKind
    int32_t gate;    // gate index (single-gate steps), -1 otherwise
    int32_t r0, r1;
    uint32_t ctrl;   // U1CTRL/U2: control mask (ext positions). DIAG: skip mask (bit j: phase j == 1)
    int32_t mat;     // offset in the team's matrix area (double2 units)
    int32_t swap;    // U2: 1 when the gate's first qubit sits on r1 (matrix must be permuted)
    int32_t pat;     // DIAG: (ra + 1) | (rb + 1) << 2, ra / rb = register slot of the 1st / 2nd qubit or -1
};
struct MatRec {      // one gate matrix to build per (sample, tile)
    int32_t gate, kind, swap, mat;
};
struct TabGate {     // a diagonal gate folded into a phase table
    int32_t m4;      // offset of its four phases (built as a MAT_DIAG MatRec)
    int32_t p0, p1;  // ext positions of its qubits (p1 = -1 for one-qubit gates)
};
struct Tab {
    int32_t mat, lo, nbits, g0, g1;  // table offset, first local bit, width, TabGate range
    int32_t vb[2];                   // outer ext positions the table reads (-1: none); > 2 of them: legacy = 1
    int32_t legacy;                  // rebuilt per tile (reads more than two outer bits)
    int32_t carrier;                 // carries the rotation scale of the pass
};
struct Round {
    int32_t nb;      // number of register bits (0..3)
    int32_t rb[3];   // their tile-local positions, ascending
    int32_t s0, s1;  // step range
};

struct Pass {
    int32_t T;               // tile bits
    uint32_t L;              // tile qubits
    uint32_t a_in, a_out;    // active-qubit sets of the input / output layout
    int32_t last;
    int32_t r0, r1;          // round range
    int32_t s0, s1;          // step range
    int32_t m0, m1;          // MatRec range
    int32_t t0, t1;          // Tab range
    int32_t i0, i1;          // folded (init) gate range in Plan::init_gates
    int32_t mat_elems;       // double2 elements of matrix + table storage per team
    int32_t rot_carrier;     // 0 none, 1 low product table, 2 a phase table
    int32_t tile_qubits[32]; // the T tile qubits (ascending)
    uint8_t extq[32];        // ext position -> qubit
};

struct Plan {
    int n = 0;
    int T = 0;
    std::vector<b200sq_gate> gates;
    std::vector<Pass> passes;
    std::vector<Round> rounds;
    std::vector<Step> steps;
    std::vector<MatRec> mats;
    std::vector<Tab> tabs;
    std::vector<TabGate> tabgates;
    std::vector<int32_t> init_gates;  // folded 1-qubit gates, grouped per pass (Pass::i0..i1)
};

namespace detail {

struct GMask {
    uint32_t all = 0;  // every qubit of the gate
    uint32_t tgt = 0;  // qubits acted on non-diagonally
};

inline GMask gate_masks(const b200sq_gate& g, int n) {
    GateInfo gi = gate_info(g.kind);
    if (gi.nq == 0) throw std::invalid_argument("unknown gate kind " + std::to_string(g.kind));
    GMask m;
    for (int k = 0; k < gi.nq; ++k) {
        int q = g.q[k];
        if (q < 0 || q >= n) throw std::invalid_argument("gate qubit out of range");
        if (m.all & (1u << q)) throw std::invalid_argument("gate uses a qubit twice");
        m.all |= 1u << q;
        if (!gi.diag && k >= gi.nctrl) m.tgt |= 1u << q;
    }
    return m;
}

inline int popc(uint32_t v) { return __builtin_popcount(v); }
inline int rank_in(uint32_t set, int q) { return popc(set & ((1u << q) - 1u)); }

}  // namespace detail

// low_bits: number of lowest qubits every tile must contain (keeps global accesses in
// contiguous runs of 2^low_bits * 16 bytes).
inline Plan make_plan(int n, const std::vector<b200sq_gate>& gates_in, int tile_bits,
                      int low_bits = 4, bool fold_prefix = true, bool lift_rotations = false) {
    using namespace detail;
    if (n < 1 || n > B200SQ_MAX_QUBITS) throw std::invalid_argument("n_qubits out of range");
    Plan plan;
    plan.n = n;
    // default tile: the whole state when it fits 64 KiB (n <= 12); otherwise 2^11 amplitudes so that
    // two tile buffers (double buffering) still leave three CTAs per SM
    int T = tile_bits <= 0 ? (n <= 12 ? 12 : 11) : tile_bits;
    if (T > kMaxTileBits) throw std::invalid_argument("tile_bits must be <= 13");
    T = std::min(T, n);
    plan.T = T;
    plan.gates = gates_in;
    const int G = (int)gates_in.size();
    if (G > 65535) throw std::invalid_argument("at most 65535 gates per program");
    std::vector<GMask> gm(G);
    for (int i = 0; i < G; ++i) gm[i] = gate_masks(gates_in[i], n);
    // leave room for two freely chosen target qubits; with T == n the tile is the whole state
    low_bits = (T == n) ? T : std::max(0, std::min(low_bits, T - 2));
    const uint32_t mandatory = (low_bits >= 32 ? ~0u : (1u << low_bits) - 1u);
    const uint32_t full = (n >= 32 ? ~0u : (1u << n) - 1u);

    std::vector<int> pending;
    for (int i = 0; i < G; ++i)
        if (gates_in[i].kind != B200SQ_ID) pending.push_back(i);

    uint32_t active = 0;
    bool first = true;
    // A circuit without gates still needs one pass to write |0..0>.
    while (!pending.empty() || first) {
        // ---- choose the tile qubit set by first demand
        uint32_t L = mandatory;
        {
            uint32_t blockedT = 0, blockedD = 0;
            for (int gi : pending) {
                const GMask& m = gm[gi];
                bool free_ = !(m.tgt & blockedT) && !(m.all & blockedD);
                if (free_ && popc(L | m.tgt) <= T) {
                    L |= m.tgt;
                } else {
                    blockedT |= m.all;
                    blockedD |= m.tgt;
                }
            }
            // fill up: prefer qubits that are already active (keeps the layouts small), then the rest
            for (int q = 0; q < n && popc(L) < T; ++q) if (active >> q & 1u) L |= 1u << q;
            for (int q = 0; q < n && popc(L) < T; ++q) L |= 1u << q;
        }
        Pass ps;
        std::memset(&ps, 0, sizeof(ps));
        ps.T = T;
        ps.L = L;
        ps.a_in = active;
        const uint32_t fresh = L & ~active;

        // ---- fold the 1-qubit prefix of the fresh tile qubits into their start vectors
        ps.i0 = (int)plan.init_gates.size();
        if (fold_prefix) {
            std::vector<int> kept;
            uint32_t touched = 0;
            for (int gi : pending) {
                GateInfo info = gate_info(gates_in[gi].kind);
                if (info.nq == 1 && (gm[gi].all & fresh) && !(gm[gi].all & touched) &&
                    (int)plan.init_gates.size() - ps.i0 < kMaxInit) {
                    plan.init_gates.push_back(gi);
                } else {
                    touched |= gm[gi].all;
                    kept.push_back(gi);
                }
            }
            pending.swap(kept);
        }
        ps.i1 = (int)plan.init_gates.size();
        active |= L;

        // ---- absorb gates
        std::vector<int> absorbed, rest;
        {
            uint32_t blockedT = 0, blockedD = 0;
            for (int gi : pending) {
                const GMask& m = gm[gi];
                bool free_ = !(m.tgt & blockedT) && !(m.all & blockedD);
                if (free_ && (m.tgt & ~L) == 0 && (int)absorbed.size() < kMaxPassSteps - 8) {
                    absorbed.push_back(gi);
                } else {
                    rest.push_back(gi);
                    blockedT |= m.all;
                    blockedD |= m.tgt;
                }
            }
        }
        if (absorbed.empty() && !pending.empty() && ps.i1 == ps.i0)
            throw std::runtime_error("planner made no progress (gate needs more target qubits than a tile)");

        // ---- ext positions
        int pos[32];
        {
            int c = 0;
            for (int q = 0; q < n; ++q) if (L >> q & 1u) pos[q] = c++;
            for (int q = 0; q < n; ++q) if (!(L >> q & 1u)) pos[q] = c++;
            c = 0;
            for (int q = 0; q < n; ++q) if (L >> q & 1u) ps.tile_qubits[c++] = q;
            for (int q = 0; q < n; ++q) ps.extq[pos[q]] = (uint8_t)q;
        }
        ps.r0 = (int)plan.rounds.size();
        ps.s0 = (int)plan.steps.size();
        ps.m0 = (int)plan.mats.size();
        ps.t0 = (int)plan.tabs.size();
        int mat_off = 0;

        // ---- step emission helpers
        auto add_mat = [&](int gate, int kind, int swap, int elems) {
            MatRec mr{gate, kind, swap, mat_off};
            plan.mats.push_back(mr);
            mat_off += elems;
            return mr.mat;
        };
        auto diag_skip = [&](int kind) -> uint32_t {  // phases that are exactly 1: index = bit(q0) | bit(q1) << 1
            switch (kind) {
                case B200SQ_Z: case B200SQ_S: case B200SQ_SDG: case B200SQ_T: case B200SQ_TDG:
                case B200SQ_P: return 0x5;                                  // bit(q0) = 0
                case B200SQ_CZ: case B200SQ_CS: case B200SQ_CP: return 0x7;
                case B200SQ_CRZ: return 0x5;                                // control = 0
                default: return 0;
            }
        };
        // Emit one round's ordered member list as steps.
        auto emit_members = [&](const std::vector<int>& members, const Round& rd) {
            auto regbit = [&](int q) {
                for (int c = 0; c < rd.nb; ++c) if (rd.rb[c] == pos[q]) return c;
                throw std::logic_error("target qubit is not a register bit");
            };
            size_t i = 0;
            int layer_step = -1;  // index of the open L_* step
            while (i < members.size()) {
                const int gi = members[i];
                const b200sq_gate& g = gates_in[gi];
                const GateInfo info = gate_info(g.kind);
                if (info.diag) {
                    layer_step = -1;
                    size_t j = i;
                    while (j < members.size() && gate_info(gates_in[members[j]].kind).diag) ++j;
                    std::vector<int> run(members.begin() + i, members.begin() + j);
                    i = j;
                    // --- phase tables over one or two groups of local bits
                    std::vector<int> stray;
                    if (run.size() >= 3) {
                        int split = T;  // single group
                        if (T > 7) {
                            int best = -1, best_cross = 1 << 30;
                            for (int s = std::max(1, T - 7); s <= std::min(7, T - 1); ++s) {
                                int cross = 0;
                                for (int r : run) {
                                    bool lo = false, hi = false;
                                    const GateInfo ri = gate_info(gates_in[r].kind);
                                    for (int k = 0; k < ri.nq; ++k) {
                                        int p = pos[gates_in[r].q[k]];
                                        if (p < s) lo = true; else if (p < T) hi = true;
                                    }
                                    cross += lo && hi;
                                }
                                if (cross < best_cross) { best_cross = cross; best = s; }
                            }
                            split = best;
                        }
                        std::vector<int> grp[2];
                        for (int r : run) {
                            bool lo = false, hi = false;
                            const GateInfo ri = gate_info(gates_in[r].kind);
                            for (int k = 0; k < ri.nq; ++k) {
                                int p = pos[gates_in[r].q[k]];
                                if (p < split) lo = true; else if (p < T) hi = true;
                            }
                            if (lo && hi) stray.push_back(r);
                            else grp[hi ? 1 : 0].push_back(r);
                        }
                        for (int gidx = 0; gidx < 2; ++gidx) {
                            if (grp[gidx].size() < 2 || (int)plan.mats.size() - ps.m0 + (int)grp[gidx].size() > kMaxPassMats) {
                                stray.insert(stray.end(), grp[gidx].begin(), grp[gidx].end());
                                continue;
                            }
                            Tab tb;
                            tb.lo = gidx == 0 ? 0 : split;
                            tb.nbits = gidx == 0 ? std::min(split, T) : T - split;
                            tb.g0 = (int)plan.tabgates.size();
                            tb.vb[0] = tb.vb[1] = -1;
                            tb.legacy = 0;
                            tb.carrier = 0;
                            int nvb = 0;
                            for (int r : grp[gidx]) {
                                const GateInfo ri = gate_info(gates_in[r].kind);
                                TabGate tg;
                                tg.m4 = add_mat(r, MAT_DIAG, 0, 4);
                                tg.p0 = pos[gates_in[r].q[0]];
                                tg.p1 = ri.nq > 1 ? pos[gates_in[r].q[1]] : -1;
                                plan.tabgates.push_back(tg);
                                for (int pp : {tg.p0, tg.p1}) {   // outer index bits this table reads
                                    if (pp < T) continue;
                                    if (nvb > 0 && tb.vb[0] == pp) continue;
                                    if (nvb > 1 && tb.vb[1] == pp) continue;
                                    if (nvb < 2) tb.vb[nvb] = pp;
                                    ++nvb;
                                }
                            }
                            tb.g1 = (int)plan.tabgates.size();
                            tb.mat = mat_off;
                            // one variant per value of the (<= 2) outer bits, all built once per state; a table that
                            // reads more outer bits (or would not fit) is rebuilt per tile
                            if (nvb > 2 || mat_off + ((1 << std::min(nvb, 2)) << tb.nbits) > kMaxPassMatElems - 1024) {
                                tb.legacy = nvb > 0 ? 1 : 0;
                                tb.vb[0] = tb.vb[1] = -1;
                                nvb = 0;
                            }
                            mat_off += (1 << nvb) << tb.nbits;
                            plan.tabs.push_back(tb);
                            Step st;
                            std::memset(&st, 0, sizeof(st));
                            st.kind = ST_DTAB;
                            st.gate = -1;
                            st.r0 = tb.lo;
                            st.r1 = tb.nbits;
                            st.mat = tb.mat;
                            st.ctrl = (uint32_t)(tb.vb[0] < 0 ? kNoBit : tb.vb[0]) | ((uint32_t)(tb.vb[1] < 0 ? kNoBit : tb.vb[1]) << 8);
                            plan.steps.push_back(st);
                        }
                        std::sort(stray.begin(), stray.end());
                    } else {
                        stray = run;
                    }
                    for (int r : stray) {
                        const b200sq_gate& gg = gates_in[r];
                        const GateInfo ri = gate_info(gg.kind);
                        Step st;
                        std::memset(&st, 0, sizeof(st));
                        st.kind = ST_DIAG;
                        st.gate = r;
                        st.r0 = pos[gg.q[0]];
                        st.r1 = ri.nq > 1 ? pos[gg.q[1]] : -1;
                        st.ctrl = diag_skip(gg.kind);
                        int ra = -1, rb = -1;
                        for (int c = 0; c < rd.nb; ++c) {
                            if (rd.rb[c] == st.r0) ra = c;
                            if (st.r1 >= 0 && rd.rb[c] == st.r1) rb = c;
                        }
                        st.pat = (ra + 1) | ((rb + 1) << 2);
                        st.mat = add_mat(r, MAT_DIAG, 0, 4);
                        plan.steps.push_back(st);
                    }
                    continue;
                }
                ++i;
                if (info.nq - info.nctrl == 1) {
                    const int slot = regbit(g.q[info.nctrl]);
                    if (info.nctrl > 0) {
                        layer_step = -1;
                        Step st;
                        std::memset(&st, 0, sizeof(st));
                        st.kind = ST_U1CTRL;
                        st.gate = gi;
                        st.r0 = slot;
                        for (int k = 0; k < info.nctrl; ++k) st.ctrl |= 1u << pos[g.q[k]];
                        st.mat = add_mat(gi, MAT_U1, 0, 4);
                        plan.steps.push_back(st);
                        continue;
                    }
                    // rx / ry as scaled rotations (one FMA per component) need a carrier for the factored-out scale:
                    // see the conversion after the pass has been planned (MAT_ROT)
                    const int fam = g.kind == B200SQ_RX ? ST_LRX
                                  : g.kind == B200SQ_RY ? ST_LROTY   // provisional family of its own: pure-ry layers
                                  : (g.kind == B200SQ_H || g.kind == B200SQ_X) ? ST_LREAL
                                  : ST_LGEN;
                    const int mkind = MAT_U1;
                    if (layer_step >= 0 && plan.steps[layer_step].kind == fam &&
                        !(plan.steps[layer_step].r0 >> slot & 1)) {
                        Step& st = plan.steps[layer_step];
                        st.r0 |= 1 << slot;
                        MatRec mr{gi, mkind, 0, st.mat + 4 * slot};
                        plan.mats.push_back(mr);
                    } else {
                        Step st;
                        std::memset(&st, 0, sizeof(st));
                        st.kind = fam;
                        st.gate = -1;
                        st.r0 = 1 << slot;
                        st.mat = mat_off;
                        mat_off += 12;  // one 2x2 per register slot
                        MatRec mr{gi, mkind, 0, st.mat + 4 * slot};
                        plan.mats.push_back(mr);
                        layer_step = (int)plan.steps.size();
                        plan.steps.push_back(st);
                    }
                } else {
                    layer_step = -1;
                    Step st;
                    std::memset(&st, 0, sizeof(st));
                    st.kind = ST_U2;
                    st.gate = gi;
                    int ra = regbit(g.q[info.nctrl]), rb = regbit(g.q[info.nctrl + 1]);
                    st.r0 = std::max(ra, rb);
                    st.r1 = std::min(ra, rb);
                    st.swap = ra < rb ? 1 : 0;
                    for (int k = 0; k < info.nctrl; ++k) st.ctrl |= 1u << pos[g.q[k]];
                    st.mat = add_mat(gi, MAT_U2, st.swap, 16);
                    plan.steps.push_back(st);
                }
            }
        };

        // ---- pack absorbed gates into rounds: list scheduling with the same hop-over rule.  Free
        // diagonal gates are DEFERRED (they commute with everything that does not act non-diagonally
        // on their qubits) and flushed together right before the first gate that needs them, so that
        // whole entangling layers end up in one phase-table step.
        std::vector<char> done(absorbed.size(), 0);
        size_t n_done = 0;
        std::vector<int> lazy;  // deferred diagonal gates (gate indices), in circuit order
        uint32_t lazy_mask = 0;
        bool overflow = false;
        while (n_done < absorbed.size()) {
            if ((int)plan.rounds.size() - ps.r0 >= kMaxPassRounds - 1 ||
                (int)plan.steps.size() - ps.s0 >= kMaxPassSteps - 40 ||
                (int)plan.mats.size() - ps.m0 >= kMaxPassMats - 48 ||
                mat_off >= kMaxPassMatElems - 1024) {   // matrices + tables live in shared memory
                overflow = true;
                break;
            }
            Round rd;
            std::memset(&rd, 0, sizeof(rd));
            uint32_t bits = 0;  // local positions of this round's register bits
            uint32_t blockedT = 0, blockedD = 0;
            std::vector<int> members;
            for (size_t a = 0; a < absorbed.size(); ++a) {
                if (done[a]) continue;
                const int gi = absorbed[a];
                const GMask& m = gm[gi];
                bool free_ = !(m.tgt & blockedT) && !(m.all & blockedD);
                if (!free_) {
                    blockedT |= m.all;
                    blockedD |= m.tgt;
                    continue;
                }
                if (m.tgt == 0) {  // diagonal: defer
                    lazy.push_back(gi);
                    lazy_mask |= m.all;
                    done[a] = 1;
                    ++n_done;
                    continue;
                }
                uint32_t tl = 0;
                for (int q = 0; q < n; ++q) if (m.tgt >> q & 1u) tl |= 1u << pos[q];
                if (popc(bits | tl) <= 3) {
                    if (m.tgt & lazy_mask) {  // needs the deferred phases first: flush them all
                        members.insert(members.end(), lazy.begin(), lazy.end());
                        lazy.clear();
                        lazy_mask = 0;
                    }
                    bits |= tl;
                    members.push_back(gi);
                    done[a] = 1;
                    ++n_done;
                } else {
                    blockedT |= m.all;
                    blockedD |= m.tgt;
                }
            }
            if (n_done == absorbed.size() && !lazy.empty()) {  // end of the pass: flush what is left
                members.insert(members.end(), lazy.begin(), lazy.end());
                lazy.clear();
                lazy_mask = 0;
            }
            if (members.empty()) continue;  // only deferrals happened in this scan
            // give rounds without non-diagonal work (pure phase rounds) three low bits so that
            // each thread still handles eight amplitudes
            for (int b = 0; b < T && popc(bits) < std::min(3, T); ++b) bits |= 1u << b;
            rd.nb = popc(bits);
            {
                int c = 0;
                for (int b = 0; b < T; ++b) if (bits >> b & 1u) rd.rb[c++] = b;
            }
            rd.s0 = (int)plan.steps.size();
            emit_members(members, rd);
            rd.s1 = (int)plan.steps.size();
            plan.rounds.push_back(rd);
        }
        if (overflow) {
            // the pass is full: flush deferred phases into a last round and hand the remaining gates back
            if (!lazy.empty()) {
                Round rd;
                std::memset(&rd, 0, sizeof(rd));
                uint32_t bits = 0;
                for (int b = 0; b < T && popc(bits) < std::min(3, T); ++b) bits |= 1u << b;
                rd.nb = popc(bits);
                int c = 0;
                for (int b = 0; b < T; ++b) if (bits >> b & 1u) rd.rb[c++] = b;
                rd.s0 = (int)plan.steps.size();
                emit_members(lazy, rd);
                rd.s1 = (int)plan.steps.size();
                plan.rounds.push_back(rd);
                lazy.clear();
            }
            for (size_t a = 0; a < absorbed.size(); ++a)
                if (!done[a]) rest.push_back(absorbed[a]);
            std::sort(rest.begin(), rest.end());
        }
        ps.r1 = (int)plan.rounds.size();
        ps.s1 = (int)plan.steps.size();
        ps.m1 = (int)plan.mats.size();
        ps.t1 = (int)plan.tabs.size();
        ps.mat_elems = mat_off;
        // ---- scaled rotations: rx / ry layer steps become one-FMA-per-component steps when the pass has a carrier
        // for the product of the factored-out scales (the low product table of the fresh qubits, else a phase table
        // that is built once per state); otherwise ry layers fall back to the real-matrix family
        {
            int carrier = 0;
            if (lift_rotations) {
                if (fresh != 0) {
                    carrier = 1;
                } else {
                    for (int t = ps.t0; t < ps.t1 && !carrier; ++t)
                        if (!plan.tabs[t].legacy) { plan.tabs[t].carrier = 1; carrier = 2; }
                }
            }
            ps.rot_carrier = carrier;
            for (int si = ps.s0; si < ps.s1; ++si) {
                Step& st = plan.steps[si];
                if (st.kind != ST_LRX && st.kind != ST_LROTY) continue;
                if (!carrier) {
                    if (st.kind == ST_LROTY) st.kind = ST_LREAL;
                    continue;
                }
                if (st.kind == ST_LRX) st.kind = ST_LROTX;
                for (int m = ps.m0; m < ps.m1; ++m)
                    if (plan.mats[m].mat >= st.mat && plan.mats[m].mat < st.mat + 12 && plan.mats[m].kind == MAT_U1 &&
                        gate_info(gates_in[plan.mats[m].gate].kind).nq == 1 &&
                        (gates_in[plan.mats[m].gate].kind == B200SQ_RX || gates_in[plan.mats[m].gate].kind == B200SQ_RY))
                        plan.mats[m].kind = MAT_ROT;
            }
        }
        ps.last = rest.empty() ? 1 : 0;
        ps.a_out = ps.last ? full : active;
        if (ps.s1 - ps.s0 > kMaxPassSteps || ps.r1 - ps.r0 > kMaxPassRounds || ps.m1 - ps.m0 > kMaxPassMats)
            throw std::runtime_error("internal: pass exceeds its capacity");
        plan.passes.push_back(ps);
        pending.swap(rest);
        first = false;
    }
    return plan;
}

// ---------------------------------------------------------------------------------------------
// Pack pass `ip` into its device blob.

inline std::vector<uint8_t> build_pass_blob(const Plan& plan, int ip) {
    using namespace detail;
    const Pass& ps = plan.passes[ip];
    const int n = plan.n, T = ps.T;
    DPass h;
    std::memset(&h, 0, sizeof(h));
    h.T = T;
    h.n_rounds = ps.r1 - ps.r0;
    h.n_steps = ps.s1 - ps.s0;
    h.n_mats = ps.m1 - ps.m0;
    h.n_tabs = ps.t1 - ps.t0;
    h.n_init = ps.i1 - ps.i0;
    h.mat_elems = ps.mat_elems;
    h.has_src = ps.a_in != 0;
    h.last = (uint8_t)ps.last;
    h.rot_carrier = (uint8_t)ps.rot_carrier;
    h.in_bits = (uint8_t)popc(ps.a_in);
    h.out_bits = (uint8_t)popc(ps.a_out);
    std::memset(h.l_in, kNoBit, sizeof(h.l_in));
    std::memset(h.l_out, kNoBit, sizeof(h.l_out));
    std::memset(h.l_plo, kNoBit, sizeof(h.l_plo));
    std::memset(h.l_phi, kNoBit, sizeof(h.l_phi));
    std::memset(h.o_in, kNoBit, sizeof(h.o_in));
    std::memset(h.o_out, kNoBit, sizeof(h.o_out));
    std::memset(h.o_ext, kNoBit, sizeof(h.o_ext));
    std::memset(h.fresh_q, kNoBit, sizeof(h.fresh_q));
    const uint32_t fresh = ps.L & ~ps.a_in;
    h.n_fresh = (uint8_t)popc(fresh);
    h.nlo = (uint8_t)std::min<int>(h.n_fresh, 6);
    h.nhi = (uint8_t)(h.n_fresh - h.nlo);
    int nf = 0;
    for (int i = 0; i < T; ++i) {
        const int q = ps.tile_qubits[i];
        h.l_out[i] = (uint8_t)rank_in(ps.a_out, q);
        if (ps.a_in >> q & 1u) {
            h.l_in[i] = (uint8_t)rank_in(ps.a_in, q);
        } else {
            h.fresh_q[nf] = (uint8_t)q;
            if (nf < h.nlo) h.l_plo[i] = (uint8_t)nf;
            else h.l_phi[i] = (uint8_t)(nf - h.nlo);
            ++nf;
        }
    }
    int nob = 0, orank = 0;
    for (int q = 0; q < n; ++q) {
        if (ps.L >> q & 1u) continue;
        if (ps.a_out >> q & 1u) {
            h.o_out[nob] = (uint8_t)rank_in(ps.a_out, q);
            if (ps.a_in >> q & 1u) h.o_in[nob] = (uint8_t)rank_in(ps.a_in, q);
            h.o_ext[nob] = (uint8_t)(T + orank);
            ++nob;
        }
        ++orank;
    }
    h.n_obits = (uint8_t)nob;
    for (int i = 0; i < 32; ++i) h.extq[i] = ps.extq[i];

    std::vector<DRound> rounds;
    for (int r = ps.r0; r < ps.r1; ++r) {
        const Round& rd = plan.rounds[r];
        DRound d{};
        d.nb = (uint8_t)rd.nb;
        for (int c = 0; c < 3; ++c) d.rb[c] = (uint8_t)rd.rb[c];
        d.s0 = (uint16_t)(rd.s0 - ps.s0);
        d.s1 = (uint16_t)(rd.s1 - ps.s0);
        for (int k = 0; k < 8; ++k) {
            uint32_t ko = 0;
            for (int c = 0; c < rd.nb; ++c)
                if (k >> c & 1) ko |= 1u << rd.rb[c];
            d.ksw[k] = (uint16_t)(ko ^ ((ko >> 3) & 7u));
        }
        rounds.push_back(d);
    }
    std::vector<DStep> steps;
    for (int s = ps.s0; s < ps.s1; ++s) {
        const Step& u = plan.steps[s];
        DStep d{};
        d.kind = (uint8_t)u.kind;
        d.r0 = (uint8_t)u.r0;
        d.r1 = (int8_t)u.r1;
        d.swap = (uint8_t)u.swap;
        d.gate = (uint16_t)(u.gate < 0 ? 0xffff : u.gate);
        d.ctrl = u.ctrl;
        d.mat = (uint16_t)u.mat;
        d.pat = (uint16_t)u.pat;
        steps.push_back(d);
    }
    std::vector<DMat> mats;
    for (int m = ps.m0; m < ps.m1; ++m) {
        const MatRec& mr = plan.mats[m];
        mats.push_back(DMat{(uint16_t)mr.gate, (uint8_t)mr.kind, (uint8_t)mr.swap, (uint16_t)mr.mat, 0});
    }
    std::vector<DTab> tabs;
    std::vector<DTabGate> tabgates;
    for (int t = ps.t0; t < ps.t1; ++t) {
        const Tab& tb = plan.tabs[t];
        DTab d{};
        d.mat = (uint16_t)tb.mat;
        d.lo = (uint8_t)tb.lo;
        d.nbits = (uint8_t)tb.nbits;
        if (tb.legacy) d.nbits |= 0x80;  // reads more than two outer index bits: rebuilt per tile
        d.vb[0] = (uint8_t)(tb.vb[0] < 0 ? kNoBit : tb.vb[0]);
        d.vb[1] = (uint8_t)(tb.vb[1] < 0 ? kNoBit : tb.vb[1]);
        d.carrier = (uint8_t)tb.carrier;
        d.g0 = (uint16_t)tabgates.size();
        for (int g = tb.g0; g < tb.g1; ++g) {
            const TabGate& tg = plan.tabgates[g];
            tabgates.push_back(DTabGate{(uint16_t)tg.m4, (uint8_t)tg.p0, (uint8_t)(tg.p1 < 0 ? kNoBit : tg.p1)});
        }
        d.g1 = (uint16_t)tabgates.size();
        tabs.push_back(d);
    }
    h.n_tabgates = (int32_t)tabgates.size();
    std::vector<DInit> inits;
    for (int i = ps.i0; i < ps.i1; ++i) {
        const int gi = plan.init_gates[i];
        inits.push_back(DInit{(uint16_t)gi, (uint8_t)plan.gates[gi].q[0], 0});
    }

    auto align16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    size_t off = align16(sizeof(DPass));
    h.off_rounds = (int32_t)off; off = align16(off + rounds.size() * sizeof(DRound));
    h.off_steps = (int32_t)off; off = align16(off + steps.size() * sizeof(DStep));
    h.off_mats = (int32_t)off; off = align16(off + mats.size() * sizeof(DMat));
    h.off_tabs = (int32_t)off; off = align16(off + tabs.size() * sizeof(DTab));
    h.off_tabgates = (int32_t)off; off = align16(off + tabgates.size() * sizeof(DTabGate));
    h.off_inits = (int32_t)off; off = align16(off + inits.size() * sizeof(DInit));
    h.blob_bytes = (int32_t)off;
    std::vector<uint8_t> blob(off, 0);
    std::memcpy(blob.data(), &h, sizeof(h));
    if (!rounds.empty()) std::memcpy(blob.data() + h.off_rounds, rounds.data(), rounds.size() * sizeof(DRound));
    if (!steps.empty()) std::memcpy(blob.data() + h.off_steps, steps.data(), steps.size() * sizeof(DStep));
    if (!mats.empty()) std::memcpy(blob.data() + h.off_mats, mats.data(), mats.size() * sizeof(DMat));
    if (!tabs.empty()) std::memcpy(blob.data() + h.off_tabs, tabs.data(), tabs.size() * sizeof(DTab));
    if (!tabgates.empty()) std::memcpy(blob.data() + h.off_tabgates, tabgates.data(), tabgates.size() * sizeof(DTabGate));
    if (!inits.empty()) std::memcpy(blob.data() + h.off_inits, inits.data(), inits.size() * sizeof(DInit));
    return blob;
}

}  // namespace b200sq
