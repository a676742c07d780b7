// Gate-list -> pass / round / micro-op plan for the tile kernel (host-only C++, no CUDA).
//
// A PASS sweeps every state once: each thread team stages a TILE of 2^T amplitudes (the
// amplitudes that differ only in the T "tile qubits") in shared memory, applies every gate the
// tile can absorb, and writes the tile back.  Inside a pass, gates are grouped into ROUNDS: in a
// round every thread holds the 2^nb (nb <= 3) amplitudes spanned by the round's register bits
// and applies the round's MICRO-OPS to them in registers; rounds are separated by one team
// barrier.  A gate can be absorbed by a pass when all qubits it acts on NON-diagonally are tile
// qubits (controls and diagonal gates only read index bits, so they may sit anywhere).
//
// Index spaces:  "local" = position inside the tile (0..T-1, tile qubits in ascending order);
// "ext" = local bits followed by the remaining ("outer") qubits in ascending order, so
// ext = local | outer_index << T.  All masks / bit positions in micro-ops are ext positions.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200sq.h"

namespace b200sq {

// Micro-op kinds.  The U1 family differs only in how much FP64 work the 2x2 product needs:
//   U1RX   [[c, -is], [-is, c]]  (rx)            8 FP64 instructions per amplitude pair
//   U1REAL real 2x2 (ry, h, x)                   8
//   U1G    general complex 2x2 (y, sx, u)       16
//   U1CTRL general 2x2 under a control mask (cx, cy, ch, csx, crx, cry, ccx)
enum UopKind : int32_t { UOP_U1G = 0, UOP_U2 = 1, UOP_DIAG = 2, UOP_U1RX = 3, UOP_U1REAL = 4,
                         UOP_U1CTRL = 5 };

struct Uop {         // host-side micro-op (the device gets the packed DUop below)
    int32_t kind;    // UopKind
    int32_t gate;    // index into the gate table (kind, angle coefficients)
    int32_t r0, r1;  // U1*: r0 = register bit. U2: r0 > r1 register bits (r0 = matrix MSB).
                     // DIAG: ext positions of the first / second qubit (r1 = -1 for 1-qubit)
    uint32_t ctrl;   // U1CTRL/U2: control mask (ext positions). DIAG: skip mask (bit j: phase j == 1)
    int32_t mat;     // offset of this op's matrix in the team's matrix area (double2 units)
    int32_t swap;    // U2: 1 when the gate's first qubit sits on r1 (matrix must be permuted)
    int32_t pat;     // DIAG: (ra + 1) | (rb + 1) << 2, ra / rb = register slot of the 1st / 2nd qubit or -1
};

struct Round {       // host-side round
    int32_t nb;      // number of register bits (0..3)
    int32_t rb[3];   // their tile-local positions, ascending
    int32_t u0, u1;  // micro-op range
};

// Packed forms that travel to the device inside the kernel parameters (constant bank).
struct DUop {        // 16 bytes
    uint8_t kind, r0, swap, pad0;
    int8_t r1, pad1;
    uint16_t gate;
    uint32_t ctrl;
    uint16_t mat;
    uint16_t pat;
};
struct DRound {      // 24 bytes
    uint8_t nb, rb[3];
    uint16_t u0, u1;  // micro-op range relative to the pass
    uint16_t ksw[8];  // swizzled SMEM offset of register slot k: ko[k] ^ ((ko[k] >> 3) & 7)
};
struct DPass {       // device view of a pass
    int32_t T, first, n_rounds, n_uops, mat_elems, n_lruns, n_oruns;
    uint32_t lmask[8];
    int32_t lshift[8];
    uint32_t omask[8];
    int32_t oshift[8];
    uint8_t extq[32];  // ext position -> qubit (tile qubits first, then outer qubits)
};
struct DInit {       // a 1-qubit gate folded into the product-state start of the first pass
    uint16_t gate;
    uint8_t qubit, pad;
};
constexpr int kMaxPassUops = 96;     // per pass, so that a pass fits the 4 KiB kernel-parameter space
constexpr int kMaxPassRounds = 48;
constexpr int kMaxInit = 96;

inline DUop pack_uop(const Uop& u, int pass_u0) {
    (void)pass_u0;
    DUop d{};
    d.kind = (uint8_t)u.kind;
    d.r0 = (uint8_t)u.r0;
    d.r1 = (int8_t)u.r1;
    d.swap = (uint8_t)u.swap;
    d.gate = (uint16_t)u.gate;
    d.ctrl = u.ctrl;
    d.mat = (uint16_t)u.mat;
    d.pat = (uint16_t)u.pat;
    return d;
}
inline DRound pack_round(const Round& r, int pass_u0) {
    DRound d{};
    d.nb = (uint8_t)r.nb;
    d.rb[0] = (uint8_t)r.rb[0];
    d.rb[1] = (uint8_t)r.rb[1];
    d.rb[2] = (uint8_t)r.rb[2];
    d.u0 = (uint16_t)(r.u0 - pass_u0);
    d.u1 = (uint16_t)(r.u1 - pass_u0);
    for (int k = 0; k < 8; ++k) {
        uint32_t ko = 0;
        for (int c = 0; c < r.nb; ++c)
            if (k >> c & 1) ko |= 1u << r.rb[c];
        d.ksw[k] = (uint16_t)(ko ^ ((ko >> 3) & 7u));
    }
    return d;
}

constexpr int kMaxRuns = 8;

struct Pass {
    int32_t T;               // tile bits
    int32_t first;           // 1: tile starts as |0..0> (no load)
    int32_t r0, r1;          // round range
    int32_t u0, u1;          // micro-op range
    int32_t mat_elems;       // double2 elements of matrix storage per team
    int32_t n_lruns, n_oruns;
    uint32_t lmask[kMaxRuns];  // local-index bit runs -> global amplitude index
    int32_t lshift[kMaxRuns];
    uint32_t omask[kMaxRuns];  // outer-index bit runs -> global amplitude index
    int32_t oshift[kMaxRuns];
    int32_t tile_qubits[32];   // the T tile qubits (ascending), for introspection / tests
    uint8_t extq[32];          // ext position -> qubit
};

struct GateInfo {
    int nq;          // number of qubits
    int nctrl;       // leading control qubits
    bool diag;       // acts diagonally on all its qubits
    bool has_angle;
};

#if defined(__CUDACC__)
__host__ __device__
#endif
inline GateInfo gate_info(int kind) {
    switch (kind) {
        case B200SQ_ID: return {1, 0, true, false};
        case B200SQ_H: case B200SQ_X: case B200SQ_Y: case B200SQ_SX: return {1, 0, false, false};
        case B200SQ_Z: case B200SQ_S: case B200SQ_SDG: case B200SQ_T: case B200SQ_TDG:
            return {1, 0, true, false};
        case B200SQ_RX: case B200SQ_RY: case B200SQ_U: return {1, 0, false, true};
        case B200SQ_RZ: case B200SQ_P: return {1, 0, true, true};
        case B200SQ_CX: case B200SQ_CY: case B200SQ_CH: case B200SQ_CSX: return {2, 1, false, false};
        case B200SQ_CZ: case B200SQ_CS: return {2, 1, true, false};
        case B200SQ_CP: case B200SQ_CRZ: return {2, 1, true, true};
        case B200SQ_CRX: case B200SQ_CRY: return {2, 1, false, true};
        case B200SQ_SWAP: case B200SQ_ISWAP: case B200SQ_ECR: return {2, 0, false, false};
        case B200SQ_RXX: case B200SQ_RYY: case B200SQ_RZX: return {2, 0, false, true};
        case B200SQ_RZZ: return {2, 0, true, true};
        case B200SQ_CCX: return {3, 2, false, false};
        case B200SQ_CSWAP: return {3, 1, false, false};
        default: return {0, 0, false, false};  // unknown kind (rejected by gate_masks)
    }
}

struct Plan {
    int n = 0;
    int T = 0;
    std::vector<b200sq_gate> gates;
    std::vector<Pass> passes;
    std::vector<Round> rounds;
    std::vector<Uop> uops;
    // Product-state prefix: 1-qubit gates that act on a qubit before any multi-qubit gate touches
    // it.  U_k ... U_1 |0> stays a tensor product, so the first pass starts from
    // (x)_q v_q with v_q = (folded gates of q)|0> instead of applying these gates amplitude by
    // amplitude (circuit order is preserved per qubit).
    std::vector<int32_t> init_gates;
    uint32_t first_tile_mask = 0;  // the first pass must use exactly this tile (folded qubits are local)
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

// Build the bit-run tables that scatter a compact index over the set bits of `qmask`.
inline int make_runs(uint32_t qmask, uint32_t* masks, int32_t* shifts) {
    int nruns = 0, compact = 0, q = 0;
    while (q < 32) {
        if (!(qmask >> q & 1u)) { ++q; continue; }
        int start = q, len = 0;
        while (q < 32 && (qmask >> q & 1u)) { ++q; ++len; }
        if (nruns == kMaxRuns) throw std::runtime_error("tile qubit set has too many runs");
        masks[nruns] = ((len >= 32 ? 0u : (1u << len)) - 1u) << compact;
        shifts[nruns] = start - compact;
        ++nruns;
        compact += len;
    }
    return nruns;
}

}  // namespace detail

// low_bits: number of lowest qubits every tile must contain (keeps global accesses in
// contiguous runs of 2^low_bits * 16 bytes).
inline Plan make_plan(int n, const std::vector<b200sq_gate>& gates_in, int tile_bits,
                      int low_bits = 4, bool fold_prefix = true) {
    using namespace detail;
    if (n < 1 || n > B200SQ_MAX_QUBITS) throw std::invalid_argument("n_qubits out of range");
    Plan plan;
    plan.n = n;
    // default tile: the whole state when it fits 64 KiB (n <= 12); otherwise 2^11 amplitudes so that
    // two tile buffers (double buffering) still leave three CTAs per SM
    int T = tile_bits <= 0 ? (n <= 12 ? 12 : 11) : tile_bits;
    if (T > 13) throw std::invalid_argument("tile_bits must be <= 13");
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

    // Gates of the first pass' tile qubits that can be folded into the product-state start.  Only
    // TILE qubits are folded: an outer qubit that stays |0> keeps 2^(n-T) - 1 of every 2^(n-T) tiles
    // identically zero during the first pass (they are skipped), which is worth more than folding.
    auto first_tile = [&](const std::vector<int>& gates_all) {
        uint32_t L = mandatory, blockedT = 0, blockedD = 0;
        for (int gi : gates_all) {
            const GMask& m = gm[gi];
            bool free_ = !(m.tgt & blockedT) && !(m.all & blockedD);
            if (free_ && popc(L | m.tgt) <= T) {
                L |= m.tgt;
            } else {
                blockedT |= m.all;
                blockedD |= m.tgt;
            }
        }
        for (int q = 0; q < n && popc(L) < T; ++q) L |= 1u << q;
        return L;
    };
    std::vector<int> pending;
    {
        std::vector<int> all;
        for (int i = 0; i < G; ++i)
            if (gates_in[i].kind != B200SQ_ID) all.push_back(i);
        const uint32_t L0 = first_tile(all);
        uint32_t entangled = 0;  // qubits already touched by a multi-qubit gate (or a kept gate)
        for (int i : all) {
            GateInfo gi = gate_info(gates_in[i].kind);
            if (fold_prefix && gi.nq == 1 && !(entangled & gm[i].all) && (gm[i].all & L0) &&
                (int)plan.init_gates.size() < kMaxInit) {
                plan.init_gates.push_back(i);
                continue;
            }
            entangled |= gm[i].all;
            pending.push_back(i);
        }
        plan.first_tile_mask = L0;
    }
    bool first = true;
    // A circuit without gates still needs one pass to write |0..0>.
    while (!pending.empty() || first) {
        // ---- choose the tile qubit set by first demand
        uint32_t L = mandatory;
        if (first) {
            L = plan.first_tile_mask;
        } else {
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
            for (int q = 0; q < n && popc(L) < T; ++q) L |= 1u << q;
        }
        // ---- absorb gates
        std::vector<int> absorbed, rest;
        {
            uint32_t blockedT = 0, blockedD = 0;
            for (int gi : pending) {
                const GMask& m = gm[gi];
                bool free_ = !(m.tgt & blockedT) && !(m.all & blockedD);
                if (free_ && (m.tgt & ~L) == 0 && (int)absorbed.size() < kMaxPassUops) {
                    absorbed.push_back(gi);
                } else {
                    rest.push_back(gi);
                    blockedT |= m.all;
                    blockedD |= m.tgt;
                }
            }
        }
        if (absorbed.empty() && !pending.empty() && !first)  // the first pass may only write the product start
            throw std::runtime_error("planner made no progress (gate needs more target qubits than a tile)");

        // ---- ext positions
        int pos[32];
        {
            int c = 0;
            for (int q = 0; q < n; ++q) if (L >> q & 1u) pos[q] = c++;
            for (int q = 0; q < n; ++q) if (!(L >> q & 1u)) pos[q] = c++;
        }
        Pass ps;
        std::memset(&ps, 0, sizeof(ps));
        ps.T = T;
        ps.first = first ? 1 : 0;
        ps.n_lruns = make_runs(L, ps.lmask, ps.lshift);
        uint32_t full = (n >= 32 ? ~0u : (1u << n) - 1u);
        ps.n_oruns = make_runs(full & ~L, ps.omask, ps.oshift);
        {
            int c = 0;
            for (int q = 0; q < n; ++q) if (L >> q & 1u) ps.tile_qubits[c++] = q;
            for (int q = 0; q < n; ++q) ps.extq[pos[q]] = (uint8_t)q;
        }
        ps.r0 = (int)plan.rounds.size();
        ps.u0 = (int)plan.uops.size();

        // ---- pack absorbed gates into rounds (list scheduling with the same hop-over rule)
        std::vector<char> done(absorbed.size(), 0);
        size_t n_done = 0;
        int mat_off = 0;
        while (n_done < absorbed.size()) {
            if ((int)plan.rounds.size() - ps.r0 >= kMaxPassRounds) {
                // the pass is full: hand the remaining gates back (gate order == circuit order)
                for (size_t a = 0; a < absorbed.size(); ++a)
                    if (!done[a]) rest.push_back(absorbed[a]);
                std::sort(rest.begin(), rest.end());
                break;
            }
            Round rd;
            std::memset(&rd, 0, sizeof(rd));
            rd.u0 = (int)plan.uops.size();
            uint32_t bits = 0;  // local positions of this round's register bits
            uint32_t blockedT = 0, blockedD = 0;
            std::vector<int> members;
            for (size_t a = 0; a < absorbed.size(); ++a) {
                if (done[a]) continue;
                const GMask& m = gm[absorbed[a]];
                uint32_t tl = 0;
                for (int q = 0; q < n; ++q) if (m.tgt >> q & 1u) tl |= 1u << pos[q];
                bool free_ = !(m.tgt & blockedT) && !(m.all & blockedD);
                if (free_ && popc(bits | tl) <= 3) {
                    bits |= tl;
                    members.push_back((int)a);
                    done[a] = 1;
                    ++n_done;
                } else {
                    blockedT |= m.all;
                    blockedD |= m.tgt;
                }
            }
            // give rounds without non-diagonal work (pure phase rounds) three low bits so that
            // each thread still handles eight amplitudes
            for (int b = 0; b < T && popc(bits) < std::min(3, T); ++b) bits |= 1u << b;
            rd.nb = popc(bits);
            {
                int c = 0;
                for (int b = 0; b < T; ++b) if (bits >> b & 1u) rd.rb[c++] = b;
            }
            auto regbit = [&](int q) {
                for (int c = 0; c < rd.nb; ++c) if (rd.rb[c] == pos[q]) return c;
                throw std::logic_error("target qubit is not a register bit");
            };
            for (int a : members) {
                int gi = absorbed[a];
                const b200sq_gate& g = gates_in[gi];
                GateInfo info = gate_info(g.kind);
                Uop u;
                std::memset(&u, 0, sizeof(u));
                u.gate = gi;
                u.mat = mat_off;
                if (info.diag) {
                    u.kind = UOP_DIAG;
                    u.r0 = pos[g.q[0]];
                    u.r1 = info.nq > 1 ? pos[g.q[1]] : -1;
                    // phases that are exactly 1: index = bit(q0) | bit(q1) << 1
                    switch (g.kind) {
                        case B200SQ_Z: case B200SQ_S: case B200SQ_SDG: case B200SQ_T:
                        case B200SQ_TDG: case B200SQ_P: u.ctrl = 0x5; break;        // bit(q0)=0
                        case B200SQ_CZ: case B200SQ_CS: case B200SQ_CP: u.ctrl = 0x7; break;
                        case B200SQ_CRZ: u.ctrl = 0x5; break;                       // control = 0
                        default: u.ctrl = 0; break;
                    }
                    // register slot (or -1) of each of the gate's qubits: pat = (ra + 1) | (rb + 1) << 2
                    int ra = -1, rb = -1;
                    for (int c = 0; c < rd.nb; ++c) {
                        if (rd.rb[c] == u.r0) ra = c;
                        if (u.r1 >= 0 && rd.rb[c] == u.r1) rb = c;
                    }
                    u.pat = (ra + 1) | ((rb + 1) << 2);
                    mat_off += 4;
                } else if (info.nq - info.nctrl == 1) {
                    u.r0 = regbit(g.q[info.nctrl]);
                    for (int k = 0; k < info.nctrl; ++k) u.ctrl |= 1u << pos[g.q[k]];
                    if (info.nctrl > 0) u.kind = UOP_U1CTRL;
                    else if (g.kind == B200SQ_RX) u.kind = UOP_U1RX;
                    else if (g.kind == B200SQ_RY || g.kind == B200SQ_H || g.kind == B200SQ_X) u.kind = UOP_U1REAL;
                    else u.kind = UOP_U1G;
                    mat_off += 4;
                } else {
                    u.kind = UOP_U2;
                    int ra = regbit(g.q[info.nctrl]), rb = regbit(g.q[info.nctrl + 1]);
                    u.r0 = std::max(ra, rb);
                    u.r1 = std::min(ra, rb);
                    u.swap = ra < rb ? 1 : 0;
                    for (int k = 0; k < info.nctrl; ++k) u.ctrl |= 1u << pos[g.q[k]];
                    mat_off += 16;
                }
                plan.uops.push_back(u);
            }
            rd.u1 = (int)plan.uops.size();
            plan.rounds.push_back(rd);
        }
        ps.r1 = (int)plan.rounds.size();
        ps.u1 = (int)plan.uops.size();
        ps.mat_elems = mat_off;
        plan.passes.push_back(ps);
        pending.swap(rest);
        first = false;
    }
    return plan;
}

}  // namespace b200sq
