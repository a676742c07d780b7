// Micro-op interpreter shared by the tile kernel (device) and the host emulator used by the CPU
// tests (tests/hostemu).  Everything here is per-thread code: one call handles the 2^nb
// amplitudes a thread owns in a round; barriers live in the callers.
#pragma once

#include <math.h>
#include <stdint.h>
#include <vector_types.h>

#include "plan.hpp"

#if defined(__CUDACC__)
#define B200SQ_HD __host__ __device__ __forceinline__
#else
#define B200SQ_HD inline
#endif

namespace b200sq {

// Shared-memory swizzle of the tile: XOR the 16-byte slot index inside a 128-byte line with the
// next three index bits, so that threads whose amplitudes are 8 slots apart (register bits
// 0,1,2) still hit eight different bank groups.
B200SQ_HD uint32_t tile_sw(uint32_t l) { return l ^ ((l >> 3) & 7u); }

B200SQ_HD double2 cmul(double2 a, double2 b) {
    return double2{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
B200SQ_HD double2 cfma(double2 a, double2 b, double2 c) {  // a*b + c
    return double2{fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y))};
}

// Scatter a compact index over bit runs (local index -> global amplitude index).
B200SQ_HD uint32_t scatter_runs(uint32_t v, int nruns, const uint32_t* mask, const int32_t* shift) {
    uint32_t out = 0;
    for (int r = 0; r < nruns; ++r) out |= (v & mask[r]) << shift[r];
    return out;
}

B200SQ_HD double gate_angle(const b200sq_gate& g, const double* x_row, const double* table,
                            int64_t n_samples, int64_t sample) {
    switch (g.fn) {
        case B200SQ_FN_LINEAR: return fma(g.c1, x_row[g.feat], g.c0);
        case B200SQ_FN_ARCCOS: return fma(g.c1, acos(x_row[g.feat]), g.c0);
        case B200SQ_FN_ARCTAN: return fma(g.c1, atan(x_row[g.feat]), g.c0);
        case B200SQ_FN_TABLE: return table[(int64_t)g.slot * n_samples + sample];
        default: return g.c0;
    }
}

// Fill the matrix / phase table of one micro-op.
//   U1:   m[0..3]  = 2x2 row-major
//   U2:   m[0..15] = 4x4 row-major, index = bit(r0) << 1 | bit(r1)
//   DIAG: m[0..3]  = phases, index = bit(q0) | bit(q1) << 1
B200SQ_HD void build_matrix(const DUop& u, const b200sq_gate& g, double theta, double2* m) {
    const double kInvSqrt2 = 0.70710678118654752440;
    double s, c;
    sincos(0.5 * theta, &s, &c);
    if (u.kind == UOP_DIAG) {
        double2 one{1.0, 0.0};
        double2 em{c, -s}, ep{c, s};  // e^{-i theta/2}, e^{+i theta/2}
        m[0] = m[1] = m[2] = m[3] = one;
        switch (g.kind) {
            case B200SQ_Z: m[1] = m[3] = double2{-1.0, 0.0}; break;
            case B200SQ_S: m[1] = m[3] = double2{0.0, 1.0}; break;
            case B200SQ_SDG: m[1] = m[3] = double2{0.0, -1.0}; break;
            case B200SQ_T: m[1] = m[3] = double2{kInvSqrt2, kInvSqrt2}; break;
            case B200SQ_TDG: m[1] = m[3] = double2{kInvSqrt2, -kInvSqrt2}; break;
            case B200SQ_RZ: m[0] = m[2] = em; m[1] = m[3] = ep; break;
            case B200SQ_P: { double sl, cl; sincos(theta, &sl, &cl); m[1] = m[3] = double2{cl, sl}; break; }
            case B200SQ_CZ: m[3] = double2{-1.0, 0.0}; break;
            case B200SQ_CS: m[3] = double2{0.0, 1.0}; break;
            case B200SQ_CP: { double sl, cl; sincos(theta, &sl, &cl); m[3] = double2{cl, sl}; break; }
            case B200SQ_CRZ: m[1] = em; m[3] = ep; break;        // control = bit 0 of the index
            case B200SQ_RZZ: m[0] = m[3] = em; m[1] = m[2] = ep; break;
            default: break;
        }
        return;
    }
    if (u.kind != UOP_U2) {  // the U1 family
        switch (g.kind) {
            case B200SQ_H: case B200SQ_CH:
                m[0] = m[1] = m[2] = double2{kInvSqrt2, 0.0}; m[3] = double2{-kInvSqrt2, 0.0}; break;
            case B200SQ_X: case B200SQ_CX: case B200SQ_CCX:
                m[0] = m[3] = double2{0.0, 0.0}; m[1] = m[2] = double2{1.0, 0.0}; break;
            case B200SQ_Y: case B200SQ_CY:
                m[0] = m[3] = double2{0.0, 0.0}; m[1] = double2{0.0, -1.0}; m[2] = double2{0.0, 1.0}; break;
            case B200SQ_SX: case B200SQ_CSX:
                m[0] = m[3] = double2{0.5, 0.5}; m[1] = m[2] = double2{0.5, -0.5}; break;
            case B200SQ_RX: case B200SQ_CRX:
                m[0] = m[3] = double2{c, 0.0}; m[1] = m[2] = double2{0.0, -s}; break;
            case B200SQ_RY: case B200SQ_CRY:
                m[0] = m[3] = double2{c, 0.0}; m[1] = double2{-s, 0.0}; m[2] = double2{s, 0.0}; break;
            case B200SQ_U: {
                double sp, cp, sl, cl, spl, cpl;
                sincos(g.c2, &sp, &cp);
                sincos(g.c3, &sl, &cl);
                sincos(g.c2 + g.c3, &spl, &cpl);
                m[0] = double2{c, 0.0};
                m[1] = double2{-cl * s, -sl * s};
                m[2] = double2{cp * s, sp * s};
                m[3] = double2{cpl * c, spl * c};
                break;
            }
            default:
                m[0] = m[3] = double2{1.0, 0.0}; m[1] = m[2] = double2{0.0, 0.0}; break;
        }
        return;
    }
    // U2: build in gate order (index = bit(first qubit) << 1 | bit(second qubit)), then permute
    double2 t[16];
    for (int i = 0; i < 16; ++i) t[i] = double2{0.0, 0.0};
    switch (g.kind) {
        case B200SQ_SWAP: case B200SQ_CSWAP:
            t[0] = t[15] = t[6] = t[9] = double2{1.0, 0.0}; break;
        case B200SQ_ISWAP:
            t[0] = t[15] = double2{1.0, 0.0}; t[6] = t[9] = double2{0.0, 1.0}; break;
        case B200SQ_ECR:
            // (X (x) I - Y (x) X) / sqrt(2), first qubit = most significant
            t[2] = double2{kInvSqrt2, 0.0}; t[3] = double2{0.0, kInvSqrt2};
            t[6] = double2{0.0, kInvSqrt2}; t[7] = double2{kInvSqrt2, 0.0};
            t[8] = double2{kInvSqrt2, 0.0}; t[9] = double2{0.0, -kInvSqrt2};
            t[12] = double2{0.0, -kInvSqrt2}; t[13] = double2{kInvSqrt2, 0.0};
            break;
        case B200SQ_RXX:  // cos I - i sin X(x)X
            t[0] = t[5] = t[10] = t[15] = double2{c, 0.0};
            t[3] = t[6] = t[9] = t[12] = double2{0.0, -s}; break;
        case B200SQ_RYY:  // cos I - i sin Y(x)Y ; Y(x)Y = antidiag(-1, 1, 1, -1)
            t[0] = t[5] = t[10] = t[15] = double2{c, 0.0};
            t[3] = t[12] = double2{0.0, s}; t[6] = t[9] = double2{0.0, -s}; break;
        case B200SQ_RZX:  // cos I - i sin Z(x)X ; Z on the first qubit
            t[0] = t[5] = t[10] = t[15] = double2{c, 0.0};
            t[1] = t[4] = double2{0.0, -s}; t[11] = t[14] = double2{0.0, s}; break;
        default:
            t[0] = t[5] = t[10] = t[15] = double2{1.0, 0.0}; break;
    }
    if (u.swap) {
        // first qubit sits on the LOWER register bit: exchange the two index bits
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                int ii = ((i & 1) << 1) | (i >> 1), jj = ((j & 1) << 1) | (j >> 1);
                m[ii * 4 + jj] = t[i * 4 + j];
            }
    } else {
        for (int i = 0; i < 16; ++i) m[i] = t[i];
    }
}

// ---------------------------------------------------------------------------------------------
// Product-state start of the first pass.

// v = (folded 1-qubit gates of `qubit`) |0>, in circuit order.
B200SQ_HD void fold_qubit_vector(int qubit, const DInit* inits, int n_init, const b200sq_gate* gates,
                                 const double* xrow, const double* table, int64_t n_total,
                                 int64_t sample, int vgate, double vshift, double2* out2) {
    double2 v0{1.0, 0.0}, v1{0.0, 0.0};
    for (int i = 0; i < n_init; ++i) {
        if (inits[i].qubit != qubit) continue;
        const b200sq_gate g = gates[inits[i].gate];
        double theta = gate_angle(g, xrow, table, n_total, sample);
        if ((int)inits[i].gate == vgate) theta += vshift;
        DUop u{};
        double2 m[4];
        const GateInfo gi = gate_info(g.kind);
        if (gi.diag) {
            u.kind = UOP_DIAG;
            build_matrix(u, g, theta, m);  // m[0], m[1] = phases of |0>, |1>
            v0 = cmul(m[0], v0);
            v1 = cmul(m[1], v1);
        } else {
            u.kind = UOP_U1G;
            build_matrix(u, g, theta, m);
            const double2 w0 = cfma(m[1], v1, cmul(m[0], v0));
            const double2 w1 = cfma(m[3], v1, cmul(m[2], v0));
            v0 = w0;
            v1 = w1;
        }
    }
    out2[0] = v0;
    out2[1] = v1;
}

// prod_b qv[qubits[b]][bit b of idx]
B200SQ_HD double2 product_amp(uint32_t idx, int nbits, const uint8_t* qubits, const double2* qv) {
    double2 r{1.0, 0.0};
    for (int b = 0; b < nbits; ++b) r = cmul(r, qv[2 * qubits[b] + ((idx >> b) & 1u)]);
    return r;
}

// ---------------------------------------------------------------------------------------------
// Register micro-op interpreter.  Everything is templated on NB (register bits of the round) so
// that the amplitude array a[2^NB] is indexed with compile-time constants only.

template <int NB, int R>
B200SQ_HD void u1_general(double2 (&a)[1 << NB], const double2* __restrict__ m) {
    const double2 m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma unroll
    for (int k0 = 0; k0 < (1 << NB); ++k0) {
        if (k0 & (1 << R)) continue;
        const int k1 = k0 | (1 << R);
        const double2 v0 = a[k0], v1 = a[k1];
        a[k0] = cfma(m01, v1, cmul(m00, v0));
        a[k1] = cfma(m11, v1, cmul(m10, v0));
    }
}

// [[c, -i s], [-i s, c]]
template <int NB, int R>
B200SQ_HD void u1_rx(double2 (&a)[1 << NB], const double2* __restrict__ m) {
    const double c = m[0].x, s = -m[1].y;
#pragma unroll
    for (int k0 = 0; k0 < (1 << NB); ++k0) {
        if (k0 & (1 << R)) continue;
        const int k1 = k0 | (1 << R);
        const double2 v0 = a[k0], v1 = a[k1];
        a[k0] = double2{fma(s, v1.y, c * v0.x), fma(-s, v1.x, c * v0.y)};
        a[k1] = double2{fma(s, v0.y, c * v1.x), fma(-s, v0.x, c * v1.y)};
    }
}

// real 2x2
template <int NB, int R>
B200SQ_HD void u1_real(double2 (&a)[1 << NB], const double2* __restrict__ m) {
    const double m00 = m[0].x, m01 = m[1].x, m10 = m[2].x, m11 = m[3].x;
#pragma unroll
    for (int k0 = 0; k0 < (1 << NB); ++k0) {
        if (k0 & (1 << R)) continue;
        const int k1 = k0 | (1 << R);
        const double2 v0 = a[k0], v1 = a[k1];
        a[k0] = double2{fma(m01, v1.x, m00 * v0.x), fma(m01, v1.y, m00 * v0.y)};
        a[k1] = double2{fma(m11, v1.x, m10 * v0.x), fma(m11, v1.y, m10 * v0.y)};
    }
}

template <int NB, int R>
B200SQ_HD void u1_ctrl(double2 (&a)[1 << NB], const double2* __restrict__ m, uint32_t e,
                       const uint32_t (&ko)[1 << NB], uint32_t cm) {
    const double2 m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma unroll
    for (int k0 = 0; k0 < (1 << NB); ++k0) {
        if (k0 & (1 << R)) continue;
        const int k1 = k0 | (1 << R);
        if (((e | ko[k0]) & cm) == cm) {
            const double2 v0 = a[k0], v1 = a[k1];
            a[k0] = cfma(m01, v1, cmul(m00, v0));
            a[k1] = cfma(m11, v1, cmul(m10, v0));
        }
    }
}

// 4x4 on register bits (RH > RL), matrix index = bit(RH) << 1 | bit(RL)
template <int NB, int RH, int RL>
B200SQ_HD void u2_general(double2 (&a)[1 << NB], const double2* __restrict__ m, uint32_t e,
                          const uint32_t (&ko)[1 << NB], uint32_t cm) {
#pragma unroll
    for (int k = 0; k < (1 << NB); ++k) {
        if (k & ((1 << RH) | (1 << RL))) continue;
        if (((e | ko[k]) & cm) != cm) continue;
        const int k00 = k, k01 = k | (1 << RL), k10 = k | (1 << RH), k11 = k | (1 << RH) | (1 << RL);
        const double2 v0 = a[k00], v1 = a[k01], v2 = a[k10], v3 = a[k11];
        a[k00] = cfma(m[3], v3, cfma(m[2], v2, cfma(m[1], v1, cmul(m[0], v0))));
        a[k01] = cfma(m[7], v3, cfma(m[6], v2, cfma(m[5], v1, cmul(m[4], v0))));
        a[k10] = cfma(m[11], v3, cfma(m[10], v2, cfma(m[9], v1, cmul(m[8], v0))));
        a[k11] = cfma(m[15], v3, cfma(m[14], v2, cfma(m[13], v1, cmul(m[12], v0))));
    }
}

// Diagonal gate.  RA / RB = register slot that carries the gate's first / second qubit, or -1 when
// that qubit is not a register bit of this round (then its bit comes from the thread's index e).
// With both known at compile time the phase index of every register amplitude is a constant, so a
// DIAG costs <= 4 phase loads + one (predicated) complex multiply per amplitude and no per-amplitude
// integer work.
template <int NB, int RA, int RB>
B200SQ_HD void diag_case(double2 (&a)[1 << NB], const DUop& u, const double2* __restrict__ m,
                         uint32_t e) {
    uint32_t base = 0;
    if (RA < 0) base |= (e >> u.r0) & 1u;
    if (RB < 0 && u.r1 >= 0) base |= ((e >> u.r1) & 1u) << 1;
    const uint32_t skip = u.ctrl;
    double2 ph[4];
    bool on[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        if ((RA < 0 && (v & 1)) || (RB < 0 && (v & 2))) continue;
        const uint32_t idx = base | (uint32_t)v;
        on[v] = !((skip >> idx) & 1u);
        ph[v] = m[idx];
    }
#pragma unroll
    for (int k = 0; k < (1 << NB); ++k) {
        const int v = (RA >= 0 ? ((k >> (RA >= 0 ? RA : 0)) & 1) : 0) |
                      (RB >= 0 ? (((k >> (RB >= 0 ? RB : 0)) & 1) << 1) : 0);
        if (on[v]) a[k] = cmul(a[k], ph[v]);
    }
}

template <int NB>
B200SQ_HD void diag_apply(double2 (&a)[1 << NB], const DUop& u, const double2* __restrict__ m,
                          uint32_t e) {
    // u.pat = (RA + 1) | (RB + 1) << 2
#define B200SQ_DC(RA, RB)                                                              \
    case ((RA) + 1) | (((RB) + 1) << 2):                                               \
        if ((RA) < NB && (RB) < NB) diag_case<NB, ((RA) < NB ? (RA) : -1), ((RB) < NB ? (RB) : -1)>(a, u, m, e); \
        break;
    switch (u.pat) {
        B200SQ_DC(-1, -1) B200SQ_DC(0, -1) B200SQ_DC(1, -1) B200SQ_DC(2, -1)
        B200SQ_DC(-1, 0) B200SQ_DC(-1, 1) B200SQ_DC(-1, 2)
        B200SQ_DC(0, 1) B200SQ_DC(0, 2) B200SQ_DC(1, 0) B200SQ_DC(1, 2) B200SQ_DC(2, 0) B200SQ_DC(2, 1)
        default: break;
    }
#undef B200SQ_DC
}

template <int NB>
B200SQ_HD void apply_uop(double2 (&a)[1 << NB], const DUop& u, const double2* __restrict__ m,
                         uint32_t e, const uint32_t (&ko)[1 << NB]) {
#define B200SQ_BY_R(FN, ...)                                        \
    switch (u.r0) {                                                  \
        case 0: FN<NB, 0>(__VA_ARGS__); break;                      \
        case 1: if (NB > 1) FN<NB, (NB > 1 ? 1 : 0)>(__VA_ARGS__); break; \
        default: if (NB > 2) FN<NB, (NB > 2 ? 2 : 0)>(__VA_ARGS__); break; \
    }
    switch (u.kind) {
        case UOP_U1RX: B200SQ_BY_R(u1_rx, a, m) break;
        case UOP_U1REAL: B200SQ_BY_R(u1_real, a, m) break;
        case UOP_U1G: B200SQ_BY_R(u1_general, a, m) break;
        case UOP_U1CTRL: B200SQ_BY_R(u1_ctrl, a, m, e, ko, u.ctrl) break;
        case UOP_DIAG: diag_apply<NB>(a, u, m, e); break;
        default:  // UOP_U2
            if (NB >= 2) {
                if (u.r0 == 1) {
                    u2_general<NB, (NB >= 2 ? 1 : 0), 0>(a, m, e, ko, u.ctrl);
                } else if (NB >= 3) {
                    if (u.r1 == 0) u2_general<NB, (NB >= 3 ? 2 : 1), 0>(a, m, e, ko, u.ctrl);
                    else u2_general<NB, (NB >= 3 ? 2 : 1), (NB >= 3 ? 1 : 0)>(a, m, e, ko, u.ctrl);
                }
            }
            break;
    }
#undef B200SQ_BY_R
}

// One thread's share of a round: groups g = tid, tid + nthreads, ...
// SMEM address of register slot k of group g:  sw(base | ko[k]) = (base ^ ((base >> 3) & 7)) ^ ksw[k]
// because base and ko[k] have disjoint bits (ksw[k] is precomputed on the host, see pack_round).
template <int NB>
B200SQ_HD void exec_round_nb(double2* tile, const DRound& rd, const DUop* uops, const double2* mats,
                             uint32_t ext_base, int T, int tid, int nthreads) {
    constexpr int NK = 1 << NB;
    const int ngroups = 1 << (T - NB);
    uint32_t ko[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        uint32_t o = 0;
        if (NB > 0 && (k & 1)) o |= 1u << rd.rb[0];
        if (NB > 1 && (k & 2)) o |= 1u << rd.rb[1];
        if (NB > 2 && (k & 4)) o |= 1u << rd.rb[2];
        ko[k] = o;
    }
    for (int g = tid; g < ngroups; g += nthreads) {
        // insert zero bits at the register-bit positions (ascending)
        uint32_t base = (uint32_t)g;
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            const uint32_t low = base & ((1u << rd.rb[c]) - 1u);
            base = ((base >> rd.rb[c]) << (rd.rb[c] + 1)) | low;
        }
        const uint32_t bsw = tile_sw(base);
        double2 a[NK];
#pragma unroll
        for (int k = 0; k < NK; ++k) a[k] = tile[bsw ^ rd.ksw[k]];
        const uint32_t e = ext_base | base;
        for (int ui = rd.u0; ui < rd.u1; ++ui) {
            const DUop u = uops[ui];
            apply_uop<NB>(a, u, mats + u.mat, e, ko);
        }
#pragma unroll
        for (int k = 0; k < NK; ++k) tile[bsw ^ rd.ksw[k]] = a[k];
    }
}

B200SQ_HD void exec_round_thread(double2* tile, const DRound& rd, const DUop* uops,
                                 const double2* mats, uint32_t ext_base, int T, int tid,
                                 int nthreads) {
    switch (rd.nb) {
        case 3: exec_round_nb<3>(tile, rd, uops, mats, ext_base, T, tid, nthreads); break;
        case 2: exec_round_nb<2>(tile, rd, uops, mats, ext_base, T, tid, nthreads); break;
        default: exec_round_nb<1>(tile, rd, uops, mats, ext_base, T, tid, nthreads); break;
    }
}

// <P> contribution helper: returns sign * Re(i^ny conj(p) l).
B200SQ_HD double pauli_term(double2 p, double2 l, uint32_t idx, uint32_t zm, int ny) {
    double A = p.x * l.x + p.y * l.y;   // Re conj(p) l
    double B = p.x * l.y - p.y * l.x;   // Im conj(p) l
    double v;
    switch (ny & 3) {
        case 0: v = A; break;
        case 1: v = -B; break;
        case 2: v = -A; break;
        default: v = B; break;
    }
#if defined(__CUDA_ARCH__)
    return (__popc(idx & zm) & 1) ? -v : v;
#else
    return (__builtin_popcount(idx & zm) & 1) ? -v : v;
#endif
}

}  // namespace b200sq
