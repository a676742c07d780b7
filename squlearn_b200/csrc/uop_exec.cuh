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
B200SQ_HD void build_matrix(const Uop& u, const b200sq_gate& g, double theta, double2* m) {
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
    if (u.kind == UOP_U1) {
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

// Apply one micro-op to the register amplitudes a[0..2^nb).  `e` is the ext index of a[0]
// (register bits cleared); ko[k] is the ext/local offset of register slot k.
B200SQ_HD void apply_uop(double2 (&a)[8], const Uop& u, const double2* __restrict__ m, uint32_t e,
                         const uint32_t (&ko)[8], int nk) {
    if (u.kind == UOP_DIAG) {
        const int qa = u.r0, qb = u.r1;
        const uint32_t skip = u.ctrl;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k < nk) {
                uint32_t ek = e | ko[k];
                uint32_t idx = (ek >> qa) & 1u;
                if (qb >= 0) idx |= ((ek >> qb) & 1u) << 1;
                if (!((skip >> idx) & 1u)) a[k] = cmul(a[k], m[idx]);
            }
        }
        return;
    }
    if (u.kind == UOP_U1) {
        const double2 m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
        const uint32_t cm = u.ctrl;
#define B200SQ_PAIR(k0, k1)                                                       \
    if ((k1) < nk && (((e | ko[k0]) & cm) == cm)) {                              \
        double2 v0 = a[k0], v1 = a[k1];                                          \
        a[k0] = cfma(m01, v1, cmul(m00, v0));                                    \
        a[k1] = cfma(m11, v1, cmul(m10, v0));                                    \
    }
        switch (u.r0) {
            case 0: B200SQ_PAIR(0, 1) B200SQ_PAIR(2, 3) B200SQ_PAIR(4, 5) B200SQ_PAIR(6, 7) break;
            case 1: B200SQ_PAIR(0, 2) B200SQ_PAIR(1, 3) B200SQ_PAIR(4, 6) B200SQ_PAIR(5, 7) break;
            default: B200SQ_PAIR(0, 4) B200SQ_PAIR(1, 5) B200SQ_PAIR(2, 6) B200SQ_PAIR(3, 7) break;
        }
#undef B200SQ_PAIR
        return;
    }
    // U2 on register bits (r0 > r1); matrix index = bit(r0) << 1 | bit(r1)
    {
        const uint32_t cm = u.ctrl;
#define B200SQ_QUAD(k00, k01, k10, k11)                                           \
    if ((k11) < nk && (((e | ko[k00]) & cm) == cm)) {                            \
        double2 v0 = a[k00], v1 = a[k01], v2 = a[k10], v3 = a[k11];              \
        a[k00] = cfma(m[3], v3, cfma(m[2], v2, cfma(m[1], v1, cmul(m[0], v0)))); \
        a[k01] = cfma(m[7], v3, cfma(m[6], v2, cfma(m[5], v1, cmul(m[4], v0)))); \
        a[k10] = cfma(m[11], v3, cfma(m[10], v2, cfma(m[9], v1, cmul(m[8], v0)))); \
        a[k11] = cfma(m[15], v3, cfma(m[14], v2, cfma(m[13], v1, cmul(m[12], v0)))); \
    }
        if (u.r0 == 1) {            // bits (1, 0); free bit 2
            B200SQ_QUAD(0, 1, 2, 3) B200SQ_QUAD(4, 5, 6, 7)
        } else if (u.r1 == 0) {     // bits (2, 0); free bit 1
            B200SQ_QUAD(0, 1, 4, 5) B200SQ_QUAD(2, 3, 6, 7)
        } else {                    // bits (2, 1); free bit 0
            B200SQ_QUAD(0, 2, 4, 6) B200SQ_QUAD(1, 3, 5, 7)
        }
#undef B200SQ_QUAD
    }
}

// One thread's share of a round: groups g = tid, tid + nthreads, ...
B200SQ_HD void exec_round_thread(double2* tile, const Round& rd, const Uop* uops,
                                 const double2* mats, uint32_t ext_base, int T, int tid,
                                 int nthreads) {
    const int nb = rd.nb;
    const int nk = 1 << nb;
    const int ngroups = 1 << (T - nb);
    uint32_t ko[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t o = 0;
        if (nb > 0 && (k & 1)) o |= 1u << rd.rb[0];
        if (nb > 1 && (k & 2)) o |= 1u << rd.rb[1];
        if (nb > 2 && (k & 4)) o |= 1u << rd.rb[2];
        ko[k] = o;
    }
    for (int g = tid; g < ngroups; g += nthreads) {
        // insert zero bits at the register-bit positions (ascending)
        uint32_t base = (uint32_t)g;
        for (int c = 0; c < nb; ++c) {
            uint32_t low = base & ((1u << rd.rb[c]) - 1u);
            base = ((base >> rd.rb[c]) << (rd.rb[c] + 1)) | low;
        }
        double2 a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < nk) a[k] = tile[tile_sw(base | ko[k])];
        const uint32_t e = ext_base | base;
        for (int ui = rd.u0; ui < rd.u1; ++ui) apply_uop(a, uops[ui], mats + uops[ui].mat, e, ko, nk);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < nk) tile[tile_sw(base | ko[k])] = a[k];
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
