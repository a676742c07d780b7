// Per-thread building blocks of the tile kernel, shared by the device kernel (sim_kernels.cu) and the
// host emulator used by the CPU tests (tests/hostemu).  Everything here is per-thread code: the
// phases of a work item (build matrices -> build tables -> stage tile -> rounds -> store) are
// separated by team barriers that live in the callers.
#pragma once

#if !defined(__CUDACC_RTC__)
#include <math.h>
#include <stdint.h>
#include <vector_types.h>
#endif

#include "plan_dev.h"

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

// out bit pos[i] = bit i of v, for i in [b0, b1); pos[i] == kNoBit drops the bit.
B200SQ_HD uint32_t map_bits(uint32_t v, const uint8_t* pos, int b0, int b1) {
    uint32_t out = 0;
    for (int i = b0; i < b1; ++i)
        if (((v >> i) & 1u) && pos[i] != kNoBit) out |= 1u << pos[i];
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

// Fill the matrix / phase quadruple of one gate.
//   MAT_U1:   m[0..3]  = 2x2 row-major
//   MAT_U2:   m[0..15] = 4x4 row-major, index = bit(r0) << 1 | bit(r1)
//   MAT_DIAG: m[0..3]  = phases, index = bit(q0) | bit(q1) << 1
B200SQ_HD void build_matrix(int mkind, int swap, const b200sq_gate& g, double theta, double2* m) {
    const double kInvSqrt2 = 0.70710678118654752440;
    double s, c;
    sincos(0.5 * theta, &s, &c);
    if (mkind == MAT_ROT) {  // rx / ry as a scaled rotation: the larger of |c|, |s| is factored out (plan_dev.h)
        if (fabs(c) >= fabs(s)) {
            m[0] = double2{s / c, 0.0};
            m[1] = double2{c, 0.0};
        } else {
            m[0] = double2{c / s, 1.0};
            m[1] = double2{s, 0.0};
        }
        return;
    }
    if (mkind == MAT_DIAG) {
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
    if (mkind != MAT_U2) {  // the U1 family
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
    if (swap) {
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
// Views and per-team / per-thread state.

struct PassView {  // pointers into a pass blob (its shared-memory copy on the device)
    const DPass* P;
    const DRound* rounds;
    const DStep* steps;
    const DMat* mats;
    const DTab* tabs;
    const DTabGate* tabgates;
    const DInit* inits;
};
B200SQ_HD PassView view_pass(const unsigned char* blob) {
    PassView v;
    v.P = reinterpret_cast<const DPass*>(blob);
    v.rounds = reinterpret_cast<const DRound*>(blob + v.P->off_rounds);
    v.steps = reinterpret_cast<const DStep*>(blob + v.P->off_steps);
    v.mats = reinterpret_cast<const DMat*>(blob + v.P->off_mats);
    v.tabs = reinterpret_cast<const DTab*>(blob + v.P->off_tabs);
    v.tabgates = reinterpret_cast<const DTabGate*>(blob + v.P->off_tabgates);
    v.inits = reinterpret_cast<const DInit*>(blob + v.P->off_inits);
    return v;
}

// Team scratch behind the tile buffer(s): [mats: mat_elems][qv: 32][plo: 64][phi: 128][jt: 4 * J uint32]
constexpr int kScratchElems = 32 + 64 + 128;
struct TeamMem {
    double2* tile;
    double2* mats;
    double2* qv;    // start vectors of the fresh qubits: qv[2 * f + bit]
    double2* plo;   // partial products over the first nlo fresh qubits
    double2* phi;   // ... over the remaining nhi
    uint32_t* jt;   // per-j address parts: jt[0..J) input, [J..2J) output, [2J..3J) plo index, [3J..4J) phi index
    int J;          // tile elements per thread
};

struct ThreadMap {  // the thread's share (low log2(TS) bits of the local index) of every index
    uint32_t in_lo, out_lo, plo_lo, phi_lo;
};

struct ItemCtx {
    const b200sq_gate* gates;
    const double* xrow;
    const double* table;
    int64_t n_total, sample;
    int vgate, vgate2;       // gates whose angle is shifted in this variant (-1: none); a second shift serves the
    double vshift, vshift2;  // second-order derivatives (two gates, or twice the same gate: the shifts add)
    uint32_t ext_base;
};

// Decode the outer bits of a work item.  Returns true when the tile is identically zero (an outer
// qubit that is still |0> in the input layout has its bit set).
B200SQ_HD bool item_outer(const DPass& P, uint32_t o, uint32_t& ext_base, uint32_t& obase_in,
                          uint32_t& obase_out) {
    ext_base = 0;
    obase_in = 0;
    obase_out = 0;
    bool zero = false;
    for (int b = 0; b < P.n_obits; ++b) {
        if (!((o >> b) & 1u)) continue;
        ext_base |= 1u << P.o_ext[b];
        obase_out |= 1u << P.o_out[b];
        if (P.o_in[b] == kNoBit) zero = true;
        else obase_in |= 1u << P.o_in[b];
    }
    return zero;
}

B200SQ_HD void setup_thread(const PassView& pv, const TeamMem& tm, int tid, int TS, int log2ts,
                            ThreadMap& map) {
    const DPass& P = *pv.P;
    const int nl = log2ts < P.T ? log2ts : P.T;
    map.in_lo = map_bits((uint32_t)tid, P.l_in, 0, nl);
    map.out_lo = map_bits((uint32_t)tid, P.l_out, 0, nl);
    map.plo_lo = map_bits((uint32_t)tid, P.l_plo, 0, nl);
    map.phi_lo = map_bits((uint32_t)tid, P.l_phi, 0, nl);
    for (int j = tid; j < tm.J; j += TS) {
        const uint32_t hb = (uint32_t)j << log2ts;
        tm.jt[j] = map_bits(hb, P.l_in, nl, P.T);
        tm.jt[tm.J + j] = map_bits(hb, P.l_out, nl, P.T);
        tm.jt[2 * tm.J + j] = map_bits(hb, P.l_plo, nl, P.T);
        tm.jt[3 * tm.J + j] = map_bits(hb, P.l_phi, nl, P.T);
    }
}

// ---------------------------------------------------------------------------------------------
// Phase A: gate matrices and the start vectors of the fresh qubits.

// v = (folded 1-qubit gates of `qubit`) |0>, in circuit order.
B200SQ_HD void fold_qubit_vector(int qubit, const DInit* inits, int n_init, const ItemCtx& ic, double2* out2) {
    double2 v0{1.0, 0.0}, v1{0.0, 0.0};
    for (int i = 0; i < n_init; ++i) {
        if (inits[i].qubit != qubit) continue;
        const b200sq_gate g = ic.gates[inits[i].gate];
        double theta = gate_angle(g, ic.xrow, ic.table, ic.n_total, ic.sample);
        if ((int)inits[i].gate == ic.vgate) theta += ic.vshift;
        if ((int)inits[i].gate == ic.vgate2) theta += ic.vshift2;
        double2 m[4];
        if (gate_info(g.kind).diag) {
            build_matrix(MAT_DIAG, 0, g, theta, m);  // m[0], m[1] = phases of |0>, |1>
            v0 = cmul(m[0], v0);
            v1 = cmul(m[1], v1);
        } else {
            build_matrix(MAT_U1, 0, g, theta, m);
            const double2 w0 = cfma(m[1], v1, cmul(m[0], v0));
            const double2 w1 = cfma(m[3], v1, cmul(m[2], v0));
            v0 = w0;
            v1 = w1;
        }
    }
    out2[0] = v0;
    out2[1] = v1;
}

B200SQ_HD void build_mats_thread(const PassView& pv, const TeamMem& tm, const ItemCtx& ic, int tid, int TS) {
    const DPass& P = *pv.P;
    for (int i = tid; i < P.n_mats; i += TS) {
        const DMat r = pv.mats[i];
        const b200sq_gate g = ic.gates[r.gate];
        double theta = gate_angle(g, ic.xrow, ic.table, ic.n_total, ic.sample);
        if ((int)r.gate == ic.vgate) theta += ic.vshift;
        if ((int)r.gate == ic.vgate2) theta += ic.vshift2;
        build_matrix(r.kind, r.swap, g, theta, tm.mats + r.mat);
    }
    for (int f = tid; f < P.n_fresh; f += TS)
        fold_qubit_vector(P.fresh_q[f], pv.inits, P.n_init, ic, tm.qv + 2 * f);
}

// Phase B: product tables of the fresh qubits and the phase tables of the DTAB steps.
// `new_sample`: the matrices were rebuilt, so every table is stale; otherwise only the phase tables that
// read outer index bits (DTab::nbits & 0x80) change from tile to tile of the same state.
B200SQ_HD void build_tables_thread(const PassView& pv, const TeamMem& tm, const ItemCtx& ic, int tid, int TS,
                                   bool new_sample) {
    const DPass& P = *pv.P;
    // product of the factored-out scales of the scaled rotations of this pass (MAT_ROT), carried by the low
    // product table or by one phase table
    double rot_scale = 1.0;
    if (new_sample && P.rot_carrier)
        for (int i = 0; i < P.n_mats; ++i)
            if (pv.mats[i].kind == MAT_ROT) rot_scale *= tm.mats[pv.mats[i].mat + 1].x;
    const double plo_scale = P.rot_carrier == 1 ? rot_scale : 1.0;
    for (int t = tid; t < (1 << P.nlo) && P.nlo > 0 && new_sample; t += TS) {
        double2 r{plo_scale, 0.0};
        for (int b = 0; b < P.nlo; ++b) r = cmul(r, tm.qv[2 * b + ((t >> b) & 1)]);
        tm.plo[t] = r;
    }
    for (int t = tid; t < (1 << P.nhi) && P.nhi > 0 && new_sample; t += TS) {
        double2 r{1.0, 0.0};
        for (int b = 0; b < P.nhi; ++b) r = cmul(r, tm.qv[2 * (P.nlo + b) + ((t >> b) & 1)]);
        tm.phi[t] = r;
    }
    for (int it = 0; it < P.n_tabs; ++it) {
        const DTab tb = pv.tabs[it];
        const bool legacy = tb.nbits & 0x80;   // reads more than two outer bits: rebuilt for every tile
        if (!new_sample && !legacy) continue;
        const int nbits = tb.nbits & 0x7f;
        const int nv = legacy ? 0 : (tb.vb[0] != kNoBit) + (tb.vb[1] != kNoBit);
        const double scale = (tb.carrier && P.rot_carrier == 2) ? rot_scale : 1.0;
        // every variant (value of the outer bits the table reads) is built here, once per state
        for (int t = tid; t < (1 << (nbits + nv)); t += TS) {
            const uint32_t v = (uint32_t)t >> nbits;
            uint32_t e = ((uint32_t)(t & ((1 << nbits) - 1)) << tb.lo);
            if (legacy) {
                e |= ic.ext_base;
            } else {
                if (nv > 0) e |= (v & 1u) << tb.vb[0];
                if (nv > 1) e |= ((v >> 1) & 1u) << tb.vb[1];
            }
            double2 ph{scale, 0.0};
            for (int g = tb.g0; g < tb.g1; ++g) {
                const DTabGate tg = pv.tabgates[g];
                uint32_t idx = (e >> tg.p0) & 1u;
                if (tg.p1 != kNoBit) idx |= ((e >> tg.p1) & 1u) << 1;
                ph = cmul(ph, tm.mats[tg.m4 + idx]);
            }
            tm.mats[tb.mat + t] = ph;
        }
    }
}

// Phase C (gather form): tile[l] = src[l restricted to the input layout] * prod of fresh start vectors.
// Plain passes (no fresh qubits) use the cp.async double-buffered path of the kernel instead; the
// host emulator uses this form for every pass.
B200SQ_HD void stage_thread(const PassView& pv, const TeamMem& tm, const double2* src, uint32_t obase_in,
                            const ThreadMap& map, int tid, int TS) {
    const uint32_t tile_elems = 1u << pv.P->T;
    const bool has_src = pv.P->has_src, has_lo = pv.P->nlo > 0, has_hi = pv.P->nhi > 0;
    const double2* __restrict__ sp = src + (obase_in | map.in_lo);
    const double2* __restrict__ plo = tm.plo + map.plo_lo;
    const double2* __restrict__ phi = tm.phi + map.phi_lo;
    const uint32_t* __restrict__ jin = tm.jt;
    const uint32_t* __restrict__ jlo = tm.jt + 2 * tm.J;
    const uint32_t* __restrict__ jhi = tm.jt + 3 * tm.J;
    double2* __restrict__ tile = tm.tile;
    int j = 0;
    // unrolled: the source loads of several elements are in flight together (the gather goes through L1 / L2)
#pragma unroll 8
    for (uint32_t l = (uint32_t)tid; l < tile_elems; l += (uint32_t)TS, ++j) {
        double2 v{1.0, 0.0};
        if (has_src) {
#if defined(__CUDA_ARCH__)
            v = __ldg(sp + jin[j]);
#else
            v = sp[jin[j]];
#endif
        }
        if (has_lo) {
            const double2 p = plo[jlo[j]];
            v = has_src ? cmul(v, p) : p;
        }
        if (has_hi) v = cmul(v, phi[jhi[j]]);
        tile[tile_sw(l)] = v;
    }
}

B200SQ_HD void store_thread(const PassView& pv, const TeamMem& tm, double2* dst, uint32_t obase_out,
                            const ThreadMap& map, int tid, int TS) {
    const uint32_t tile_elems = 1u << pv.P->T;
    int j = 0;
    for (uint32_t l = (uint32_t)tid; l < tile_elems; l += (uint32_t)TS, ++j)
        dst[obase_out | map.out_lo | tm.jt[tm.J + j]] = tm.tile[tile_sw(l)];
}

// ---------------------------------------------------------------------------------------------
// Register step interpreter.  Templated on NB (register bits of the round) and G (groups a thread
// processes together), so that the amplitude array a[G][2^NB] is indexed with compile-time
// constants only.  A step is decoded once and its matrix loaded once for all G * 2^NB amplitudes
// of the thread.

template <int NB, int R, int G>
B200SQ_HD void u1_general(double2 (&a)[G][1 << NB], const double2* __restrict__ m) {
    const double2 m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma unroll
    for (int gi = 0; gi < G; ++gi)
#pragma unroll
        for (int k0 = 0; k0 < (1 << NB); ++k0) {
            if (k0 & (1 << R)) continue;
            const int k1 = k0 | (1 << R);
            const double2 v0 = a[gi][k0], v1 = a[gi][k1];
            a[gi][k0] = cfma(m01, v1, cmul(m00, v0));
            a[gi][k1] = cfma(m11, v1, cmul(m10, v0));
        }
}

// [[c, -i s], [-i s, c]]
template <int NB, int R, int G>
B200SQ_HD void u1_rx(double2 (&a)[G][1 << NB], const double2* __restrict__ m) {
    const double c = m[0].x, s = -m[1].y;
#pragma unroll
    for (int gi = 0; gi < G; ++gi)
#pragma unroll
        for (int k0 = 0; k0 < (1 << NB); ++k0) {
            if (k0 & (1 << R)) continue;
            const int k1 = k0 | (1 << R);
            const double2 v0 = a[gi][k0], v1 = a[gi][k1];
            a[gi][k0] = double2{fma(s, v1.y, c * v0.x), fma(-s, v1.x, c * v0.y)};
            a[gi][k1] = double2{fma(s, v0.y, c * v1.x), fma(-s, v0.x, c * v1.y)};
        }
}

// real 2x2
template <int NB, int R, int G>
B200SQ_HD void u1_real(double2 (&a)[G][1 << NB], const double2* __restrict__ m) {
    const double m00 = m[0].x, m01 = m[1].x, m10 = m[2].x, m11 = m[3].x;
#pragma unroll
    for (int gi = 0; gi < G; ++gi)
#pragma unroll
        for (int k0 = 0; k0 < (1 << NB); ++k0) {
            if (k0 & (1 << R)) continue;
            const int k1 = k0 | (1 << R);
            const double2 v0 = a[gi][k0], v1 = a[gi][k1];
            a[gi][k0] = double2{fma(m01, v1.x, m00 * v0.x), fma(m01, v1.y, m00 * v0.y)};
            a[gi][k1] = double2{fma(m11, v1.x, m10 * v0.x), fma(m11, v1.y, m10 * v0.y)};
        }
}

// Scaled rotations (MAT_ROT, plan_dev.h): one FMA per real component; the factored-out scale lives in the
// pass's carrier table.  The form (which of |c|, |s| was factored out) is uniform over the team.

// rx(theta) / scale:  form 0: v0' = v0 - i t v1, v1' = v1 - i t v0;   form 1: v0' = u v0 - i v1, v1' = u v1 - i v0
template <int NB, int R, int G>
B200SQ_HD void u1_rotx(double2 (&a)[G][1 << NB], const double2* __restrict__ m) {
    const double t = m[0].x;
    if (m[0].y == 0.0) {
#pragma unroll
        for (int gi = 0; gi < G; ++gi)
#pragma unroll
            for (int k0 = 0; k0 < (1 << NB); ++k0) {
                if (k0 & (1 << R)) continue;
                const int k1 = k0 | (1 << R);
                const double2 v0 = a[gi][k0], v1 = a[gi][k1];
                a[gi][k0] = double2{fma(t, v1.y, v0.x), fma(-t, v1.x, v0.y)};
                a[gi][k1] = double2{fma(t, v0.y, v1.x), fma(-t, v0.x, v1.y)};
            }
    } else {
#pragma unroll
        for (int gi = 0; gi < G; ++gi)
#pragma unroll
            for (int k0 = 0; k0 < (1 << NB); ++k0) {
                if (k0 & (1 << R)) continue;
                const int k1 = k0 | (1 << R);
                const double2 v0 = a[gi][k0], v1 = a[gi][k1];
                a[gi][k0] = double2{fma(t, v0.x, v1.y), fma(t, v0.y, -v1.x)};
                a[gi][k1] = double2{fma(t, v1.x, v0.y), fma(t, v1.y, -v0.x)};
            }
    }
}

// ry(theta) / scale:  form 0: v0' = v0 - t v1, v1' = v1 + t v0;   form 1: v0' = u v0 - v1, v1' = u v1 + v0
template <int NB, int R, int G>
B200SQ_HD void u1_roty(double2 (&a)[G][1 << NB], const double2* __restrict__ m) {
    const double t = m[0].x;
    if (m[0].y == 0.0) {
#pragma unroll
        for (int gi = 0; gi < G; ++gi)
#pragma unroll
            for (int k0 = 0; k0 < (1 << NB); ++k0) {
                if (k0 & (1 << R)) continue;
                const int k1 = k0 | (1 << R);
                const double2 v0 = a[gi][k0], v1 = a[gi][k1];
                a[gi][k0] = double2{fma(-t, v1.x, v0.x), fma(-t, v1.y, v0.y)};
                a[gi][k1] = double2{fma(t, v0.x, v1.x), fma(t, v0.y, v1.y)};
            }
    } else {
#pragma unroll
        for (int gi = 0; gi < G; ++gi)
#pragma unroll
            for (int k0 = 0; k0 < (1 << NB); ++k0) {
                if (k0 & (1 << R)) continue;
                const int k1 = k0 | (1 << R);
                const double2 v0 = a[gi][k0], v1 = a[gi][k1];
                a[gi][k0] = double2{fma(t, v0.x, -v1.x), fma(t, v0.y, -v1.y)};
                a[gi][k1] = double2{fma(t, v1.x, v0.x), fma(t, v1.y, v0.y)};
            }
    }
}

template <int NB, int R, int G>
B200SQ_HD void u1_ctrl(double2 (&a)[G][1 << NB], const double2* __restrict__ m, const uint32_t (&e)[G],
                       const uint32_t (&ko)[1 << NB], uint32_t cm) {
    const double2 m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma unroll
    for (int gi = 0; gi < G; ++gi)
#pragma unroll
        for (int k0 = 0; k0 < (1 << NB); ++k0) {
            if (k0 & (1 << R)) continue;
            const int k1 = k0 | (1 << R);
            if (((e[gi] | ko[k0]) & cm) == cm) {
                const double2 v0 = a[gi][k0], v1 = a[gi][k1];
                a[gi][k0] = cfma(m01, v1, cmul(m00, v0));
                a[gi][k1] = cfma(m11, v1, cmul(m10, v0));
            }
        }
}

// 4x4 on register bits (RH > RL), matrix index = bit(RH) << 1 | bit(RL)
template <int NB, int RH, int RL, int G>
B200SQ_HD void u2_general(double2 (&a)[G][1 << NB], const double2* __restrict__ m, const uint32_t (&e)[G],
                          const uint32_t (&ko)[1 << NB], uint32_t cm) {
#pragma unroll
    for (int gi = 0; gi < G; ++gi)
#pragma unroll
        for (int k = 0; k < (1 << NB); ++k) {
            if (k & ((1 << RH) | (1 << RL))) continue;
            if (((e[gi] | ko[k]) & cm) != cm) continue;
            const int k00 = k, k01 = k | (1 << RL), k10 = k | (1 << RH), k11 = k | (1 << RH) | (1 << RL);
            const double2 v0 = a[gi][k00], v1 = a[gi][k01], v2 = a[gi][k10], v3 = a[gi][k11];
            a[gi][k00] = cfma(m[3], v3, cfma(m[2], v2, cfma(m[1], v1, cmul(m[0], v0))));
            a[gi][k01] = cfma(m[7], v3, cfma(m[6], v2, cfma(m[5], v1, cmul(m[4], v0))));
            a[gi][k10] = cfma(m[11], v3, cfma(m[10], v2, cfma(m[9], v1, cmul(m[8], v0))));
            a[gi][k11] = cfma(m[15], v3, cfma(m[14], v2, cfma(m[13], v1, cmul(m[12], v0))));
        }
}

// Single diagonal gate, generic form: the phase index of every amplitude comes from its ext index.
// (Runs of diagonal gates go through the phase tables; this path only sees stray gates.)
template <int NB, int G>
B200SQ_HD void diag_apply(double2 (&a)[G][1 << NB], const DStep& u, const double2* __restrict__ m,
                          const uint32_t (&e)[G], const uint32_t (&ko)[1 << NB]) {
    const int q0 = u.r0, q1 = u.r1 >= 0 ? u.r1 : u.r0;
    const uint32_t two = u.r1 >= 0 ? 2u : 0u;
    uint32_t kp[1 << NB];  // the register slots' share of the phase index
#pragma unroll
    for (int k = 0; k < (1 << NB); ++k) kp[k] = ((ko[k] >> q0) & 1u) | (((ko[k] >> q1) << 1) & two);
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
        const uint32_t i0 = ((e[gi] >> q0) & 1u) | (((e[gi] >> q1) << 1) & two);
#pragma unroll
        for (int k = 0; k < (1 << NB); ++k) a[gi][k] = cmul(a[gi][k], m[i0 | kp[k]]);
    }
}

// Phase table: one complex multiply per amplitude for a whole run of diagonal gates.
template <int NB, int G>
B200SQ_HD void dtab_apply(double2 (&a)[G][1 << NB], const DStep& u, const double2* __restrict__ tab,
                          const uint32_t (&e)[G], const uint32_t (&base)[G], const uint32_t (&ko)[1 << NB]) {
    const uint32_t mask = (1u << u.r1) - 1u;
    // variant of the table for this tile's value of the outer bits it reads (u.ctrl = vb0 | vb1 << 8)
    const uint32_t vb0 = u.ctrl & 0xffu, vb1 = (u.ctrl >> 8) & 0xffu;
    if (vb0 != (uint32_t)kNoBit) {
        uint32_t v = (e[0] >> vb0) & 1u;
        if (vb1 != (uint32_t)kNoBit) v |= ((e[0] >> vb1) & 1u) << 1;
        tab += v << u.r1;
    }
    uint32_t kp[1 << NB];
#pragma unroll
    for (int k = 0; k < (1 << NB); ++k) kp[k] = (ko[k] >> u.r0) & mask;
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
        const uint32_t i0 = (base[gi] >> u.r0) & mask;
#pragma unroll
        for (int k = 0; k < (1 << NB); ++k) a[gi][k] = cmul(a[gi][k], tab[i0 | kp[k]]);
    }
}

template <int NB, int G>
B200SQ_HD void apply_step(double2 (&a)[G][1 << NB], const DStep& u, const double2* __restrict__ m,
                          const uint32_t (&e)[G], const uint32_t (&base)[G], const uint32_t (&ko)[1 << NB]) {
#define B200SQ_LAYER(FN)                                                      \
    if (u.r0 & 1) FN<NB, 0, G>(a, m);                                         \
    if (NB > 1 && (u.r0 & 2)) FN<NB, (NB > 1 ? 1 : 0), G>(a, m + 4);          \
    if (NB > 2 && (u.r0 & 4)) FN<NB, (NB > 2 ? 2 : 0), G>(a, m + 8);
    switch (u.kind) {
        case ST_LRX: B200SQ_LAYER(u1_rx) break;
        case ST_LREAL: B200SQ_LAYER(u1_real) break;
        case ST_LGEN: B200SQ_LAYER(u1_general) break;
        case ST_LROTX: B200SQ_LAYER(u1_rotx) break;
        case ST_LROTY: B200SQ_LAYER(u1_roty) break;
        case ST_DTAB: dtab_apply<NB, G>(a, u, m, e, base, ko); break;
        case ST_DIAG: diag_apply<NB, G>(a, u, m, e, ko); break;
        case ST_U1CTRL:
            switch (u.r0) {
                case 0: u1_ctrl<NB, 0, G>(a, m, e, ko, u.ctrl); break;
                case 1: if (NB > 1) u1_ctrl<NB, (NB > 1 ? 1 : 0), G>(a, m, e, ko, u.ctrl); break;
                default: if (NB > 2) u1_ctrl<NB, (NB > 2 ? 2 : 0), G>(a, m, e, ko, u.ctrl); break;
            }
            break;
        default:  // ST_U2
            if (NB >= 2) {
                if (u.r0 == 1) {
                    u2_general<NB, (NB >= 2 ? 1 : 0), 0, G>(a, m, e, ko, u.ctrl);
                } else if (NB >= 3) {
                    if (u.r1 == 0) u2_general<NB, (NB >= 3 ? 2 : 1), 0, G>(a, m, e, ko, u.ctrl);
                    else u2_general<NB, (NB >= 3 ? 2 : 1), (NB >= 3 ? 1 : 0), G>(a, m, e, ko, u.ctrl);
                }
            }
            break;
    }
#undef B200SQ_LAYER
}

// One thread's share of a round: groups g = tid, tid + nthreads, ... taken G at a time.
// SMEM address of register slot k of group g:  sw(base | ko[k]) = (base ^ ((base >> 3) & 7)) ^ ksw[k]
// because base and ko[k] have disjoint bits (ksw[k] is precomputed on the host, see build_pass_blob).
// `steps(a, e, base, ko)` applies the round's steps to the register amplitudes: the interpreter loop
// of exec_round_nb, or the straight-line sequence a run-time specialised kernel (jit.cu) passes in.
template <int NB, int G, class Steps>
B200SQ_HD void exec_round_with(double2* __restrict__ tile, const DRound& rd, uint32_t ext_base, int T, int tid,
                               int nthreads, Steps steps) {
    constexpr int NK = 1 << NB;
    const int ngroups = 1 << (T - NB);
    uint32_t ko[NK], ksw[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        uint32_t o = 0;
        if (NB > 0 && (k & 1)) o |= 1u << rd.rb[0];
        if (NB > 1 && (k & 2)) o |= 1u << rd.rb[1];
        if (NB > 2 && (k & 4)) o |= 1u << rd.rb[2];
        ko[k] = o;
        ksw[k] = rd.ksw[k];
    }
    for (int g0 = tid; g0 < ngroups; g0 += G * nthreads) {
        uint32_t base[G], e[G], bsw[G];
        double2 a[G][NK];
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
            // insert zero bits at the register-bit positions (ascending)
            uint32_t b = (uint32_t)(g0 + gi * nthreads);
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                const uint32_t low = b & ((1u << rd.rb[c]) - 1u);
                b = ((b >> rd.rb[c]) << (rd.rb[c] + 1)) | low;
            }
            base[gi] = b;
            e[gi] = ext_base | b;
            bsw[gi] = tile_sw(b);
#pragma unroll
            for (int k = 0; k < NK; ++k) a[gi][k] = tile[bsw[gi] ^ ksw[k]];
        }
        steps(a, e, base, ko);
#pragma unroll
        for (int gi = 0; gi < G; ++gi)
#pragma unroll
            for (int k = 0; k < NK; ++k) tile[bsw[gi] ^ ksw[k]] = a[gi][k];
    }
}

// Interpreted round.  The round record is copied to registers first: the tile stores would otherwise
// force the compiler to reload it (same address space).
template <int NB, int G>
B200SQ_HD void exec_round_nb(double2* __restrict__ tile, const DRound* __restrict__ rdp,
                             const DStep* __restrict__ steps, const double2* __restrict__ mats,
                             uint32_t ext_base, int T, int tid, int nthreads) {
    const DRound rd = *rdp;
    const int s0 = rd.s0, s1 = rd.s1;
    exec_round_with<NB, G>(tile, rd, ext_base, T, tid, nthreads,
                           [&](double2 (&a)[G][1 << NB], const uint32_t (&e)[G], const uint32_t (&base)[G],
                               const uint32_t (&ko)[1 << NB]) {
                               for (int si = s0; si < s1; ++si) {
                                   const DStep u = steps[si];
                                   apply_step<NB, G>(a, u, mats + u.mat, e, base, ko);
                               }
                           });
}

B200SQ_HD void exec_round_thread(double2* tile, const DRound* rd, const DStep* steps,
                                 const double2* mats, uint32_t ext_base, int T, int tid,
                                 int nthreads) {
    switch (rd->nb) {
        case 3:
            if (((1 << (T - 3)) % (2 * nthreads)) == 0)
                exec_round_nb<3, 2>(tile, rd, steps, mats, ext_base, T, tid, nthreads);
            else
                exec_round_nb<3, 1>(tile, rd, steps, mats, ext_base, T, tid, nthreads);
            break;
        case 2: exec_round_nb<2, 1>(tile, rd, steps, mats, ext_base, T, tid, nthreads); break;
        default: exec_round_nb<1, 1>(tile, rd, steps, mats, ext_base, T, tid, nthreads); break;
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
