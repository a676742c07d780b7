// Device-visible part of the plan: step / matrix kinds, the packed records of a pass blob and the gate
// table.  Plain C++ without the standard library, so that it also compiles under NVRTC (the
// specialised pass kernels of jit.cu include it at run time).
#pragma once

#if defined(__CUDACC_RTC__)
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#define B200SQ_PUBLIC_API_TYPES_ONLY 1
#else
#include <stdint.h>
#endif

#include "../../include/b200sq.h"

namespace b200sq {

enum StepKind : int32_t {
    ST_U2 = 1,      // 4x4 on register slots r0 > r1 (r0 = matrix MSB), optional control mask
    ST_DIAG = 2,    // single diagonal gate; r0 / r1 = ext positions of its qubits
    ST_U1CTRL = 5,  // general 2x2 on register slot r0 under a control mask (cx, cy, ch, crx, ccx ...)
    ST_LRX = 6,     // [[c, -is], [-is, c]] on every register slot of mask r0 (matrix of slot c at mat + 4c)
    ST_LREAL = 7,   // real 2x2 (ry, h, x)
    ST_LGEN = 8,    // general complex 2x2 (y, sx, u)
    ST_DTAB = 9,    // a[k] *= table[(local_index >> r0) & ((1 << r1) - 1)], table at mat (+ variant << r1, see DTab)
    ST_LROTX = 10,  // rx as a SCALED rotation: one FMA per real component (2 per amplitude), slot mask r0
    ST_LROTY = 11,  // ry, idem
};
// MAT_ROT (scaled rotation): with c = cos(theta/2), s = sin(theta/2) the gate is written as scale * (I + t G) with
// the larger of |c|, |s| factored out, so that every output component is ONE fused multiply-add:
//   |c| >= |s|:  m[0] = (s / c, 0), m[1].x = c      rx: v0' = v0 - i t v1, v1' = v1 - i t v0;  ry: v0' = v0 - t v1, v1' = v1 + t v0
//   |s| >  |c|:  m[0] = (c / s, 1), m[1].x = s      rx: v0' = u v0 - i v1, v1' = u v1 - i v0;  ry: v0' = u v0 - v1, v1' = u v1 + v0
// (|t|, |u| <= 1: intermediates grow by at most sqrt(2) per gate).  The product of the factored-out scales of all
// MAT_ROT gates of a pass is one real number per state; it multiplies into the pass's carrier table (the product
// table of the fresh qubits, or the first phase table), which is built per state anyway.
enum MatKind : int32_t { MAT_U1 = 0, MAT_U2 = 1, MAT_DIAG = 2, MAT_ROT = 3 };

constexpr int kMaxRuns = 8;
constexpr int kMaxTileBits = 13;
constexpr int kMaxPassSteps = 160;
constexpr int kMaxPassRounds = 64;
constexpr int kMaxPassMats = 224;
constexpr int kMaxInit = 128;
constexpr int kMaxPassMatElems = 4096;  // double2 elements of matrix / table storage per team (64 KiB)
constexpr int kNoBit = 255;

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

// ---------------------------------------------------------------------------------------------
// Device view of a pass: one contiguous blob (header + arrays) that the kernel copies to shared
// memory.  The same blob drives the host emulator of the CPU tests.

struct DStep {       // 16 bytes
    uint8_t kind, r0, swap, pad0;
    int8_t r1, pad1;
    uint16_t gate;
    uint32_t ctrl;
    uint16_t mat;
    uint16_t pat;
};
struct DRound {      // 24 bytes
    uint8_t nb, rb[3];
    uint16_t s0, s1;  // step range relative to the pass
    uint16_t ksw[8];  // swizzled SMEM offset of register slot k: ko[k] ^ ((ko[k] >> 3) & 7)
};
struct DMat {        // 8 bytes
    uint16_t gate;
    uint8_t kind, swap;
    uint16_t mat, pad;
};
struct DTabGate {    // 4 bytes
    uint16_t m4;
    uint8_t p0, p1;  // ext positions (p1 = kNoBit for one-qubit gates)
};
struct DTab {        // 12 bytes
    uint16_t mat;
    uint8_t lo, nbits;  // nbits bit 7 (legacy form): the table reads > 2 outer index bits and is rebuilt per tile
    uint16_t g0, g1;
    uint8_t vb[2];      // ext positions of the (at most two) OUTER index bits the table reads, kNoBit = none:
    uint8_t carrier;    // the table is stored in 2^(#vb) variants (variant v at mat + (v << nbits)), all built once
    uint8_t pad;        // per state, and a tile picks its variant from its outer bits.  carrier: multiply by the
};                      // pass's rotation scale (see MAT_ROT)
struct DInit {       // a 1-qubit gate folded into the product start of a fresh tile qubit
    uint16_t gate;
    uint8_t qubit, pad;
};
struct DPass {
    int32_t T, n_rounds, n_steps, n_mats, n_tabs, n_tabgates, n_init, mat_elems;
    int32_t off_rounds, off_steps, off_mats, off_tabs, off_tabgates, off_inits, blob_bytes;  // byte offsets
    uint8_t has_src;    // the input layout is not empty (something to load)
    uint8_t n_fresh;    // fresh tile qubits; 0 -> plain pass (cp.async double-buffered loads)
    uint8_t nlo, nhi;   // fresh qubits in the low / high product table
    uint8_t n_obits;    // outer bits of a work item = outer qubits present in the output layout
    uint8_t in_bits, out_bits;  // log2(amplitudes per state) of the input / output layout
    uint8_t last;
    uint8_t rot_carrier;  // where the MAT_ROT scale goes: 0 none, 1 low product table, 2 a phase table (DTab::carrier)
    uint8_t l_in[16];   // local bit -> bit of the input-layout index (kNoBit: fresh)
    uint8_t l_out[16];  // local bit -> bit of the output-layout index
    uint8_t l_plo[16];  // local bit -> bit of the low product-table index (kNoBit: none)
    uint8_t l_phi[16];  // local bit -> bit of the high product-table index
    uint8_t o_in[32];   // item outer bit -> bit of the input-layout index (kNoBit: qubit still |0> there)
    uint8_t o_out[32];  // item outer bit -> bit of the output-layout index
    uint8_t o_ext[32];  // item outer bit -> ext position
    uint8_t fresh_q[16];  // fresh qubits, ascending (first nlo -> low table, then nhi -> high table)
    uint8_t extq[32];   // ext position -> qubit
};

}  // namespace b200sq
