// Internal launcher interface between api.cu and the kernel translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "plan.hpp"

namespace b200sq {

struct TileArgs {
    DPass pass;                        // this pass, its rounds and micro-ops travel in the kernel
    DRound rounds[kMaxPassRounds];     // parameters (constant bank: uniform, broadcast reads)
    DUop uops[kMaxPassUops];
    DInit inits[kMaxInit];             // product-state prefix (first pass only)
    int32_t n_init;
    const b200sq_gate* gates;          // device gate table (angle coefficients)
    double2* psi;              // states; state (v, i) lives at psi + ((v * n_total + i - psi_off) << n)
    int64_t psi_off;
    const double* x;           // [n_total][d]
    const double* table;       // [slots][n_total]
    int64_t n_total;           // samples in the whole call (row stride of table, variant stride of psi/out)
    int64_t sample0;           // first sample of this launch
    int64_t n_samples;         // samples in this launch
    int32_t n, d;
    int32_t n_variants;        // variants in this launch
    int32_t variant0;          // first variant of this launch
    const int32_t* var_gate;   // device, indexed by absolute variant; nullptr = unshifted
    const double* var_shift;
    int32_t store;             // write the tile back to psi
    int32_t n_terms;           // > 0: reduce the tile to expectation values (requires T == n)
    const uint32_t* xmask;     // device
    const uint32_t* zmask;
    double* out;               // [n_variants_total][n_total][n_terms]
    int32_t team_size, teams_per_cta, team_elems;
    int64_t total_items;
};

// Returns cudaError_t as int.
int launch_tile_pass(const TileArgs& args, cudaStream_t stream);
int tile_pass_smem_bytes(int T, int mat_elems, int first, int* team_size, int* teams_per_cta, int* team_elems);

int launch_expvals(const double2* psi, int64_t n_states, int n_qubits, int n_terms,
                   const uint32_t* xmask_dev, const uint32_t* zmask_dev, double* out,
                   int64_t out_stride, cudaStream_t stream);

int launch_gram(const double2* x, int64_t nx, const double2* y, int64_t ny, int64_t dim, double* k,
                int64_t ldk, uint32_t flags, cudaStream_t stream);

int launch_rbf(const double* fx, int64_t nx, const double* fy, int64_t ny, int nf, double gamma,
               double* k, int64_t ldk, cudaStream_t stream);

void count_launch();

}  // namespace b200sq
