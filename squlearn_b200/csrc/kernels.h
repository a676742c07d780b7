// Internal launcher interface between api.cu and the kernel translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "plan.hpp"

namespace b200sq {

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
    int32_t store;             // write the tile to dst
    int32_t n_terms;           // > 0: reduce the tile to expectation values (requires T == n)
    const uint32_t* xmask;     // device
    const uint32_t* zmask;
    double* out;               // [n_variants_total][n_total][n_terms]
    int32_t team_size, log2ts, teams_per_cta, team_elems;
    int64_t total_items;
};

// Returns cudaError_t as int.
int launch_tile_pass(const TileArgs& args, cudaStream_t stream);
void tile_pass_geometry(const DPass& pass, TileArgs* args);

int launch_expvals(const double2* psi, int64_t n_states, int n_qubits, int n_terms,
                   const uint32_t* xmask_dev, const uint32_t* zmask_dev, double* out,
                   int64_t out_stride, cudaStream_t stream);

int launch_gram(const double2* x, int64_t nx, const double2* y, int64_t ny, int64_t dim, double* k,
                int64_t ldk, uint32_t flags, cudaStream_t stream);

// Split-integer (INT8 tcgen05) Gram: see gram_i8.cu.  Falls back to nothing: callers check gram_i8_supported.
bool gram_i8_supported(int64_t dim);
int launch_gram_i8(const double2* x, int64_t nx, const double2* y, int64_t ny, int64_t dim, double* k,
                   int64_t ldk, uint32_t flags, cudaStream_t stream);

int launch_rbf(const double* fx, int64_t nx, const double* fy, int64_t ny, int nf, double gamma,
               double* k, int64_t ldk, cudaStream_t stream);

void count_launch();

}  // namespace b200sq
