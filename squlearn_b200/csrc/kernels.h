// Internal launcher interface between api.cu and the kernel translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/b200sq.h"
#include "plan.hpp"
#include "tile_kernel.cuh"

namespace b200sq {

// Returns cudaError_t as int.  `jit_kernel`: a run-time specialised kernel of this pass (jit.cu) or nullptr
// for the generic interpreting kernel.
int launch_tile_pass(const TileArgs& args, cudaStream_t stream, const void* jit_kernel = nullptr);

// Run-time specialisation (jit.cu)
const void* jit_pass_kernel(const Plan& plan, int pass, int team_size);
std::string jit_pass_source(const Plan& plan, int pass, int team_size);
long jit_compile_only(const Plan& plan, int pass, int team_size, std::string* log_out);
void tile_pass_geometry(const DPass& pass, TileArgs* args);

int launch_expvals(const double2* psi, int64_t n_states, int n_qubits, int n_terms,
                   const uint32_t* xmask_dev, const uint32_t* zmask_dev, double* out,
                   int64_t out_stride, cudaStream_t stream);

int launch_gram(const double2* x, int64_t nx, const double2* y, int64_t ny, int64_t dim, double* k,
                int64_t ldk, uint32_t flags, cudaStream_t stream);

// Split-integer (INT8 tcgen05) Gram: see gram_i8.cu.  Falls back to nothing: callers check gram_i8_supported.
bool gram_i8_supported(int64_t dim);
bool gram_crt_supported(int64_t dim);
int gram_crt_moduli(int64_t dim);   // residue planes per operand half of the CRT scheme (0: unsupported)
int gram_crt_bits(int64_t dim);     // B of the CRT scheme: a row is quantised as round(v * 2^(B - e))
int launch_slice_crt(const double2* psi, int64_t n, int64_t dim, int8_t* planes, int64_t plane_rows, int64_t row0,
                     int* expo, cudaStream_t stream, const RowStats* row_stats = nullptr, int8_t* digits = nullptr,
                     int64_t digit_rows = 0, int64_t digit_row0 = 0);
int launch_wait_flags(const uint32_t* flags, const int32_t* which, int n, uint32_t value, cudaStream_t stream);
int launch_digits_to_crt(const int8_t* digits, int64_t rows_d, int64_t row0_d, int64_t n, int64_t dim, int8_t* out,
                         int64_t rows_o, int64_t row0_o, cudaStream_t stream);
unsigned int gram_ready_timeouts();   // jobs whose ready flag never arrived (synchronises the device)
int launch_gram_i8(const double2* x, int64_t nx, const double2* y, int64_t ny, int64_t dim, double* k,
                   int64_t ldk, uint32_t flags, cudaStream_t stream, const RowStats* row_stats_x = nullptr,
                   const RowStats* row_stats_y = nullptr);

// Digit planes of n states into rows [row0, row0 + n) of `planes` ([12][plane_rows][2 dim] int8) and their
// power-of-two row scales into expo[row0 ...]; Gram from planes.
int launch_slice_i8(const double2* psi, int64_t n, int64_t dim, int8_t* planes, int64_t plane_rows, int64_t row0,
                    int* expo, cudaStream_t stream, const RowStats* row_stats = nullptr);
int launch_gram_jobs(const int8_t* px, const int* ex, int64_t rows_x, int64_t dim, int n_jobs,
                     const b200sq_gram_job* jobs, cudaStream_t stream);
int launch_gram_planes(const int8_t* px, const int* ex, int64_t rows_x, int64_t x0, int64_t nx, const int8_t* py,
                       const int* ey, int64_t rows_y, int64_t y0, int64_t ny, int64_t dim, double* k, int64_t ldk,
                       uint32_t flags, cudaStream_t stream);

int launch_rbf(const double* fx, int64_t nx, const double* fy, int64_t ny, int nf, double gamma,
               double* k, int64_t ldk, cudaStream_t stream);

void count_launch();

}  // namespace b200sq
