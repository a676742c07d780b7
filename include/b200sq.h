/*
 * b200sq — B200-native batched statevector engine: public C ABI.
 *
 * This is the drop-in boundary for sQUlearn's statevector hot path.  The reference has no
 * C/FFI interface of its own (it is 100 % Python); each entry point below states which
 * reference call it replaces (paths relative to the sQUlearn v0.11.0 source tree).
 *
 * Conventions
 *   - plain C types only; every `*_dev` pointer is a CUDA device pointer, every other pointer
 *     is host memory; `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - all device work is enqueued asynchronously on `stream`; nothing synchronises.
 *   - return value: 0 on success, non-zero error code otherwise; b200sq_last_error() gives the
 *     message of the last failure on the calling thread.
 *   - handles are not thread-safe; use one program handle per host thread.
 *   - qubit order is little-endian (qubit 0 = least significant bit of the amplitude index),
 *     i.e. Qiskit / Qulacs order.  A batch of states is `N x 2^n` complex128, one state per row,
 *     (re, im) interleaved: exactly the (N, 2^n) array FidelityKernelStatevector consumes
 *     (src/squlearn/kernel/lowlevel_kernel/fidelity_kernel_statevector.py:318-352), up to the
 *     PennyLane bit reversal documented in INTEGRATION.md.
 */
#ifndef B200SQ_PUBLIC_API_H_
#define B200SQ_PUBLIC_API_H_

#if !defined(__CUDACC_RTC__) /* NVRTC (run-time specialised kernels) brings its own fixed-width types */
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200SQ_API __attribute__((visibility("default")))
#else
#define B200SQ_API
#endif

#define B200SQ_ABI_VERSION 3
#define B200SQ_MAX_QUBITS 26

/* Gate vocabulary = the reference's gate table
 * (src/squlearn/util/pennylane/pennylane_gates.py:56-93; Qulacs subset
 * src/squlearn/util/qulacs/qulacs_gates.py:105-122).  For controlled gates the first listed
 * qubit(s) are the control(s) (pennylane_circuit.py:397-401). */
enum b200sq_gate_kind {
    B200SQ_ID = 0, B200SQ_H, B200SQ_X, B200SQ_Y, B200SQ_Z, B200SQ_S, B200SQ_SDG, B200SQ_T,
    B200SQ_TDG, B200SQ_SX,
    B200SQ_RX, B200SQ_RY, B200SQ_RZ, B200SQ_P, B200SQ_U,
    B200SQ_CX, B200SQ_CY, B200SQ_CZ, B200SQ_CH, B200SQ_CS, B200SQ_CSX,
    B200SQ_CP, B200SQ_CRX, B200SQ_CRY, B200SQ_CRZ,
    B200SQ_SWAP, B200SQ_ISWAP, B200SQ_ECR,
    B200SQ_RXX, B200SQ_RYY, B200SQ_RZZ, B200SQ_RZX,
    B200SQ_CCX, B200SQ_CSWAP,
    B200SQ_NUM_GATE_KINDS
};

/* How a gate angle depends on the data point.  theta_i = c0 + c1 * f(x[i][feat]).
 * Replaces the sympy-lambdified angle functions the reference evaluates per gate on (N,)
 * arrays (pennylane_circuit.py:361-378, :570-579).  The library circuits only use
 * p, x, p*x, p*arccos(x), p*arctan(x); anything else goes through B200SQ_FN_TABLE. */
enum b200sq_angle_fn {
    B200SQ_FN_CONST = 0, /* theta = c0                                   */
    B200SQ_FN_LINEAR,    /* theta = c0 + c1 * x[feat]                    */
    B200SQ_FN_ARCCOS,    /* theta = c0 + c1 * arccos(x[feat])            */
    B200SQ_FN_ARCTAN,    /* theta = c0 + c1 * arctan(x[feat])            */
    B200SQ_FN_TABLE      /* theta = angle_table[slot * N + i] (device)   */
};

typedef struct b200sq_gate {
    int32_t kind;    /* enum b200sq_gate_kind                                   */
    int32_t q[3];    /* qubits, unused entries = -1                             */
    int32_t fn;      /* enum b200sq_angle_fn                                    */
    int32_t feat;    /* feature column for LINEAR / ARCCOS / ARCTAN             */
    int32_t slot;    /* angle-table row for FN_TABLE                            */
    int32_t reserved;
    double c0, c1;   /* angle coefficients (functions of the circuit parameters) */
    double c2, c3;   /* phi, lambda of the U gate (constants)                   */
} b200sq_gate;

typedef struct b200sq_program b200sq_program;

/* Gram flags */
#define B200SQ_GRAM_SYMMETRIC 1u   /* psi_x == psi_y: compute lower-triangle tiles, mirror      */
#define B200SQ_GRAM_DIAG_ONE 2u    /* symmetric: force K[i][i] = 1.0 (reference default)        */
#define B200SQ_GRAM_DIAG_ZERO 4u   /* symmetric: leave K[i][i] = 0.0 (reference quirk, "all")   */
#define B200SQ_GRAM_3M 8u          /* 3-multiplication complex product (fewer flops)            */
#define B200SQ_GRAM_CLASSIC 16u    /* with 3M: use the barrier-pipelined kernel, not the          */
                                   /* warp-specialised bulk-copy one (kept for A/B measurements) */
#define B200SQ_GRAM_INT8 32u       /* split-integer FP64 emulation on the INT8 tcgen05 tensor     */
                                   /* cores (47-bit fixed point per state row; |dK| ~ 1e-12, see  */
                                   /* DESIGN.md); needs dim >= 64, otherwise the FP64 path runs   */

#define B200SQ_GRAM_CRT 64u        /* with B200SQ_GRAM_INT8: the residue-number ("Ozaki II") form of the scheme:  */
                                   /* fixed point relative to the row 2-norm (44 bits at 2^16 amplitudes), one     */
                                   /* int8 product per modulus (11 there) and a CRT reconstruction instead of 22   */
                                   /* digit-pair products; worst-case |dK| <= 2 sqrt(2 dim) 2^-B (b200sq_gram_crt_bits) */

#if !defined(__CUDACC_RTC__) /* host entry points: invisible to the run-time compiled device code */

B200SQ_API const char* b200sq_last_error(void);
B200SQ_API int b200sq_abi_version(void);

/* Number of CUDA devices visible; 0 when none (no CPU fallback exists: every compute entry
 * point then fails with an error). */
B200SQ_API int b200sq_device_count(void);

/* Compile a gate list (the IR PennyLaneCircuit.build_circuit_instructions produces,
 * pennylane_circuit.py:249-415) into a pass/round/micro-op plan for the tile kernel.
 * tile_bits = 0 picks the default (12, or n if smaller). Host-only: needs no GPU. */
B200SQ_API int b200sq_program_create(int n_qubits, int n_gates, const b200sq_gate* gates, int tile_bits,
                          b200sq_program** out);
B200SQ_API void b200sq_program_destroy(b200sq_program* prog);

/* Update the angle coefficients (new circuit parameters) without re-planning.
 * c0, c1: arrays of n_gates doubles. */
B200SQ_API int b200sq_program_set_coeffs(b200sq_program* prog, const double* c0, const double* c1);

/* Introspection (host-only; used by tests): number of passes over the state the plan makes,
 * and a dump of the plan as int32 words (see csrc/plan.hpp for the layout). */
B200SQ_API int b200sq_program_num_passes(const b200sq_program* prog);
B200SQ_API int b200sq_program_dump(const b200sq_program* prog, int32_t* buf, int capacity, int* n_written);

/* Large batches run pass kernels that are specialised for the program at run time (NVRTC, see
 * csrc/jit.cu; environment B200SQ_JIT=0 disables it).  This entry point only compiles the kernel of one
 * pass (no GPU needed) and returns the cubin size, or a negative number with the compiler log in `log`. */
B200SQ_API long b200sq_program_jit_compile(const b200sq_program* prog, int pass, char* log, int log_capacity);

/* Apply the program to N data points: psi[v][i][:] = U(x_i; shift_v)|0...0>.
 * Replaces Executor.pennylane_execute(circuit returning qml.state(), ...) /
 * qulacs_evaluate_statevector (src/squlearn/util/executor.py:908-939,
 * src/squlearn/util/qulacs/qulacs_execution.py:45-63).
 *   x_dev            [N][d] float64 features (row-major), may be NULL when d == 0
 *   angle_table_dev  [n_slots][N] float64 or NULL
 *   n_variants       >= 1; variant v adds var_shift[v] to the angle of gate var_gate[v]
 *                    (var_gate[v] < 0: unshifted).  var_* are HOST arrays (NULL => one
 *                    unshifted variant).
 *   psi_dev          [n_variants][N][2^n] complex128 output                                  */
B200SQ_API int b200sq_simulate(b200sq_program* prog, int64_t N, int d, const double* x_dev,
                    const double* angle_table_dev, int n_variants, const int32_t* var_gate,
                    const double* var_shift, void* psi_dev, void* stream);

/* Per-state statistics of a batch of states, reduced by the store epilogue of the last gate pass (atomicMax /
 * atomicAdd; zeroed by b200sq_simulate_rowmax): max_hi = the HIGH 32-bit word of max(|re|, |im|) over the row (it
 * holds the exponent: the power-of-two row scale of the digit slicing), norm2 = sum |amplitude|^2 (the row scale of
 * the CRT slicing comes from the 2-norm).  All-zero statistics mean "not available".                     */
typedef struct b200sq_row_stats {
    double norm2;
    uint32_t max_hi;
    uint32_t reserved;
} b200sq_row_stats;

/* b200sq_simulate that also fills row_stats_dev [n_variants][N], so that b200sq_gram_rowmax / b200sq_gram_slice_*
 * read the states once instead of twice.  The statistics describe psi as written by this call: pass them only with
 * those very states.                                                                                  */
B200SQ_API int b200sq_simulate_rowmax(b200sq_program* prog, int64_t N, int d, const double* x_dev,
                                      const double* angle_table_dev, int n_variants, const int32_t* var_gate,
                                      const double* var_shift, void* psi_dev, b200sq_row_stats* row_stats_dev, void* stream);

/* Simulate and reduce to Pauli expectation values without keeping the states when the state
 * fits one tile (n <= tile_bits); otherwise states go through `workspace_dev`.
 * Replaces the qml.expval stack of PennyLaneCircuit (pennylane_circuit.py:636-658) /
 * qulacs_evaluate (qulacs_execution.py:11-42) and, with variants, the shifted-circuit batch
 * of src/squlearn/util/optree/optree_derivative.py:130-158.
 *   x_mask / z_mask  HOST arrays, one Pauli string per term: bit q of x_mask set for X or Y on
 *                    qubit q, bit q of z_mask set for Z or Y on qubit q
 *   out_dev          [n_variants][N][n_terms] float64                                         */
B200SQ_API int b200sq_simulate_expvals(b200sq_program* prog, int64_t N, int d, const double* x_dev,
                            const double* angle_table_dev, int n_variants,
                            const int32_t* var_gate, const double* var_shift, int n_terms,
                            const uint32_t* x_mask, const uint32_t* z_mask, double* out_dev,
                            void* workspace_dev, size_t workspace_bytes, void* stream);

/* b200sq_simulate_expvals with a SECOND shift per variant (var_gate2 / var_shift2: HOST arrays, may be NULL):
 * variant v also adds var_shift2[v] to the angle of gate var_gate2[v] (the same gate as var_gate[v]: the shifts
 * add).  Two-gate shifted circuits are what second-order derivatives ("dfdxdx", "dfdpdx", "dfdpdp", the Laplacian:
 * src/squlearn/qnn/lowlevel_qnn/evaluation_classes.py:105-164) are assembled from.                        */
B200SQ_API int b200sq_simulate_expvals2(b200sq_program* prog, int64_t N, int d, const double* x_dev,
                                        const double* angle_table_dev, int n_variants, const int32_t* var_gate,
                                        const double* var_shift, const int32_t* var_gate2, const double* var_shift2,
                                        int n_terms, const uint32_t* x_mask, const uint32_t* z_mask,
                                        double* out_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Pauli expectation values of states already in device memory: out[i][t] = <psi_i|P_t|psi_i>. */
B200SQ_API int b200sq_expvals(const void* psi_dev, int64_t N, int n_qubits, int n_terms,
                   const uint32_t* x_mask, const uint32_t* z_mask, double* out_dev, void* stream);

/* Fidelity Gram matrix K[i][j] = |<psi_x[i] | psi_y[j]>|^2 (FP64 tensor-core contraction with the
 * modulus-square fused into the epilogue).  Replaces the O(N^2) Python pair loop of
 * FidelityKernelStatevector.evaluate_kernel_sv (fidelity_kernel_statevector.py:303-310,358-383).
 *   psi_x_dev [Nx][dim], psi_y_dev [Ny][dim] complex128;  K_dev [Nx][ldk] float64              */
B200SQ_API int b200sq_gram(const void* psi_x_dev, int64_t Nx, const void* psi_y_dev, int64_t Ny, int64_t dim,
                double* K_dev, int64_t ldk, uint32_t flags, void* stream);

/* b200sq_gram with the row statistics b200sq_simulate_rowmax produced for psi_x / psi_y (either may be NULL; only
 * the INT8 paths use them).                                                                                 */
B200SQ_API int b200sq_gram_rowmax(const void* psi_x_dev, int64_t Nx, const void* psi_y_dev, int64_t Ny, int64_t dim,
                                  double* K_dev, int64_t ldk, uint32_t flags, const b200sq_row_stats* row_stats_x_dev,
                                  const b200sq_row_stats* row_stats_y_dev, void* stream);

/* The two halves of the B200SQ_GRAM_INT8 path, for callers that reuse or exchange the digit planes (the
 * multi-GPU Gram sends planes, not complex128 states: 12 instead of 16 bytes per amplitude, sliced once).
 *   planes  [12][plane_rows][2 * dim] int8: planes 0..5 = digits of nat = (re, im), 6..11 = digits of
 *           rot = (-im, re).  Re<x|y> = nat(x).nat(y), Im<x|y> = rot(x).nat(y): the ROW operand uses all 12
 *           planes, the COLUMN operand only planes 0..5 (a column block may be a [6][rows][2 * dim] buffer).
 *   expo    [plane_rows] int32: power-of-two scale of every row
 * b200sq_gram_slice writes rows [row0, row0 + N) of the planes from N states; b200sq_gram_from_planes
 * computes K[i][j] = |<x_{x0+i} | y_{y0+j}>|^2 for Nx x Ny rows of the two plane sets (flags as b200sq_gram). */
B200SQ_API size_t b200sq_gram_planes_bytes(int64_t plane_rows, int64_t dim);
B200SQ_API int b200sq_gram_slice(const void* psi_dev, int64_t N, int64_t dim, void* planes_dev, int64_t plane_rows,
                                 int64_t row0, int32_t* expo_dev, void* stream);
B200SQ_API int b200sq_gram_slice_rowmax(const void* psi_dev, int64_t N, int64_t dim, void* planes_dev,
                                        int64_t plane_rows, int64_t row0, int32_t* expo_dev,
                                        const b200sq_row_stats* row_stats_dev, void* stream);
B200SQ_API int b200sq_gram_from_planes(const void* planes_x_dev, const int32_t* expo_x_dev, int64_t plane_rows_x,
                                       int64_t x0, int64_t Nx, const void* planes_y_dev, const int32_t* expo_y_dev,
                                       int64_t plane_rows_y, int64_t y0, int64_t Ny, int64_t dim, double* K_dev,
                                       int64_t ldk, uint32_t flags, void* stream);

/* CRT form of the plane interface (B200SQ_GRAM_CRT): planes are [2 s][plane_rows][2 * dim] int8 with
 * s = b200sq_gram_crt_moduli(dim) residue planes of nat followed by s of rot (column operands: the first s);
 * b200sq_gram_from_planes / b200sq_gram_jobs take such planes when B200SQ_GRAM_CRT is set in the (job) flags.
 * A row is quantised as round(v * 2^(B - e)) with B = b200sq_gram_crt_bits(dim) and e (expo) the smallest integer
 * that keeps 2^(B - e) |row|_2 below sqrt(P / 2), P the product of the s moduli; s is the smallest count whose B
 * holds the worst-case error 2 sqrt(2 dim) 2^-B of a unit-norm Gram entry below 1e-10 (B200SQ_CRT_TOL).      */
B200SQ_API int b200sq_gram_crt_moduli(int64_t dim);
B200SQ_API int b200sq_gram_crt_bits(int64_t dim);
B200SQ_API int b200sq_gram_slice_crt(const void* psi_dev, int64_t N, int64_t dim, void* planes_dev,
                                     int64_t plane_rows, int64_t row0, int32_t* expo_dev,
                                     const b200sq_row_stats* row_stats_dev, void* stream);

/* b200sq_gram_slice_crt that also writes the 6 nat DIGIT planes ([6][digit_rows][2 * dim]: the compact exchange
 * format, 6 instead of s bytes per component) from the same read of the states: the balanced base-256 digits of the
 * CRT integers themselves, so the residues b200sq_gram_digits_to_crt derives from them are the sender's.     */
B200SQ_API int b200sq_gram_slice_crt_digits(const void* psi_dev, int64_t N, int64_t dim, void* planes_dev,
                                            int64_t plane_rows, int64_t row0, int32_t* expo_dev,
                                            const b200sq_row_stats* row_stats_dev, void* digit_planes_dev, int64_t digit_rows,
                                            int64_t digit_row0, void* stream);

/* Residue planes of a COLUMN block from its 6 nat digit planes (what a digit exchange delivers: 6 instead of
 * s bytes per component): rows [row0_d, row0_d + N) of digit planes [>= 6][plane_rows_d][2 * dim] ->
 * rows [row0_o, row0_o + N) of residue planes [s][plane_rows_o][2 * dim] (nat residues only; the row exponents of
 * the digit planes stay valid).                                                                          */
B200SQ_API int b200sq_gram_digits_to_crt(const void* digit_planes_dev, int64_t plane_rows_d, int64_t row0_d, int64_t N,
                                         int64_t dim, void* crt_planes_dev, int64_t plane_rows_o, int64_t row0_o,
                                         void* stream);

/* Several Gram blocks that share the row operand, in ONE persistent launch (the multi-GPU Gram: the diagonal
 * block of a rank plus every block it owns against received column planes; there is no reference counterpart,
 * the reference evaluates the pair loop of fidelity_kernel_statevector.py:358-383 on one host).
 *   ready_flag_dev / ready_value   optional: the job's column planes are complete when *ready_flag_dev >=
 *           ready_value (a uint32 in device memory written by whoever delivers the planes, e.g. a peer GPU's
 *           copy engine after its push); the kernel starts the job's tiles when it observes the flag, so blocks
 *           are multiplied while later ones are still in flight.  The writer must be independent of this stream.
 *   align_*                        optional, fused into the |.|^2 epilogue (kernel target alignment,
 *           src/squlearn/kernel/loss/target_alignment.py:75-90): with row labels yx[i] (i < nx) and column
 *           labels yy[j] (j < ny) of THIS block, align_out[0] += w * sum_ij K_ij yx_i yy_j and
 *           align_out[1] += w * sum_ij K_ij^2, so a training step never re-reads K.  For a symmetric block the
 *           mirrored off-diagonal entries are counted (w = 1 on the diagonal tiles' entries, 2 elsewhere);
 *           for a rectangular block w = align_weight (2 when the block also stands for its transpose).     */
typedef struct b200sq_gram_job {
    int64_t x0, nx;                 /* rows [x0, x0 + nx) of the row planes                              */
    const void* planes_y_dev;       /* column planes ([>= 6][plane_rows_y][2 * dim] int8)                */
    const int32_t* expo_y_dev;
    int64_t plane_rows_y, y0, ny;
    double* K_dev;                  /* [nx][ldk]                                                         */
    int64_t ldk;
    uint32_t flags;                 /* B200SQ_GRAM_SYMMETRIC | DIAG_* (symmetric: column planes == row planes, y0 == x0) */
    uint32_t ready_value;
    const uint32_t* ready_flag_dev; /* NULL: no wait                                                     */
    const double* align_yx_dev;     /* NULL: no reduction                                                */
    const double* align_yy_dev;
    double* align_out_dev;          /* [2] float64, accumulated with atomics (zero it first)             */
    double align_weight;
    double* K_mirror_dev;           /* NULL, or (rectangular jobs): also store the transpose,            */
    int64_t ldk_mirror;             /* K_mirror[j][i] = K[i][j] — may be a peer mapping (NVLink stores)  */
} b200sq_gram_job;
B200SQ_API int b200sq_gram_jobs(const void* planes_x_dev, const int32_t* expo_x_dev, int64_t plane_rows_x,
                                int64_t dim, int n_jobs, const b200sq_gram_job* jobs, void* stream);

/* Enqueue a wait on `stream` until flags_dev[which[k]] >= value for every k < n (n <= 32; uint32 flags in
 * device memory written by peers, e.g. "my mirrored blocks are stored"); gives up after 20 s like the ready
 * flags of the Gram jobs (counted by b200sq_gram_ready_timeouts).                                        */
B200SQ_API int b200sq_wait_flags(const uint32_t* flags_dev, const int32_t* which, int n, uint32_t value,
                                 void* stream);

/* A job whose ready flag does not arrive within 20 s is multiplied anyway (the GPU must not hang on a lost
 * peer) and counted: this returns the count of such jobs since the library was loaded (it synchronises the
 * device; callers that use ready flags check it after a step and treat non-zero as an error).           */
B200SQ_API int64_t b200sq_gram_ready_timeouts(void);

/* Peer plumbing of the multi-GPU Gram (one process per GPU): export a device allocation to the other ranks of
 * the box (CUDA IPC), open a peer's export, and enqueue copy-engine copies (device-to-device, local or into a
 * peer mapping over NVLink) on a stream.  b200sq_ipc_export returns the 64-byte handle of the ALLOCATION that
 * contains dev_ptr and the offset of dev_ptr inside it; the peer adds the offset to the base it opens.   */
B200SQ_API int b200sq_ipc_export(const void* dev_ptr, void* handle_out_64, int64_t* offset_out);
B200SQ_API int b200sq_ipc_open(const void* handle_64, void** base_out);
B200SQ_API int b200sq_ipc_close(void* base);
B200SQ_API int b200sq_copy_async(void* dst_dev, const void* src_dev, size_t bytes, void* stream);
/* n_planes copies of `bytes` bytes each, plane p from src + p * src_pitch to dst + p * dst_pitch: a row range of
 * every plane of a plane set in one call (one copy-engine copy per plane).                                */
B200SQ_API int b200sq_copy_planes_async(void* dst_dev, size_t dst_pitch, const void* src_dev, size_t src_pitch,
                                        size_t bytes, int n_planes, void* stream);

/* Gaussian outer kernel of the projected quantum kernel: K[i][j] = exp(-gamma |F_x[i]-F_y[j]|^2)
 * (src/squlearn/kernel/lowlevel_kernel/projected_quantum_kernel.py:945-973).
 *   fx_dev [Nx][nf], fy_dev [Ny][nf] float64; K_dev [Nx][ldk] float64                          */
B200SQ_API int b200sq_rbf(const double* fx_dev, int64_t Nx, const double* fy_dev, int64_t Ny, int nf,
               double gamma, double* K_dev, int64_t ldk, void* stream);

/* Launch counter: number of kernels this library has launched in this process (bench.py's
 * `gpu_launches`), and the duration bookkeeping of the last Gram / gate-pass launches is left to
 * the caller's CUDA events. */
B200SQ_API int64_t b200sq_launch_count(void);

#endif /* !__CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* B200SQ_PUBLIC_API_H_ */
