// C ABI of libb200sq (see include/b200sq.h).  Host-side orchestration only: plan handling,
// device mirrors of the plan, launch sequencing.  There is no CPU compute path in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/b200sq.h"
#include "kernels.h"

namespace b200sq {

static thread_local std::string g_error;
static std::atomic<int64_t> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static int fail(int code, const std::string& msg) {
    g_error = msg;
    return code ? code : -1;
}
static int cuda_fail(int e, const char* what) {
    return fail(1000 + e, std::string(what) + ": " + cudaGetErrorString((cudaError_t)e));
}

}  // namespace b200sq

using namespace b200sq;

struct b200sq_program {
    Plan plan;
    // device mirrors (lazily created on the device that is current at first use)
    int device = -1;
    b200sq_gate* d_gates = nullptr;
    bool gates_dirty = true;
    std::vector<std::vector<uint8_t>> blobs;  // one device blob per pass (plan.hpp: build_pass_blob)
    unsigned char* d_blobs = nullptr;
    std::vector<size_t> blob_off;
    std::vector<const void*> jit_kernels;  // per pass: run-time specialised kernel (jit.cu), or nullptr
    std::vector<char> jit_tried;
    int32_t* d_var_gate = nullptr;   // [2][var_cap]: first / second shifted gate of every variant
    double* d_var_shift = nullptr;
    int var_cap = 0;
    bool have_var2 = false;
    uint32_t* d_masks = nullptr;
    int mask_cap = 0;

    ~b200sq_program() {
        if (device >= 0) {
            int cur = -1;
            cudaGetDevice(&cur);
            cudaSetDevice(device);
            cudaFree(d_gates);
            cudaFree(d_blobs);
            cudaFree(d_var_gate);
            cudaFree(d_var_shift);
            cudaFree(d_masks);
            if (cur >= 0) cudaSetDevice(cur);
        }
    }
};

namespace {

int ensure_device(b200sq_program* p, cudaStream_t stream) {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice (no CUDA device: b200sq has no CPU path)");
    if (p->device >= 0 && p->device != dev)
        return fail(3, "program handle was first used on another device; create one handle per device");
    if (p->device < 0) {
        const Plan& pl = p->plan;
        size_t ng = std::max<size_t>(pl.gates.size(), 1);
        if ((e = cudaMalloc(&p->d_gates, ng * sizeof(b200sq_gate))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
        size_t total = 0;
        p->blob_off.clear();
        for (const auto& b : p->blobs) {
            p->blob_off.push_back(total);
            total += (b.size() + 255) & ~(size_t)255;
        }
        if ((e = cudaMalloc(&p->d_blobs, std::max<size_t>(total, 256))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
        for (size_t i = 0; i < p->blobs.size(); ++i) {
            e = cudaMemcpyAsync(p->d_blobs + p->blob_off[i], p->blobs[i].data(), p->blobs[i].size(),
                                cudaMemcpyHostToDevice, stream);
            if (e != cudaSuccess) return cuda_fail(e, "upload pass blobs");
        }
        p->device = dev;
        p->gates_dirty = true;
        // The inter-pass workspaces come from the device's default stream-ordered pool: keep freed
        // blocks cached across synchronisation points instead of returning them to the driver.
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    if (p->gates_dirty && !p->plan.gates.empty()) {
        e = cudaMemcpyAsync(p->d_gates, p->plan.gates.data(), p->plan.gates.size() * sizeof(b200sq_gate),
                            cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return cuda_fail(e, "upload gates");
    }
    p->gates_dirty = false;
    return 0;
}

int upload_variants(b200sq_program* p, int n_variants, const int32_t* var_gate, const double* var_shift,
                    cudaStream_t stream, const int32_t* var_gate2 = nullptr, const double* var_shift2 = nullptr) {
    p->have_var2 = false;
    if (!var_gate) return var_gate2 ? fail(2, "a second shift needs a first one") : 0;
    if (!var_shift) return fail(2, "var_shift is NULL but var_gate is not");
    if (var_gate2 && !var_shift2) return fail(2, "var_shift2 is NULL but var_gate2 is not");
    for (int v = 0; v < n_variants; ++v)
        if (var_gate[v] >= (int)p->plan.gates.size() || (var_gate2 && var_gate2[v] >= (int)p->plan.gates.size()))
            return fail(2, "variant gate index out of range");
    if (n_variants > p->var_cap) {
        cudaFree(p->d_var_gate);
        cudaFree(p->d_var_shift);
        p->var_cap = 0;
        cudaError_t e;
        if ((e = cudaMalloc(&p->d_var_gate, 2 * (size_t)n_variants * sizeof(int32_t))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
        if ((e = cudaMalloc(&p->d_var_shift, 2 * (size_t)n_variants * sizeof(double))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
        p->var_cap = n_variants;
    }
    cudaMemcpyAsync(p->d_var_gate, var_gate, n_variants * sizeof(int32_t), cudaMemcpyHostToDevice, stream);
    cudaError_t e = cudaMemcpyAsync(p->d_var_shift, var_shift, n_variants * sizeof(double),
                                    cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess && var_gate2) {
        cudaMemcpyAsync(p->d_var_gate + p->var_cap, var_gate2, n_variants * sizeof(int32_t), cudaMemcpyHostToDevice, stream);
        e = cudaMemcpyAsync(p->d_var_shift + p->var_cap, var_shift2, n_variants * sizeof(double),
                            cudaMemcpyHostToDevice, stream);
        p->have_var2 = true;
    }
    if (e != cudaSuccess) return cuda_fail(e, "upload variants");
    return 0;
}

int upload_masks(b200sq_program* p, int n_terms, const uint32_t* xm, const uint32_t* zm, int n,
                 cudaStream_t stream) {
    const uint32_t full = (n >= 32) ? ~0u : ((1u << n) - 1u);
    for (int t = 0; t < n_terms; ++t)
        if ((xm[t] | zm[t]) & ~full) return fail(2, "Pauli mask addresses a qubit outside the register");
    if (2 * n_terms > p->mask_cap) {
        cudaFree(p->d_masks);
        p->mask_cap = 0;
        cudaError_t e = cudaMalloc(&p->d_masks, 2 * (size_t)n_terms * sizeof(uint32_t));
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
        p->mask_cap = 2 * n_terms;
    }
    cudaMemcpyAsync(p->d_masks, xm, n_terms * sizeof(uint32_t), cudaMemcpyHostToDevice, stream);
    cudaError_t e = cudaMemcpyAsync(p->d_masks + n_terms, zm, n_terms * sizeof(uint32_t),
                                    cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return cuda_fail(e, "upload masks");
    return 0;
}

// Run every pass of the plan on samples [s0, s0 + ns) and variants [v0, v0 + nv).
// Passes whose output layout is the full state write psi; compact inter-pass layouts go through a
// stream-ordered scratch allocation (at most two buffers, ping-pong).
int run_passes(b200sq_program* p, int64_t n_total, int64_t s0, int64_t ns, int v0, int nv, int d,
               const double* x, const double* table, bool have_variants, double2* psi, int64_t psi_off,
               bool fuse_expvals, int n_terms, double* out, cudaStream_t stream, RowStats* row_stats = nullptr) {
    const Plan& pl = p->plan;
    const int64_t n_states = (int64_t)nv * ns;
    double2* ws[2] = {nullptr, nullptr};
    int flip = 0;
    const double2* src = nullptr;
    int64_t src_vstride = 0, src_off = 0;
    int src_bits = 0;
    int rc = 0;
    for (size_t ip = 0; ip < pl.passes.size() && rc == 0; ++ip) {
        TileArgs a;
        std::memset(&a, 0, sizeof(a));
        const DPass& P = *reinterpret_cast<const DPass*>(p->blobs[ip].data());
        a.blob = p->d_blobs + p->blob_off[ip];
        a.blob_bytes = P.blob_bytes;
        const bool last = ip + 1 == pl.passes.size();
        const bool to_out = fuse_expvals && last;
        if (P.has_src && P.in_bits != src_bits) { rc = fail(5, "internal: layout chain broken"); break; }
        if (to_out) {
            a.dst = nullptr;
        } else if (P.out_bits == pl.n) {
            a.dst = psi;
            a.dst_vstride = n_total;
            a.dst_off = psi_off;
        } else if (src && P.out_bits == src_bits && P.n_fresh == 0) {
            a.dst = const_cast<double2*>(src);  // same layout: in place
            a.dst_vstride = src_vstride;
            a.dst_off = src_off;
        } else {
            if (ws[flip]) { cudaFreeAsync(ws[flip], stream); ws[flip] = nullptr; }
            cudaError_t e = cudaMallocAsync(&ws[flip], ((size_t)n_states << P.out_bits) * sizeof(double2), stream);
            if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMallocAsync (inter-pass workspace)"); break; }
            a.dst = ws[flip];
            a.dst_vstride = ns;
            a.dst_off = (int64_t)v0 * ns + s0;
            flip ^= 1;
        }
        a.src = src;
        a.src_vstride = src_vstride;
        a.src_off = src_off;
        a.gates = p->d_gates;
        a.x = x;
        a.table = table;
        a.n_total = n_total;
        a.sample0 = s0;
        a.n_samples = ns;
        a.n = pl.n;
        a.d = d;
        a.n_variants = nv;
        a.variant0 = v0;
        a.var_gate = have_variants ? p->d_var_gate : nullptr;
        a.var_shift = have_variants ? p->d_var_shift : nullptr;
        a.var_gate2 = (have_variants && p->have_var2) ? p->d_var_gate + p->var_cap : nullptr;
        a.var_shift2 = (have_variants && p->have_var2) ? p->d_var_shift + p->var_cap : nullptr;
        a.store = to_out ? 0 : 1;
        a.n_terms = to_out ? n_terms : 0;
        a.xmask = p->d_masks;
        a.zmask = p->d_masks ? p->d_masks + n_terms : nullptr;
        a.out = out;
        a.row_stats = (last && a.dst == psi && psi && P.out_bits == pl.n) ? row_stats : nullptr;  // the FINAL amplitudes only
        tile_pass_geometry(P, &a);
        a.total_items = n_states << P.n_obits;
        // Large batches run a kernel specialised for this pass at run time (compiled once per program);
        // B200SQ_JIT = 0 / 1 forces it off / on.
        const void* jit = nullptr;
        {
            const char* mode = getenv("B200SQ_JIT");  // read per call: tests and smoke() toggle it
            const bool big = ((int64_t)n_states << P.out_bits) >= ((int64_t)1 << 25);
            const bool want = (!mode || mode[0] == 'a') ? big : (mode[0] != '0');
            if (want) {
                if (p->jit_tried.size() != pl.passes.size()) {
                    p->jit_tried.assign(pl.passes.size(), 0);
                    p->jit_kernels.assign(pl.passes.size(), nullptr);
                }
                if (!p->jit_tried[ip]) {
                    p->jit_kernels[ip] = jit_pass_kernel(pl, (int)ip, a.team_size);
                    p->jit_tried[ip] = 1;
                }
                jit = p->jit_kernels[ip];
            }
        }
        int e = launch_tile_pass(a, stream, jit);
        if (e) { rc = cuda_fail(e, "tile pass launch"); break; }
        src = a.dst;
        src_vstride = a.dst_vstride;
        src_off = a.dst_off;
        src_bits = P.out_bits;
    }
    for (int i = 0; i < 2; ++i)
        if (ws[i]) cudaFreeAsync(ws[i], stream);
    return rc;
}

}  // namespace

extern "C" {

const char* b200sq_last_error(void) { return g_error.c_str(); }
int b200sq_abi_version(void) { return B200SQ_ABI_VERSION; }
int64_t b200sq_launch_count(void) { return g_launches.load(); }

int b200sq_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int b200sq_program_create(int n_qubits, int n_gates, const b200sq_gate* gates, int tile_bits,
                          b200sq_program** out) {
    if (!out) return fail(2, "out is NULL");
    *out = nullptr;
    if (n_gates < 0 || (n_gates > 0 && !gates)) return fail(2, "bad gate list");
    try {
        std::vector<b200sq_gate> g(gates, gates + n_gates);
        for (const auto& x : g) {
            if (x.fn < B200SQ_FN_CONST || x.fn > B200SQ_FN_TABLE) return fail(2, "unknown angle function");
            if ((x.fn == B200SQ_FN_LINEAR || x.fn == B200SQ_FN_ARCCOS || x.fn == B200SQ_FN_ARCTAN) && x.feat < 0)
                return fail(2, "negative feature index");
            if (x.fn == B200SQ_FN_TABLE && x.slot < 0) return fail(2, "negative angle-table slot");
        }
        b200sq_program* p = new b200sq_program();
        // rx / ry layer steps as scaled rotations (one FMA per component, plan_dev.h MAT_ROT); B200SQ_ROT_LIFT=0 disables
        const char* lift = getenv("B200SQ_ROT_LIFT");
        p->plan = make_plan(n_qubits, g, tile_bits, 4, true, !(lift && lift[0] == '0'));
        for (size_t ip = 0; ip < p->plan.passes.size(); ++ip) p->blobs.push_back(build_pass_blob(p->plan, (int)ip));
        *out = p;
        return 0;
    } catch (const std::exception& e) {
        return fail(2, e.what());
    }
}

void b200sq_program_destroy(b200sq_program* prog) { delete prog; }

int b200sq_program_set_coeffs(b200sq_program* prog, const double* c0, const double* c1) {
    if (!prog || !c0 || !c1) return fail(2, "NULL argument");
    for (size_t i = 0; i < prog->plan.gates.size(); ++i) {
        prog->plan.gates[i].c0 = c0[i];
        prog->plan.gates[i].c1 = c1[i];
    }
    prog->gates_dirty = true;
    return 0;
}

// Compile the run-time specialised kernel of pass `pass` with NVRTC without loading it (needs no GPU):
// returns the cubin size in bytes, or a negative number on failure; the compiler log goes to `log`.
long b200sq_program_jit_compile(const b200sq_program* prog, int pass, char* log, int log_capacity) {
    if (!prog || pass < 0 || pass >= (int)prog->plan.passes.size()) return -2;
    TileArgs a;
    std::memset(&a, 0, sizeof(a));
    tile_pass_geometry(*reinterpret_cast<const DPass*>(prog->blobs[pass].data()), &a);
    std::string lg;
    const long r = jit_compile_only(prog->plan, pass, a.team_size, &lg);
    if (log && log_capacity > 0) {
        std::strncpy(log, lg.c_str(), (size_t)log_capacity - 1);
        log[log_capacity - 1] = 0;
    }
    return r;
}

int b200sq_program_num_passes(const b200sq_program* prog) {
    return prog ? (int)prog->plan.passes.size() : -1;
}

// Dump layout (int32 words): [n, T, n_passes, n_rounds, n_steps] then per pass
// [T, a_in, a_out, last, r0, r1, s0, s1, mat_elems, n_fresh, n_tabs, tile_qubits[T]...], then per round
// [nb, rb0, rb1, rb2, s0, s1], then per step [kind, gate, r0, r1, ctrl, mat, swap], then the folded gates
// [count, gate...].
int b200sq_program_dump(const b200sq_program* prog, int32_t* buf, int capacity, int* n_written) {
    if (!prog || !n_written) return fail(2, "NULL argument");
    std::vector<int32_t> w;
    const Plan& pl = prog->plan;
    w.insert(w.end(), {pl.n, pl.T, (int32_t)pl.passes.size(), (int32_t)pl.rounds.size(), (int32_t)pl.steps.size()});
    for (const Pass& ps : pl.passes) {
        w.insert(w.end(), {ps.T, (int32_t)ps.a_in, (int32_t)ps.a_out, ps.last, ps.r0, ps.r1, ps.s0, ps.s1, ps.mat_elems,
                           __builtin_popcount(ps.L & ~ps.a_in), ps.t1 - ps.t0});
        for (int i = 0; i < ps.T; ++i) w.push_back(ps.tile_qubits[i]);
    }
    for (const Round& r : pl.rounds) w.insert(w.end(), {r.nb, r.rb[0], r.rb[1], r.rb[2], r.s0, r.s1});
    for (const Step& u : pl.steps) w.insert(w.end(), {u.kind, u.gate, u.r0, u.r1, (int32_t)u.ctrl, u.mat, u.swap});
    w.push_back((int32_t)pl.init_gates.size());
    for (int32_t g : pl.init_gates) w.push_back(g);
    *n_written = (int)w.size();
    if (buf && capacity >= (int)w.size()) std::memcpy(buf, w.data(), w.size() * sizeof(int32_t));
    return 0;
}

int b200sq_simulate(b200sq_program* prog, int64_t N, int d, const double* x_dev,
                    const double* angle_table_dev, int n_variants, const int32_t* var_gate,
                    const double* var_shift, void* psi_dev, void* stream_) {
    return b200sq_simulate_rowmax(prog, N, d, x_dev, angle_table_dev, n_variants, var_gate, var_shift, psi_dev,
                                  nullptr, stream_);
}

int b200sq_simulate_rowmax(b200sq_program* prog, int64_t N, int d, const double* x_dev,
                           const double* angle_table_dev, int n_variants, const int32_t* var_gate,
                           const double* var_shift, void* psi_dev, b200sq_row_stats* row_stats_dev, void* stream_) {
    if (!prog || !psi_dev) return fail(2, "NULL argument");
    if (N < 0 || n_variants < 1) return fail(2, "bad batch size");
    if (N == 0) return 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    int e;
    if ((e = ensure_device(prog, stream))) return e;
    if ((e = upload_variants(prog, n_variants, var_gate, var_shift, stream))) return e;
    if (row_stats_dev) {
        cudaError_t ce = cudaMemsetAsync(row_stats_dev, 0, (size_t)n_variants * (size_t)N * sizeof(b200sq_row_stats), stream);
        if (ce != cudaSuccess) return cuda_fail(ce, "cudaMemsetAsync (row maxima)");
    }
    return run_passes(prog, N, 0, N, 0, n_variants, d, x_dev, angle_table_dev, var_gate != nullptr,
                      (double2*)psi_dev, 0, false, 0, nullptr, stream, reinterpret_cast<RowStats*>(row_stats_dev));
}

int b200sq_simulate_expvals(b200sq_program* prog, int64_t N, int d, const double* x_dev,
                            const double* angle_table_dev, int n_variants, const int32_t* var_gate,
                            const double* var_shift, int n_terms, const uint32_t* x_mask,
                            const uint32_t* z_mask, double* out_dev, void* workspace_dev,
                            size_t workspace_bytes, void* stream_) {
    return b200sq_simulate_expvals2(prog, N, d, x_dev, angle_table_dev, n_variants, var_gate, var_shift, nullptr,
                                    nullptr, n_terms, x_mask, z_mask, out_dev, workspace_dev, workspace_bytes, stream_);
}

int b200sq_simulate_expvals2(b200sq_program* prog, int64_t N, int d, const double* x_dev,
                             const double* angle_table_dev, int n_variants, const int32_t* var_gate,
                             const double* var_shift, const int32_t* var_gate2, const double* var_shift2,
                             int n_terms, const uint32_t* x_mask, const uint32_t* z_mask, double* out_dev,
                             void* workspace_dev, size_t workspace_bytes, void* stream_) {
    if (!prog || !out_dev || !x_mask || !z_mask) return fail(2, "NULL argument");
    if (N < 0 || n_variants < 1 || n_terms < 1) return fail(2, "bad sizes");
    if (N == 0) return 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    int e;
    if ((e = ensure_device(prog, stream))) return e;
    if ((e = upload_variants(prog, n_variants, var_gate, var_shift, stream, var_gate2, var_shift2))) return e;
    if ((e = upload_masks(prog, n_terms, x_mask, z_mask, prog->plan.n, stream))) return e;
    const Plan& pl = prog->plan;
    const bool have_var = var_gate != nullptr;
    if (pl.passes.size() == 1 && pl.T == pl.n) {
        return run_passes(prog, N, 0, N, 0, n_variants, d, x_dev, angle_table_dev, have_var, nullptr, 0,
                          true, n_terms, out_dev, stream);
    }
    const size_t state_bytes = sizeof(double2) << pl.n;
    int64_t chunk = (int64_t)(workspace_bytes / state_bytes);
    if (!workspace_dev || chunk < 1) return fail(4, "workspace too small: need at least one state (16 * 2^n bytes)");
    for (int v = 0; v < n_variants; ++v) {
        for (int64_t s0 = 0; s0 < N; s0 += chunk) {
            const int64_t ns = std::min<int64_t>(chunk, N - s0);
            e = run_passes(prog, N, s0, ns, v, 1, d, x_dev, angle_table_dev, have_var,
                           (double2*)workspace_dev, (int64_t)v * N + s0, false, 0, nullptr, stream);
            if (e) return e;
            e = launch_expvals((const double2*)workspace_dev, ns, pl.n, n_terms, prog->d_masks,
                               prog->d_masks + n_terms, out_dev + ((int64_t)v * N + s0) * n_terms, n_terms,
                               stream);
            if (e) return cuda_fail(e, "expvals launch");
        }
    }
    return 0;
}

int b200sq_expvals(const void* psi_dev, int64_t N, int n_qubits, int n_terms, const uint32_t* x_mask,
                   const uint32_t* z_mask, double* out_dev, void* stream_) {
    if (!psi_dev || !out_dev || !x_mask || !z_mask) return fail(2, "NULL argument");
    if (n_qubits < 1 || n_qubits > B200SQ_MAX_QUBITS || n_terms < 1 || N < 0) return fail(2, "bad sizes");
    if (N == 0) return 0;
    const uint32_t full = (1u << n_qubits) - 1u;
    for (int t = 0; t < n_terms; ++t)
        if ((x_mask[t] | z_mask[t]) & ~full) return fail(2, "Pauli mask addresses a qubit outside the register");
    cudaStream_t stream = (cudaStream_t)stream_;
    uint32_t* d_masks = nullptr;
    cudaError_t ce = cudaMallocAsync(&d_masks, 2 * (size_t)n_terms * sizeof(uint32_t), stream);
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaMallocAsync (no CUDA device: b200sq has no CPU path)");
    cudaMemcpyAsync(d_masks, x_mask, n_terms * sizeof(uint32_t), cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(d_masks + n_terms, z_mask, n_terms * sizeof(uint32_t), cudaMemcpyHostToDevice, stream);
    int e = launch_expvals((const double2*)psi_dev, N, n_qubits, n_terms, d_masks, d_masks + n_terms,
                           out_dev, n_terms, stream);
    cudaFreeAsync(d_masks, stream);
    if (e) return cuda_fail(e, "expvals launch");
    return 0;
}

int b200sq_gram(const void* psi_x_dev, int64_t Nx, const void* psi_y_dev, int64_t Ny, int64_t dim,
                double* K_dev, int64_t ldk, uint32_t flags, void* stream_) {
    return b200sq_gram_rowmax(psi_x_dev, Nx, psi_y_dev, Ny, dim, K_dev, ldk, flags, nullptr, nullptr, stream_);
}

int b200sq_gram_rowmax(const void* psi_x_dev, int64_t Nx, const void* psi_y_dev, int64_t Ny, int64_t dim,
                       double* K_dev, int64_t ldk, uint32_t flags, const b200sq_row_stats* row_stats_x_dev,
                       const b200sq_row_stats* row_stats_y_dev, void* stream_) {
    if (!psi_x_dev || !psi_y_dev || !K_dev) return fail(2, "NULL argument");
    if (Nx < 0 || Ny < 0 || dim < 1 || ldk < Ny) return fail(2, "bad sizes");
    if ((flags & B200SQ_GRAM_SYMMETRIC) && (Nx != Ny || ldk < Nx))
        return fail(2, "symmetric Gram needs Nx == Ny");
    int e;
    if ((flags & B200SQ_GRAM_INT8) && gram_i8_supported(dim))
        e = launch_gram_i8((const double2*)psi_x_dev, Nx, (const double2*)psi_y_dev, Ny, dim, K_dev, ldk, flags,
                           (cudaStream_t)stream_, reinterpret_cast<const RowStats*>(row_stats_x_dev),
                           reinterpret_cast<const RowStats*>(row_stats_y_dev));
    else
        e = launch_gram((const double2*)psi_x_dev, Nx, (const double2*)psi_y_dev, Ny, dim, K_dev, ldk, flags,
                        (cudaStream_t)stream_);
    if (e) return cuda_fail(e, "gram launch (no CUDA device? b200sq has no CPU path)");
    return 0;
}

size_t b200sq_gram_planes_bytes(int64_t plane_rows, int64_t dim) {
    return plane_rows > 0 && dim > 0 ? (size_t)12 * (size_t)plane_rows * (size_t)(2 * dim) : 0;
}

int b200sq_gram_slice(const void* psi_dev, int64_t N, int64_t dim, void* planes_dev, int64_t plane_rows,
                      int64_t row0, int32_t* expo_dev, void* stream_) {
    return b200sq_gram_slice_rowmax(psi_dev, N, dim, planes_dev, plane_rows, row0, expo_dev, nullptr, stream_);
}

int b200sq_gram_slice_rowmax(const void* psi_dev, int64_t N, int64_t dim, void* planes_dev, int64_t plane_rows,
                             int64_t row0, int32_t* expo_dev, const b200sq_row_stats* row_stats_dev, void* stream_) {
    if (!psi_dev || !planes_dev || !expo_dev) return fail(2, "NULL argument");
    if (N < 0 || row0 < 0 || row0 + N > plane_rows) return fail(2, "rows outside the planes");
    if (!gram_i8_supported(dim)) return fail(2, "the INT8 Gram needs dim >= 64 and dim % 8 == 0");
    int e = launch_slice_i8((const double2*)psi_dev, N, dim, (int8_t*)planes_dev, plane_rows, row0, expo_dev,
                            (cudaStream_t)stream_, reinterpret_cast<const RowStats*>(row_stats_dev));
    if (e) return cuda_fail(e, "slice launch (no CUDA device? b200sq has no CPU path)");
    return 0;
}

int b200sq_gram_crt_moduli(int64_t dim) { return gram_crt_supported(dim) ? gram_crt_moduli(dim) : 0; }
int b200sq_gram_crt_bits(int64_t dim) { return gram_crt_supported(dim) ? gram_crt_bits(dim) : 0; }
static_assert(sizeof(b200sq_row_stats) == sizeof(RowStats) && sizeof(RowStats) == 16, "row statistics layout");

int b200sq_gram_slice_crt(const void* psi_dev, int64_t N, int64_t dim, void* planes_dev, int64_t plane_rows,
                          int64_t row0, int32_t* expo_dev, const b200sq_row_stats* row_stats_dev, void* stream_) {
    return b200sq_gram_slice_crt_digits(psi_dev, N, dim, planes_dev, plane_rows, row0, expo_dev, row_stats_dev, nullptr,
                                        0, 0, stream_);
}

int b200sq_gram_slice_crt_digits(const void* psi_dev, int64_t N, int64_t dim, void* planes_dev, int64_t plane_rows,
                                 int64_t row0, int32_t* expo_dev, const b200sq_row_stats* row_stats_dev, void* digit_planes_dev,
                                 int64_t digit_rows, int64_t digit_row0, void* stream_) {
    if (!psi_dev || !planes_dev || !expo_dev) return fail(2, "NULL argument");
    if (N < 0 || row0 < 0 || row0 + N > plane_rows) return fail(2, "rows outside the planes");
    if (digit_planes_dev && (digit_row0 < 0 || digit_row0 + N > digit_rows)) return fail(2, "rows outside the digit planes");
    if (!gram_crt_supported(dim)) return fail(2, "the CRT Gram does not support this state dimension");
    int e = launch_slice_crt((const double2*)psi_dev, N, dim, (int8_t*)planes_dev, plane_rows, row0, expo_dev,
                             (cudaStream_t)stream_, reinterpret_cast<const RowStats*>(row_stats_dev), (int8_t*)digit_planes_dev, digit_rows, digit_row0);
    if (e) return cuda_fail(e, "slice launch (no CUDA device? b200sq has no CPU path)");
    return 0;
}

int b200sq_gram_digits_to_crt(const void* digit_planes_dev, int64_t plane_rows_d, int64_t row0_d, int64_t N,
                              int64_t dim, void* crt_planes_dev, int64_t plane_rows_o, int64_t row0_o, void* stream_) {
    if (!digit_planes_dev || !crt_planes_dev) return fail(2, "NULL argument");
    if (N < 0 || row0_d < 0 || row0_o < 0 || row0_d + N > plane_rows_d || row0_o + N > plane_rows_o)
        return fail(2, "rows outside the planes");
    if (!gram_crt_supported(dim)) return fail(2, "the CRT Gram does not support this state dimension");
    int e = launch_digits_to_crt((const int8_t*)digit_planes_dev, plane_rows_d, row0_d, N, dim, (int8_t*)crt_planes_dev,
                                 plane_rows_o, row0_o, (cudaStream_t)stream_);
    if (e) return cuda_fail(e, "digits-to-residues launch");
    return 0;
}

int b200sq_gram_from_planes(const void* planes_x_dev, const int32_t* expo_x_dev, int64_t plane_rows_x, int64_t x0,
                            int64_t Nx, const void* planes_y_dev, const int32_t* expo_y_dev, int64_t plane_rows_y,
                            int64_t y0, int64_t Ny, int64_t dim, double* K_dev, int64_t ldk, uint32_t flags,
                            void* stream_) {
    if (!planes_x_dev || !planes_y_dev || !expo_x_dev || !expo_y_dev || !K_dev) return fail(2, "NULL argument");
    if (Nx < 0 || Ny < 0 || x0 < 0 || y0 < 0 || x0 + Nx > plane_rows_x || y0 + Ny > plane_rows_y || ldk < Ny)
        return fail(2, "bad sizes");
    if (!gram_i8_supported(dim)) return fail(2, "the INT8 Gram needs dim >= 64 and dim % 8 == 0");
    if ((flags & B200SQ_GRAM_SYMMETRIC) && (planes_x_dev != planes_y_dev || x0 != y0 || Nx != Ny))
        return fail(2, "symmetric Gram needs identical operands");
    if ((flags & B200SQ_GRAM_CRT) && !gram_crt_supported(dim)) return fail(2, "the CRT Gram does not support this state dimension");
    int e = launch_gram_planes((const int8_t*)planes_x_dev, expo_x_dev, plane_rows_x, x0, Nx,
                               (const int8_t*)planes_y_dev, expo_y_dev, plane_rows_y, y0, Ny, dim, K_dev, ldk, flags,
                               (cudaStream_t)stream_);
    if (e) return cuda_fail(e, "gram launch (no CUDA device? b200sq has no CPU path)");
    return 0;
}

int b200sq_gram_jobs(const void* planes_x_dev, const int32_t* expo_x_dev, int64_t plane_rows_x, int64_t dim,
                     int n_jobs, const b200sq_gram_job* jobs, void* stream_) {
    if (!planes_x_dev || !expo_x_dev || (n_jobs > 0 && !jobs)) return fail(2, "NULL argument");
    if (n_jobs < 0 || plane_rows_x < 0) return fail(2, "bad sizes");
    if (!gram_i8_supported(dim)) return fail(2, "the INT8 Gram needs dim >= 64 and dim % 8 == 0");
    if (n_jobs > 0 && (jobs[0].flags & B200SQ_GRAM_CRT) && !gram_crt_supported(dim))
        return fail(2, "the CRT Gram does not support this state dimension");
    for (int j = 0; j < n_jobs; ++j) {
        const b200sq_gram_job& J = jobs[j];
        if (!J.planes_y_dev || !J.expo_y_dev || !J.K_dev) return fail(2, "NULL pointer in a Gram job");
        if (J.nx < 0 || J.ny < 0 || J.x0 < 0 || J.y0 < 0 || J.x0 + J.nx > plane_rows_x || J.y0 + J.ny > J.plane_rows_y ||
            J.ldk < J.ny)
            return fail(2, "bad sizes in a Gram job");
        if ((J.flags & B200SQ_GRAM_SYMMETRIC) && (J.planes_y_dev != planes_x_dev || J.x0 != J.y0 || J.nx != J.ny))
            return fail(2, "symmetric Gram job needs identical operands");
        if (J.align_yx_dev && !J.align_out_dev) return fail(2, "align_out_dev is NULL");
        if (J.K_mirror_dev && J.ldk_mirror < J.nx) return fail(2, "ldk_mirror is smaller than the block height");
        if ((J.flags & B200SQ_GRAM_CRT) != (jobs[0].flags & B200SQ_GRAM_CRT)) return fail(2, "jobs of one launch must use one scheme");
    }
    int e = launch_gram_jobs((const int8_t*)planes_x_dev, expo_x_dev, plane_rows_x, dim, n_jobs, jobs,
                             (cudaStream_t)stream_);
    if (e) return cuda_fail(e, "gram launch (no CUDA device? b200sq has no CPU path)");
    return 0;
}

int b200sq_wait_flags(const uint32_t* flags_dev, const int32_t* which, int n, uint32_t value, void* stream_) {
    if (n <= 0) return 0;
    if (!flags_dev || !which) return fail(2, "NULL argument");
    if (n > 32) return fail(2, "at most 32 flags per wait");
    int e = launch_wait_flags(flags_dev, which, n, value, (cudaStream_t)stream_);
    if (e) return cuda_fail(e, "wait-flags launch");
    return 0;
}

int64_t b200sq_gram_ready_timeouts(void) { return (int64_t)gram_ready_timeouts(); }

int b200sq_ipc_export(const void* dev_ptr, void* handle_out_64, int64_t* offset_out) {
    if (!dev_ptr || !handle_out_64 || !offset_out) return fail(2, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    // the handle names the whole allocation: find its base (driver API through the runtime's entry-point query)
    typedef int (*GetRangeFn)(unsigned long long*, size_t*, unsigned long long);
    static GetRangeFn get_range = nullptr;
    if (!get_range) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess || !p)
            return fail(6, "cuMemGetAddressRange is unavailable");
        get_range = (GetRangeFn)p;
    }
    unsigned long long base = 0;
    size_t size = 0;
    if (get_range(&base, &size, (unsigned long long)(uintptr_t)dev_ptr) != 0 || !base)
        return fail(6, "cuMemGetAddressRange failed (not a device allocation?)");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return cuda_fail(e, "cudaIpcGetMemHandle (allocation not exportable: pool / expandable-segment memory?)");
    }
    std::memcpy(handle_out_64, &h, 64);
    *offset_out = (int64_t)((unsigned long long)(uintptr_t)dev_ptr - base);
    return 0;
}

int b200sq_ipc_open(const void* handle_64, void** base_out) {
    if (!handle_64 || !base_out) return fail(2, "NULL argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle_64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return cuda_fail(e, "cudaIpcOpenMemHandle");
    }
    *base_out = p;
    return 0;
}

int b200sq_ipc_close(void* base) {
    if (!base) return 0;
    cudaError_t e = cudaIpcCloseMemHandle(base);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return cuda_fail(e, "cudaIpcCloseMemHandle");
    }
    return 0;
}

int b200sq_copy_async(void* dst_dev, const void* src_dev, size_t bytes, void* stream_) {
    if (bytes == 0) return 0;
    if (!dst_dev || !src_dev) return fail(2, "NULL argument");
    cudaError_t e = cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream_);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return cuda_fail(e, "cudaMemcpyAsync (device to device)");
    }
    return 0;
}

int b200sq_copy_planes_async(void* dst_dev, size_t dst_pitch, const void* src_dev, size_t src_pitch, size_t bytes,
                             int n_planes, void* stream_) {
    if (bytes == 0 || n_planes <= 0) return 0;
    if (!dst_dev || !src_dev) return fail(2, "NULL argument");
    for (int p = 0; p < n_planes; ++p) {
        cudaError_t e = cudaMemcpyAsync((char*)dst_dev + (size_t)p * dst_pitch, (const char*)src_dev + (size_t)p * src_pitch,
                                        bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream_);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return cuda_fail(e, "cudaMemcpyAsync (device to device, plane)");
        }
    }
    return 0;
}

int b200sq_rbf(const double* fx_dev, int64_t Nx, const double* fy_dev, int64_t Ny, int nf, double gamma,
               double* K_dev, int64_t ldk, void* stream_) {
    if (!fx_dev || !fy_dev || !K_dev) return fail(2, "NULL argument");
    if (Nx < 0 || Ny < 0 || nf < 0 || ldk < Ny) return fail(2, "bad sizes");
    int e = launch_rbf(fx_dev, Nx, fy_dev, Ny, nf, gamma, K_dev, ldk, (cudaStream_t)stream_);
    if (e) return cuda_fail(e, "rbf launch");
    return 0;
}

}  // extern "C"
