// C ABI of libb200sq (see include/b200sq.h).  Host-side orchestration only: plan handling,
// device mirrors of the plan, launch sequencing.  There is no CPU compute path in this library.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
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
    int32_t* d_var_gate = nullptr;
    double* d_var_shift = nullptr;
    int var_cap = 0;
    uint32_t* d_masks = nullptr;
    int mask_cap = 0;

    ~b200sq_program() {
        if (device >= 0) {
            int cur = -1;
            cudaGetDevice(&cur);
            cudaSetDevice(device);
            cudaFree(d_gates);
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
        p->device = dev;
        p->gates_dirty = true;
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
                    cudaStream_t stream) {
    if (!var_gate) return 0;
    if (!var_shift) return fail(2, "var_shift is NULL but var_gate is not");
    for (int v = 0; v < n_variants; ++v)
        if (var_gate[v] >= (int)p->plan.gates.size()) return fail(2, "variant gate index out of range");
    if (n_variants > p->var_cap) {
        cudaFree(p->d_var_gate);
        cudaFree(p->d_var_shift);
        p->var_cap = 0;
        cudaError_t e;
        if ((e = cudaMalloc(&p->d_var_gate, n_variants * sizeof(int32_t))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
        if ((e = cudaMalloc(&p->d_var_shift, n_variants * sizeof(double))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
        p->var_cap = n_variants;
    }
    cudaMemcpyAsync(p->d_var_gate, var_gate, n_variants * sizeof(int32_t), cudaMemcpyHostToDevice, stream);
    cudaError_t e = cudaMemcpyAsync(p->d_var_shift, var_shift, n_variants * sizeof(double),
                                    cudaMemcpyHostToDevice, stream);
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
int run_passes(b200sq_program* p, int64_t n_total, int64_t s0, int64_t ns, int v0, int nv, int d,
               const double* x, const double* table, bool have_variants, double2* psi, int64_t psi_off,
               bool fuse_expvals, int n_terms, double* out, cudaStream_t stream) {
    const Plan& pl = p->plan;
    for (size_t ip = 0; ip < pl.passes.size(); ++ip) {
        TileArgs a;
        std::memset(&a, 0, sizeof(a));
        const Pass& ps = pl.passes[ip];
        a.pass.T = ps.T;
        a.pass.first = ps.first;
        a.pass.n_rounds = ps.r1 - ps.r0;
        a.pass.n_uops = ps.u1 - ps.u0;
        a.pass.mat_elems = ps.mat_elems;
        a.pass.n_lruns = ps.n_lruns;
        a.pass.n_oruns = ps.n_oruns;
        for (int i = 0; i < kMaxRuns; ++i) {
            a.pass.lmask[i] = ps.lmask[i];
            a.pass.lshift[i] = ps.lshift[i];
            a.pass.omask[i] = ps.omask[i];
            a.pass.oshift[i] = ps.oshift[i];
        }
        for (int i = 0; i < 32; ++i) a.pass.extq[i] = ps.extq[i];
        a.n_init = 0;
        if (ps.first) {
            a.n_init = (int)pl.init_gates.size();
            for (int i = 0; i < a.n_init; ++i) {
                a.inits[i].gate = (uint16_t)pl.init_gates[i];
                a.inits[i].qubit = (uint8_t)pl.gates[pl.init_gates[i]].q[0];
            }
        }
        if (a.pass.n_rounds > kMaxPassRounds || a.pass.n_uops > kMaxPassUops)
            return fail(5, "internal: pass exceeds the kernel-parameter capacity");
        for (int r = ps.r0; r < ps.r1; ++r) a.rounds[r - ps.r0] = pack_round(pl.rounds[r], ps.u0);
        for (int u = ps.u0; u < ps.u1; ++u) a.uops[u - ps.u0] = pack_uop(pl.uops[u], ps.u0);
        a.gates = p->d_gates;
        a.psi = psi;
        a.psi_off = psi_off;
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
        const bool last = ip + 1 == pl.passes.size();
        a.store = (fuse_expvals && last) ? 0 : 1;
        a.n_terms = (fuse_expvals && last) ? n_terms : 0;
        a.xmask = p->d_masks;
        a.zmask = p->d_masks ? p->d_masks + n_terms : nullptr;
        a.out = out;
        tile_pass_smem_bytes(a.pass.T, a.pass.mat_elems, a.pass.first, &a.team_size, &a.teams_per_cta, &a.team_elems);
        a.total_items = (int64_t)nv * ns * ((int64_t)1 << (pl.n - a.pass.T));
        int e = launch_tile_pass(a, stream);
        if (e) return cuda_fail(e, "tile pass launch");
    }
    return 0;
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
        p->plan = make_plan(n_qubits, g, tile_bits);
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

int b200sq_program_num_passes(const b200sq_program* prog) {
    return prog ? (int)prog->plan.passes.size() : -1;
}

// Dump layout (int32 words): [n, T, n_passes, n_rounds, n_uops] then per pass
// [T, first, r0, r1, u0, u1, mat_elems, tile_qubits[T]...], then per round [nb, rb0, rb1, rb2, u0, u1],
// then per uop [kind, gate, r0, r1, ctrl, mat, swap].
int b200sq_program_dump(const b200sq_program* prog, int32_t* buf, int capacity, int* n_written) {
    if (!prog || !n_written) return fail(2, "NULL argument");
    std::vector<int32_t> w;
    const Plan& pl = prog->plan;
    w.insert(w.end(), {pl.n, pl.T, (int32_t)pl.passes.size(), (int32_t)pl.rounds.size(), (int32_t)pl.uops.size()});
    for (const Pass& ps : pl.passes) {
        w.insert(w.end(), {ps.T, ps.first, ps.r0, ps.r1, ps.u0, ps.u1, ps.mat_elems});
        for (int i = 0; i < ps.T; ++i) w.push_back(ps.tile_qubits[i]);
    }
    for (const Round& r : pl.rounds) w.insert(w.end(), {r.nb, r.rb[0], r.rb[1], r.rb[2], r.u0, r.u1});
    for (const Uop& u : pl.uops) w.insert(w.end(), {u.kind, u.gate, u.r0, u.r1, (int32_t)u.ctrl, u.mat, u.swap});
    w.push_back((int32_t)pl.init_gates.size());
    for (int32_t g : pl.init_gates) w.push_back(g);
    *n_written = (int)w.size();
    if (buf && capacity >= (int)w.size()) std::memcpy(buf, w.data(), w.size() * sizeof(int32_t));
    return 0;
}

int b200sq_simulate(b200sq_program* prog, int64_t N, int d, const double* x_dev,
                    const double* angle_table_dev, int n_variants, const int32_t* var_gate,
                    const double* var_shift, void* psi_dev, void* stream_) {
    if (!prog || !psi_dev) return fail(2, "NULL argument");
    if (N < 0 || n_variants < 1) return fail(2, "bad batch size");
    if (N == 0) return 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    int e;
    if ((e = ensure_device(prog, stream))) return e;
    if ((e = upload_variants(prog, n_variants, var_gate, var_shift, stream))) return e;
    return run_passes(prog, N, 0, N, 0, n_variants, d, x_dev, angle_table_dev, var_gate != nullptr,
                      (double2*)psi_dev, 0, false, 0, nullptr, stream);
}

int b200sq_simulate_expvals(b200sq_program* prog, int64_t N, int d, const double* x_dev,
                            const double* angle_table_dev, int n_variants, const int32_t* var_gate,
                            const double* var_shift, int n_terms, const uint32_t* x_mask,
                            const uint32_t* z_mask, double* out_dev, void* workspace_dev,
                            size_t workspace_bytes, void* stream_) {
    if (!prog || !out_dev || !x_mask || !z_mask) return fail(2, "NULL argument");
    if (N < 0 || n_variants < 1 || n_terms < 1) return fail(2, "bad sizes");
    if (N == 0) return 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    int e;
    if ((e = ensure_device(prog, stream))) return e;
    if ((e = upload_variants(prog, n_variants, var_gate, var_shift, stream))) return e;
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
    if (!psi_x_dev || !psi_y_dev || !K_dev) return fail(2, "NULL argument");
    if (Nx < 0 || Ny < 0 || dim < 1 || ldk < Ny) return fail(2, "bad sizes");
    if ((flags & B200SQ_GRAM_SYMMETRIC) && (Nx != Ny || ldk < Nx))
        return fail(2, "symmetric Gram needs Nx == Ny");
    int e = launch_gram((const double2*)psi_x_dev, Nx, (const double2*)psi_y_dev, Ny, dim, K_dev, ldk, flags,
                        (cudaStream_t)stream_);
    if (e) return cuda_fail(e, "gram launch (no CUDA device? b200sq has no CPU path)");
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
