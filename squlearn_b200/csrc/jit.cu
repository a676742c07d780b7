// Run-time specialisation of the pass kernel (the batched counterpart of the reference's per-gate replay of
// its gate-list IR, src/squlearn/util/pennylane/pennylane_circuit.py:541-681).
//
// The generic kernel k_tile_pass interprets the rounds / steps of a pass; ncu shows it bound by the
// interpreter's own instructions (step decode, register moves at the switch joins), not by the FP64
// pipe.  For large batches the library therefore compiles, once per program and pass, the SAME
// kernel body (tile_kernel.cuh) with the interpreter replaced by straight-line code: every round
// becomes a literal DRound / DStep sequence passed through the same primitives, so that all kinds,
// register slots, bit positions and matrix offsets are compile-time constants.  NVRTC
// (libnvrtc.so.12, loaded with dlopen) produces an sm_100a cubin that is loaded with
// cudaLibraryLoadData.  If NVRTC or the headers are not available the generic kernel runs: this is a
// speed option of the CUDA path, not a fallback to any other implementation.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "kernels.h"

namespace b200sq {

namespace {

struct Nvrtc {
    void* handle = nullptr;
    decltype(&nvrtcCreateProgram) create = nullptr;
    decltype(&nvrtcCompileProgram) compile = nullptr;
    decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
    decltype(&nvrtcGetCUBIN) cubin = nullptr;
    decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
    decltype(&nvrtcGetProgramLog) log = nullptr;
    decltype(&nvrtcDestroyProgram) destroy = nullptr;
    bool ok = false;
};

Nvrtc& nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"};
        for (const char* nm : names) {
            n.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (n.handle) break;
        }
        if (!n.handle) return;
#define B200SQ_SYM(field, sym) n.field = (decltype(n.field))dlsym(n.handle, #sym)
        B200SQ_SYM(create, nvrtcCreateProgram);
        B200SQ_SYM(compile, nvrtcCompileProgram);
        B200SQ_SYM(cubin_size, nvrtcGetCUBINSize);
        B200SQ_SYM(cubin, nvrtcGetCUBIN);
        B200SQ_SYM(log_size, nvrtcGetProgramLogSize);
        B200SQ_SYM(log, nvrtcGetProgramLog);
        B200SQ_SYM(destroy, nvrtcDestroyProgram);
#undef B200SQ_SYM
        n.ok = n.create && n.compile && n.cubin_size && n.cubin && n.log_size && n.log && n.destroy;
    });
    return n;
}

// Directory of the csrc headers: next to this shared library (the library is built and used in-tree).
std::string csrc_dir() {
    if (const char* e = getenv("B200SQ_CSRC")) return e;
    Dl_info info;
    if (dladdr((void*)&csrc_dir, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t k = p.find_last_of('/');
        return (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/csrc";
    }
    return "csrc";
}

}  // namespace

// Straight-line source of pass `ip` (public for tests / tools through b200sq_program_jit_source).
std::string jit_pass_source(const Plan& plan, int ip, int team_size) {
    const Pass& ps = plan.passes[ip];
    const std::vector<uint8_t> blob = build_pass_blob(plan, ip);
    const DPass& P = *reinterpret_cast<const DPass*>(blob.data());
    const DRound* rounds = reinterpret_cast<const DRound*>(blob.data() + P.off_rounds);
    const DStep* steps = reinterpret_cast<const DStep*>(blob.data() + P.off_steps);
    const int T = ps.T;
    std::ostringstream o;
    o << "#include \"tile_kernel.cuh\"\n"
         "using namespace b200sq;\n"
         "struct JitRounds {\n"
         "    __device__ __forceinline__ void operator()(double2* __restrict__ tile, const PassView&, const double2* __restrict__ mats,\n"
         "                                               uint32_t ext_base, int, int tid, int, int team) const {\n"
         "        constexpr int T = " << T << ", TS = " << team_size << ";\n";
    for (int r = 0; r < P.n_rounds; ++r) {
        const DRound& rd = rounds[r];
        const int nb = rd.nb;
        const int g = (nb == 3 && ((1 << (T - 3)) % (2 * team_size)) == 0) ? 2 : 1;
        o << "        {   // round " << r << "\n"
          << "            const DRound rd = {" << (int)rd.nb << ", {" << (int)rd.rb[0] << ", " << (int)rd.rb[1] << ", "
          << (int)rd.rb[2] << "}, 0, 0, {";
        for (int k = 0; k < 8; ++k) o << rd.ksw[k] << (k < 7 ? ", " : "");
        o << "}};\n"
          << "            exec_round_with<" << nb << ", " << g << ">(tile, rd, ext_base, T, tid, TS,\n"
          << "                [&](double2 (&a)[" << g << "][" << (1 << nb) << "], const uint32_t (&e)[" << g
          << "], const uint32_t (&base)[" << g << "], const uint32_t (&ko)[" << (1 << nb) << "]) {\n";
        for (int s = rd.s0; s < rd.s1; ++s) {
            const DStep& u = steps[s];
            o << "                    { const DStep u = {" << (int)u.kind << ", " << (int)u.r0 << ", " << (int)u.swap << ", 0, "
              << (int)u.r1 << ", 0, " << u.gate << ", " << u.ctrl << "u, " << u.mat << ", " << u.pat << "}; "
              << "apply_step<" << nb << ", " << g << ">(a, u, mats + " << u.mat << ", e, base, ko); }\n";
        }
        o << "                });\n"
          << "            team_sync(TS, team);\n"
          << "        }\n";
    }
    o << "    }\n};\n"
         "extern \"C\" __global__ void __launch_bounds__(kCtaThreads, B200SQ_TILE_MINBLOCKS)\n"
         "k_tile_pass_jit(const __grid_constant__ TileArgs A) { tile_pass_body(A, JitRounds{}); }\n";
    return o.str();
}

// Compile (or fetch from the process-wide cache) the specialised kernel of a pass.  Returns nullptr when
// run-time compilation is unavailable or fails (the reason goes to stderr once).
const void* jit_pass_kernel(const Plan& plan, int ip, int team_size) {
    static std::mutex mu;
    static std::unordered_map<std::string, const void*> cache;
    static bool warned = false;
    const std::string src = jit_pass_source(plan, ip, team_size);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(src);
    if (it != cache.end()) return it->second;
    const void* result = nullptr;
    auto fail = [&](const std::string& why) {
        if (!warned) fprintf(stderr, "b200sq: run-time kernel specialisation unavailable (%s); using the generic pass kernel\n", why.c_str());
        warned = true;
        cache[src] = nullptr;
        return (const void*)nullptr;
    };
    Nvrtc& rt = nvrtc();
    if (!rt.ok) return fail("libnvrtc.so.12 not found");
    nvrtcProgram prog;
    if (rt.create(&prog, src.c_str(), "b200sq_pass_jit.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) return fail("nvrtcCreateProgram");
    const std::string inc = "--include-path=" + csrc_dir();
    static const std::string minb = std::string("-DB200SQ_TILE_MINBLOCKS=") + (getenv("B200SQ_JIT_MINBLOCKS") ? getenv("B200SQ_JIT_MINBLOCKS") : "2");
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", inc.c_str(),
                          "--include-path=/usr/local/cuda/include", minb.c_str()};
    const nvrtcResult cr = rt.compile(prog, 6, opts);
    if (cr != NVRTC_SUCCESS) {
        size_t ls = 0;
        rt.log_size(prog, &ls);
        std::string log(ls, '\0');
        if (ls) rt.log(prog, &log[0]);
        rt.destroy(&prog);
        if (log.size() > 1500) log.resize(1500);
        return fail("nvrtc compile error: " + log);
    }
    size_t cs = 0;
    rt.cubin_size(prog, &cs);
    std::vector<char> cubin(cs);
    rt.cubin(prog, cubin.data());
    rt.destroy(&prog);
    cudaLibrary_t lib;
    cudaError_t e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) return fail(std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e));
    cudaKernel_t kern;
    e = cudaLibraryGetKernel(&kern, lib, "k_tile_pass_jit");
    if (e != cudaSuccess) return fail(std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(e));
    result = (const void*)kern;
    cache[src] = result;
    return result;
}

// Compile only (no GPU needed): returns the cubin size or -1, log in `log`.  Used by the CPU test-suite
// to check that the generated source is valid CUDA.
long jit_compile_only(const Plan& plan, int ip, int team_size, std::string* log_out) {
    Nvrtc& rt = nvrtc();
    if (!rt.ok) {
        if (log_out) *log_out = "libnvrtc.so.12 not found";
        return -1;
    }
    const std::string src = jit_pass_source(plan, ip, team_size);
    nvrtcProgram prog;
    if (rt.create(&prog, src.c_str(), "b200sq_pass_jit.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) return -1;
    const std::string inc = "--include-path=" + csrc_dir();
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", inc.c_str(),
                          "--include-path=/usr/local/cuda/include"};
    const nvrtcResult cr = rt.compile(prog, 5, opts);
    size_t ls = 0;
    rt.log_size(prog, &ls);
    std::string log(ls, '\0');
    if (ls) rt.log(prog, &log[0]);
    if (log_out) *log_out = log;
    long size = -1;
    if (cr == NVRTC_SUCCESS) {
        size_t cs = 0;
        rt.cubin_size(prog, &cs);
        size = (long)cs;
        if (const char* dump = getenv("B200SQ_JIT_DUMP")) {
            std::vector<char> cubin(cs);
            rt.cubin(prog, cubin.data());
            if (FILE* f = fopen(dump, "wb")) {
                fwrite(cubin.data(), 1, cs, f);
                fclose(f);
            }
        }
    }
    rt.destroy(&prog);
    return size;
}

}  // namespace b200sq
