"""Device timings (CUDA events) of the gate passes and the Gram kernel at bench size."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import squlearn_b200 as sq

n = int(os.environ.get("QUBITS", 16)); N = int(os.environ.get("POINTS", 4096)); reps = 3
ex = sq.Executor("b200"); eng = ex.engine
X = np.random.default_rng(3).uniform(-0.95, 0.95, (N, 4))
enc = sq.ChebyshevPQC(n, 2)
p = enc.generate_initial_parameters(4, seed=0)
xd = torch.from_numpy(X).cuda()
def timeit(f):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = f(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts), r
out = {"lib": os.environ.get("B200SQ_LIB", "default"), "n": n, "N": N}
for T in [int(t) for t in os.environ.get("TILES", "12").split(",")]:
    cp = sq.B200Circuit(enc.get_program(4), tile_bits=T).compiled(eng, p)
    ms, psi = timeit(lambda: eng.simulate(cp, xd))
    sb = 16.0 * N * (1 << n)
    out[f"sim_T{T}"] = {"ms": ms, "passes": cp.num_passes, "GBps": sb * (2 * cp.num_passes - 1) / ms / 1e6}
for name, mode in (("gram_fp64_4m", False), ("gram_fp64_3m_ws", True), ("gram_fp64_3m_classic", "classic")):
    ms, k = timeit(lambda: eng.gram(psi, None, three_m=mode, precision="fp64"))
    out[name] = {"ms": ms, "alg_TFLOPs": 8.0 * (1 << n) * N * (N + 1) / 2 / ms / 1e9}
ms, k = timeit(lambda: eng.gram(psi, None, precision="int8"))
out["gram_int8_tcgen05"] = {"ms": ms, "fp64_equiv_TFLOPs": 8.0 * (1 << n) * N * (N + 1) / 2 / ms / 1e9}
ms, k = timeit(lambda: eng.gram(psi[: N // 8], psi[N // 8: 5 * N // 8], precision="int8"))
out["gram_rect_eighth_int8"] = {"ms": ms, "fp64_equiv_TFLOPs": 8.0 * (1 << n) * (N // 8) * (N // 2) / ms / 1e9}
print(json.dumps(out))
