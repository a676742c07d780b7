"""INT8 split-integer Gram vs the FP64 DMMA Gram and torch (GPU): correctness + timing."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import squlearn_b200 as sq
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ex = sq.Executor("b200"); eng = ex.engine
X = np.random.default_rng(3).uniform(-0.95, 0.95, (N, 4))
enc = sq.ChebyshevPQC(n, 2)
p = enc.generate_initial_parameters(4, seed=0)
cp = sq.B200Circuit(enc.get_program(4)).compiled(eng, p)
psi = eng.simulate(cp, torch.from_numpy(X).cuda())
torch.cuda.synchronize()
def timeit(f, reps=3):
    f(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = f(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts), r
out = {"N": N, "n": n}
t64, k64 = timeit(lambda: eng.gram(psi, None, precision="fp64"))
t8, k8 = timeit(lambda: eng.gram(psi, None, precision="int8"))
out["sym"] = {"ms_fp64": t64, "ms_int8": t8, "max_abs_diff": float((k8 - k64).abs().max()),
              "asym": float((k8 - k8.T).abs().max())}
if N <= 2048:
    kref = (psi.conj() @ psi.T).abs() ** 2
    kref.fill_diagonal_(1.0)
    out["sym"]["vs_torch_int8"] = float((k8 - kref).abs().max())
    out["sym"]["vs_torch_fp64"] = float((k64 - kref).abs().max())
h = N // 2
t64, r64 = timeit(lambda: eng.gram(psi[:h - 3], psi[h:], precision="fp64"))
t8, r8 = timeit(lambda: eng.gram(psi[:h - 3], psi[h:], precision="int8"))
out["rect"] = {"ms_fp64": t64, "ms_int8": t8, "max_abs_diff": float((r8 - r64).abs().max())}
print(json.dumps(out))
