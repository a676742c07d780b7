import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import squlearn_b200 as sq
ex = sq.Executor("b200"); eng = ex.engine
N, D = 4096, int(os.environ.get("DIM", 65536))
psi = torch.randn(N, D, dtype=torch.complex128, device="cuda"); psi /= psi.norm(dim=1, keepdim=True)
def timeit(f, reps=3):
    f(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); f(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts)
ms = timeit(lambda: eng.gram(psi, None))
k = eng.gram(psi, None); k4 = eng.gram(psi, None, three_m=False)
print(json.dumps({"warps": os.environ.get("B200SQ_GRAM_WARPS", "8"), "ms": ms, "alg_TF": 8.0 * D * N * (N + 1) / 2 / ms / 1e9,
                  "maxdiff_vs_4m": float((k - k4).abs().max())}))
