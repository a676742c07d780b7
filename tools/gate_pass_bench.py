"""Memory-path ceiling of the tile kernel: passes with very few gates (HBM-bound by construction)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import squlearn_b200 as sq
from squlearn_b200 import Program, GateOp, Angle

n = int(os.environ.get("QUBITS", 16)); N = int(os.environ.get("POINTS", 4096))
ex = sq.Executor("b200"); eng = ex.engine
X = np.random.default_rng(3).uniform(-0.95, 0.95, (N, 4)); xd = torch.from_numpy(X).cuda()
def timeit(f, reps=5):
    f(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); f(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts)
out = {}
sb = 16.0 * N * (1 << n)
cases = {
    "h_all": [GateOp("h", (q,)) for q in range(n)],
    "rx_all": [GateOp("rx", (q,), Angle(feature=q % 4, fn="arccos")) for q in range(n)],
    "rx_all_crz_ring": [GateOp("rx", (q,), Angle(feature=q % 4, fn="arccos")) for q in range(n)]
                        + [GateOp("crz", (q, (q + 1) % n), Angle(scale=0.3 + 0.1 * q)) for q in range(n)],
}
for name, gates in cases.items():
    for T in (11, 12):
        cp = sq.B200Circuit(Program(n, gates, 0, 4), tile_bits=T).compiled(eng, None)
        ms = timeit(lambda: eng.simulate(cp, xd))
        out[f"{name}_T{T}"] = {"ms": round(ms, 3), "passes": cp.num_passes, "GBps_2passes": round(sb * (2 * cp.num_passes - 1) / ms / 1e6)}
# plain device copy of the same buffer for reference
psi = eng.simulate(cp, xd); dst = torch.empty_like(psi)
ms = timeit(lambda: dst.copy_(psi))
out["torch_copy_same_size"] = {"ms": round(ms, 3), "GBps": round(2 * sb / ms / 1e6)}
print(json.dumps(out))
