"""One step of the bench workload (simulate + Gram at C3 size) for ncu; prints nothing timed."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import squlearn_b200 as sq

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ex = sq.Executor("b200")
X = np.random.default_rng(3).uniform(-0.95, 0.95, (N, 4))
fk = sq.FidelityKernel(sq.ChebyshevPQC(n, 2), ex)
for _ in range(steps):
    k = fk.evaluate_device(X)
torch.cuda.synchronize()
print("ok", float(k[1, 0]))
