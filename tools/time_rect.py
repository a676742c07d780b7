import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import squlearn_b200 as sq
ex = sq.Executor("b200"); eng = ex.engine
N, D = 4096, 65536
psi = torch.randn(N, D, dtype=torch.complex128, device="cuda"); psi /= psi.norm(dim=1, keepdim=True)
def timeit(f, reps=3):
    f(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); f(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts)
out = {}
for sl in sys.argv[1:]:
    os.environ["B200SQ_GRAM_SLICES"] = sl
    out["rect512x2048_s" + sl] = timeit(lambda: eng.gram(psi[:512], psi[512:2560]))
    out["rect512x1792_s" + sl] = timeit(lambda: eng.gram(psi[:512], psi[512:2304]))
    out["sym512_s" + sl] = timeit(lambda: eng.gram(psi[:512], None))
print(json.dumps(out))
