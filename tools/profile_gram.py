"""Gram kernel alone on random unit-norm states (for ncu). usage: profile_gram.py N D [precision]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import squlearn_b200 as sq
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
D = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
prec = sys.argv[3] if len(sys.argv) > 3 else "auto"
ex = sq.Executor("b200")
psi = torch.randn(N, D, dtype=torch.complex128, device="cuda")
psi /= psi.norm(dim=1, keepdim=True)
for _ in range(2):
    k = ex.engine.gram(psi, None, precision=prec)
torch.cuda.synchronize()
print("ok", float(k[1, 0]))
