import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import squlearn_b200 as sq
from squlearn_b200 import Program, GateOp
n, N, T = 16, 2048, int(sys.argv[1]) if len(sys.argv) > 1 else 11
ex = sq.Executor("b200"); eng = ex.engine
xd = torch.zeros((N, 1), dtype=torch.float64, device="cuda")
cp = sq.B200Circuit(Program(n, [GateOp("h", (q,)) for q in range(n)], 0, 1), tile_bits=T).compiled(eng, None)
for _ in range(2):
    psi = eng.simulate(cp, xd)
torch.cuda.synchronize(); print("ok", complex(psi[0, 0]))
