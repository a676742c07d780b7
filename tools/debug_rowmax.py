import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import squlearn_b200 as sq
ex = sq.Executor("b200"); eng = ex.engine
for n, N, tb in ((12, 300, 0), (12, 300, 7), (6, 40, 0), (16, 64, 0)):
    X = np.random.default_rng(12).uniform(-0.95, 0.95, (N, 3))
    enc = sq.ChebyshevPQC(n, 2)
    cp = sq.B200Circuit(enc.get_program(3), tile_bits=tb).compiled(eng, enc.generate_initial_parameters(3, seed=0))
    psi = eng.simulate(cp, X)
    rm = eng._row_max_of(psi)
    comp = psi.view(torch.float64).abs().amax(dim=1)
    want = (comp.view(torch.int64) >> 32).to(torch.int32)
    mod = psi.abs().amax(dim=1)
    got = torch.zeros_like(comp).view(torch.int64); got[:] = rm.to(torch.int64) << 32
    gotf = got.view(torch.float64)
    print(n, N, tb, "equal", bool(torch.equal(rm, want)), "got/comp", float((gotf / comp).min()), float((gotf / comp).max()),
          "got/mod", float((gotf / mod).min()), float((gotf / mod).max()), "passes", cp.num_passes)
