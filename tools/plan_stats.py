"""Print the pass / round / micro-op structure the planner produces for a circuit (host only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import squlearn_b200 as sq
from squlearn_b200.engine import CompiledProgram

def stats(enc, d, tile_bits=0, params=None):
    prog = enc.get_program(d)
    p = enc.generate_initial_parameters(d, seed=0) if params is None else params
    cp = CompiledProgram(prog, p, tile_bits)
    w = cp.dump()
    n, T, npass, nr, nu = w[:5]
    pos = 5
    passes = []
    for _ in range(npass):
        Tp, first, r0, r1, u0, u1, mat = w[pos:pos + 7]
        tq = list(w[pos + 7:pos + 7 + Tp]); pos += 7 + Tp
        passes.append((Tp, first, r0, r1, u0, u1, mat, tq))
    rounds = [tuple(w[pos + 6 * i: pos + 6 * i + 6]) for i in range(nr)]; pos += 6 * nr
    uops = [tuple(w[pos + 7 * i: pos + 7 * i + 7]) for i in range(nu)]
    print(f"n={n} T={T} gates={len(prog)} passes={npass} rounds={nr} uops={nu}")
    for i, (Tp, first, r0, r1, u0, u1, mat, tq) in enumerate(passes):
        print(f" pass {i}: first={first} tile_qubits={tq} rounds={r1 - r0} uops={u1 - u0} mat_elems={mat}")
        for r in range(r0, r1):
            nb, b0, b1, b2, ru0, ru1 = rounds[r]
            kinds = [("U1", "U2", "DG", "RX", "RE", "CT")[uops[u][0]] for u in range(ru0, ru1)]
            print(f"   round {r}: bits={[b0, b1, b2][:nb]} uops={ru1 - ru0} " + " ".join(
                f"{k}{prog.gates[uops[u][1]].name}{list(prog.gates[uops[u][1]].qubits)}" for k, u in zip(kinds, range(ru0, ru1))))

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    stats(sq.ChebyshevPQC(n, 2), 4, T)
