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
        Tp, a_in, a_out, last, r0, r1, s0, s1, mat, nfresh, ntabs = w[pos:pos + 11]
        tq = [int(q) for q in w[pos + 11:pos + 11 + Tp]]; pos += 11 + Tp
        passes.append((Tp, a_in, a_out, last, r0, r1, s0, s1, mat, nfresh, ntabs, tq))
    rounds = [tuple(w[pos + 6 * i: pos + 6 * i + 6]) for i in range(nr)]; pos += 6 * nr
    steps = [tuple(w[pos + 7 * i: pos + 7 * i + 7]) for i in range(nu)]; pos += 7 * nu
    ninit = w[pos]
    print(f"n={n} T={T} gates={len(prog)} passes={npass} rounds={nr} steps={nu} folded={ninit}")
    names = {1: "U2", 2: "DIAG", 5: "CTRL", 6: "L_RX", 7: "L_REAL", 8: "L_GEN", 9: "DTAB", 10: "S_RX", 11: "S_RY"}
    qs = lambda m: [q for q in range(n) if m >> q & 1]
    for i, (Tp, a_in, a_out, last, r0, r1, s0, s1, mat, nfresh, ntabs, tq) in enumerate(passes):
        print(f" pass {i}: tile={tq} in={len(qs(a_in))}q out={len(qs(a_out))}q fresh={nfresh} last={last} "
              f"rounds={r1 - r0} steps={s1 - s0} tables={ntabs} mat_elems={mat}")
        for r in range(r0, r1):
            nb, b0, b1, b2, rs0, rs1 = rounds[r]
            desc = []
            for u in range(rs0, rs1):
                kind, gate, a, b, ctrl, m, sw = steps[u]
                if kind in (6, 7, 8, 10, 11):
                    desc.append(f"{names[kind]}[slots {[c for c in range(3) if a >> c & 1]}]")
                elif kind == 9:
                    desc.append(f"DTAB[bits {a}..{a + b - 1}]")
                else:
                    desc.append(f"{names[kind]}:{prog.gates[gate].name}{list(prog.gates[gate].qubits)}")
            print(f"   round {r}: bits={[int(b0), int(b1), int(b2)][:nb]} " + " ".join(desc))

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    stats(sq.ChebyshevPQC(n, 2), 4, T)
