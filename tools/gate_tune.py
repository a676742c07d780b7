"""Gate-pass tuning sweep at bench size: one subprocess per (B200SQ_TILE_G, B200SQ_JIT_MINBLOCKS, B200SQ_ROT_LIFT,
tile bits) setting (the library reads these once per process); prints the simulate time of ChebyshevPQC(n, 2)."""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, json
sys.path.insert(0, %r)
import numpy as np, torch
import squlearn_b200 as sq
n = int(os.environ.get("QUBITS", 16)); N = int(os.environ.get("POINTS", 4096)); T = int(os.environ.get("TILE", 0))
ex = sq.Executor("b200"); eng = ex.engine
X = np.random.default_rng(3).uniform(-0.95, 0.95, (N, 4)); xd = torch.from_numpy(X).cuda()
enc = sq.ChebyshevPQC(n, 2); p = enc.generate_initial_parameters(4, seed=0)
cp = sq.B200Circuit(enc.get_program(4), tile_bits=T).compiled(eng, p)
for _ in range(3): eng.simulate(cp, xd)
torch.cuda.synchronize(); ts = []
for _ in range(5):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); eng.simulate(cp, xd); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
print(json.dumps({"ms_min": min(ts), "ms_med": sorted(ts)[2], "passes": cp.num_passes}))
''' % ROOT

def run(**env):
    e = dict(os.environ); e.update({k: str(v) for k, v in env.items()})
    r = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True, timeout=300)
    try:
        out = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        out = {"error": (r.stderr or r.stdout)[-300:]}
    out.update(env)
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    for g in (2, 1):
        for mb in (2, 3, 4):
            run(B200SQ_TILE_G=g, B200SQ_JIT_MINBLOCKS=mb)
    run(B200SQ_TILE_G=1, B200SQ_JIT_MINBLOCKS=3, B200SQ_ROT_LIFT=1)
    run(B200SQ_TILE_G=2, B200SQ_JIT_MINBLOCKS=2, TILE=12)
    run(B200SQ_TILE_G=1, B200SQ_JIT_MINBLOCKS=2, TILE=12)
    run(B200SQ_TILE_G=2, B200SQ_JIT_MINBLOCKS=2, TILE=10)
    run(B200SQ_JIT=0)
