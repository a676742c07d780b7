"""Timing (1 GPU) of the pieces the multi-GPU CRT path adds per received block: digits -> residues conversion alone and
next to a running Gram launch, and the fused CRT + digit slicing of a row block."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import squlearn_b200 as sq
n, rows = 16, int(os.environ.get("ROWS", 512))
ex = sq.Executor("b200"); eng = ex.engine
dim = 1 << n
X = np.random.default_rng(3).uniform(-0.95, 0.95, (rows, 4))
enc = sq.ChebyshevPQC(n, 2)
cp = sq.B200Circuit(enc.get_program(4)).compiled(eng, enc.generate_initial_parameters(4, seed=0))
psi = eng.simulate(cp, X)
s = eng.plane_count(dim, "crt")
_, cpl, cex = eng.alloc_plane_block(rows, dim, "crt")
digits = torch.empty((6, rows, 2 * dim), dtype=torch.int8, device="cuda")
conv = torch.empty((s, rows, 2 * dim), dtype=torch.int8, device="cuda")
def timeit(f, reps=5):
    f(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
out = {"rows": rows}
out["slice_crt_ms"] = timeit(lambda: eng.slice_states(psi.clone(), cpl, cex, 0, "crt"))
out["slice_crt_digits_ms"] = timeit(lambda: eng.slice_states(psi.clone(), cpl, cex, 0, "crt", digits=digits))
out["convert_alone_ms"] = timeit(lambda: eng.digits_to_crt(digits, rows, 0, rows, dim, conv, rows, 0))
# next to a Gram launch: the symmetric block of these rows runs on the main stream, the conversion on a side stream
k = torch.empty((rows, rows), dtype=torch.float64, device="cuda")
side = torch.cuda.Stream(priority=-1)
def both():
    job = [{"x0": 0, "nx": rows, "planes_y": cpl, "expo_y": cex, "rows_y": rows, "y0": 0, "ny": rows, "out": k,
            "symmetric": True, "diagonal": "one"}]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    side.wait_stream(torch.cuda.current_stream())
    eng.gram_jobs(cpl, cex, job)
    with torch.cuda.stream(side):
        e0.record(side); eng.digits_to_crt(digits, rows, 0, rows, dim, conv, rows, 0); e1.record(side)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)
both()
out["convert_beside_gram_ms"] = min(both() for _ in range(5))
out["gram_block_ms"] = timeit(lambda: eng.gram_jobs(cpl, cex, [{"x0": 0, "nx": rows, "planes_y": cpl, "expo_y": cex, "rows_y": rows,
                                                               "y0": 0, "ny": rows, "out": k, "symmetric": True, "diagonal": "one"}]))
print(json.dumps(out))
