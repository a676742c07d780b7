"""GPU timeline of ShardedFidelityKernel.evaluate_rows under torchrun WITHOUT synchronising between the phases:
every phase mark records a CUDA event on the stream that carries the phase (main, push or conversion stream) plus
the host clock; after the step the event times (relative to the step's first event) and the host issue times are
printed for every rank.  usage: torchrun ... tools/dist_timeline.py   (QUBITS / POINTS from the environment;
VARIANTS="1:full,4:full,4:slice" runs several settings of B200SQ_DIST_SUBBLOCKS:B200SQ_DIST_PIPELINE[:exchange[:order[:push streams]]]
in one process group)"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import squlearn_b200 as sq
from squlearn_b200 import distributed as D

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n, N = int(os.environ.get("QUBITS", 16)), int(os.environ.get("POINTS", 4096))
X = np.random.default_rng(3).uniform(-0.95, 0.95, (N, 4))
ex = sq.Executor("b200", device=lr)
fk = sq.FidelityKernel(sq.ChebyshevPQC(n, 2), ex)
marks = []
def mark(name, stream=None):
    ev = torch.cuda.Event(enable_timing=True)
    ev.record(stream if stream is not None else torch.cuda.current_stream(dev))
    marks.append((name, time.perf_counter(), ev))
D._mark = mark
for variant in os.environ.get("VARIANTS", "").split(",") if os.environ.get("VARIANTS") else [""]:
    if variant:
        f = variant.split(":")
        os.environ["B200SQ_DIST_SUBBLOCKS"] = f[0]
        os.environ["B200SQ_DIST_PIPELINE"] = f[1] if len(f) > 1 else "full"
        os.environ["B200SQ_DIST_CRT_EXCHANGE"] = f[2] if len(f) > 2 else "auto"
        os.environ["B200SQ_DIST_ORDER"] = f[3] if len(f) > 3 else "ring"
        os.environ["B200SQ_DIST_PUSH_STREAMS"] = f[4] if len(f) > 4 else "1"
    sh = D.ShardedFidelityKernel(fk)
    out, totals = None, []
    for it in range(6):
        marks.clear()
        dist.barrier(); torch.cuda.synchronize(dev)
        mark("start")
        sh._borrow_rows = True
        rows, _ = sh.evaluate_rows(X)
        mark("rows_done")
        torch.cuda.synchronize(dev)
        t_end = time.perf_counter()
        t0, e0 = marks[0][1], marks[0][2]
        totals.append(round((t_end - t0) * 1e3, 2))
        out = {"rank": rank, "host_total_ms": totals[-1],
               "marks": [(nm, round((t - t0) * 1e3, 2), round(e0.elapsed_time(ev), 2)) for nm, t, ev in marks]}
    out["totals_ms"] = totals
    # parity of this variant's result: 64 local rows x 128 columns spread over all ranks vs the FP64 DMMA Gram
    r0 = D.row_partition(N, world)[rank][0]
    idx = np.arange(5, N, N // 128)[:128]
    k64 = ex.engine.gram(fk.states(X[r0:r0 + 64]), fk.states(X[idx]), precision="fp64")
    out["max_err_vs_fp64"] = float((rows[:64][:, torch.as_tensor(idx, device=dev)] - k64).abs().max())
    assert out["max_err_vs_fp64"] < 1e-10, out["max_err_vs_fp64"]
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print(f"# variant '{variant}' world {world}: (phase, host issue time ms, GPU completion time ms) of the last of 6 steps")
        for g in gathered[:2] + gathered[-1:]:
            print(json.dumps(g))
        print("# step totals (ms) per rank:", [g["totals_ms"][2:] for g in gathered][:3],
              "max |K - K_fp64| over ranks: %.2e" % max(g["max_err_vs_fp64"] for g in gathered), flush=True)
    del sh
    torch.cuda.synchronize(dev)
    dist.barrier()
dist.destroy_process_group()
