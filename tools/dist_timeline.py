"""GPU timeline of ShardedFidelityKernel.evaluate_rows under torchrun WITHOUT synchronising between the phases:
every phase mark records a CUDA event on the stream that carries the phase (main, push or conversion stream) plus
the host clock; after the step the event times (relative to the step's first event) and the host issue times are
printed for every rank.  usage: torchrun ... tools/dist_timeline.py   (QUBITS / POINTS from the environment)"""
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
sh = D.ShardedFidelityKernel(fk)
marks = []
def mark(name, stream=None):
    ev = torch.cuda.Event(enable_timing=True)
    ev.record(stream if stream is not None else torch.cuda.current_stream(dev))
    marks.append((name, time.perf_counter(), ev))
D._mark = mark
out = None
for it in range(5):
    marks.clear()
    dist.barrier(); torch.cuda.synchronize(dev)
    mark("start")
    sh._borrow_rows = True
    rows, _ = sh.evaluate_rows(X)
    mark("rows_done")
    torch.cuda.synchronize(dev)
    t_end = time.perf_counter()
    t0, e0 = marks[0][1], marks[0][2]
    out = {"rank": rank, "host_total_ms": round((t_end - t0) * 1e3, 2),
           "marks": [(nm, round((t - t0) * 1e3, 2), round(e0.elapsed_time(ev), 2)) for nm, t, ev in marks]}
gathered = [None] * world
dist.all_gather_object(gathered, out)
if rank == 0:
    print("# (phase, host issue time ms, GPU completion time ms) of the last of 5 steps")
    for g in gathered:
        print(json.dumps(g))
dist.barrier()
dist.destroy_process_group()
