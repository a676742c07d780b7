"""Phase timing of ShardedFidelityKernel.evaluate under torchrun (wall clock with syncs, rank 0 prints)."""
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
def mark(name):
    torch.cuda.synchronize(dev)
    marks.append((name, time.perf_counter()))
D._mark = mark
for it in range(4):
    marks.clear()
    dist.barrier(); torch.cuda.synchronize(dev)
    mark("start")
    rows, _ = sh.evaluate_rows(X)
    mark("rows_done")
    dist.barrier(); mark("barrier")
    k = sh.evaluate(X)
    mark("evaluate_total")
if rank == 0:
    t0 = marks[0][1]
    print(json.dumps([(nm, round((t - t0) * 1e3, 2)) for nm, t in marks]))
dist.barrier()
dist.destroy_process_group()
