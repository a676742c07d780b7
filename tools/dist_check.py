"""Sharded Gram (all exchange modes) against the single-GPU Gram; rank 0 prints the max differences."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import squlearn_b200 as sq
from squlearn_b200.distributed import ShardedFidelityKernel

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n, N = int(os.environ.get("QUBITS", 12)), int(os.environ.get("POINTS", 1024))
X = np.random.default_rng(3).uniform(-0.95, 0.95, (N, 4))
ex = sq.Executor("b200", device=lr)
fk = sq.FidelityKernel(sq.ChebyshevPQC(n, 2), ex)
out = {}
for mode in ("p2p", "allgather"):
    os.environ["B200SQ_DIST_EXCHANGE"] = mode
    k = ShardedFidelityKernel(fk, precision="int8").evaluate(X)
    if rank == 0:
        ref = fk.evaluate(X)
        out[mode] = {"max_abs_diff_vs_single_gpu": float(np.max(np.abs(k - ref))), "symmetric": bool(np.array_equal(k, k.T))}
if rank == 0:
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
