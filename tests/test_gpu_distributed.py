"""Multi-GPU parity (needs >= 2 CUDA devices, skipped otherwise): the sharded Gram over NCCL equals
the single-GPU Gram bit for bit where the same kernel computed it, and the oracle within tolerance."""

import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    import squlearn_b200 as sq
    from squlearn_b200.distributed import ShardedFidelityKernel

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        X = np.random.default_rng(0).uniform(-0.95, 0.95, (301, 3))
        ex = sq.Executor("b200", device=rank)
        fk = sq.FidelityKernel(sq.ChebyshevPQC(9, 2), ex)
        full = ShardedFidelityKernel(fk).evaluate(X)
        planes = ShardedFidelityKernel(fk, precision="int8").evaluate(X)  # digit planes travel, not states
        if rank == 0:
            np.save(os.path.join(out_dir, "k.npy"), full)
            np.save(os.path.join(out_dir, "k8.npy"), planes)
            np.save(os.path.join(out_dir, "k1.npy"), fk.evaluate(X))
    finally:
        dist.destroy_process_group()


def test_sharded_gram_nccl(tmp_path):
    import torch

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    import torch.multiprocessing as mp

    import oracle

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got, single = np.load(tmp_path / "k.npy"), np.load(tmp_path / "k1.npy")
    X = np.random.default_rng(0).uniform(-0.95, 0.95, (301, 3))
    want = oracle.fidelity_kernel(oracle.chebyshev_pqc(9, 2, 3), 9, X, None,
                                  oracle.chebyshev_pqc_initial_parameters(9, 2, 3, seed=0), pairloop=False)
    assert np.max(np.abs(got - want)) < 1e-11
    assert np.max(np.abs(got - single)) < 1e-13
    assert np.array_equal(got, got.T)
    got8 = np.load(tmp_path / "k8.npy")
    assert np.max(np.abs(got8 - want)) < 1e-10
    assert np.array_equal(got8, got8.T)
