"""Multi-GPU parity (needs >= 2 CUDA devices, skipped otherwise): the sharded Gram over NCCL / NVLink equals the
single-GPU Gram and the oracle within tolerance — FP64 state exchange, and the INT8 digit-plane path with each of
its exchange modes (copy-engine pushes over CUDA IPC with in-kernel ready flags, NCCL point-to-point, NCCL
all-gather), repeated calls (epochs, buffer reuse), the shared pinned host gather, a C3-size check on 256 points,
and the parameter-shift batch sharded by shifted variant.

Run on hardware with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_distributed.py -m gpu` (logs under
profiles/)."""

import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MODES = ("ipc", "p2p", "allgather")


def _data(seed, n_points, d):
    return np.random.default_rng(seed).uniform(-0.95, 0.95, (n_points, d))


def _worker(rank, world, port, out_dir, big):
    import torch
    import torch.distributed as dist

    import squlearn_b200 as sq
    from squlearn_b200.distributed import ShardedFidelityKernel, sharded_parameter_shift

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ex = sq.Executor("b200", device=rank)
        save = lambda name, arr: np.save(os.path.join(out_dir, name + ".npy"), arr) if rank == 0 else None
        if not big:
            fk = sq.FidelityKernel(sq.ChebyshevPQC(9, 2), ex)
            X = _data(0, 301, 3)
            save("k_fp64", ShardedFidelityKernel(fk, precision="fp64").evaluate(X))
            save("k_single", fk.evaluate(X))
            for mode in MODES:
                os.environ["B200SQ_DIST_EXCHANGE"] = mode
                sh = ShardedFidelityKernel(fk, precision="int8")
                # three different data sets through ONE object: epochs advance, arenas and planes are reused
                for rep, n_points in enumerate((301, 301, 304)):
                    k = sh.evaluate(_data(10 + rep, n_points, 3))
                    if rank == 0:
                        save(f"k_{mode}_{rep}", np.array(k))   # copy: the shared host buffer is reused
                del sh
            os.environ.pop("B200SQ_DIST_EXCHANGE", None)
            # CRT arithmetic over the ipc exchange: residue planes pushed whole, pushed in sub-blocks while the later
            # sub-blocks are still being simulated, and digit planes converted by the receiver
            for tag, env in (("res", {}), ("sub", {"B200SQ_DIST_SUBBLOCKS": "2"}),
                             ("slc", {"B200SQ_DIST_SUBBLOCKS": "2", "B200SQ_DIST_PIPELINE": "slice"}), ("dig", {"B200SQ_DIST_CRT_EXCHANGE": "digits"})):
                os.environ.update(env)
                sh = ShardedFidelityKernel(fk, precision="crt")
                for rep, n_points in enumerate((301, 301, 304)):
                    k = sh.evaluate(_data(10 + rep, n_points, 3))
                    if rank == 0:
                        save(f"k_crt{tag}_{rep}", np.array(k))
                del sh
                for key in env:
                    os.environ.pop(key, None)
            # parameter shift sharded by shifted variant (BASELINE configs[3] shape, fewer samples)
            enc = sq.ChebyshevRx(8, 4, closed=False)
            circ = sq.B200Circuit(enc.get_program(2), sq.SummedPaulis(8))
            Xq = _data(4, 96, 2)
            p = enc.generate_initial_parameters(2, seed=0)
            g = sharded_parameter_shift(circ, ex, p, Xq, np.ones(9))
            save("grad_sharded", g)
            save("grad_single", sq.b200_gradient(circ, ex, param=p, x=Xq, param_obs=np.ones(9)))
        else:
            n, n_points = 16, 4096
            fk = sq.FidelityKernel(sq.ChebyshevPQC(n, 2), ex)
            X = _data(3, n_points, 4)
            sh = ShardedFidelityKernel(fk)                      # default policy: INT8 planes, ipc pushes
            k = sh.evaluate(X)
            k2 = sh.evaluate(X)
            if rank == 0:
                idx = np.arange(0, n_points, n_points // 256)
                save("big_sub", k[np.ix_(idx, idx)])
                save("big_sym", np.array([float(np.max(np.abs(k - k.T))), float(np.max(np.abs(np.diag(k) - 1.0))),
                                          float(np.max(np.abs(k - k2)))]))
                psi = fk.states(X[:1024])
                k64 = ex.engine.gram(psi, None, precision="fp64").cpu().numpy()
                save("big_vs_fp64", np.array([float(np.max(np.abs(k[:1024, :1024] - k64)))]))
        torch.cuda.synchronize()
        assert ex.engine.ready_timeouts() == 0      # every pushed block arrived (no flag wait gave up)
    finally:
        dist.destroy_process_group()


def _spawn(tmp_path, big):
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path), big), nprocs=world, join=True)
    return world


def test_sharded_gram_nccl(tmp_path):
    import oracle

    _spawn(tmp_path, False)
    gates, p = oracle.chebyshev_pqc(9, 2, 3), oracle.chebyshev_pqc_initial_parameters(9, 2, 3, seed=0)
    X = _data(0, 301, 3)
    want = oracle.fidelity_kernel(gates, 9, X, None, p, pairloop=False)
    got, single = np.load(tmp_path / "k_fp64.npy"), np.load(tmp_path / "k_single.npy")
    assert np.max(np.abs(got - want)) < 1e-11
    assert np.max(np.abs(got - single)) < 1e-13
    assert np.array_equal(got, got.T)
    for mode in MODES:
        for rep, n_points in enumerate((301, 301, 304)):
            want = oracle.fidelity_kernel(gates, 9, _data(10 + rep, n_points, 3), None, p, pairloop=False)
            got8 = np.load(tmp_path / f"k_{mode}_{rep}.npy")
            assert got8.shape == want.shape
            assert np.max(np.abs(got8 - want)) < 1e-10, (mode, rep)
            assert np.array_equal(got8, got8.T), (mode, rep)
            assert np.all(np.diag(got8) == 1.0)
    for rep, n_points in enumerate((301, 301, 304)):
        want = oracle.fidelity_kernel(gates, 9, _data(10 + rep, n_points, 3), None, p, pairloop=False)
        ks = [np.load(tmp_path / f"k_crt{tag}_{rep}.npy") for tag in ("res", "sub", "slc", "dig")]
        for k in ks:
            assert k.shape == want.shape and np.max(np.abs(k - want)) < 1e-10, rep
            assert np.array_equal(k, k.T) and np.all(np.diag(k) == 1.0)
        # the exchange format does not change a single bit: every rank works with the same integers
        assert all(np.array_equal(ks[0], k) for k in ks[1:])
    g, g1 = np.load(tmp_path / "grad_sharded.npy"), np.load(tmp_path / "grad_single.npy")
    assert g.shape == g1.shape == (96, 1, 64)
    assert np.max(np.abs(g - g1)) < 1e-12


def test_sharded_gram_c3_size_against_oracle_on_256_points(tmp_path):
    """BASELINE configs[2] at full size (16 qubits, N = 4096) over all GPUs of the box: a 256-point sub-matrix
    against the oracle, a 1024-point block against the single-GPU FP64 Gram, exact symmetry, unit diagonal and
    run-to-run agreement."""
    import oracle

    _spawn(tmp_path, True)
    n, n_points = 16, 4096
    X = _data(3, n_points, 4)
    idx = np.arange(0, n_points, n_points // 256)
    p = oracle.chebyshev_pqc_initial_parameters(n, 2, 4, seed=0)
    want = oracle.fidelity_kernel(oracle.chebyshev_pqc(n, 2, 4), n, X[idx], None, p, pairloop=False)
    got = np.load(tmp_path / "big_sub.npy")
    err = float(np.max(np.abs(got - want)))
    sym, diag, rerun = np.load(tmp_path / "big_sym.npy")
    vs64 = float(np.load(tmp_path / "big_vs_fp64.npy")[0])
    print(f"C3 sharded: |K - oracle|(256 pts) = {err:.2e}, |K - K_fp64|(1024 block) = {vs64:.2e}, "
          f"asym = {sym:.1e}, rerun = {rerun:.1e}")
    assert err < 1e-10
    assert vs64 < 1e-10
    assert sym == 0.0 and diag == 0.0
    assert rerun < 1e-12
