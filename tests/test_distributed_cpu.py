"""world_size-2 (and 3) gloo tests of the multi-GPU sharding logic on the CPU.

The sharding code is engine-agnostic; here a stand-in engine backed by the ORACLE supplies states
and Gram blocks (tests may use the oracle; the product never does), so what is tested is the row
partition, the symmetric half-ring block schedule, the mirror exchange and the final gather."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from tests.helpers import i8_gram, i8_slice
from squlearn_b200.distributed import (ShardedFidelityKernel, gram_jobs, row_partition,
                                       sharded_parameter_shift, symmetric_block_schedule)


def test_row_partition_and_schedule():
    assert row_partition(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert row_partition(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    for parts in range(1, 10):
        sched = symmetric_block_schedule(parts)
        pairs = set()
        for r, cols in enumerate(sched):
            for c in cols:
                key = (min(r, c), max(r, c))
                assert key not in pairs          # every unordered pair exactly once
                pairs.add(key)
        assert len(pairs) == parts * (parts + 1) // 2
        loads = [len(c) for c in sched]
        assert max(loads) - min(loads) <= 1      # balanced


def test_gram_jobs_cover_every_pair_once_and_balance():
    for world in range(1, 9):
        for n in (world * 8, world * 8 + 3, 5):
            parts = row_partition(n, world)
            jobs = gram_jobs(parts)
            cover = np.zeros((n, n), dtype=int)
            loads = []
            for r, js in enumerate(jobs):
                load = 0.0
                for j in js:
                    (a, b), (c, d) = j["rows"], j["cols"]
                    assert parts[r][0] <= a <= b <= parts[r][1]          # rows are local
                    assert parts[j["owner"]][0] <= c <= d <= parts[j["owner"]][1]
                    if j["diag"]:
                        cover[a:b, c:d] += 1
                        load += 0.5 * (b - a) * (d - c)
                    else:
                        cover[a:b, c:d] += 1
                        cover[c:d, a:b] += 1
                        load += (b - a) * (d - c)
                loads.append(load)
            assert np.all(cover == 1), (world, n)
            if n % (2 * world) == 0:
                assert max(loads) == min(loads)                          # perfectly balanced


class _OracleEngine:
    device = "cpu"

    def gram(self, a, b=None, diagonal="one"):
        an = a.numpy()
        if b is None:
            k = oracle.fidelity_gram_zgemm(an, an, True, "off_diagonal" if diagonal == "one" else "all")
            if diagonal == "computed":
                k = np.abs(an.conj() @ an.T) ** 2
            return torch.from_numpy(k)
        return torch.from_numpy(np.abs(an.conj() @ b.numpy().T) ** 2)


class _PlanesEngine(_OracleEngine):
    """Stand-in with the digit-plane entry points of the INT8 path (NumPy restatement in tests/helpers.py):
    exercises the plane exchange of ShardedFidelityKernel under gloo."""

    @staticmethod
    def gram_precision(nx, ny, dim, precision=None):
        return "int8"

    def alloc_planes(self, plane_rows, dim):
        return (torch.zeros((12, plane_rows, 2 * dim), dtype=torch.int8), torch.zeros((plane_rows,), dtype=torch.int32))

    def alloc_plane_block(self, rows, dim):
        nb = 12 * rows * 2 * dim
        buf = torch.zeros((nb + 4 * rows,), dtype=torch.int8)
        return buf, buf[:nb].view(12, rows, 2 * dim), buf[nb:].view(torch.int32)

    def alloc_bytes(self, nbytes):
        return torch.zeros((int(nbytes),), dtype=torch.int8)

    def gram_jobs(self, px, ex, jobs):
        for j in jobs:
            py, ey = j["planes_y"], j["expo_y"]
            assert not isinstance(py, int)          # raw pointers only appear on the CUDA ipc path
            if j.get("symmetric"):
                k = self.gram_from_planes(px, ex, j["x0"], j["nx"], diagonal=j.get("diagonal", "one"))
            else:
                assert py.shape[0] >= 6 and py.shape[1] == j["rows_y"]
                k = self.gram_from_planes(px, ex, j["x0"], j["nx"], py, ey, j["y0"], j["ny"])
            j["out"].copy_(k)

    def slice_states(self, psi, planes, expo, row0=0):
        if psi.shape[0] == 0:
            return
        p, e = i8_slice(psi.numpy())
        planes[:, row0:row0 + psi.shape[0]] = torch.from_numpy(p)
        expo[row0:row0 + psi.shape[0]] = torch.from_numpy(e)

    def gram_from_planes(self, px, ex, x0, nx, py=None, ey=None, y0=0, ny=0, diagonal="one", out=None):
        sym = py is None
        if sym:
            py, ey, y0, ny = px, ex, x0, nx
        k = i8_gram(px[:, x0:x0 + nx].numpy(), ex[x0:x0 + nx].numpy(), py[:, y0:y0 + ny].numpy(), ey[y0:y0 + ny].numpy())
        if sym:
            k = np.tril(k) + np.tril(k, -1).T
            if diagonal == "one":
                np.fill_diagonal(k, 1.0)
            elif diagonal == "zero":
                np.fill_diagonal(k, 0.0)
        return torch.from_numpy(k)


class _OracleExecutor:
    engine = _OracleEngine()


class _OracleKernel:
    def __init__(self, n, layers, d):
        self.n, self.d = n, d
        self.gates = oracle.chebyshev_pqc(n, layers, d)
        self.p = oracle.chebyshev_pqc_initial_parameters(n, layers, d, seed=0)
        self._executor = _OracleExecutor()

    def states(self, x):
        if len(x) == 0:
            return torch.zeros((0, 2 ** self.n), dtype=torch.complex128)
        return torch.from_numpy(oracle.simulate(self.gates, self.n, x, self.p))


def _worker(rank, world, port, n_points, out_dir, planes=False, exchange="auto"):
    os.environ["B200SQ_DIST_EXCHANGE"] = exchange
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        X = rng.uniform(-0.95, 0.95, (n_points, 2))
        kern = _OracleKernel(4, 2, 2)
        if planes:
            kern._executor = type("E", (), {"engine": _PlanesEngine()})()
        full = ShardedFidelityKernel(kern).evaluate(X)
        if rank == 0:
            np.save(os.path.join(out_dir, f"k_{world}_{n_points}.npy"), full)
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world,n_points", [(2, 10), (3, 7), (2, 1), (4, 13)])
def test_sharded_gram_gloo(tmp_path, world, n_points):
    mp.spawn(_worker, args=(world, _free_port(), n_points, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / f"k_{world}_{n_points}.npy")
    rng = np.random.default_rng(0)
    X = rng.uniform(-0.95, 0.95, (n_points, 2))
    want = oracle.fidelity_kernel(oracle.chebyshev_pqc(4, 2, 2), 4, X, None,
                                  oracle.chebyshev_pqc_initial_parameters(4, 2, 2, seed=0))
    assert got.shape == (n_points, n_points)
    assert np.max(np.abs(got - want)) < 1e-13
    assert np.array_equal(got, got.T)


@pytest.mark.parametrize("world,n_points,exchange", [(2, 10, "auto"), (3, 7, "auto"), (4, 13, "auto"),
                                                     (3, 12, "allgather"), (4, 16, "allgather"), (2, 8, "allgather")])
def test_sharded_gram_gloo_digit_planes(tmp_path, world, n_points, exchange):
    """The INT8 path exchanges digit planes instead of states: same result within the scheme's error.
    Point-to-point messages (default) and one all-gather of equal-sized plane blocks (B200SQ_DIST_EXCHANGE)."""
    mp.spawn(_worker, args=(world, _free_port(), n_points, str(tmp_path), True, exchange), nprocs=world, join=True)
    got = np.load(tmp_path / f"k_{world}_{n_points}.npy")
    rng = np.random.default_rng(0)
    X = rng.uniform(-0.95, 0.95, (n_points, 2))
    want = oracle.fidelity_kernel(oracle.chebyshev_pqc(4, 2, 2), 4, X, None,
                                  oracle.chebyshev_pqc_initial_parameters(4, 2, 2, seed=0))
    assert np.max(np.abs(got - want)) < 1e-10
    assert np.array_equal(got, got.T)


def test_split_integer_scheme_error_on_circuit_states():
    """NumPy restatement of the INT8 Gram (digit planes, classes c <= 5 + (3,3)) against complex128."""
    n = 10
    rng = np.random.default_rng(1)
    X = rng.uniform(-0.95, 0.95, (40, 3))
    p = oracle.chebyshev_pqc_initial_parameters(n, 2, 3, seed=0)
    sv = oracle.simulate(oracle.chebyshev_pqc(n, 2, 3), n, X, p)
    sv = np.concatenate([sv, sv[:2], np.eye(1, 2 ** n, 3, dtype=complex)])   # duplicates and a basis state
    planes, expo = i8_slice(sv)
    assert planes[0].min() >= -65 and planes[0].max() <= 65
    got = i8_gram(planes, expo, planes, expo)
    want = np.abs(sv.conj() @ sv.T) ** 2
    assert np.max(np.abs(got - want)) < 2e-11


# ------------------------------------------------------------------------------------------------
# Parameter shift sharded by shifted variant (north star; optree_derivative.py:130-158): every rank evaluates a
# slice of the variant list on the full batch, the partial gradients are all-reduced.  The engine is the host
# emulator (the product's own plan + per-thread kernel code on the CPU).
def _shift_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import squlearn_b200 as sq
        from tests.emu_engine import EmuExecutor

        ex = EmuExecutor()
        enc = sq.ChebyshevRx(4, 2, closed=False)
        circ = sq.B200Circuit(enc.get_program(2), [sq.SummedPaulis(4), sq.SinglePauli(4, 1, "X")])
        X = np.random.default_rng(4).uniform(-0.9, 0.9, (7, 2))
        p = enc.generate_initial_parameters(2, seed=0)
        pobs = np.linspace(0.2, 1.0, 5)
        g = sharded_parameter_shift(circ, ex, p, X, pobs)
        np.save(os.path.join(out_dir, f"g_{rank}.npy"), g)
        if rank == 0:
            np.save(os.path.join(out_dir, "g_single.npy"),
                    sq.b200_gradient(circ, ex, param=p, x=X, param_obs=pobs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_parameter_shift_gloo(tmp_path, world):
    mp.spawn(_shift_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    want = np.load(tmp_path / "g_single.npy")
    assert want.shape == (7, 2, 16)
    for r in range(world):                                   # every rank holds the complete gradient
        assert np.max(np.abs(np.load(tmp_path / f"g_{r}.npy") - want)) < 1e-13
    # and it is the oracle's gradient (generator route, independent of the shift rule)
    X = np.random.default_rng(4).uniform(-0.9, 0.9, (7, 2))
    p = oracle.chebyshev_rx_initial_parameters(4, 2, 2, seed=0)
    ref = oracle.qnn_evaluate(oracle.chebyshev_rx(4, 2, 2, closed=False), 4, X, p,
                              [("summed_paulis", {}), ("single_pauli", {"qubit": 1, "op_str": "X"})],
                              np.linspace(0.2, 1.0, 5), ("dfdp",))["dfdp"]
    assert np.max(np.abs(want - ref)) < 1e-12


def test_crt_scheme_error_on_circuit_states():
    """NumPy restatement of the CRT (residue-number) form of the INT8 Gram against complex128: the only error is the
    B-bit quantisation of the inputs relative to the row norm (no product terms are dropped), and K is exactly
    symmetric; the worst-case bound 2 sqrt(K') 2^-B of a unit-norm entry stays below 1e-10."""
    from tests.helpers import crt_gram, crt_num_moduli, crt_slice

    from tests.helpers import crt_plan

    assert [crt_num_moduli(1 << n) for n in (6, 12, 16, 20)] == [10, 11, 11, 12]
    assert [crt_plan(1 << n)[1] for n in (6, 12, 16, 20)] == [39, 43, 43, 46]
    for n_ in (6, 12, 16, 18, 20):
        assert 2.0 * np.sqrt(2.0 ** (n_ + 1)) * 2.0 ** -crt_plan(1 << n_)[1] <= 1e-10
    n = 10
    rng = np.random.default_rng(1)
    X = rng.uniform(-0.95, 0.95, (40, 3))
    p = oracle.chebyshev_pqc_initial_parameters(n, 2, 3, seed=0)
    sv = oracle.simulate(oracle.chebyshev_pqc(n, 2, 3), n, X, p)
    sv = np.concatenate([sv, sv[:2], np.eye(1, 2 ** n, 3, dtype=complex)])
    planes, expo = crt_slice(sv)
    assert planes.shape[0] == 22 and np.array_equal(expo[:40], np.zeros(40, dtype=np.int32))
    got = crt_gram(planes, expo, planes, expo)
    want = np.abs(sv.conj() @ sv.T) ** 2
    assert np.max(np.abs(got - want)) < 5e-12
    assert np.array_equal(got, got.T)
    # rows of any scale: the exponent follows the row norm
    sv2 = sv * np.array([3.7, 1e-3, 0.49, 17.0] * 11)[:sv.shape[0], None]
    pl2, ex2 = crt_slice(sv2)
    got2 = crt_gram(pl2, ex2, pl2, ex2)
    want2 = np.abs(sv2.conj() @ sv2.T) ** 2
    n2 = np.sum(np.abs(sv2) ** 2, axis=1)
    assert np.max(np.abs(got2 - want2) / (n2[:, None] * n2[None, :])) < 1e-11
    # the exchange digits are the digits of the same integers
    from tests.helpers import crt_digits

    dg, ex3 = crt_digits(sv2)
    q = sum(dg[s_].astype(np.int64) << (8 * (5 - s_)) for s_ in range(6))
    assert np.array_equal(ex3, ex2)
    for mi, p_ in enumerate(__import__("tests.helpers", fromlist=["CRT_MODULI"]).CRT_MODULI[:pl2.shape[0] // 2]):
        r = np.mod(q, p_)
        assert np.array_equal(np.where(r > (p_ - 1) // 2, r - p_, r).astype(np.int8), pl2[mi])


def test_push_schedule_covers_every_row_once_and_in_order():
    """Sub-block pipelining of the residue exchange: every push's rows go out exactly once, in row order, never
    before the sub-block that holds them has been sliced, and the arrival flag follows the last chunk."""
    import random

    from squlearn_b200.distributed import push_schedule, row_partition

    rnd = random.Random(7)
    for _ in range(300):
        n, ns = rnd.randint(64, 700), rnd.randint(1, 6)
        pushes = []
        for _k in range(rnd.randint(1, 5)):
            a = rnd.randint(0, n - 1)
            pushes.append({"a": a, "b": rnd.randint(a + 1, n)})
        bounds = [ab for ab in row_partition(n, ns) if ab[1] > ab[0]]
        stages = push_schedule(pushes, bounds, fill=rnd.randint(0, 3))
        assert len(stages) == len(bounds)
        cov, finished = {pi: [] for pi in range(len(pushes))}, set()
        for s_, stage in enumerate(stages):
            for pi, lo, hi, last in stage:
                assert hi <= bounds[s_][1] and pi not in finished
                cov[pi].append((lo, hi))
                if last:
                    finished.add(pi)
        for pi, p in enumerate(pushes):
            assert cov[pi] == sorted(cov[pi]) and cov[pi][0][0] == p["a"] and cov[pi][-1][1] == p["b"]
            assert all(b == c for (_a, b), (c, _d) in zip(cov[pi], cov[pi][1:])) and pi in finished
    # the nearest consumer's block completes right after the last sub-block, not after all earlier pushes
    st = push_schedule([{"a": 0, "b": 512}] * 4, row_partition(512, 4), fill=1)
    assert st[3][0] == (0, 384, 512, True) and [len(x) for x in st] == [2, 2, 2, 10]
