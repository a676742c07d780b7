"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI / the
reference-shaped Python surface, against the oracle on the same seeded inputs.

Tolerances: the north star asks for 1e-10 absolute in complex128; the tests hold the engine to
1e-12 (states, expectation values) and 1e-11 (Gram entries, gradients).
"""

import numpy as np
import pytest

import oracle
from tests.helpers import random_circuit

pytestmark = pytest.mark.gpu

STATE_TOL = 1e-12
GRAM_TOL = 1e-11


@pytest.fixture(scope="module")
def sq():
    import squlearn_b200

    return squlearn_b200


@pytest.fixture(scope="module")
def executor(sq):
    return sq.Executor("b200")


def _states(sq, executor, program, p, X, tile_bits=0, variants=None):
    circ = sq.B200Circuit(program, tile_bits=tile_bits)
    eng = executor.engine
    return eng.simulate(circ.compiled(eng, p), X, variants=variants).cpu().numpy()


# ------------------------------------------------------------------------------ gate application
@pytest.mark.parametrize("n,layers,d,N,tile_bits", [
    (4, 2, 2, 100, 0),      # config C1
    (8, 2, 2, 33, 0),       # one warp-team per state
    (10, 2, 3, 17, 0),
    (12, 2, 4, 9, 0),       # whole state in one 64 KiB tile
    (13, 1, 4, 5, 13),      # 128 KiB tile
    (14, 2, 4, 6, 0),       # two passes, 4 outer tiles
    (9, 2, 4, 11, 5),       # forced small tiles: many passes, scattered tile qubits
    (16, 2, 4, 8, 0),       # config C3 shape, subset of points
])
def test_chebyshev_pqc_states(sq, executor, n, layers, d, N, tile_bits):
    rng = np.random.default_rng(3)
    X = rng.uniform(-0.95, 0.95, (N, d))
    enc = sq.ChebyshevPQC(n, layers)
    p = enc.generate_initial_parameters(d, seed=0)
    got = _states(sq, executor, enc.get_program(d), p, X, tile_bits)
    want = oracle.simulate(oracle.chebyshev_pqc(n, layers, d), n, X, p)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) < STATE_TOL


def test_hubregtsen_and_chebyshev_rx_states(sq, executor):
    rng = np.random.default_rng(2)
    X = rng.uniform(-np.pi, np.pi, (50, 10))
    enc = sq.HubregtsenEncodingCircuit(10, 1)
    p = enc.generate_initial_parameters(10, seed=0)
    got = _states(sq, executor, enc.get_program(10), p, X)
    want = oracle.simulate(oracle.hubregtsen(10, 1, 10), 10, X, p)
    assert np.max(np.abs(got - want)) < STATE_TOL
    X = rng.uniform(-0.95, 0.95, (1024, 2))
    enc = sq.ChebyshevRx(8, 4)
    p = enc.generate_initial_parameters(2, seed=0)
    got = _states(sq, executor, enc.get_program(2), p, X)
    want = oracle.simulate(oracle.chebyshev_rx(8, 4, 2), 8, X, p)
    assert np.max(np.abs(got - want)) < STATE_TOL


@pytest.mark.parametrize("seed", range(10))
def test_random_circuits_all_gates(sq, executor, seed):
    rng = np.random.default_rng(500 + seed)
    n = int(rng.integers(1, 12))
    tile_bits = 0 if n < 4 else int(rng.integers(4, n + 1))
    prog, ogates = random_circuit(n, 60, 5, 3, rng)
    X = rng.uniform(-0.9, 0.9, (7, 3))
    p = rng.uniform(-2, 2, 5)
    got = _states(sq, executor, prog, p, X, tile_bits)
    want = oracle.simulate(ogates, n, X, p)
    assert np.max(np.abs(got - want)) < STATE_TOL


def test_edge_cases_empty_circuit_single_qubit_and_table_angles(sq, executor):
    prog = sq.Program(3, [], 0, 1)
    got = _states(sq, executor, prog, None, np.zeros((4, 1)))
    want = np.zeros((4, 8), dtype=complex)
    want[:, 0] = 1
    assert np.array_equal(got, want)
    prog = sq.Program(1, [sq.GateOp("h", (0,)), sq.GateOp("rz", (0,), sq.Angle(feature=0))], 0, 1)
    X = np.linspace(-1, 1, 5).reshape(-1, 1)
    from oracle.statevector import Gate
    want = oracle.simulate([Gate("h", (0,), ()), Gate("rz", (0,), (lambda x, p: x[:, 0],))], 1, X)
    assert np.max(np.abs(_states(sq, executor, prog, None, X) - want)) < STATE_TOL
    prog = sq.Program(2, [sq.GateOp("ry", (1,), sq.Angle(table=lambda x, p: np.sin(x[:, 0]) * p[0])),
                          sq.GateOp("cx", (1, 0))], 1, 1)
    want = oracle.simulate([Gate("ry", (1,), (lambda x, p: np.sin(x[:, 0]) * p[0],)), Gate("cx", (1, 0), ())],
                           2, X, np.array([0.7]))
    assert np.max(np.abs(_states(sq, executor, prog, np.array([0.7]), X) - want)) < STATE_TOL


def test_large_state_subset_n20(sq, executor):
    """Config C5 shape (20 qubits) on a 3-point subset: three passes over 16 MiB states."""
    n, d = 20, 4
    rng = np.random.default_rng(5)
    X = rng.uniform(-0.95, 0.95, (3, d))
    enc = sq.ChebyshevPQC(n, 2)
    p = enc.generate_initial_parameters(d, seed=0)
    got = _states(sq, executor, enc.get_program(d), p, X)
    want = oracle.simulate(oracle.chebyshev_pqc(n, 2, d), n, X, p)
    assert np.max(np.abs(got - want)) < STATE_TOL


# ------------------------------------------------------------------------------ Gram
@pytest.mark.parametrize("three_m", [False, True])
@pytest.mark.parametrize("nx,ny,dim", [(1, 1, 2), (5, 3, 8), (37, 130, 64), (129, 65, 1024),
                                        (200, 200, 48), (64, 257, 4096), (700, 520, 2048)])
def test_gram_rectangular(executor, nx, ny, dim, three_m):
    import torch

    rng = np.random.default_rng(nx * 7 + ny)
    a = rng.normal(size=(nx, dim)) + 1j * rng.normal(size=(nx, dim))
    b = rng.normal(size=(ny, dim)) + 1j * rng.normal(size=(ny, dim))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    eng = executor.engine
    got = eng.gram(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), three_m=three_m).cpu().numpy()
    want = oracle.fidelity_gram_zgemm(a, b, False)
    assert got.shape == (nx, ny)
    assert np.max(np.abs(got - want)) < GRAM_TOL
    if nx * ny <= 400:
        assert np.max(np.abs(got - oracle.fidelity_gram_pairloop(a, b, False))) < GRAM_TOL


@pytest.mark.parametrize("three_m", [False, True])
@pytest.mark.parametrize("n,dim", [(1, 4), (63, 16), (128, 256), (300, 512), (513, 64), (1100, 1024)])
def test_gram_symmetric_and_diagonal_policies(executor, n, dim, three_m):
    import torch

    rng = np.random.default_rng(n)
    a = rng.normal(size=(n, dim)) + 1j * rng.normal(size=(n, dim))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    t = torch.from_numpy(a).cuda()
    eng = executor.engine
    k_one = eng.gram(t, None, diagonal="one", three_m=three_m).cpu().numpy()
    k_zero = eng.gram(t, None, diagonal="zero", three_m=three_m).cpu().numpy()
    k_cmp = eng.gram(t, None, diagonal="computed", three_m=three_m).cpu().numpy()
    assert np.array_equal(k_one, eng.gram(t, None, diagonal="one", three_m=three_m).cpu().numpy())  # deterministic
    want = oracle.fidelity_gram_pairloop(a, a, True, "off_diagonal") if n <= 130 else \
        oracle.fidelity_gram_zgemm(a, a, True, "off_diagonal")
    assert np.max(np.abs(k_one - want)) < GRAM_TOL
    assert np.array_equal(k_one, k_one.T)                 # exactly symmetric (mirrored)
    assert np.all(np.diag(k_one) == 1.0)
    assert np.all(np.diag(k_zero) == 0.0)                 # reference quirk for "all"
    assert np.max(np.abs(np.diag(k_cmp) - 1.0)) < GRAM_TOL
    off = ~np.eye(n, dtype=bool)
    assert np.array_equal(k_one[off], k_zero[off]) and np.array_equal(k_one[off], k_cmp[off])


# ---- split-integer Gram on the INT8 tcgen05 tensor cores (B200SQ_GRAM_INT8): north-star bound 1e-10
INT8_TOL = 1e-10


@pytest.mark.parametrize("scheme", ["int8", "crt"])
@pytest.mark.parametrize("nx,ny,dim", [(1, 1, 64), (5, 3, 64), (37, 130, 128), (129, 65, 1024),
                                        (200, 260, 72), (64, 257, 4096), (700, 520, 16384)])
def test_gram_int8_rectangular(executor, nx, ny, dim, scheme):
    import torch

    rng = np.random.default_rng(nx * 11 + ny)
    a = rng.normal(size=(nx, dim)) + 1j * rng.normal(size=(nx, dim))
    b = rng.normal(size=(ny, dim)) + 1j * rng.normal(size=(ny, dim))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    got = executor.engine.gram(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), precision=scheme).cpu().numpy()
    want = oracle.fidelity_gram_zgemm(a, b, False)
    assert got.shape == (nx, ny)
    assert np.max(np.abs(got - want)) < INT8_TOL


@pytest.mark.parametrize("scheme", ["int8", "crt"])
@pytest.mark.parametrize("n,dim", [(1, 64), (63, 128), (128, 256), (300, 512), (513, 64), (1100, 2048)])
def test_gram_int8_symmetric_and_diagonal_policies(executor, n, dim, scheme):
    import torch

    rng = np.random.default_rng(n)
    a = rng.normal(size=(n, dim)) + 1j * rng.normal(size=(n, dim))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    t = torch.from_numpy(a).cuda()
    eng = executor.engine
    k_one = eng.gram(t, None, diagonal="one", precision=scheme).cpu().numpy()
    k_zero = eng.gram(t, None, diagonal="zero", precision=scheme).cpu().numpy()
    k_cmp = eng.gram(t, None, diagonal="computed", precision=scheme).cpu().numpy()
    want = oracle.fidelity_gram_zgemm(a, a, True, "off_diagonal")
    assert np.max(np.abs(k_one - want)) < INT8_TOL
    assert np.array_equal(k_one, k_one.T)                 # exactly symmetric (mirrored)
    assert np.all(np.diag(k_one) == 1.0)
    assert np.all(np.diag(k_zero) == 0.0)
    assert np.max(np.abs(np.diag(k_cmp) - 1.0)) < INT8_TOL
    if n > 1:
        off = ~np.eye(n, dtype=bool)
        assert np.max(np.abs(k_one[off] - k_cmp[off])) < INT8_TOL


@pytest.mark.parametrize("scheme", ["int8", "crt"])
def test_gram_int8_adversarial_states(executor, scheme):
    """Rows with very different scales: basis states (one amplitude of modulus 1), exact duplicates,
    near-duplicates, a uniform superposition and states with a few large and many tiny amplitudes."""
    import torch

    dim = 4096
    rng = np.random.default_rng(7)
    rows = []
    e0 = np.zeros(dim, complex); e0[5] = 1.0
    e1 = np.zeros(dim, complex); e1[77] = 1j
    rows += [e0, e1, np.full(dim, 1 / np.sqrt(dim), complex)]
    g = rng.normal(size=dim) + 1j * rng.normal(size=dim)
    g /= np.linalg.norm(g)
    rows += [g, g.copy(), g * np.exp(0.3j)]
    near = g + 1e-7 * (rng.normal(size=dim) + 1j * rng.normal(size=dim))
    rows.append(near / np.linalg.norm(near))
    spiky = 1e-6 * (rng.normal(size=dim) + 1j * rng.normal(size=dim))
    spiky[:3] = [0.6, 0.5j, -0.62]
    rows.append(spiky / np.linalg.norm(spiky))
    for _ in range(140):
        v = rng.normal(size=dim) * np.exp(rng.uniform(-6, 0, dim)) + 1j * rng.normal(size=dim)
        rows.append(v / np.linalg.norm(v))
    a = np.array(rows)
    got = executor.engine.gram(torch.from_numpy(a).cuda(), None, diagonal="computed", precision=scheme).cpu().numpy()
    want = oracle.fidelity_gram_zgemm(a, a, True, "all")
    np.fill_diagonal(want, np.abs(np.einsum("ij,ij->i", a.conj(), a)) ** 2)
    assert np.max(np.abs(got - want)) < INT8_TOL
    assert abs(got[3, 4] - 1.0) < INT8_TOL and abs(got[3, 5] - 1.0) < INT8_TOL


def test_gram_int8_matches_fp64_on_circuit_states(sq, executor):
    import torch

    n, d, N = 13, 4, 700
    X = np.random.default_rng(5).uniform(-0.95, 0.95, (N, d))
    enc = sq.ChebyshevPQC(n, 2)
    p = enc.generate_initial_parameters(d, seed=0)
    psi = _states(sq, executor, enc.get_program(d), p, X)
    t = torch.from_numpy(psi).cuda()
    k64 = executor.engine.gram(t, None, precision="fp64")
    for scheme in ("int8", "crt"):
        k8 = executor.engine.gram(t, None, precision=scheme)
        assert float((k64 - k8).abs().max()) < 2e-11, scheme
        assert torch.equal(k8, k8.T)
    # the digit planes of the CRT scheme equal the NumPy restatement bit for bit
    from tests.helpers import crt_slice

    eng = executor.engine
    _, planes, expo = eng.alloc_plane_block(64, 1 << n, "crt")
    eng.slice_states(t[:64].clone(), planes, expo, 0, "crt")
    want_planes, want_expo = crt_slice(psi[:64])
    assert np.array_equal(expo.cpu().numpy(), want_expo)
    assert np.array_equal(planes.cpu().numpy(), want_planes)


def test_row_maxima_from_the_gate_pass_feed_the_digit_slicing(sq, executor):
    """simulate() tags its output with the per-state statistics the last gate pass reduced (b200sq_simulate_rowmax:
    b200sq_row_stats = {norm2, max_hi, reserved}); gram() / slice_states() of that very tensor then skip their own
    sweep.  Same planes, same exponents as the two-sweep slicing of an untagged copy (digit and CRT form); an
    in-place write invalidates the tag."""
    import torch

    n, d, N = 12, 3, 300
    X = np.random.default_rng(12).uniform(-0.95, 0.95, (N, d))
    enc = sq.ChebyshevPQC(n, 2)
    eng = executor.engine
    for tile_bits in (0, 7):                                         # one pass / several passes
        cp = sq.B200Circuit(enc.get_program(d), tile_bits=tile_bits).compiled(eng, enc.generate_initial_parameters(d, seed=0))
        psi = eng.simulate(cp, X)
        rm = eng._row_max_of(psi)
        assert rm is not None
        want_hi = (psi.view(torch.float64).abs().amax(dim=1).view(torch.int64) >> 32).to(torch.int32)
        assert rm.shape == (N, 4) and torch.equal(rm[:, 2], want_hi)
        norm2 = rm[:, :2].contiguous().view(torch.float64).reshape(N)
        assert float((norm2 - 1.0).abs().max()) < 1e-13
        for scheme in ("int8", "crt"):
            _, p1, e1 = eng.alloc_plane_block(N, 1 << n, scheme)
            _, p2, e2 = eng.alloc_plane_block(N, 1 << n, scheme)
            eng.slice_states(psi, p1, e1, 0, scheme)
            eng.slice_states(psi.clone(), p2, e2, 0, scheme)         # untagged: two sweeps
            assert torch.equal(e1, e2) and torch.equal(p1, p2), scheme
        k_tag = eng.gram(psi, None, precision="int8")
        k_ref = eng.gram(psi.clone(), None, precision="int8")
        assert float((k_tag - k_ref).abs().max()) < 1e-14
    psi.mul_(0.5)                                                    # in-place write: the tag is stale
    assert eng._row_max_of(psi) is None
    k_half = eng.gram(psi, None, diagonal="computed", precision="int8")
    assert abs(float(k_half[3, 3]) - 1.0 / 16.0) < 1e-12


def test_config_c5_size_int8_gram_on_16_points(sq, executor):
    """BASELINE configs[4] state size (20 qubits, K' = 2^21 int8 per row): the INT8 Gram of 16 + 5 points against
    the oracle; the measured maximum error is printed (the scheme's bound grows with K')."""
    import torch

    n, d = 20, 4
    X = np.random.default_rng(5).uniform(-0.95, 0.95, (16, d))
    enc = sq.ChebyshevPQC(n, 2)
    p = enc.generate_initial_parameters(d, seed=0)
    want_sv = oracle.simulate(oracle.chebyshev_pqc(n, 2, d), n, X, oracle.chebyshev_pqc_initial_parameters(n, 2, d, seed=0))
    eng = executor.engine
    psi = eng.simulate(sq.B200Circuit(enc.get_program(d)).compiled(eng, p), X)
    err_sv = float(np.max(np.abs(psi.cpu().numpy() - want_sv)))
    want = np.abs(want_sv.conj() @ want_sv.T) ** 2
    assert err_sv < STATE_TOL
    for scheme in ("int8", "crt"):
        k8 = eng.gram(psi, None, diagonal="computed", precision=scheme).cpu().numpy()
        k8r = eng.gram(psi[:5], psi[3:], precision=scheme).cpu().numpy()
        err, err_r = float(np.max(np.abs(k8 - want))), float(np.max(np.abs(k8r - want[:5, 3:])))
        print(f"C5-size (n=20) {scheme} Gram on 16 points: |K - oracle| = {err:.2e} (rect {err_r:.2e}), states {err_sv:.2e}")
        assert err < INT8_TOL and err_r < INT8_TOL


# ------------------------------------------------------------------------------ expectation values
def test_expvals_fused_and_from_memory(sq, executor):
    rng = np.random.default_rng(11)
    eng = executor.engine
    for n, tile_bits in ((5, 0), (10, 0), (9, 5)):
        prog, ogates = random_circuit(n, 40, 4, 2, rng)
        X = rng.uniform(-0.9, 0.9, (9, 2))
        p = rng.uniform(-1, 1, 4)
        strings = oracle.random_pauli_list(n, 50, seed=3) + oracle.xyz_measurement_strings(n)
        sv = oracle.simulate(ogates, n, X, p)
        want = np.stack([oracle.pauli_expectation(sv, s) for s in strings], axis=1)
        circ = sq.B200Circuit(prog, tile_bits=tile_bits)
        cp = circ.compiled(eng, p)
        got = eng.simulate_expvals(cp, X, strings).cpu().numpy()       # fused (or workspace) path
        assert np.max(np.abs(got - want)) < STATE_TOL
        got2 = eng.expvals(eng.simulate(cp, X), n, strings).cpu().numpy()   # states in HBM
        assert np.max(np.abs(got2 - want)) < STATE_TOL


# ------------------------------------------------------------------------------ reference-shaped API
def test_fidelity_kernel_config_c1_and_duplicate_policies(sq, executor):
    rng = np.random.default_rng(1)
    X = rng.uniform(-0.95, 0.95, (100, 2))
    gates = oracle.chebyshev_pqc(4, 2, 2)
    p = oracle.chebyshev_pqc_initial_parameters(4, 2, 2, seed=0)
    for mode in ("off_diagonal", "all", "none"):
        k = sq.FidelityKernel(sq.ChebyshevPQC(4, 2), executor, evaluate_duplicates=mode).evaluate(X)
        want = oracle.fidelity_kernel(gates, 4, X, None, p, evaluate_duplicates=mode)
        assert k.shape == (100, 100)
        assert np.max(np.abs(k - want)) < GRAM_TOL, mode
    fk = sq.FidelityKernel(sq.ChebyshevPQC(4, 2), executor)
    Y = rng.uniform(-0.95, 0.95, (7, 2))
    assert np.max(np.abs(fk.evaluate(X, Y) - oracle.fidelity_kernel(gates, 4, X, Y, p))) < GRAM_TOL
    assert np.isclose(fk.evaluate_pairwise(X[0], X[1]), fk.evaluate(X)[0, 1], atol=1e-14)
    with pytest.raises(ValueError):
        fk.evaluate(np.zeros((3, 5, 2)))


def test_fidelity_kernel_regularization_mitigation_and_none_policy(sq, executor):
    """fidelity_kernel.py:215-231, 264-330 and regularization.py:5-46 restated with NumPy / SciPy on the oracle's
    kernel matrix; evaluate() and evaluate_device() (what QKRR / QGPR / the kernel losses consume) agree."""
    import scipy.linalg
    import torch

    n, d = 4, 2
    rng = np.random.default_rng(8)
    X = rng.uniform(-0.95, 0.95, (40, d))
    X[7] = X[3]                                   # an exact duplicate pair for the "none" policy
    gates = oracle.chebyshev_pqc(n, 2, d)
    p = oracle.chebyshev_pqc_initial_parameters(n, 2, d, seed=0)
    k0 = oracle.fidelity_kernel(gates, n, X, None, p)
    inv = 2.0 ** -n
    surv = np.sqrt((np.diag(k0) - inv) / (1 - inv))
    want = {"msplit": (k0 - inv * (1 - np.outer(surv, surv))) / np.outer(surv, surv),
            "mmean": (k0 - inv * (1 - surv.mean() ** 2)) / surv.mean() ** 2}
    for mit in ("msplit", "mmean"):
        fk = sq.FidelityKernel(sq.ChebyshevPQC(n, 2), executor, mit_depol_noise=mit)
        assert np.max(np.abs(fk.evaluate(X) - want[mit])) < GRAM_TOL
        with pytest.raises(ValueError):
            fk.evaluate(X, X[:5])
    evals, evecs = scipy.linalg.eig(k0)
    thr = np.real(evecs @ np.diag(evals.clip(min=0)) @ evecs.T)
    for reg, ref in (("thresholding", thr), ("tikhonov", k0)):   # K is PSD: tikhonov leaves it alone
        fk = sq.FidelityKernel(sq.ChebyshevPQC(n, 2), executor, regularization=reg)
        got = fk.evaluate(X)
        assert np.max(np.abs(got - ref)) < 1e-10, reg
        assert np.max(np.abs(fk.evaluate_device(X).cpu().numpy() - got)) == 0.0
        # rectangular matrices are never regularised (:228-231)
        assert np.max(np.abs(fk.evaluate(X, X[:5]) - oracle.fidelity_kernel(gates, n, X, X[:5], p))) < GRAM_TOL
    # an indefinite symmetric matrix through the device regularisers against regularization.py:5-46
    a = rng.normal(size=(30, 30))
    a = 0.5 * (a + a.T)
    fk = sq.FidelityKernel(sq.ChebyshevPQC(n, 2), executor, regularization="thresholding")
    dev = executor.engine.to_device(a, torch.float64)
    ev, evc = scipy.linalg.eig(a)
    assert np.max(np.abs(fk._regularize_device(dev).cpu().numpy() - np.real(evc @ np.diag(ev.clip(min=0)) @ evc.T))) < 1e-10
    fk.set_params(regularization="tikhonov")
    shifted = a - (np.min(np.real(scipy.linalg.eigvals(a))) - 1e-14) * np.identity(30)
    assert np.max(np.abs(fk._regularize_device(dev.clone()).cpu().numpy() - shifted)) < 1e-10
    # "none": identical feature rows give exactly 1.0, also through evaluate_device and for one 1-D sample
    fk = sq.FidelityKernel(sq.ChebyshevPQC(n, 2), executor, evaluate_duplicates="none")
    kd = fk.evaluate_device(X).cpu().numpy()
    assert kd[3, 7] == 1.0 and kd[7, 3] == 1.0 and np.all(np.diag(kd) == 1.0)
    assert np.max(np.abs(kd - oracle.fidelity_kernel(gates, n, X, None, p, evaluate_duplicates="none"))) < GRAM_TOL
    assert fk.evaluate(X[3], X[7])[0, 0] == 1.0
    with pytest.raises(ValueError):
        fk.set_params(evaluate_duplicates="sometimes")
    fk.set_params(evaluate_duplicates="ALL")
    assert fk.get_params()["evaluate_duplicates"] == "all"


def test_closed_form_goldens_through_the_product(sq, executor, goldens):
    g = goldens["fqk_closed_form"]["multi"]
    lay = sq.LayeredEncodingCircuit(2)
    lay.Rx("p")
    lay.Rx("x")
    fk = sq.FidelityKernel(lay, executor, initial_parameters=np.array(g["p"]), evaluate_duplicates="all")
    x0, x1 = g["x"]
    y0, y1 = g["y"]
    expect = (0.25 * np.cos(x0 - y0) + 0.25 * np.cos(x1 - y1) + 0.125 * np.cos(x0 - x1 - y0 + y1)
              + 0.125 * np.cos(x0 + x1 - y0 - y1) + 0.25)
    assert abs(fk.evaluate(np.array([g["x"]]), np.array([g["y"]]))[0, 0] - expect) < 1e-12
    g = goldens["pqk_closed_form"]["single"]
    lay = sq.LayeredEncodingCircuit(1)
    lay.Ry("p")
    lay.Rx("x")
    pqk = sq.ProjectedQuantumKernel(lay, executor, gamma=g["gamma"], initial_parameters=np.array([g["p"]]))
    expect = np.exp(-2 * g["gamma"] * (1 - np.cos(g["x"] - g["y"])) * np.cos(g["p"]) ** 2)
    assert abs(pqk.evaluate(np.array([[g["x"]]]), np.array([[g["y"]]]))[0, 0] - expect) < 1e-12


def test_projected_quantum_kernel_config_c2(sq, executor):
    rng = np.random.default_rng(2)
    X = rng.uniform(-np.pi, np.pi, (1000, 10))
    pqk = sq.ProjectedQuantumKernel(sq.HubregtsenEncodingCircuit(10, 1), executor, gamma=1.0)
    k = pqk.evaluate(X)
    f = pqk.evaluate_qnn(X)
    p = oracle.hubregtsen_initial_parameters(10, 1, seed=0)
    gates = oracle.hubregtsen(10, 1, 10)
    sub = np.arange(0, 1000, 9)
    f_want = oracle.projected_kernel_features(gates, 10, X[sub], p)
    assert f.shape == (1000, 30)
    assert np.max(np.abs(f[sub] - f_want)) < STATE_TOL
    k_want = oracle.gaussian_outer_kernel(f_want, f_want, 1.0)
    assert np.max(np.abs(k[np.ix_(sub, sub)] - k_want)) < 1e-12


def test_qnn_goldens_and_gradients(sq, executor, goldens, regression_data):
    X, y = regression_data
    qnn = sq.LowLevelQNN(sq.ChebyshevRx(2, 1), sq.SummedPaulis(2), executor, 1)
    f = qnn.evaluate(X, np.linspace(0.1, 0.4, 4), np.linspace(0.1, 0.3, 3), "f")["f"]
    assert f.shape == (6,)
    assert np.allclose(f, goldens["qnnr_predict"]["expected"])
    # config C4 shape on a subset of the batch: f, dfdp, dfdop, var
    rng = np.random.default_rng(4)
    Xb = rng.uniform(-0.95, 0.95, (64, 2))
    enc = sq.ChebyshevRx(8, 4)
    p = enc.generate_initial_parameters(2, seed=0)
    pop = np.ones(9)
    qnn = sq.LowLevelQNN(enc, sq.SummedPaulis(8), executor, 2)
    r = qnn.evaluate(Xb, p, pop, "f", "dfdp", "dfdop", "var")
    w = oracle.qnn_evaluate(oracle.chebyshev_rx(8, 4, 2), 8, Xb, p, ("summed_paulis", {}), pop,
                            ("f", "dfdp", "dfdop", "var"))
    assert r["f"].shape == (64,) and r["dfdp"].shape == (64, 64) and r["dfdop"].shape == (64, 9)
    for key in ("f", "dfdp", "dfdop", "var"):
        assert np.max(np.abs(r[key] - w[key])) < GRAM_TOL, key
    # multiple outputs + controlled rotations (four-term rule) + single sample
    enc = sq.ChebyshevPQC(4, 1)
    p = rng.uniform(-1, 1, enc.num_parameters)
    obs = [sq.SummedPaulis(4), sq.SinglePauli(4, 1, "X")]
    oobs = [("summed_paulis", {}), ("single_pauli", {"qubit": 1, "op_str": "X"})]
    pop = rng.uniform(-1, 1, 5)
    qnn = sq.LowLevelQNN(enc, obs, executor, 2)
    r = qnn.evaluate(Xb[:5], p, pop, "f", "dfdp", "dfdop")
    w = oracle.qnn_evaluate(oracle.chebyshev_pqc(4, 1, 2), 4, Xb[:5], p, oobs, pop, ("f", "dfdp", "dfdop"))
    assert r["dfdp"].shape == (5, 2, len(p))
    for key in ("f", "dfdp", "dfdop"):
        assert np.max(np.abs(r[key] - w[key])) < GRAM_TOL, key
    single = qnn.evaluate(Xb[0], p, pop, "f")["f"]
    assert single.shape == (2,) and np.allclose(single, w["f"][0])
    with pytest.raises(NotImplementedError):
        qnn.evaluate(Xb[:2], p, pop, "dfdxdxdp")      # third order in the circuit angles is not offered


def test_optree_gradient_golden_through_the_product(sq, executor, goldens):
    g = goldens["optree_gradient"]
    A = sq.Angle
    prog = sq.Program(2, [
        sq.GateOp("rx", (0,), A(param=0, feature=0)), sq.GateOp("rx", (1,), A(param=1, feature=0)),
        sq.GateOp("ry", (0,), A(param=2)), sq.GateOp("ry", (1,), A(param=3)),
        sq.GateOp("rxx", (0, 1), A(param=0, feature=0))], 4, 1)
    circ = sq.B200Circuit(prog, sq.CustomObservable(2, g["observable"]))
    grad = executor.b200_execute(sq.b200_gradient, circ, param=np.array(g["p"]), x=np.array([[g["x"]]]))
    assert np.allclose(grad[0, 0], g["reference_grad"])


def test_qrc_and_highdim_goldens(sq, executor, goldens, regression_data):
    X, y = regression_data
    g = goldens["qrc_regressor_linear"]
    qrc = sq.QRCRegressor(sq.HubregtsenEncodingCircuit(2, num_features=1, num_layers=1), executor,
                          ml_model="linear")
    qrc.fit(X, y)
    assert np.allclose(qrc.predict(X), g["expected"], atol=g["atol"])
    g = goldens["highdim_all_zero"]
    qnn = sq.LowLevelQNN(sq.HighDimEncodingCircuit(2, num_layers=1), sq.SinglePauli(2, 0, "Z"), executor, 1)
    f = qnn.evaluate(np.array(g["X"]).reshape(-1, 1), [], [], "f")["f"]
    assert np.allclose(f, g["expected"], atol=g["atol"])


# ------------------------------------------------------------------------------ full BASELINE size
def test_config_c3_full_size_properties(sq, executor):
    """n=16, N=4096 (BASELINE configs[2]): size-independent properties of the full Gram plus an
    oracle check of a sub-block."""
    import torch

    n, d, N = 16, 4, 4096
    rng = np.random.default_rng(3)
    X = rng.uniform(-0.95, 0.95, (N, d))
    enc = sq.ChebyshevPQC(n, 2)
    fk = sq.FidelityKernel(enc, executor)
    k = fk.evaluate_device(X)
    psi = fk.last_states
    assert psi.shape == (N, 1 << n)
    norms = (psi.real ** 2 + psi.imag ** 2).sum(dim=1)
    assert float((norms - 1).abs().max()) < 1e-12                 # unitarity
    kc = executor.engine.gram(psi, None, diagonal="computed")
    assert float((torch.diagonal(kc) - 1).abs().max()) < GRAM_TOL   # <psi|psi> = 1
    assert torch.equal(k, k.T)                                      # exact symmetry
    assert float(k.min()) >= 0.0 and float(k.max()) <= 1.0 + 1e-12
    # rectangular kernel of a slice equals the corresponding block of the symmetric one
    blk = executor.engine.gram(psi[1000:1130], psi[37:300], precision="fp64")
    assert float((blk - k[1000:1130, 37:300]).abs().max()) < 2e-11  # the full Gram runs on the INT8 path
    k64 = executor.engine.gram(psi, None, precision="fp64")
    assert float((k64 - k).abs().max()) < 2e-11
    # oracle on a 24-point subset
    sub = rng.choice(N, 24, replace=False)
    p = oracle.chebyshev_pqc_initial_parameters(n, 2, d, seed=0)
    sv = oracle.simulate(oracle.chebyshev_pqc(n, 2, d), n, X[sub], p)
    assert np.max(np.abs(psi[torch.from_numpy(sub).cuda()].cpu().numpy() - sv)) < STATE_TOL
    want = oracle.fidelity_gram_zgemm(sv, sv, True)
    got = k[torch.from_numpy(sub).cuda()][:, torch.from_numpy(sub).cuda()].cpu().numpy()
    assert np.max(np.abs(got - want)) < GRAM_TOL


# ------------------------------------------------------------------------------ next row f1: QKRR on device
def test_qkrr_device_solver_matches_scipy(sq, executor):
    """kernel/qkrr.py:158-169: (K + alpha I) c = y by Cholesky, predict = K(X, X_train) c."""
    import scipy.linalg

    n, d = 6, 2
    rng = np.random.default_rng(11)
    X = rng.uniform(-0.9, 0.9, (60, d))
    Xt = rng.uniform(-0.9, 0.9, (17, d))
    y = np.sin(X[:, 0]) + 0.3 * X[:, 1]
    model = sq.QKRR(sq.FidelityKernel(sq.ChebyshevPQC(n, 2), executor), alpha=1e-4).fit(X, y)
    gates = oracle.chebyshev_pqc(n, 2, d)
    p = oracle.chebyshev_pqc_initial_parameters(n, 2, d, seed=0)
    k = oracle.fidelity_kernel(gates, n, X, None, p) + 1e-4 * np.eye(60)
    c = scipy.linalg.cho_solve((scipy.linalg.cholesky(k, lower=True), True), y)
    assert np.max(np.abs(model.dual_coeff_ - c)) < 1e-6 * np.max(np.abs(c))
    want = oracle.fidelity_kernel(gates, n, Xt, X, p) @ c
    assert np.max(np.abs(model.predict(Xt) - want)) < 1e-7


# ------------------------------------------------------------------------------ next row f3: d/dx (first order)
def test_input_derivative_golden_and_finite_differences(sq, executor, goldens):
    """The reference's d/dx golden (tests/util/optree/test_optree_derivative.py:49-96) and "dfdx" of a QNN
    against central differences; same shifted-variant batch as dfdp, chain-rule weights of the encodings."""
    g = goldens["optree_gradient"]
    A = sq.Angle
    prog = sq.Program(2, [
        sq.GateOp("rx", (0,), A(param=0, feature=0)), sq.GateOp("rx", (1,), A(param=1, feature=0)),
        sq.GateOp("ry", (0,), A(param=2)), sq.GateOp("ry", (1,), A(param=3)),
        sq.GateOp("rxx", (0, 1), A(param=0, feature=0))], 4, 1)
    circ = sq.B200Circuit(prog, sq.CustomObservable(2, g["observable"]))
    dx = executor.b200_execute(sq.b200_gradient, circ, param=np.array(g["p"]), x=np.array([[g["x"]]]), wrt="x")
    assert np.allclose(dx[0, 0], g["reference_dx"])
    n, d = 5, 2
    rng = np.random.default_rng(12)
    X = rng.uniform(-0.8, 0.8, (6, d))
    enc = sq.ChebyshevPQC(n, 2)
    p = rng.uniform(-1, 1, enc.num_parameters)
    pop = rng.uniform(-1, 1, n + 1)
    qnn = sq.LowLevelQNN(enc, sq.SummedPaulis(n), executor, d)
    got = qnn.evaluate(X, p, pop, "dfdx")["dfdx"]
    assert got.shape == (6, d)
    gates = oracle.chebyshev_pqc(n, 2, d)
    h = 1e-6
    for j in range(d):
        e = np.zeros(d)
        e[j] = h
        fp = oracle.qnn_evaluate(gates, n, X + e, p, ("summed_paulis", {}), pop, ("f",))["f"]
        fm = oracle.qnn_evaluate(gates, n, X - e, p, ("summed_paulis", {}), pop, ("f",))["f"]
        assert np.max(np.abs(got[:, j] - (fp - fm) / (2 * h))) < 1e-7


# ------------------------------------------------------------------------------ QNN training (row a14)
def test_qnn_regressor_reference_golden_and_fit(sq, executor, goldens, regression_data):
    """tests/qnn/test_qnnr.py:20-43,179-198: the reference's predict golden through QNNRegressor; then the same
    model is fitted with SLSQP(maxiter=2) like test_fit (:75-95) and the loss must go down; SquaredLoss gradient
    against central differences of its value (squared_loss.py:35-61,92-143)."""
    X, y = regression_data
    rng = np.random.default_rng(seed=30)
    pqc, op = sq.ChebyshevRx(2, 1), sq.SummedPaulis(2)
    make = lambda loss, opt: sq.QNNRegressor(pqc, op, executor, loss, opt, rng.random(pqc.num_parameters),
                                             rng.random(op.num_parameters))
    reg = make(sq.MeanSquaredError(), sq.SLSQP(options={"maxiter": 2}))
    with pytest.raises(RuntimeError, match="The model is not fitted."):
        reg.predict(X)
    reg._initialize_lowlevel_qnn(1)
    reg._param, reg._param_op, reg._is_fitted = np.linspace(0.1, 0.4, 4), np.linspace(0.1, 0.3, 3), True
    pred = reg.predict(X)
    assert pred.shape == y.shape
    assert np.allclose(pred, goldens["qnnr_predict"]["expected"], atol=1e-8)

    reg = make(sq.MeanSquaredError(), sq.SLSQP(options={"maxiter": 2}))
    reg.fit(X, y)
    assert reg._is_fitted
    assert not np.allclose(reg.param, reg.param_ini) and not np.allclose(reg.param_op, reg.param_op_ini)
    qnn = reg._qnn
    mse = lambda p, po: float(np.mean((qnn.evaluate(X, p, po, "f")["f"] - y) ** 2))
    assert mse(reg.param, reg.param_op) < mse(reg.param_ini, reg.param_op_ini)

    loss = sq.SquaredLoss()
    loss.set_opt_param_op(True)
    p0, po0 = np.linspace(0.3, 0.9, 4), np.array([0.2, -0.4, 0.7])
    w = np.linspace(0.5, 1.5, len(y))
    vals = qnn.evaluate(X, p0, po0, *loss.gradient_args_tuple)
    g_p, g_op = loss.gradient(vals, ground_truth=y, weights=w)
    theta0 = np.concatenate([p0, po0])
    num = np.zeros_like(theta0)
    for k in range(len(theta0)):
        e = np.zeros_like(theta0)
        e[k] = 1e-5
        f = lambda t: loss.value(qnn.evaluate(X, t[:4], t[4:], "f"), ground_truth=y, weights=w)
        num[k] = (f(theta0 + e) - f(theta0 - e)) / 2e-5
    assert np.max(np.abs(np.concatenate([g_p, g_op]) - num)) < 1e-4 * max(1.0, np.max(np.abs(num)))


def test_qnn_regressor_config_c4_shape_trains(sq, executor):
    """BASELINE configs[3]: ChebyshevRx(8 qubits, 4 layers) + SummedPaulis(8), batch 1024, parameter-shift gradients
    (64 circuit parameters -> 129 shifted variants x 1024 samples per gradient): a few Adam steps on a smooth
    target lower the squared loss, and one loss gradient equals the oracle's generator-route gradient."""
    rng = np.random.default_rng(4)
    X = rng.uniform(-0.95, 0.95, (1024, 2))
    y = np.sin(2.0 * X[:, 0]) * np.cos(X[:, 1])
    enc, op = sq.ChebyshevRx(8, 4, closed=False), sq.SummedPaulis(8)
    reg = sq.QNNRegressor(enc, op, executor, sq.SquaredLoss(), sq.Adam({"lr": 0.05, "maxiter": 6}),
                          param_op_ini=np.ones(9) * 0.1)
    reg._initialize_lowlevel_qnn(2)
    p0, po0 = reg.param.copy(), reg.param_op.copy()
    loss0 = float(np.sum((reg._qnn.evaluate(X, p0, po0, "f")["f"] - y) ** 2))
    reg.partial_fit(X, y)
    loss1 = float(np.sum((reg.predict(X) - y) ** 2))
    assert loss1 < loss0
    vals = reg._qnn.evaluate(X[:64], p0, po0, "f", "dfdp", "dfdop")
    want = oracle.qnn_evaluate(oracle.chebyshev_rx(8, 4, 2, closed=False), 8, X[:64], p0, ("summed_paulis", {}), po0,
                               ("f", "dfdp", "dfdop"))
    for key in ("f", "dfdp", "dfdop"):
        assert np.max(np.abs(vals[key] - want[key])) < GRAM_TOL, key


# ------------------------------------------------------------------------------ second-order derivatives (row f3)
def test_second_order_derivatives_against_finite_differences(sq, executor):
    """"dfdxdx", "dfdpdx", "dfdxdp", "dfdpdp", "laplace", "dfdopdx", "dfdopdxdx", "laplace_dop", "dvardx"
    (evaluation_classes.py:105-164) from doubly shifted circuits, against central differences of the first-order /
    value results, through the non-linear arccos encoding; axes follow the key ("dfdpdx" -> (N, P, d))."""
    d = 2
    enc = sq.ChebyshevPQC(3, 1)
    qnn = sq.LowLevelQNN(enc, [sq.SummedPaulis(3), sq.SinglePauli(3, 1, "X")], executor, d, caching=False)
    rng = np.random.default_rng(0)
    X = rng.uniform(-0.8, 0.8, (5, d))
    P = enc.num_parameters
    p, po = rng.uniform(-1, 1, P), rng.uniform(0.2, 1, 4)
    r = qnn.evaluate(X, p, po, "f", "dfdx", "dfdp", "dfdxdx", "dfdpdx", "dfdxdp", "dfdpdp", "laplace", "dfdopdx",
                     "dfdopdxdx", "laplace_dop", "dfdopdp", "dfdpdop", "dfdopdop", "dvardx", "dfccdxdx")
    assert r["dfdxdx"].shape == (5, 2, d, d) and r["dfdpdx"].shape == (5, 2, P, d) and r["dfdxdp"].shape == (5, 2, d, P)
    assert r["dfdopdxdx"].shape == (5, 2, 4, d, d) and r["laplace_dop"].shape == (5, 2, 4) and r["laplace"].shape == (5, 2)
    h = 1e-4
    first = lambda key, X_, p_: qnn.evaluate(X_, p_, po, key)[key]
    hx = np.stack([(first("dfdx", X + h * np.eye(d)[j], p) - first("dfdx", X - h * np.eye(d)[j], p)) / (2 * h)
                   for j in range(d)], axis=-1)                                   # (N, obs, d, d): [i][j] = d/dx_j dfdx_i
    assert np.max(np.abs(r["dfdxdx"] - hx)) < 1e-6
    assert np.max(np.abs(r["laplace"] - np.trace(hx, axis1=-2, axis2=-1))) < 1e-6
    px = np.stack([(first("dfdp", X + h * np.eye(d)[j], p) - first("dfdp", X - h * np.eye(d)[j], p)) / (2 * h)
                   for j in range(d)], axis=-1)                                   # (N, obs, P, d)
    assert np.max(np.abs(r["dfdpdx"] - px)) < 1e-6
    assert np.max(np.abs(r["dfdxdp"] - np.swapaxes(px, -1, -2))) < 1e-6
    pp = np.stack([(first("dfdp", X, p + h * np.eye(P)[j]) - first("dfdp", X, p - h * np.eye(P)[j])) / (2 * h)
                   for j in range(P)], axis=-1)
    assert np.max(np.abs(r["dfdpdp"] - pp)) < 1e-6
    # observable-parameter derivatives are exact differences (the observables are linear in param_obs)
    for key, base in (("dfdopdx", "dfdx"), ("dfdopdp", "dfdp"), ("dfdopdxdx", "dfdxdx")):
        g = np.stack([(qnn.evaluate(X, p, po + np.eye(4)[j], base)[base] - qnn.evaluate(X, p, po - np.eye(4)[j], base)[base]) / 2
                      for j in range(4)], axis=2)
        assert np.max(np.abs(r[key] - g)) < 1e-12, key
    assert np.max(np.abs(r["dfdpdop"] - np.swapaxes(r["dfdopdp"], -1, -2))) == 0.0
    assert np.all(r["dfdopdop"] == 0.0) and r["dfdopdop"].shape == (5, 2, 4, 4)
    var = lambda X_: qnn.evaluate(X_, p, po, "var")["var"]
    dv = np.stack([(var(X + h * np.eye(d)[j]) - var(X - h * np.eye(d)[j])) / (2 * h) for j in range(d)], axis=-1)
    assert np.max(np.abs(r["dvardx"] - dv)) < 1e-6
    fcc = lambda X_: qnn.evaluate(X_, p, po, "dfccdx")["dfccdx"]
    hcc = np.stack([(fcc(X + h * np.eye(d)[j]) - fcc(X - h * np.eye(d)[j])) / (2 * h) for j in range(d)], axis=-1)
    assert np.max(np.abs(r["dfccdxdx"] - hcc)) < 1e-5
    with pytest.raises(NotImplementedError):
        qnn.evaluate(X, p, po, "dfdxdxdp")
