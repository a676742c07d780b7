"""CPU tests of the host logic: circuit builders, the pass/round/micro-op planner and the per-thread
micro-op interpreter (the same C++ the CUDA kernel runs, driven by tests/hostemu) against the oracle."""

import numpy as np
import pytest

import oracle
from squlearn_b200 import (ChebyshevPQC, ChebyshevRx, HighDimEncodingCircuit,
                           HubregtsenEncodingCircuit, LayeredEncodingCircuit)
from squlearn_b200.observables import pauli_masks
from tests import hostemu
from tests.helpers import random_circuit, program_to_oracle

TOL = 1e-12


@pytest.mark.parametrize("n,layers,d,tile_bits", [(4, 2, 2, 0), (6, 2, 3, 3), (7, 1, 2, 4),
                                                   (9, 2, 4, 5), (3, 1, 1, 0), (2, 1, 1, 0)])
def test_chebyshev_pqc_matches_oracle(n, layers, d, tile_bits):
    rng = np.random.default_rng(n * 10 + layers)
    X = rng.uniform(-0.95, 0.95, (5, d))
    enc = ChebyshevPQC(n, layers)
    p = enc.generate_initial_parameters(d, seed=0)
    assert np.array_equal(p, oracle.chebyshev_pqc_initial_parameters(n, layers, d, seed=0))
    want = oracle.simulate(oracle.chebyshev_pqc(n, layers, d), n, X, p)
    got, passes = hostemu.simulate(enc.get_program(d), p, X, tile_bits=tile_bits)
    assert np.max(np.abs(got - want)) < TOL
    if tile_bits and tile_bits < n:
        assert passes >= 2


@pytest.mark.parametrize("entangling_gate,closed,nonlin", [("rzz", True, "arctan"), ("crz", False, "arccos")])
def test_chebyshev_pqc_options(entangling_gate, closed, nonlin):
    n, layers, d = 4, 2, 2
    rng = np.random.default_rng(5)
    X = rng.uniform(-0.9, 0.9, (4, d))
    enc = ChebyshevPQC(n, layers, closed=closed, entangling_gate=entangling_gate, nonlinearity=nonlin)
    p = rng.uniform(-1, 1, enc.num_parameters)
    assert enc.num_parameters == oracle.chebyshev_pqc_num_parameters(n, layers, closed)
    want = oracle.simulate(oracle.chebyshev_pqc(n, layers, d, closed, entangling_gate, nonlin), n, X, p)
    got, _ = hostemu.simulate(enc.get_program(d), p, X, tile_bits=3)
    assert np.max(np.abs(got - want)) < TOL


@pytest.mark.parametrize("n,layers,d,final", [(10, 1, 10, False), (4, 2, 6, True), (2, 1, 1, False), (5, 1, 3, False)])
def test_hubregtsen_matches_oracle(n, layers, d, final):
    if layers * n < d:
        pytest.skip("more features than encoding slots")
    rng = np.random.default_rng(n)
    X = rng.uniform(-np.pi, np.pi, (3, d))
    enc = HubregtsenEncodingCircuit(n, layers, final_encoding=final)
    p = enc.generate_initial_parameters(d, seed=0)
    assert np.array_equal(p, oracle.hubregtsen_initial_parameters(n, layers, seed=0))
    want = oracle.simulate(oracle.hubregtsen(n, layers, d, final_encoding=final), n, X, p)
    got, _ = hostemu.simulate(enc.get_program(d), p, X, tile_bits=6)
    assert np.max(np.abs(got - want)) < TOL


@pytest.mark.parametrize("n,layers,d,closed", [(8, 4, 2, False), (2, 1, 1, False), (5, 2, 2, True)])
def test_chebyshev_rx_matches_oracle(n, layers, d, closed):
    rng = np.random.default_rng(n)
    X = rng.uniform(-0.95, 0.95, (4, d))
    enc = ChebyshevRx(n, layers, closed=closed)
    p = enc.generate_initial_parameters(d, seed=0)
    assert np.array_equal(p, oracle.chebyshev_rx_initial_parameters(n, layers, d, seed=0))
    want = oracle.simulate(oracle.chebyshev_rx(n, layers, d, closed=closed), n, X, p)
    got, _ = hostemu.simulate(enc.get_program(d), p, X, tile_bits=5)
    assert np.max(np.abs(got - want)) < TOL


def test_highdim_and_layered_match_oracle():
    X = np.arange(-0.2, 0.3, 0.1).reshape(-1, 1)
    enc = HighDimEncodingCircuit(2, num_layers=1)
    got, _ = hostemu.simulate(enc.get_program(1), None, X)
    want = oracle.simulate(oracle.highdim(2, 1, num_layers=1), 2, X)
    assert np.max(np.abs(got - want)) < TOL
    lay = LayeredEncodingCircuit(2)
    lay.Ry("p")
    lay.Rx("x")
    p = np.array([-0.63, -0.63])
    X2 = np.array([[0.79, 0.9], [-0.31, -1.31]])
    got, _ = hostemu.simulate(lay.get_program(2), p, X2)
    want = oracle.simulate(oracle.layered_ry_rx(2, 2), 2, X2, p)
    assert np.max(np.abs(got - want)) < TOL


@pytest.mark.parametrize("seed", range(12))
def test_random_circuits_all_gates(seed):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(1, 8))
    tile_bits = int(rng.integers(max(1, min(n, 3)), n + 1)) if n >= 3 else 0
    num_params, d = 5, 3
    prog, ogates = random_circuit(n, 40, num_params, d, rng)
    X = rng.uniform(-0.9, 0.9, (3, d))
    p = rng.uniform(-2, 2, num_params)
    want = oracle.simulate(ogates, n, X, p)
    try:
        got, _ = hostemu.simulate(prog, p, X, tile_bits=tile_bits, team_size=int(rng.choice([1, 8, 32])))
    except RuntimeError as e:
        if "more target qubits" in str(e):
            pytest.skip("tile too small for a multi-target gate")
        raise
    assert np.max(np.abs(got - want)) < TOL


DIAG_HEAVY = ["rz", "p", "cz", "crz", "cp", "rzz", "s", "t", "z", "cs", "rx", "ry", "h", "cx", "rzz", "crz"]


@pytest.mark.parametrize("seed", range(16))
def test_random_diagonal_heavy_circuits(seed):
    """Phase-table steps (runs of commuting diagonal gates, one or two bit groups), deferred phases,
    layer steps and the compact active-qubit layouts between passes."""
    rng = np.random.default_rng(500 + seed)
    n = int(rng.integers(3, 11))
    tile_bits = int(rng.integers(3, n + 1))
    num_params, d = 6, 3
    prog, ogates = random_circuit(n, 60, num_params, d, rng, kinds=DIAG_HEAVY)
    X = rng.uniform(-0.9, 0.9, (2, d))
    p = rng.uniform(-2, 2, num_params)
    want = oracle.simulate(ogates, n, X, p)
    got, _ = hostemu.simulate(prog, p, X, tile_bits=tile_bits, team_size=int(rng.choice([1, 4, 32, 256])))
    assert np.max(np.abs(got - want)) < TOL


@pytest.mark.parametrize("n,tile_bits", [(10, 4), (11, 5), (12, 8), (13, 9), (12, 12)])
def test_layered_circuits_multi_pass_layouts(n, tile_bits):
    """ChebyshevPQC / Hubregtsen at sizes where several passes with growing active sets are needed."""
    rng = np.random.default_rng(n)
    X = rng.uniform(-0.95, 0.95, (2, 3))
    for enc, og in ((ChebyshevPQC(n, 2), oracle.chebyshev_pqc(n, 2, 3)),
                    (HubregtsenEncodingCircuit(n, 1), oracle.hubregtsen(n, 1, 3))):
        p = enc.generate_initial_parameters(3, seed=0)
        want = oracle.simulate(og, n, X, p)
        got, passes = hostemu.simulate(enc.get_program(3), p, X, tile_bits=tile_bits, team_size=32)
        assert np.max(np.abs(got - want)) < TOL
        if tile_bits < n:
            assert passes >= 2


def test_untouched_qubits_stay_zero():
    """Qubits no gate acts on never become active: the last pass must still write the full layout."""
    from squlearn_b200.program import Angle, GateOp, Program
    from oracle.statevector import Gate
    n = 7
    prog = Program(n, [GateOp("h", (1,)), GateOp("cx", (1, 5)), GateOp("rz", (6,), Angle(scale=0.3))], 0, 1)
    og = [Gate("h", (1,), ()), Gate("cx", (1, 5), ()), Gate("rz", (6,), (lambda x, p: 0.3 + 0 * x[:, 0],))]
    X = np.zeros((2, 1))
    for tb in (3, 4, 7):
        got, _ = hostemu.simulate(prog, None, X, tile_bits=tb)
        assert np.max(np.abs(got - oracle.simulate(og, n, X))) < TOL


def test_general_angle_callable_goes_through_table():
    from squlearn_b200.program import Angle, GateOp, Program

    prog = Program(2, [GateOp("h", (0,)), GateOp("ry", (1,), Angle(table=lambda x, p: np.sin(x[:, 0]) * p[0] ** 2)),
                       GateOp("cx", (0, 1))], 1, 1)
    X = np.linspace(-1, 1, 5).reshape(-1, 1)
    p = np.array([0.7])
    from oracle.statevector import Gate
    og = [Gate("h", (0,), ()), Gate("ry", (1,), (lambda x, p: np.sin(x[:, 0]) * p[0] ** 2,)), Gate("cx", (0, 1), ())]
    got, _ = hostemu.simulate(prog, p, X)
    assert np.max(np.abs(got - oracle.simulate(og, 2, X, p))) < TOL


def test_shift_variants_match_oracle_spec():
    n, d = 4, 2
    rng = np.random.default_rng(9)
    X = rng.uniform(-0.9, 0.9, (3, d))
    enc = ChebyshevPQC(n, 1)
    prog = enc.get_program(d)
    p = rng.uniform(-1, 1, enc.num_parameters)
    gate_idx, shift, weight, pidx = prog.shift_variants(X)
    ref = oracle.parameter_shift_variants(oracle.chebyshev_pqc(n, 1, d), X, p)
    assert len(ref) == len(gate_idx)
    for (k, s, w, j), gk, gs, gw, gj in zip(ref, gate_idx, shift, weight, pidx):
        assert k == gk and j == gj and np.isclose(s, gs) and np.allclose(w, gw)
    # the shifted states themselves
    got, _ = hostemu.simulate(prog, p, X, variants=(gate_idx[:6], shift[:6]))
    from oracle.qnn import _shifted
    for v in range(6):
        want = oracle.simulate(_shifted(oracle.chebyshev_pqc(n, 1, d), int(gate_idx[v]), shift[v]), n, X, p)
        assert np.max(np.abs(got[v] - want)) < TOL


def test_expval_helper_matches_oracle():
    rng = np.random.default_rng(2)
    n = 5
    prog, ogates = random_circuit(n, 30, 4, 2, rng)
    X = rng.uniform(-0.9, 0.9, (4, 2))
    p = rng.uniform(-1, 1, 4)
    sv = oracle.simulate(ogates, n, X, p)
    strings = oracle.random_pauli_list(n, 40, seed=1) + oracle.xyz_measurement_strings(n)
    masks = [pauli_masks(s) for s in strings]
    got = hostemu.expvals(sv, n, [m[0] for m in masks], [m[1] for m in masks])
    want = np.stack([oracle.pauli_expectation(sv, s) for s in strings], axis=1)
    assert np.max(np.abs(got - want)) < 1e-13


def test_config_circuit_pass_counts():
    """The fused plan needs far fewer sweeps over the state than there are gates."""
    import ctypes
    from squlearn_b200 import _lib
    lib = _lib.load()
    for n, expect_max in ((16, 2), (20, 3), (12, 1), (4, 1)):
        enc = ChebyshevPQC(n, 2)
        prog = enc.get_program(4 if n >= 4 else 2)
        h = ctypes.c_void_p()
        p = enc.generate_initial_parameters(4, seed=0)
        _lib.check(lib.b200sq_program_create(n, len(prog), prog.c_gates(p), 0, ctypes.byref(h)))
        passes = lib.b200sq_program_num_passes(h)
        lib.b200sq_program_destroy(h)
        assert 1 <= passes <= expect_max, (n, passes)


def test_runtime_specialised_pass_kernels_compile():
    """The straight-line pass kernels jit.cu generates for the config circuits are valid sm_100a CUDA
    (NVRTC compiles to a cubin without a GPU)."""
    import ctypes
    from squlearn_b200 import _lib
    lib = _lib.load()
    for enc, d in ((ChebyshevPQC(13, 2), 4), (HubregtsenEncodingCircuit(10, 1), 10), (ChebyshevRx(8, 4), 2)):
        prog = enc.get_program(d)
        p = enc.generate_initial_parameters(d, seed=0)
        h = ctypes.c_void_p()
        _lib.check(lib.b200sq_program_create(prog.num_qubits, len(prog), prog.c_gates(p), 0, ctypes.byref(h)))
        try:
            for ip in range(lib.b200sq_program_num_passes(h)):
                log = ctypes.create_string_buffer(4000)
                size = lib.b200sq_program_jit_compile(h, ip, log, 4000)
                if size == -1 and b"libnvrtc" in log.value:
                    pytest.skip("NVRTC not available")
                assert size > 0, log.value.decode()
        finally:
            lib.b200sq_program_destroy(h)


def test_long_two_qubit_circuits_split_into_passes_that_fit_shared_memory():
    """Hundreds of 4x4 gates: the planner closes a pass before its matrix storage outgrows shared memory."""
    import ctypes
    from squlearn_b200 import _lib
    rng = np.random.default_rng(42)
    n = 7
    prog, ogates = random_circuit(n, 700, 4, 2, rng, kinds=["swap", "iswap", "rxx", "ryy", "rzx", "ecr", "rzz", "crz", "h"])
    X = rng.uniform(-0.9, 0.9, (2, 2))
    p = rng.uniform(-1, 1, 4)
    got, passes = hostemu.simulate(prog, p, X, tile_bits=7, team_size=32)
    assert np.max(np.abs(got - oracle.simulate(ogates, n, X, p))) < 1e-11
    assert passes >= 2
    lib = _lib.load()
    h = ctypes.c_void_p()
    _lib.check(lib.b200sq_program_create(n, len(prog), prog.c_gates(p), 7, ctypes.byref(h)))
    try:
        cnt = ctypes.c_int(0)
        _lib.check(lib.b200sq_program_dump(h, None, 0, ctypes.byref(cnt)))
        buf = (ctypes.c_int32 * cnt.value)()
        _lib.check(lib.b200sq_program_dump(h, buf, cnt.value, ctypes.byref(cnt)))
        w = list(buf)
        pos = 5
        for _ in range(w[2]):
            assert w[pos + 8] <= 4096          # mat_elems of the pass
            pos += 11 + w[pos]
    finally:
        lib.b200sq_program_destroy(h)


@pytest.mark.parametrize("seed", range(6))
def test_lifted_rotations_option(seed, monkeypatch):
    """B200SQ_ROT_LIFT=1 (experimental, default off): rx / ry layer steps as in-place lifted rotations
    (3 FMAs per 2-D rotation), the signs of the angle reduction carried by the product table.  Large angles
    exercise the sign flips."""
    monkeypatch.setenv("B200SQ_ROT_LIFT", "1")
    rng = np.random.default_rng(900 + seed)
    n = int(rng.integers(3, 10))
    tile_bits = int(rng.integers(3, n + 1))
    prog, ogates = random_circuit(n, 50, 5, 3, rng, kinds=["rx", "ry", "rx", "ry", "rz", "crz", "cz", "h", "cx", "rzz", "p"])
    X = rng.uniform(-0.9, 0.9, (2, 3))
    p = rng.uniform(-9, 9, 5)
    got, _ = hostemu.simulate(prog, p, X, tile_bits=tile_bits, team_size=32)
    assert np.max(np.abs(got - oracle.simulate(ogates, n, X, p))) < TOL
    enc = ChebyshevPQC(9, 2)
    pp = enc.generate_initial_parameters(3, seed=0)
    got, _ = hostemu.simulate(enc.get_program(3), pp, X, tile_bits=5)
    assert np.max(np.abs(got - oracle.simulate(oracle.chebyshev_pqc(9, 2, 3), 9, X, pp))) < TOL
