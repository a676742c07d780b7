"""The PennyLane device of the b200 backend against a stub of PennyLane's tape interface, with the
host emulator as the engine: a broadcast tape is ONE engine call, the 2P shifted tapes of a parameter-shift batch
are ONE engine call, states come back in PennyLane's wire order, expectation values of Pauli arithmetic match the
oracle.  (With PennyLane installed the same class subclasses ``pennylane.devices.Device`` and plugs into
``squlearn.Executor(device)``: src/squlearn/util/executor.py:555-573.)"""

import numpy as np
import pytest

import oracle
from squlearn_b200.pennylane_device import B200QubitDevice
from tests import pennylane_stub as qml
from tests.emu_engine import EmuExecutor

_NAMES = {"h": "Hadamard", "x": "PauliX", "y": "PauliY", "z": "PauliZ", "s": "S", "t": "T", "sx": "SX",
          "sdg": "Adjoint(S)", "tdg": "Adjoint(T)", "rx": "RX", "ry": "RY", "rz": "RZ", "p": "PhaseShift",
          "cx": "CNOT", "cy": "CY", "cz": "CZ", "ch": "CH", "cs": "C(S)", "csx": "C(SX)", "cp": "ControlledPhaseShift",
          "crx": "CRX", "cry": "CRY", "crz": "CRZ", "swap": "SWAP", "iswap": "ISWAP", "ecr": "ECR", "ccx": "Toffoli",
          "cswap": "CSWAP", "id": "Identity"}
_ROT = {"rxx": "XX", "ryy": "YY", "rzz": "ZZ", "rzx": "ZX"}


def _tape(gates, x, p, measurements, shift=None):
    """PennyLane-style tape of an oracle gate list: data-dependent angles as (N,) broadcast arrays, the others as
    scalars — what a QNode called with broadcast arguments records (pennylane_circuit.py:570-579)."""
    ops = []
    for k, g in enumerate(gates):
        params = []
        if g.angles:
            a0 = g.angles[0]
            theta = np.asarray(a0(x, p) if callable(a0) else a0, dtype=float)
            theta = np.broadcast_to(theta, (x.shape[0],)).copy()
            if shift is not None and shift[0] == k:
                theta = theta + shift[1]
            params = [theta if np.ptp(theta) > 0 else float(theta[0])]
            if g.name == "u":
                params += [g.angles[1], g.angles[2]]
        if g.name in _ROT:
            ops.append(qml.Op("PauliRot", g.qubits, params, pauli_word=_ROT[g.name]))
        elif g.name == "u":
            ops.append(qml.Op("U3", g.qubits, params))
        else:
            ops.append(qml.Op(_NAMES[g.name], g.qubits, params))
    return qml.QuantumScript(ops, measurements)


def test_broadcast_tape_is_one_engine_call_and_states_are_in_pennylane_order():
    n, d, N = 5, 2, 9
    X = np.random.default_rng(0).uniform(-0.9, 0.9, (N, d))
    gates = oracle.chebyshev_pqc(n, 2, d)
    p = oracle.chebyshev_pqc_initial_parameters(n, 2, d, seed=0)
    dev = B200QubitDevice(wires=n, executor=EmuExecutor())
    got = dev.execute(_tape(gates, X, p, [qml.StateMP()]))
    want = oracle.to_pennylane_order(oracle.simulate(gates, n, X, p), n)
    assert dev.num_executions == 1
    assert got.shape == (N, 2 ** n)
    assert np.max(np.abs(got - want)) < 1e-12
    # the plan is cached by gate structure: new data, same compiled circuit
    X2 = np.random.default_rng(1).uniform(-0.9, 0.9, (4, d))
    got2 = dev.execute([_tape(gates, X2, p, [qml.StateMP()])])[0]
    assert len(dev._cache) == 1
    assert np.max(np.abs(got2 - oracle.to_pennylane_order(oracle.simulate(gates, n, X2, p), n))) < 1e-12


def test_expectation_values_of_pauli_arithmetic():
    n, d, N = 4, 2, 6
    X = np.random.default_rng(2).uniform(-0.9, 0.9, (N, d))
    gates = oracle.hubregtsen(n, 1, d)
    p = oracle.hubregtsen_initial_parameters(n, 1, seed=0)
    obs = [qml.PauliZ(0), qml.Prod(qml.PauliX(1), qml.PauliY(3)),
           qml.Sum(qml.SProd(0.5, qml.PauliZ(2)), qml.SProd(-1.5, qml.Prod(qml.PauliZ(0), qml.PauliZ(1))), qml.Identity(0))]
    dev = B200QubitDevice(wires=n, executor=EmuExecutor())
    got = dev.execute(_tape(gates, X, p, [qml.ExpectationMP(o) for o in obs]))
    assert dev.num_executions == 1 and len(got) == 3
    psi = oracle.simulate(gates, n, X, p)
    # Qiskit-style strings: rightmost character = qubit 0 = PennyLane wire 0 of the reference's circuits
    ev = lambda s: oracle.pauli_expectation(psi, s)
    want = [ev("IIIZ"), ev("YIXI"), 0.5 * ev("IZII") - 1.5 * ev("IIZZ") + 1.0]
    for g, w in zip(got, want):
        assert g.shape == (N,) and np.max(np.abs(g - w)) < 1e-12


def test_shifted_tapes_of_a_gradient_batch_are_one_engine_call():
    """What PennyLane's parameter-shift transform hands a device without its own derivatives: 2P tapes that differ
    only in one angle.  The device stacks them along the batch axis; the recombined gradient is the oracle's."""
    n, d, N = 3, 1, 5
    X = np.random.default_rng(3).uniform(-0.9, 0.9, (N, d))
    gates = oracle.chebyshev_rx(n, 1, d)
    p = oracle.chebyshev_rx_initial_parameters(n, 1, d, seed=0)
    z_sum = qml.Sum(*[qml.PauliZ(w) for w in range(n)])
    variants = oracle.parameter_shift_variants(gates, X, p)       # [(gate index, shift, weight (N,), parameter)]
    shifted = [_tape(gates, X, p, [qml.ExpectationMP(z_sum)], shift=(k, s)) for k, s, _, _ in variants]
    weights, pidx = [v[2] for v in variants], [v[3] for v in variants]
    dev = B200QubitDevice(wires=n, executor=EmuExecutor())
    res = dev.execute(shifted)
    assert dev.num_executions == 1 and len(res) == len(shifted)
    grad = np.zeros((N, len(p)))
    for r, w, j in zip(res, weights, pidx):
        grad[:, j] += w * r
    want = oracle.qnn_evaluate(gates, n, X, p, ("summed_paulis", {"include_identity": False}), np.ones(n), ("dfdp",))["dfdp"]
    assert np.max(np.abs(grad - want)) < 1e-12


def test_unsupported_operations_and_shots_are_rejected():
    dev = B200QubitDevice(wires=2, executor=EmuExecutor())
    with pytest.raises(NotImplementedError):
        dev.execute(qml.QuantumScript([qml.Op("QubitUnitary", [0], [np.eye(2)])], [qml.StateMP()]))
    with pytest.raises(ValueError):
        B200QubitDevice(wires=2, shots=100)
