"""A PennyLane device backed by the B200 engine.

With PennyLane installed, ``Executor(B200QubitDevice(wires=n))`` makes the reference's own PennyLane path
(src/squlearn/util/executor.py:555-573 accepts any ``pennylane.devices.Device``; :908-939, :941-1000 build QNodes /
tapes on it) run on this engine without touching sQUlearn: ``pennylane_execute`` calls the QNode with broadcast
parameters (one (N,) array per data-dependent angle, pennylane_circuit.py:570-579), PennyLane hands the device one
tape whose gate parameters carry the batch, and ``execute`` maps

* a broadcast tape               -> ONE ``b200sq_simulate`` / ``b200sq_simulate_expvals`` call over the N samples
                                    (per-sample angles travel as rows of the device angle table, B200SQ_FN_TABLE);
* the 2P shifted tapes of a parameter-shift gradient transform (same gates, other angles), which arrive in one
  ``execute`` call -> stacked along the batch axis into ONE engine call as well.

PennyLane is not in this image: the module imports it lazily and otherwise only duck-types tapes
(``operations`` with ``name`` / ``wires`` / ``parameters`` / ``hyperparameters``, ``measurements`` with ``obs``), so the
CPU suite drives it with a stub (tests/pennylane_stub.py).  Wire w of an n-wire register is engine qubit n-1-w, which
makes the engine's little-endian amplitude index equal to PennyLane's (wire 0 = most significant bit).
"""

from typing import List, Optional, Tuple

import numpy as np

from .executor import B200Circuit, Executor
from .program import Angle, GateOp, Program

# PennyLane operation name -> gate of the engine's table (the names the reference emits:
# src/squlearn/util/pennylane/pennylane_gates.py:56-93)
_OPS = {
    "Identity": "id", "Hadamard": "h", "PauliX": "x", "PauliY": "y", "PauliZ": "z", "S": "s", "T": "t", "SX": "sx",
    "Adjoint(S)": "sdg", "Adjoint(T)": "tdg", "RX": "rx", "RY": "ry", "RZ": "rz", "PhaseShift": "p", "U3": "u",
    "CNOT": "cx", "CY": "cy", "CZ": "cz", "CH": "ch", "C(S)": "cs", "C(SX)": "csx", "ControlledPhaseShift": "cp",
    "CRX": "crx", "CRY": "cry", "CRZ": "crz", "SWAP": "swap", "ISWAP": "iswap", "ECR": "ecr",
    "IsingXX": "rxx", "IsingYY": "ryy", "IsingZZ": "rzz", "Toffoli": "ccx", "CSWAP": "cswap",
}
_PAULI_ROT = {"XX": "rxx", "YY": "ryy", "ZZ": "rzz", "ZX": "rzx"}
_PAULI_NAMES = {"PauliX": "X", "PauliY": "Y", "PauliZ": "Z", "Identity": "I"}


def _device_base():
    try:
        import pennylane as qml

        return qml.devices.Device
    except Exception:   # PennyLane absent (this image): a plain class with the same execute() contract
        return object


def _wires(obj) -> List[int]:
    w = obj.wires
    return [int(v) for v in (w.tolist() if hasattr(w, "tolist") else list(w))]


def _pauli_terms(obs, n: int, wire_to_qubit) -> List[Tuple[float, str]]:
    """An observable as [(coefficient, Pauli string)] — Qiskit-style text, rightmost character = qubit 0."""
    name = getattr(obs, "name", type(obs).__name__)
    if name in _PAULI_NAMES:
        chars = ["I"] * n
        if name != "Identity":
            chars[n - 1 - wire_to_qubit(_wires(obs)[0])] = _PAULI_NAMES[name]
        return [(1.0, "".join(chars))]
    if name in ("Prod", "Tensor"):
        factors = getattr(obs, "operands", None) or getattr(obs, "obs")
        terms = [(1.0, "I" * n)]
        for f in factors:
            nxt = []
            for c1, s1 in terms:
                for c2, s2 in _pauli_terms(f, n, wire_to_qubit):
                    merged = []
                    for a, b in zip(s1, s2):
                        if a != "I" and b != "I":
                            raise NotImplementedError("products of Pauli operators on the same wire")
                        merged.append(a if a != "I" else b)
                    nxt.append((c1 * c2, "".join(merged)))
            terms = nxt
        return terms
    if name == "SProd":
        return [(float(obs.scalar) * c, s) for c, s in _pauli_terms(obs.base, n, wire_to_qubit)]
    if name in ("Sum", "Hamiltonian", "LinearCombination"):
        if hasattr(obs, "terms") and name != "Sum":
            coeffs, ops = obs.terms()
            return [(float(c) * cc, s) for c, o in zip(coeffs, ops) for cc, s in _pauli_terms(o, n, wire_to_qubit)]
        return [t for o in obs.operands for t in _pauli_terms(o, n, wire_to_qubit)]
    raise NotImplementedError(f"observable {name} is not supported by the b200 device")


class B200QubitDevice(_device_base()):
    """``qml.device``-style statevector device on the B200 engine (analytic: ``shots`` must be None)."""

    name = "b200.qubit"
    short_name = "b200.qubit"

    def __init__(self, wires=None, shots=None, executor: Optional[Executor] = None):
        if shots is not None:
            raise ValueError("b200.qubit is a statevector device: shots must be None")
        base = _device_base()
        if base is not object:
            super().__init__(wires=wires, shots=None)
        else:
            self._wires = list(range(wires)) if isinstance(wires, int) else (list(wires) if wires is not None else None)
        self._executor = executor
        self._cache = {}
        self.num_executions = 0          # engine calls made (the tests assert the batching through it)

    @property
    def executor(self) -> Executor:
        if self._executor is None:
            self._executor = Executor("b200")     # raises without a CUDA device: there is no CPU path
        return self._executor

    # ------------------------------------------------------------------ tape -> program
    def _num_wires(self, tapes) -> int:
        own = getattr(self, "wires", None) if _device_base() is not object else self._wires
        if own is not None and len(own) > 0:
            return len(own)
        return 1 + max(w for t in tapes for op in t.operations for w in _wires(op))

    @staticmethod
    def _structure(tape, n):
        """Gate structure of a tape (what the planner sees) and its angle values."""
        gates, angles = [], []
        for op in tape.operations:
            name = op.name
            qubits = tuple(n - 1 - w for w in _wires(op))
            params = list(getattr(op, "parameters", None) or getattr(op, "data", ()) or ())
            if name == "PauliRot":
                word = op.hyperparameters["pauli_word"]
                if word not in _PAULI_ROT:
                    raise NotImplementedError(f"PauliRot({word}) is not in the gate table")
                gate = _PAULI_ROT[word]
            elif name in _OPS:
                gate = _OPS[name]
            else:
                raise NotImplementedError(f"operation {name} is not supported by the b200 device")
            extra = (0.0, 0.0)
            theta = None
            if gate == "u":
                theta, extra = params[0], (float(np.asarray(params[1])), float(np.asarray(params[2])))
            elif params:
                theta = params[0]
            gates.append((gate, qubits, theta is not None, extra))
            angles.append(None if theta is None else np.asarray(theta, dtype=np.float64))
        return tuple(gates), angles

    def _circuit(self, structure, n):
        """Compiled-circuit cache entry of a gate structure.  Every angle reads its row of the device angle table
        (B200SQ_FN_TABLE), so the plan depends on the gates only and is reused across calls; ``rows`` is the
        mutable list the table is filled from."""
        entry = self._cache.get((structure, n))
        if entry is None:
            rows: List[Optional[np.ndarray]] = [None] * len(structure)
            ops = []
            for k, (gate, qubits, has_angle, extra) in enumerate(structure):
                angle = Angle(table=lambda x, p, k=k: rows[k]) if has_angle else None
                ops.append(GateOp(gate, qubits, angle, extra))
            entry = (B200Circuit(Program(n, ops, 0, 0)), rows)
            self._cache[(structure, n)] = entry
        return entry

    # ------------------------------------------------------------------ execution
    def execute(self, circuits, execution_config=None):
        tapes = list(circuits) if isinstance(circuits, (list, tuple)) else [circuits]
        n = self._num_wires(tapes)
        results = [None] * len(tapes)
        # group tapes with identical gates and measurements: a shifted-tape batch becomes one engine call
        groups = {}
        for i, t in enumerate(tapes):
            structure, angles = self._structure(t, n)
            key = (structure, tuple(self._measurement_key(m) for m in t.measurements))
            groups.setdefault(key, []).append((i, t, angles))
        eng = self.executor.engine
        for (structure, _), members in groups.items():
            sizes = []
            for _, t, angles in members:
                bs = {a.shape[0] for a in angles if a is not None and a.ndim == 1}
                if len(bs) > 1:
                    raise ValueError("inconsistent broadcast sizes in one tape")
                sizes.append(bs.pop() if bs else 1)
            total = int(sum(sizes))
            circ, rows = self._circuit(structure, n)
            for k in range(len(structure)):
                if structure[k][2]:
                    rows[k] = np.concatenate([np.broadcast_to(angles[k], (sz,))
                                              for (_, _, angles), sz in zip(members, sizes)])
            outs = self._measure(eng, circ.compiled(eng, None), np.zeros((total, 0)), members[0][1].measurements, n)
            self.num_executions += 1
            off = 0
            for (i, t, angles), sz in zip(members, sizes):
                batched = any(a is not None and a.ndim == 1 for a in angles)
                vals = [o[off:off + sz] if batched else o[off] for o in outs]
                results[i] = vals[0] if len(vals) == 1 else tuple(vals)
                off += sz
        return tuple(results) if isinstance(circuits, (list, tuple)) else results[0]

    @staticmethod
    def _measurement_key(mp):
        kind = type(mp).__name__
        obs = getattr(mp, "obs", None)
        return (kind, None if obs is None else repr(obs))

    def _measure(self, eng, cprog, x, measurements, n):
        """One engine call for all measurements of a tape group: states if any qml.state() is asked for,
        otherwise the fused simulate + Pauli reduction."""
        kinds = [type(m).__name__ for m in measurements]
        if any(k == "StateMP" for k in kinds):
            psi = eng.simulate(cprog, x).cpu().numpy()
            outs = []
            for m, k in zip(measurements, kinds):
                if k == "StateMP":
                    outs.append(psi)
                elif k == "ExpectationMP":
                    terms = _pauli_terms(m.obs, n, lambda w: n - 1 - w)
                    ev = eng.expvals(eng.to_device(psi), n, [s for _, s in terms]).cpu().numpy()
                    outs.append(ev @ np.array([c for c, _ in terms]))
                else:
                    raise NotImplementedError(f"measurement {k} is not supported by the b200 device")
            return outs
        strings, index, per_mp = [], {}, []
        for m, k in zip(measurements, kinds):
            if k != "ExpectationMP":
                raise NotImplementedError(f"measurement {k} is not supported by the b200 device")
            entries = []
            for c, s in _pauli_terms(m.obs, n, lambda w: n - 1 - w):
                if s not in index:
                    index[s] = len(strings)
                    strings.append(s)
                entries.append((index[s], c))
            per_mp.append(entries)
        ev = eng.simulate_expvals(cprog, x, strings).cpu().numpy()      # (N, n_terms)
        return [sum(c * ev[:, t] for t, c in entries) for entries in per_mp]
