"""Gate matrices of the oracle (TEST INFRASTRUCTURE — see oracle/__init__.py).

Gate names are the Qiskit names of the reference's gate table
(src/squlearn/util/pennylane/pennylane_gates.py:56-93, Qulacs subset
src/squlearn/util/qulacs/qulacs_gates.py:105-122).  The matrices themselves live in
third-party engines (PennyLane / Qiskit, not vendored); they are restated here from
their published definitions:

  RX/RY/RZ(t) = exp(-i t sigma / 2),  P(l) = diag(1, e^{il}),
  U(t, p, l)  = [[cos t/2, -e^{il} sin t/2], [e^{ip} sin t/2, e^{i(p+l)} cos t/2]],
  CRZ(t; c, t) = |0><0| (x) I + |1><1| (x) RZ(t),   RZZ(t) = exp(-i t/2 Z(x)Z), ...

Matrix index convention of THIS module: for a k-qubit gate applied to qubits
(q0, ..., q_{k-1}) the row/column index is the bit string b_{q0} b_{q1} ... with the
FIRST listed qubit as the most significant bit.  For controlled gates the first
listed qubit is the control (pennylane_circuit.py:397-401 passes Qiskit's qubit
order straight through as wires).
"""

import numpy as np

_I2 = np.eye(2, dtype=np.complex128)
_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
_Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
_H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2.0)
_S = np.diag([1, 1j]).astype(np.complex128)
_T = np.diag([1, np.exp(0.25j * np.pi)]).astype(np.complex128)
_SX = 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]], dtype=np.complex128)

# number of qubits / number of angle arguments per gate name
GATE_ARITY = {
    "id": (1, 0), "h": (1, 0), "x": (1, 0), "y": (1, 0), "z": (1, 0), "s": (1, 0),
    "sdg": (1, 0), "t": (1, 0), "tdg": (1, 0), "sx": (1, 0),
    "rx": (1, 1), "ry": (1, 1), "rz": (1, 1), "p": (1, 1), "u": (1, 3),
    "cx": (2, 0), "cy": (2, 0), "cz": (2, 0), "ch": (2, 0), "cs": (2, 0), "csx": (2, 0),
    "swap": (2, 0), "iswap": (2, 0), "ecr": (2, 0),
    "cp": (2, 1), "crx": (2, 1), "cry": (2, 1), "crz": (2, 1),
    "rxx": (2, 1), "ryy": (2, 1), "rzz": (2, 1), "rzx": (2, 1),
    "ccx": (3, 0), "cswap": (3, 0),
}

DIAGONAL_GATES = {"id", "z", "s", "sdg", "t", "tdg", "rz", "p", "cz", "cs", "cp", "crz", "rzz"}


def _as_batch(theta):
    theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
    return theta


def _rot(theta, pauli):
    """exp(-i theta/2 P) for a (batch of) angle(s): shape (B, dim, dim)."""
    theta = _as_batch(theta)
    c = np.cos(0.5 * theta)[:, None, None]
    s = np.sin(0.5 * theta)[:, None, None]
    eye = np.eye(pauli.shape[0], dtype=np.complex128)
    return c * eye[None] - 1j * s * pauli[None]


def _controlled(u):
    """|0><0| (x) I + |1><1| (x) U with the control as the most significant bit."""
    u = np.asarray(u, dtype=np.complex128)
    if u.ndim == 2:
        u = u[None]
    b, d, _ = u.shape
    out = np.zeros((b, 2 * d, 2 * d), dtype=np.complex128)
    out[:, :d, :d] = np.eye(d)
    out[:, d:, d:] = u
    return out


def gate_matrix(name, *angles):
    """Return the (B, 2^k, 2^k) complex128 matrix of gate ``name``.

    ``angles`` are scalars or (N,) arrays (one angle per data point); B is 1 for
    constant gates and N for per-sample angles.
    """
    name = name.lower()
    if name not in GATE_ARITY:
        raise NotImplementedError(f"gate {name!r} is not in the reference gate table")
    nq, na = GATE_ARITY[name]
    if len(angles) != na:
        raise ValueError(f"gate {name} takes {na} angle(s), got {len(angles)}")

    if name == "id":
        return _I2[None]
    if name == "h":
        return _H[None]
    if name == "x":
        return _X[None]
    if name == "y":
        return _Y[None]
    if name == "z":
        return _Z[None]
    if name == "s":
        return _S[None]
    if name == "sdg":
        return _S.conj().T[None]
    if name == "t":
        return _T[None]
    if name == "tdg":
        return _T.conj().T[None]
    if name == "sx":
        return _SX[None]
    if name == "rx":
        return _rot(angles[0], _X)
    if name == "ry":
        return _rot(angles[0], _Y)
    if name == "rz":
        return _rot(angles[0], _Z)
    if name == "p":
        lam = _as_batch(angles[0])
        out = np.zeros((lam.shape[0], 2, 2), dtype=np.complex128)
        out[:, 0, 0] = 1.0
        out[:, 1, 1] = np.exp(1j * lam)
        return out
    if name == "u":
        th, ph, lam = np.broadcast_arrays(*[_as_batch(a) for a in angles])
        out = np.zeros((th.shape[0], 2, 2), dtype=np.complex128)
        c, s = np.cos(0.5 * th), np.sin(0.5 * th)
        out[:, 0, 0] = c
        out[:, 0, 1] = -np.exp(1j * lam) * s
        out[:, 1, 0] = np.exp(1j * ph) * s
        out[:, 1, 1] = np.exp(1j * (ph + lam)) * c
        return out
    if name == "cx":
        return _controlled(_X)
    if name == "cy":
        return _controlled(_Y)
    if name == "cz":
        return _controlled(_Z)
    if name == "ch":
        return _controlled(_H)
    if name == "cs":
        return _controlled(_S)
    if name == "csx":
        return _controlled(_SX)
    if name == "cp":
        return _controlled(gate_matrix("p", angles[0]))
    if name == "crx":
        return _controlled(_rot(angles[0], _X))
    if name == "cry":
        return _controlled(_rot(angles[0], _Y))
    if name == "crz":
        return _controlled(_rot(angles[0], _Z))
    if name == "swap":
        m = np.zeros((4, 4), dtype=np.complex128)
        m[0, 0] = m[3, 3] = m[1, 2] = m[2, 1] = 1.0
        return m[None]
    if name == "iswap":
        m = np.zeros((4, 4), dtype=np.complex128)
        m[0, 0] = m[3, 3] = 1.0
        m[1, 2] = m[2, 1] = 1j
        return m[None]
    if name == "ecr":
        # (X_{q0} - Y_{q0} X_{q1}) / sqrt(2), q0 = first listed qubit (most significant here)
        m = (np.kron(_X, _I2) - np.kron(_Y, _X)) / np.sqrt(2.0)
        return m[None]
    if name == "rxx":
        return _rot(angles[0], np.kron(_X, _X))
    if name == "ryy":
        return _rot(angles[0], np.kron(_Y, _Y))
    if name == "rzz":
        return _rot(angles[0], np.kron(_Z, _Z))
    if name == "rzx":
        # Z on the first listed qubit, X on the second
        return _rot(angles[0], np.kron(_Z, _X))
    if name == "ccx":
        return _controlled(_controlled(_X)[0])
    if name == "cswap":
        return _controlled(gate_matrix("swap")[0])
    raise NotImplementedError(name)
