"""Batched statevector simulation of a gate list (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates what the reference's statevector path does once the gate list exists:
``PennyLaneCircuit.pennylane_circuit`` replays the list, evaluating every angle
function on arrays of shape (N,) (src/squlearn/util/pennylane/pennylane_circuit.py:541-626),
and PennyLane ``default.qubit`` (third party, not vendored; pinned ``pennylane>=0.34.0``
in pyproject.toml:37) applies each gate to the ``(N, 2, ..., 2)`` complex128 array with
one einsum/tensordot per gate under parameter broadcasting.  ``simulate_reference_style``
is that algorithm (used as the timed CPU baseline); ``simulate`` is the same arithmetic
with the batch axis moved so NumPy's matmul does the work a little faster (used as the
parity checker at larger sizes).  Both return little-endian ``(N, 2**n)`` arrays.
"""

from collections import namedtuple

import numpy as np

from .gates import gate_matrix, GATE_ARITY

# angles: tuple whose entries are floats or callables (x, p) -> scalar or (N,) array,
# x being the (N, d) feature batch and p the circuit parameter vector.
Gate = namedtuple("Gate", ["name", "qubits", "angles"])


def _eval_angles(gate, x, p, n_samples):
    vals = []
    for a in gate.angles:
        v = a(x, p) if callable(a) else a
        v = np.asarray(v, dtype=np.float64)
        if v.ndim == 0:
            v = v.reshape(1)
        elif v.shape[0] not in (1, n_samples):
            raise ValueError("angle batch does not match the number of samples")
        vals.append(v)
    return vals


def _axis(n, q):
    """Axis of qubit q in an (N, 2, ..., 2) array whose C-order flattening is little-endian."""
    return 1 + (n - 1 - q)


def simulate_reference_style(gates, n_qubits, x, p=None, initial_state=None):
    """Per-gate einsum over the (N, 2, ..., 2) array — the reference's hot loop A.

    Args:
        gates: list of ``Gate``.
        n_qubits: number of qubits n.
        x: (N, d) float64 features (N data points).
        p: circuit parameter vector (or None).
    Returns:
        (N, 2**n) complex128, little-endian.
    """
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[None, :]
    n_samples = x.shape[0]
    n = n_qubits
    if initial_state is None:
        psi = np.zeros((n_samples,) + (2,) * n, dtype=np.complex128)
        psi[(slice(None),) + (0,) * n] = 1.0
    else:
        psi = np.array(initial_state, dtype=np.complex128).reshape((n_samples,) + (2,) * n)
    letters = "abcdefghijklmnopqrstuvwxyz"
    for g in gates:
        k, _ = GATE_ARITY[g.name]
        if len(g.qubits) != k:
            raise ValueError(f"gate {g.name} expects {k} qubits")
        mat = gate_matrix(g.name, *_eval_angles(g, x, p, n_samples))
        mat = mat.reshape((mat.shape[0],) + (2,) * (2 * k))
        # einsum:  U[N?, o1..ok, i1..ik] * psi[N, ..., i1, ..., ik, ...]
        state_idx = ["N"] + [letters[j] for j in range(n)]
        out_idx = list(state_idx)
        in_letters, out_letters = [], []
        for j, q in enumerate(g.qubits):
            ax = _axis(n, q)
            in_letters.append(state_idx[ax])
            new = letters[n + j]
            out_letters.append(new)
            out_idx[ax] = new
        u_idx = (["N"] if mat.shape[0] != 1 else []) + out_letters + in_letters
        if mat.shape[0] == 1:
            mat = mat[0]
        expr = "".join(u_idx) + "," + "".join(state_idx) + "->" + "".join(out_idx)
        psi = np.einsum(expr, mat, psi)
    return np.ascontiguousarray(psi.reshape(n_samples, 2**n))


def simulate(gates, n_qubits, x, p=None, initial_state=None):
    """Same arithmetic as ``simulate_reference_style`` through batched matmul (faster checker)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[None, :]
    n_samples = x.shape[0]
    n = n_qubits
    dim = 2**n
    if initial_state is None:
        psi = np.zeros((n_samples, dim), dtype=np.complex128)
        psi[:, 0] = 1.0
    else:
        psi = np.array(initial_state, dtype=np.complex128).reshape(n_samples, dim)
    psi = psi.reshape((n_samples,) + (2,) * n)
    for g in gates:
        k, _ = GATE_ARITY[g.name]
        mat = gate_matrix(g.name, *_eval_angles(g, x, p, n_samples))  # (B, 2^k, 2^k)
        axes = [_axis(n, q) for q in g.qubits]
        dest = list(range(n + 1 - k, n + 1))
        psi = np.moveaxis(psi, axes, dest)
        shp = psi.shape
        flat = psi.reshape(n_samples, -1, 2**k)
        flat = np.matmul(flat, np.swapaxes(mat, 1, 2))  # out[n,r,i] = sum_j flat[n,r,j] U[n,i,j]
        psi = np.moveaxis(flat.reshape(shp), dest, axes)
    return np.ascontiguousarray(psi.reshape(n_samples, dim))


def to_pennylane_order(states, n_qubits):
    """Bit-reverse the amplitude index: little-endian -> PennyLane's wire-0-is-MSB order."""
    states = np.asarray(states)
    n = n_qubits
    lead = states.shape[:-1]
    t = states.reshape(lead + (2,) * n)
    t = np.transpose(t, list(range(len(lead))) + [len(lead) + n - 1 - j for j in range(n)])
    return np.ascontiguousarray(t.reshape(lead + (2**n,)))
