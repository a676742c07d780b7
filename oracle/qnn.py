"""QNN outputs and gradients (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates the numerical targets of ``LowLevelQNN*.evaluate`` for the requests on the hot path
(src/squlearn/qnn/lowlevel_qnn/lowlevel_qnn_pennylane.py:190-370, request vocabulary
src/squlearn/qnn/lowlevel_qnn/evaluation_classes.py:77-229):

  "f"     <O_k>(x_i)                    "dfdp"   d f / d param
  "dfdop" d f / d param_obs             "fcc"    <O_k^2>       "var" = fcc - f^2

Result axis order: [samples (if multi x)] [outputs (if the observable is a list)] [derivative]
(lowlevel_qnn_pennylane.py:331-359).

Two independent gradient routes are provided so they can check each other:
  * ``_grad_generator``: exact derivative by inserting dU/dtheta at every gate occurrence
    (what autograd backprop — the reference's PennyLane route, :563-581 — evaluates);
  * ``parameter_shift_variants``: the shifted-circuit batch of
    src/squlearn/util/optree/optree_derivative.py:130-158 (theta_k +- pi/2, weights
    +-1/2 d theta_k/d p), extended by the four-term rule for controlled rotations, which the
    reference reaches by transpiling them to single-qubit rotations first (:352-370).
The reference asserts all its routes equal (tests/qnn/lowlevel_qnn/test_lowlevel_qnn.py:36-65).
"""

import numpy as np

from .gates import gate_matrix, GATE_ARITY
from .statevector import Gate, simulate
from .observables import (
    pauli_expectation,
    single_pauli_string,
    summed_paulis,
    pauli_sum_squared,
)

_TWO_TERM = {"rx", "ry", "rz", "p", "cp", "rxx", "ryy", "rzz", "rzx"}  # one frequency
_FOUR_TERM = {"crx", "cry", "crz"}  # frequencies 1/2 and 1


# ----------------------------------------------------------------------------- observables
def observable_num_parameters(obs, n):
    kind, kw = obs
    if kind == "summed_paulis":
        op_str = kw.get("op_str", "Z")
        num = 1 if kw.get("include_identity", True) else 0
        return num + (len(op_str) * n if kw.get("full_sum", True) else len(op_str))
    if kind == "single_pauli":
        return 1 if kw.get("parameterized", False) else 0
    if kind == "custom":
        ops = kw["operator_string"]
        ops = [ops] if isinstance(ops, str) else list(ops)
        return len(ops) if kw.get("parameterized", False) else 0
    raise ValueError(kind)


def observable_terms(obs, n, params):
    """Return (strings, coefficient values, index of the parameter each coefficient is, or -1)."""
    kind, kw = obs
    if kind == "summed_paulis":
        idx_params = np.arange(observable_num_parameters(obs, n), dtype=np.float64)
        strings, idx = summed_paulis(n, idx_params, **{k: v for k, v in kw.items()})
        idx = idx.astype(int)
        return strings, np.asarray(params, dtype=np.float64)[idx], idx
    if kind == "single_pauli":
        s = [single_pauli_string(n, kw["qubit"], kw.get("op_str", "Z"))]
        if kw.get("parameterized", False):
            return s, np.array([params[0]]), np.array([0])
        return s, np.array([1.0]), np.array([-1])
    if kind == "custom":
        ops = kw["operator_string"]
        ops = [ops] if isinstance(ops, str) else list(ops)
        if kw.get("parameterized", False):
            return ops, np.asarray(params, dtype=np.float64)[: len(ops)], np.arange(len(ops))
        return ops, np.ones(len(ops)), -np.ones(len(ops), dtype=int)
    raise ValueError(kind)


# ----------------------------------------------------------------------------- gradients
def _dtheta_dp(angle, x, p, j):
    """d angle / d p_j for an angle that is linear in p (exact: central difference, step 1)."""
    if not callable(angle):
        return 0.0
    e = np.zeros_like(p)
    e[j] = 1.0
    return 0.5 * (np.asarray(angle(x, p + e)) - np.asarray(angle(x, p - e)))


def _dmatrix(name, theta):
    """dU/dtheta for single-angle gates, shape (B, 2^k, 2^k)."""
    u = gate_matrix(name, theta)
    x_ = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    y_ = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
    z_ = np.array([[1, 0], [0, -1]], dtype=np.complex128)
    gen = {"rx": x_, "ry": y_, "rz": z_, "rxx": np.kron(x_, x_), "ryy": np.kron(y_, y_),
           "rzz": np.kron(z_, z_), "rzx": np.kron(z_, x_)}
    if name in gen:
        return -0.5j * np.einsum("ij,bjk->bik", gen[name], u)
    if name in ("crx", "cry", "crz"):
        g = np.zeros((4, 4), dtype=np.complex128)
        g[2:, 2:] = gen[name[1:]]
        return -0.5j * np.einsum("ij,bjk->bik", g, u)
    if name == "p":
        out = np.zeros_like(u)
        out[:, 1, 1] = 1j * u[:, 1, 1]
        return out
    if name == "cp":
        out = np.zeros_like(u)
        out[:, 3, 3] = 1j * u[:, 3, 3]
        return out
    raise NotImplementedError(f"no derivative rule for gate {name}")


def _apply(psi, n, qubits, mat):
    n_samples = psi.shape[0]
    k = len(qubits)
    t = psi.reshape((n_samples,) + (2,) * n)
    axes = [1 + (n - 1 - q) for q in qubits]
    dest = list(range(n + 1 - k, n + 1))
    t = np.moveaxis(t, axes, dest)
    shp = t.shape
    flat = np.matmul(t.reshape(n_samples, -1, 2**k), np.swapaxes(mat, 1, 2))
    return np.ascontiguousarray(np.moveaxis(flat.reshape(shp), dest, axes)).reshape(n_samples, -1)


def _apply_pauli_sum(psi, strings, coeffs):
    from .observables import _masks

    dim = psi.shape[1]
    idx = np.arange(dim)
    out = np.zeros_like(psi)
    for s, c in zip(strings, coeffs):
        xm, zm, ny = _masks(s)
        par = np.zeros(dim, dtype=np.int64)
        t = idx & zm
        while np.any(t):
            par ^= t & 1
            t >>= 1
        tmp = np.zeros_like(psi)
        tmp[:, idx ^ xm] = (1j**ny) * (1.0 - 2.0 * par)[None, :] * psi
        out += c * tmp
    return out


def _grad_generator(gates, n, x, p, term_lists):
    """d<O_k>/dp_j for every observable k: array (N, n_obs, P)."""
    n_samples = x.shape[0]
    num_p = len(p)
    psi = simulate(gates, n, x, p)
    opsi = [_apply_pauli_sum(psi, s, c) for (s, c) in term_lists]
    grad = np.zeros((n_samples, len(term_lists), num_p))
    for k, g in enumerate(gates):
        if not g.angles or not any(callable(a) for a in g.angles):
            continue
        if len(g.angles) != 1:
            raise NotImplementedError("multi-angle gates are not differentiated by the oracle")
        fac = np.stack([np.broadcast_to(_dtheta_dp(g.angles[0], x, p, j), (n_samples,))
                        for j in range(num_p)], axis=1)  # (N, P)
        if not np.any(fac):
            continue
        theta = np.asarray(g.angles[0](x, p), dtype=np.float64)
        before = simulate(gates[:k], n, x, p)
        dpsi = _apply(before, n, g.qubits, _dmatrix(g.name, theta))
        dpsi = simulate(gates[k + 1:], n, x, p, initial_state=dpsi)
        for o, op in enumerate(opsi):
            d = 2.0 * np.real(np.sum(op.conj() * dpsi, axis=1))  # (N,)
            grad[:, o, :] += d[:, None] * fac
    return grad


def parameter_shift_variants(gates, x, p):
    """The shifted batch: list of (gate index, shift, weight[(N,)], parameter index).

    For every occurrence k of parameter p_j in a gate angle theta_k (linear in p_j):
    two-term rule (theta_k +- pi/2, weights +-1/2 dtheta_k/dp_j) for Pauli rotations,
    four-term rule for controlled rotations.
    """
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[None]
    n_samples = x.shape[0]
    p = np.asarray(p, dtype=np.float64)
    out = []
    for k, g in enumerate(gates):
        if len(g.angles) != 1 or not callable(g.angles[0]):
            continue
        for j in range(len(p)):
            fac = np.broadcast_to(np.asarray(_dtheta_dp(g.angles[0], x, p, j), dtype=np.float64),
                                  (n_samples,))
            if not np.any(fac):
                continue
            if g.name in _TWO_TERM:
                out.append((k, +0.5 * np.pi, 0.5 * fac, j))
                out.append((k, -0.5 * np.pi, -0.5 * fac, j))
            elif g.name in _FOUR_TERM:
                cp = (np.sqrt(2.0) + 1.0) / (4.0 * np.sqrt(2.0))
                cm = (np.sqrt(2.0) - 1.0) / (4.0 * np.sqrt(2.0))
                out.append((k, +0.5 * np.pi, +cp * fac, j))
                out.append((k, -0.5 * np.pi, -cp * fac, j))
                out.append((k, +1.5 * np.pi, -cm * fac, j))
                out.append((k, -1.5 * np.pi, +cm * fac, j))
            else:
                raise NotImplementedError(f"no shift rule for gate {g.name}")
    return out


def _shifted(gates, k, shift):
    g = gates[k]
    a = g.angles[0]
    new = Gate(g.name, g.qubits, ((lambda x, p, a=a, s=shift: a(x, p) + s),))
    return gates[:k] + [new] + gates[k + 1:]


def _grad_parameter_shift(gates, n, x, p, term_lists):
    n_samples = x.shape[0]
    grad = np.zeros((n_samples, len(term_lists), len(p)))
    for k, shift, weight, j in parameter_shift_variants(gates, x, p):
        psi = simulate(_shifted(gates, k, shift), n, x, p)
        for o, (s, c) in enumerate(term_lists):
            val = sum(ci * pauli_expectation(psi, si) for si, ci in zip(s, c))
            grad[:, o, j] += weight * val
    return grad


# ----------------------------------------------------------------------------- evaluate
def qnn_evaluate(gates, n_qubits, x, param, observables, param_obs, values=("f",),
                 gradient="generator"):
    """Oracle for LowLevelQNN.evaluate on a batch ``x`` of shape (N, d).

    ``observables``: one observable spec or a list of them (list => multiple outputs);
    a spec is ``(kind, kwargs)`` with kind in {"summed_paulis", "single_pauli", "custom"}.
    ``param_obs``: concatenated observable parameters in list order.
    Returns a dict with the requested keys.
    """
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    param = np.asarray(param, dtype=np.float64)
    param_obs = np.asarray(param_obs, dtype=np.float64)
    multiple_output = isinstance(observables, list)
    obs_list = observables if multiple_output else [observables]
    n = n_qubits

    term_lists, index_lists, offs = [], [], []
    off = 0
    for obs in obs_list:
        npo = observable_num_parameters(obs, n)
        s, c, idx = observable_terms(obs, n, param_obs[off:off + npo])
        term_lists.append((s, c))
        index_lists.append(idx)
        offs.append(off)
        off += npo
    num_pop = off

    psi = simulate(gates, n, x, param)
    out = {}
    wanted = set(values)
    if "var" in wanted:
        wanted |= {"f", "fcc"}

    def _shape(arr):  # (N, n_obs, ...) -> drop output axis if single observable
        return arr if multiple_output else arr[:, 0]

    exp_cache = {}

    def _exp(s):
        if s not in exp_cache:
            exp_cache[s] = pauli_expectation(psi, s)
        return exp_cache[s]

    if "f" in wanted:
        f = np.stack([sum(ci * _exp(si) for si, ci in zip(s, c)) for (s, c) in term_lists], axis=1)
        out["f"] = _shape(f)
    if "fcc" in wanted:
        fcc = []
        for (s, c) in term_lists:
            s2, c2 = pauli_sum_squared(s, c)
            fcc.append(sum(np.real(ci) * _exp(si) for si, ci in zip(s2, c2)))
        out["fcc"] = _shape(np.stack(fcc, axis=1))
    if "var" in wanted:
        out["var"] = out["fcc"] - np.square(out["f"])
    if "dfdop" in wanted:
        d = np.zeros((x.shape[0], len(obs_list), num_pop))
        for o, ((s, c), idx) in enumerate(zip(term_lists, index_lists)):
            for si, ii in zip(s, idx):
                if ii >= 0:
                    d[:, o, offs[o] + ii] += _exp(si)
        out["dfdop"] = _shape(d)
    if "dfdp" in wanted:
        fn = _grad_generator if gradient == "generator" else _grad_parameter_shift
        out["dfdp"] = _shape(fn(gates, n, x, param, term_lists))
    return {k: out[k] for k in values}
