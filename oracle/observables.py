"""Pauli observables and expectation values (TEST INFRASTRUCTURE — see oracle/__init__.py).

String conventions restated from the reference:
  * Qiskit little-endian text: the RIGHTMOST character acts on qubit 0
    (SinglePauli.get_pauli  src/squlearn/observables/observable_implemented/single_pauli.py:103-118;
     SummedPaulis.get_pauli  .../summed_paulis.py:113-148;
     CustomObservable.get_pauli .../custom_observable.py:98-128).
  * ProjectedQuantumKernel "XYZ" measurement: for each of X, Y, Z, for each qubit i a SinglePauli
    -> feature order [X_0..X_{n-1}, Y_0.., Z_0..]
    (src/squlearn/kernel/lowlevel_kernel/projected_quantum_kernel.py:861-877).
  * QRC random strings: src/squlearn/qrc/util/random_pauli_list.py:6-38.
The expectation value itself (qml.expval of a Pauli word,
src/squlearn/util/pennylane/pennylane_circuit.py:636-658) is restated from its definition
<psi|P|psi>.
"""

import numpy as np


def single_pauli_string(n, qubit, op):
    if not 0 <= qubit < n:
        raise ValueError("Specified qubit out of range")
    h = "I" * n
    return h[(qubit + 1):] + op + h[:qubit]


def xyz_measurement_strings(n, measurement="XYZ"):
    return [single_pauli_string(n, i, m) for m in measurement for i in range(n)]


def summed_paulis(n, params, op_str="Z", full_sum=True, include_identity=True):
    """Return (strings, coefficients) of SummedPaulis for the parameter vector ``params``."""
    params = np.asarray(params, dtype=np.float64)
    nparam = len(params)
    ioff = 0
    strings, coeffs = [], []
    if include_identity:
        strings.append("I" * n)
        coeffs.append(params[ioff % nparam])
        ioff += 1
    for op in op_str:
        for i in range(n):
            strings.append(single_pauli_string(n, i, op))
            coeffs.append(params[ioff % nparam])
            if full_sum:
                ioff += 1
        if not full_sum:
            ioff += 1
    return strings, np.array(coeffs)


def _mul_pauli_chars(a, b):
    """Product of two single-qubit Paulis: returns (phase, char)."""
    if a == "I":
        return 1.0, b
    if b == "I":
        return 1.0, a
    if a == b:
        return 1.0, "I"
    table = {("X", "Y"): (1j, "Z"), ("Y", "X"): (-1j, "Z"), ("Y", "Z"): (1j, "X"),
             ("Z", "Y"): (-1j, "X"), ("Z", "X"): (1j, "Y"), ("X", "Z"): (-1j, "Y")}
    return table[(a, b)]


def pauli_sum_squared(strings, coeffs):
    """(sum_k c_k P_k)^2 simplified — what SparsePauliOp.power(2).simplify() yields
    (used for "fcc"/"var": src/squlearn/qnn/lowlevel_qnn/lowlevel_qnn_pennylane.py:381,395-397)."""
    acc = {}
    for s1, c1 in zip(strings, coeffs):
        for s2, c2 in zip(strings, coeffs):
            ph, out = 1.0 + 0j, []
            for a, b in zip(s1, s2):
                p, ch = _mul_pauli_chars(a, b)
                ph *= p
                out.append(ch)
            key = "".join(out)
            acc[key] = acc.get(key, 0.0) + ph * c1 * c2
    keys = [k for k, v in acc.items() if abs(v) > 0.0]
    return keys, np.real_if_close(np.array([acc[k] for k in keys]))


def summed_paulis_squared(n, params, **kw):
    s, c = summed_paulis(n, params, **kw)
    return pauli_sum_squared(s, c)


def _masks(string):
    n = len(string)
    xm = zm = 0
    ny = 0
    for pos, ch in enumerate(string):
        q = n - 1 - pos
        if ch == "X":
            xm |= 1 << q
        elif ch == "Y":
            xm |= 1 << q
            zm |= 1 << q
            ny += 1
        elif ch == "Z":
            zm |= 1 << q
        elif ch != "I":
            raise ValueError("Only Pauli operators I, X, Y, Z are allowed.")
    return xm, zm, ny


def pauli_expectation(states, string):
    """<psi_i| P |psi_i> for every row of the little-endian (N, 2**n) batch ``states``."""
    states = np.asarray(states, dtype=np.complex128)
    if states.ndim == 1:
        states = states[None]
    dim = states.shape[1]
    xm, zm, ny = _masks(string)
    idx = np.arange(dim)
    # P|j> = i^{ny} (-1)^{popcount(j & zm)} |j ^ xm>
    par = np.zeros(dim, dtype=np.int64)
    t = idx & zm
    while np.any(t):
        par ^= t & 1
        t >>= 1
    sign = 1.0 - 2.0 * par
    ppsi = np.zeros_like(states)
    ppsi[:, idx ^ xm] = (1j**ny) * sign[None, :] * states
    return np.real(np.sum(states.conj() * ppsi, axis=1))


def random_pauli_list(pauli_string_length, num_samples, seed=42):
    total = 4**pauli_string_length
    if num_samples > total:
        raise ValueError(f"num_samples={num_samples} exceeds total possible {total}")
    sym = np.array(["I", "X", "Y", "Z"])
    rng = np.random.default_rng(seed)
    picks = rng.choice(total, size=num_samples, replace=False)
    out = []
    for index in picks:
        chars, t = [], index
        for _ in range(pauli_string_length):
            chars.append(sym[t % 4])
            t //= 4
        out.append("".join(reversed(chars)))
    return out
