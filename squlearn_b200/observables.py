"""Pauli observables of the hot path (string / coefficient conventions of the reference).

  SinglePauli       src/squlearn/observables/observable_implemented/single_pauli.py:103-118
  SummedPaulis      .../summed_paulis.py:113-148
  CustomObservable  .../custom_observable.py:98-128
  base behaviour    src/squlearn/observables/observable_base.py:71-89 (default parameters = ones)

Strings are Qiskit little-endian text: the RIGHTMOST character acts on qubit 0.  The engine only
ever sees (x_mask, z_mask, real coefficient) triples; the algebra needed for "fcc"/"var"
(SparsePauliOp.power(2).simplify(), lowlevel_qnn_pennylane.py:381,395-397) is done here on the host.
"""

from typing import List, Sequence, Tuple, Union

import numpy as np


def pauli_masks(string: str) -> Tuple[int, int]:
    """(x_mask, z_mask) of a little-endian Pauli string; Y sets both bits."""
    n = len(string)
    xm = zm = 0
    for pos, ch in enumerate(string):
        q = n - 1 - pos
        if ch == "X":
            xm |= 1 << q
        elif ch == "Y":
            xm |= 1 << q
            zm |= 1 << q
        elif ch == "Z":
            zm |= 1 << q
        elif ch != "I":
            raise ValueError("Only Pauli operators I, X, Y, Z are allowed.")
    return xm, zm


_MUL = {("X", "Y"): (1j, "Z"), ("Y", "X"): (-1j, "Z"), ("Y", "Z"): (1j, "X"),
        ("Z", "Y"): (-1j, "X"), ("Z", "X"): (1j, "Y"), ("X", "Z"): (-1j, "Y")}


def pauli_sum_squared(strings: Sequence[str], coeffs: Sequence[float]):
    """(sum_k c_k P_k)^2 with like terms collected: (strings, real coefficients)."""
    acc = {}
    for s1, c1 in zip(strings, coeffs):
        for s2, c2 in zip(strings, coeffs):
            ph, chars = 1.0 + 0.0j, []
            for a, b in zip(s1, s2):
                if a == "I":
                    chars.append(b)
                elif b == "I" or a == b:
                    chars.append(a if b == "I" else "I")
                else:
                    p, ch = _MUL[(a, b)]
                    ph *= p
                    chars.append(ch)
            key = "".join(chars)
            acc[key] = acc.get(key, 0.0) + ph * c1 * c2
    keys = [k for k, v in acc.items() if abs(v) > 0.0]
    vals = np.array([acc[k] for k in keys])
    if np.max(np.abs(vals.imag), initial=0.0) > 1e-12 * max(1.0, np.max(np.abs(vals), initial=0.0)):
        raise ValueError("squared observable is not Hermitian")
    return keys, vals.real


class ObservableBase:
    def __init__(self, num_qubits: int):
        self._num_qubits = int(num_qubits)

    @property
    def num_qubits(self) -> int:
        return self._num_qubits

    @property
    def num_parameters(self) -> int:
        return 0

    def generate_initial_parameters(self, seed=None, ones: bool = True) -> np.ndarray:
        """observable_base.py:71-89: ones by default, uniform(-pi, pi) otherwise."""
        if self.num_parameters == 0:
            return np.array([])
        if ones:
            return np.ones(self.num_parameters)
        return np.random.RandomState(seed).uniform(-np.pi, np.pi, self.num_parameters)

    def get_pauli_terms(self, parameters) -> Tuple[List[str], np.ndarray, np.ndarray]:
        """(strings, coefficient values, index into ``parameters`` of each coefficient or -1)."""
        raise NotImplementedError

    def get_params(self, deep: bool = True) -> dict:
        return {"num_qubits": self._num_qubits}


class SinglePauli(ObservableBase):
    def __init__(self, num_qubits: int, qubit: int = 0, op_str: str = "Z",
                 parameterized: bool = False):
        super().__init__(num_qubits)
        if op_str not in ("I", "X", "Y", "Z"):
            raise ValueError("Only Pauli operators I, X, Y, Z are allowed.")
        self._qubit, self._op_str, self._parameterized = qubit, op_str, parameterized

    qubit = property(lambda self: self._qubit)
    op_str = property(lambda self: self._op_str)
    parameterized = property(lambda self: self._parameterized)

    @property
    def num_parameters(self) -> int:
        return 1 if self._parameterized else 0

    def get_pauli_terms(self, parameters=None):
        i, n = self._qubit, self.num_qubits
        if i < 0 or n <= i:
            raise ValueError("Specified qubit out of range")
        h = "I" * n
        s = [h[(i + 1):] + self._op_str + h[:i]]
        if self._parameterized:
            return s, np.array([float(parameters[0])]), np.array([0])
        return s, np.array([1.0]), np.array([-1])

    def get_params(self, deep=True):
        p = super().get_params()
        p.update(qubit=self._qubit, op_str=self._op_str, parameterized=self._parameterized)
        return p


class SummedPaulis(ObservableBase):
    def __init__(self, num_qubits: int, op_str: Union[str, tuple] = "Z", full_sum: bool = True,
                 include_identity: bool = True):
        super().__init__(num_qubits)
        for s in op_str:
            if s not in ("I", "X", "Y", "Z"):
                raise ValueError("Only Pauli operators I, X, Y, Z are allowed.")
        self._op_str, self._full_sum, self._include_identity = op_str, full_sum, include_identity

    op_str = property(lambda self: self._op_str)
    full_sum = property(lambda self: self._full_sum)
    include_identity = property(lambda self: self._include_identity)

    @property
    def num_parameters(self) -> int:
        num = 1 if self._include_identity else 0
        return num + (len(self._op_str) * self.num_qubits if self._full_sum else len(self._op_str))

    def get_pauli_terms(self, parameters):
        n = self.num_qubits
        parameters = np.asarray(parameters, dtype=np.float64)
        nparam = len(parameters)
        ioff, strings, idx = 0, [], []
        if self._include_identity:
            strings.append("I" * n)
            idx.append(ioff % nparam)
            ioff += 1
        for op in self._op_str:
            for i in range(n):
                h = "I" * n
                strings.append(h[i + 1:] + op + h[:i])
                idx.append(ioff % nparam)
                if self._full_sum:
                    ioff += 1
            if not self._full_sum:
                ioff += 1
        idx = np.asarray(idx, dtype=np.int64)
        return strings, parameters[idx], idx

    def get_params(self, deep=True):
        p = super().get_params()
        p.update(op_str=self._op_str, full_sum=self._full_sum,
                 include_identity=self._include_identity)
        return p


class CustomObservable(ObservableBase):
    def __init__(self, num_qubits: int, operator_string: Union[str, list, tuple],
                 parameterized: bool = False):
        super().__init__(num_qubits)
        ops = [operator_string] if isinstance(operator_string, str) else list(operator_string)
        for s in ops:
            if len(s) != num_qubits:
                raise ValueError("Supplied string has not the same size as the number of qubits, "
                                 f"please add missing identities as 'I'")
            for ch in s:
                if ch not in ("I", "X", "Y", "Z"):
                    raise ValueError("Only Pauli operators I, X, Y, Z are allowed.")
        self._operator_string, self._ops, self._parameterized = operator_string, ops, parameterized

    operator_string = property(lambda self: self._operator_string)
    parameterized = property(lambda self: self._parameterized)

    @property
    def num_parameters(self) -> int:
        return len(self._ops) if self._parameterized else 0

    def get_pauli_terms(self, parameters=None):
        if self._parameterized:
            parameters = np.asarray(parameters, dtype=np.float64)
            idx = np.arange(len(self._ops)) % len(parameters)
            return list(self._ops), parameters[idx], idx
        return list(self._ops), np.ones(len(self._ops)), -np.ones(len(self._ops), dtype=np.int64)

    def get_params(self, deep=True):
        p = super().get_params()
        p.update(operator_string=self._operator_string, parameterized=self._parameterized)
        return p
