"""Encoding circuits on the hot path, restated without Qiskit as gate-list builders.

Each class mirrors the constructor, properties and parameter initialisation of the reference
class of the same name and emits a :class:`~squlearn_b200.program.Program` instead of a
``qiskit.QuantumCircuit`` (``get_program(num_features)`` replaces ``get_circuit(features,
parameters)``; the symbolic angle expressions are kept in closed form, see program.py).

  ChebyshevPQC                src/squlearn/encoding_circuit/circuit_library/chebyshev_pqc.py
  HubregtsenEncodingCircuit   .../hubregtsen_encoding_circuit.py
  ChebyshevRx                 .../chebyshev_rx.py
  HighDimEncodingCircuit      .../highdim_encoding_circuit.py   (cx entangler)
  LayeredEncodingCircuit      src/squlearn/encoding_circuit/layered_encoding_circuit.py (core layer ops)
  base behaviour              src/squlearn/encoding_circuit/encoding_circuit_base.py:74-92 (init),
                              :94-... (num_features handling)
"""

from typing import Optional, Union

import numpy as np

from .program import Angle, GateOp, Program


class EncodingCircuitBase:
    """Common surface of the reference's EncodingCircuitBase that the hot path touches."""

    def __init__(self, num_qubits: int, num_features: Optional[int] = None):
        self._num_qubits = int(num_qubits)
        self._num_features = num_features

    # --- properties -----------------------------------------------------------------
    @property
    def num_qubits(self) -> int:
        return self._num_qubits

    @property
    def num_features(self) -> Optional[int]:
        return self._num_features

    @property
    def num_parameters(self) -> int:
        raise NotImplementedError

    @property
    def parameter_bounds(self) -> np.ndarray:
        return np.array([[-np.pi, np.pi]] * self.num_parameters).reshape(-1, 2)

    @property
    def feature_bounds(self) -> np.ndarray:
        return np.array([-np.pi, np.pi])

    @property
    def num_encoding_slots(self):
        raise NotImplementedError

    def generate_initial_parameters(self, num_features: int, seed: Union[int, None] = None):
        """encoding_circuit_base.py:74-92: RandomState(seed).uniform over parameter_bounds."""
        if self.num_parameters == 0:
            return np.array([])
        r = np.random.RandomState(seed)
        bounds = self.parameter_bounds
        return r.uniform(low=bounds[:, 0], high=bounds[:, 1])

    def _check_feature_encoding_slots(self, num_features: int):
        slots = self.num_encoding_slots
        if num_features > slots:
            raise ValueError(
                f"The number of features ({num_features}) exceeds the number of encoding slots "
                f"({slots}) of the encoding circuit.")

    def get_program(self, num_features: int) -> Program:
        raise NotImplementedError

    def get_params(self, deep: bool = True) -> dict:
        return {"num_qubits": self._num_qubits, "num_features": self._num_features}

    def set_params(self, **params):
        valid = self.get_params()
        for key, value in params.items():
            if key not in valid:
                raise ValueError(
                    f"Invalid parameter {key!r}. Valid parameters are {sorted(valid)!r}.")
            setattr(self, "_" + key, value)
        return self


def _check_nonlinearity(name):
    if name not in ("arccos", "arctan"):
        raise ValueError(
            f"Unknown value for nonlinearity: {name}. Possible values are 'arccos' and 'arctan'")


class ChebyshevPQC(EncodingCircuitBase):
    """chebyshev_pqc.py: ry basis change, layers of rx(p * arccos x) + crz/rzz ring, ry."""

    def __init__(self, num_qubits, num_layers=1, num_features=None, closed=True,
                 entangling_gate="crz", alpha=4.0, nonlinearity="arccos"):
        super().__init__(num_qubits, num_features)
        self._num_layers = num_layers
        self._closed = closed
        self._entangling_gate = entangling_gate
        self._alpha = alpha
        self._nonlinearity = nonlinearity
        if entangling_gate not in ("crz", "rzz"):
            raise ValueError("Unknown value for entangling_gate: ", entangling_gate)
        _check_nonlinearity(nonlinearity)

    num_layers = property(lambda self: self._num_layers)
    closed = property(lambda self: self._closed)
    entangling_gate = property(lambda self: self._entangling_gate)
    alpha = property(lambda self: self._alpha)
    nonlinearity = property(lambda self: self._nonlinearity)

    @property
    def num_parameters(self) -> int:
        n, L = self.num_qubits, self._num_layers
        num = 2 * n + n * L
        num += n * L if (n > 2 and self._closed) else (n - 1) * L
        return num

    def _ent_pairs(self):
        n, c = self.num_qubits, int(bool(self._closed))
        first = list(range(0, n + c - 1, 2))
        second = list(range(1, n + c - 1, 2)) if n > 2 else []
        return first, second

    @property
    def parameter_bounds(self) -> np.ndarray:
        n = self.num_qubits
        first, second = self._ent_pairs()
        b = [[-np.pi, np.pi]] * n
        for _ in range(self._num_layers):
            b += [[0.0, self._alpha]] * n
            b += [[-2.0 * np.pi, 2.0 * np.pi]] * (len(first) + len(second))
        b += [[-np.pi, np.pi]] * n
        return np.array(b, dtype=np.float64).reshape(-1, 2)

    @property
    def feature_bounds(self) -> np.ndarray:
        return np.array([-1.0, 1.0]) if self._nonlinearity == "arccos" else np.array([-np.inf, np.inf])

    @property
    def num_encoding_slots(self) -> int:
        return self.num_qubits * self._num_layers

    def get_cheb_indices(self, flatten: bool = True):
        n = self.num_qubits
        first, second = self._ent_pairs()
        out, off = [], n
        for _ in range(self._num_layers):
            layer = list(range(off, off + n))
            off += n + len(first) + len(second)
            out = out + layer if flatten else out + [layer]
        return out

    def generate_initial_parameters(self, num_features, seed=None):
        param = super().generate_initial_parameters(num_features, seed)
        if len(param) > 0:
            fpq = int(np.ceil(self.num_qubits / num_features))
            lin = np.linspace(0.01, self._alpha, fpq)
            for layer in self.get_cheb_indices(False):
                for i, ii in enumerate(layer):
                    param[ii] = lin[i % fpq]
        return param

    def get_program(self, num_features: int) -> Program:
        self._check_feature_encoding_slots(num_features)
        n, npar, d = self.num_qubits, self.num_parameters, num_features
        first, second = self._ent_pairs()
        gates, k, f = [], 0, 0
        for i in range(n):
            gates.append(GateOp("ry", (i,), Angle(param=k % npar)))
            k += 1
        for _ in range(self._num_layers):
            for i in range(n):
                gates.append(GateOp("rx", (i,), Angle(param=k % npar, feature=f % d,
                                                      fn=self._nonlinearity)))
                k += 1
                f += 1
            for i in first + second:
                gates.append(GateOp(self._entangling_gate, (i, (i + 1) % n), Angle(param=k % npar)))
                k += 1
        for i in range(n):
            gates.append(GateOp("ry", (i,), Angle(param=k % npar)))
            k += 1
        return Program(n, gates, npar, d)

    def get_params(self, deep=True):
        p = super().get_params()
        p.update(num_layers=self._num_layers, closed=self._closed,
                 entangling_gate=self._entangling_gate, alpha=self._alpha,
                 nonlinearity=self._nonlinearity)
        return p

    def set_params(self, **kwargs):
        if "nonlinearity" in kwargs:
            _check_nonlinearity(kwargs["nonlinearity"])
        return super().set_params(**kwargs)


class HubregtsenEncodingCircuit(EncodingCircuitBase):
    """hubregtsen_encoding_circuit.py: H, alternating rz(x)/rx(x) encodings, ry(p), crz ring."""

    def __init__(self, num_qubits, num_layers=1, num_features=None, closed=True,
                 final_encoding=False):
        super().__init__(num_qubits, num_features)
        self._num_layers = num_layers
        self._closed = closed
        self._final_encoding = final_encoding

    num_layers = property(lambda self: self._num_layers)
    closed = property(lambda self: self._closed)
    final_encoding = property(lambda self: self._final_encoding)

    def _n_ent(self):
        n = self.num_qubits
        return (n if self._closed else n - 1) if n > 2 else 0

    @property
    def num_parameters(self) -> int:
        return (self.num_qubits + self._n_ent()) * self._num_layers

    @property
    def parameter_bounds(self) -> np.ndarray:
        b = []
        for _ in range(self._num_layers):
            b += [[-np.pi, np.pi]] * self.num_qubits
            b += [[-2.0 * np.pi, 2.0 * np.pi]] * self._n_ent()
        return np.array(b, dtype=np.float64).reshape(-1, 2)

    @property
    def num_encoding_slots(self) -> int:
        return self.num_qubits * self._num_layers

    def get_program(self, num_features: int) -> Program:
        self._check_feature_encoding_slots(num_features)
        n, npar, d = self.num_qubits, self.num_parameters, num_features
        gates, k = [GateOp("h", (i,)) for i in range(n)], 0
        loops = int(np.ceil(d / n))
        for _ in range(self._num_layers):
            for i in range(loops * n):
                name = "rz" if (i // n) % 2 == 0 else "rx"
                gates.append(GateOp(name, (i % n,), Angle(feature=i % d)))
            for i in range(n):
                gates.append(GateOp("ry", (i,), Angle(param=k % npar)))
                k += 1
            for i in range(self._n_ent()):
                gates.append(GateOp("crz", (i, (i + 1) % n), Angle(param=k % npar)))
                k += 1
        if self._final_encoding:
            for i in range(loops * n):
                # the reference uses ceil(i / n) here but i // n above (:182-188); kept as is
                name = "rz" if int(np.ceil(i / n)) % 2 == 0 else "rx"
                gates.append(GateOp(name, (i % n,), Angle(feature=i % d)))
        return Program(n, gates, npar, d)

    def get_params(self, deep=True):
        p = super().get_params()
        p.update(num_layers=self._num_layers, closed=self._closed,
                 final_encoding=self._final_encoding)
        return p


class ChebyshevRx(EncodingCircuitBase):
    """chebyshev_rx.py: layers of rx(p * arccos x), rx(p), cx ladder (open by default)."""

    def __init__(self, num_qubits, num_layers=1, num_features=None, closed=False, alpha=4.0,
                 nonlinearity="arccos"):
        super().__init__(num_qubits, num_features)
        self._num_layers = num_layers
        self._closed = closed
        self._alpha = alpha
        self._nonlinearity = nonlinearity
        _check_nonlinearity(nonlinearity)

    num_layers = property(lambda self: self._num_layers)
    closed = property(lambda self: self._closed)
    alpha = property(lambda self: self._alpha)
    nonlinearity = property(lambda self: self._nonlinearity)

    @property
    def num_parameters(self) -> int:
        return 2 * self.num_qubits * self._num_layers

    @property
    def parameter_bounds(self) -> np.ndarray:
        b = []
        for _ in range(self._num_layers):
            b += [[0.0, self._alpha]] * self.num_qubits
            b += [[-np.pi, np.pi]] * self.num_qubits
        return np.array(b, dtype=np.float64).reshape(-1, 2)

    @property
    def feature_bounds(self) -> np.ndarray:
        return np.array([-1.0, 1.0]) if self._nonlinearity == "arccos" else np.array([-np.inf, np.inf])

    @property
    def num_encoding_slots(self) -> int:
        return self.num_qubits * self._num_layers

    def get_cheb_indices(self, flatten: bool = True):
        n = self.num_qubits
        out = []
        for layer in range(self._num_layers):
            idx = list(range(2 * n * layer, 2 * n * layer + n))
            out = out + idx if flatten else out + [idx]
        return out

    def generate_initial_parameters(self, num_features, seed=None):
        param = super().generate_initial_parameters(num_features, seed)
        if len(param) > 0:
            fpq = int(np.ceil(self.num_qubits / num_features))
            lin = np.linspace(0.01, self._alpha, fpq)
            for layer in self.get_cheb_indices(False):
                for i, ii in enumerate(layer):
                    param[ii] = lin[i % fpq]
        return param

    def get_program(self, num_features: int) -> Program:
        self._check_feature_encoding_slots(num_features)
        n, npar, d = self.num_qubits, self.num_parameters, num_features
        c = int(bool(self._closed))
        gates, k, f = [], 0, 0
        for _ in range(self._num_layers):
            for i in range(n):
                gates.append(GateOp("rx", (i,), Angle(param=k % npar, feature=f % d,
                                                      fn=self._nonlinearity)))
                k += 1
                f += 1
            for i in range(n):
                gates.append(GateOp("rx", (i,), Angle(param=k % npar)))
                k += 1
            for i in range(0, n + c - 1, 2):
                gates.append(GateOp("cx", (i, (i + 1) % n)))
            if n > 2:
                for i in range(1, n + c - 1, 2):
                    gates.append(GateOp("cx", (i, (i + 1) % n)))
        return Program(n, gates, npar, d)

    def get_params(self, deep=True):
        p = super().get_params()
        p.update(num_layers=self._num_layers, closed=self._closed, alpha=self._alpha,
                 nonlinearity=self._nonlinearity)
        return p

    def set_params(self, **kwargs):
        if "nonlinearity" in kwargs:
            _check_nonlinearity(kwargs["nonlinearity"])
        return super().set_params(**kwargs)


class HighDimEncodingCircuit(EncodingCircuitBase):
    """highdim_encoding_circuit.py with the cx entangler (the iswap variant needs a custom
    sqrt-iSWAP unitary and is not on the hot path)."""

    def __init__(self, num_qubits, num_features=None, cycling=True, cycling_type="saw",
                 num_layers=None, layer_type="rows", entangling_gate="iswap"):
        super().__init__(num_qubits, num_features)
        if cycling_type not in ("saw", "hat"):
            raise ValueError("Unknown cycling type:", cycling_type)
        if layer_type not in ("columns", "rows"):
            raise ValueError("Unknown layer type:", layer_type)
        if entangling_gate not in ("cx", "iswap"):
            raise ValueError("Unknown entangling gate:", entangling_gate)
        self._cycling = cycling
        self._cycling_type = cycling_type
        self._num_layers = num_layers
        self._layer_type = layer_type
        self._entangling_gate = entangling_gate

    @property
    def num_parameters(self) -> int:
        return 0

    @property
    def num_encoding_slots(self):
        return 3 * self.num_qubits * self._num_layers if self._num_layers is not None else np.inf

    def get_program(self, num_features: int) -> Program:
        n, d = self.num_qubits, num_features
        layers = self._num_layers
        if layers is None:
            layers = max(int(d / (n * 3)), 2)
        if layers * n * 3 < d:
            raise RuntimeError("Not all features are represented in the encoding circuit!")
        rows = self._layer_type == "rows"
        gates = [GateOp("h", (i,)) for i in range(n)]
        off = 0
        for layer in range(layers):
            if layer != 0:
                if self._entangling_gate != "cx":
                    raise NotImplementedError("sqrt-iSWAP entangler is not supported by the b200 engine")
                for i in range(0, n - 1, 2):
                    gates.append(GateOp("cx", (i, i + 1)))
                for i in range(1, n - 1, 2):
                    gates.append(GateOp("cx", (i, i + 1)))
            for i in range(3 * n):
                iq = int(i / 3) if rows else i % n
                ii = off + i
                if self._cycling:
                    if self._cycling_type == "saw":
                        ii = ii % d
                    else:
                        itest = ii % max(d + d - 2, 1)
                        ii = d + d - 2 - itest if itest >= d else itest
                if iq >= n or ii >= d:
                    break
                if rows:
                    name = "ry" if i % 3 == 1 else "rz"
                else:
                    name = "ry" if int(i / n) == 1 else "rz"
                gates.append(GateOp(name, (iq,), Angle(feature=ii)))
            off += n * 3
            if self._cycling is False and off >= d:
                off = 0
        return Program(n, gates, 0, d)

    def get_params(self, deep=True):
        p = super().get_params()
        p.update(cycling=self._cycling, cycling_type=self._cycling_type,
                 num_layers=self._num_layers, layer_type=self._layer_type,
                 entangling_gate=self._entangling_gate)
        return p


class LayeredEncodingCircuit(EncodingCircuitBase):
    """The layer-builder core of layered_encoding_circuit.py: single-qubit layers applied to every
    qubit (H, X, Y, Z, S, T, Rx/Ry/Rz/P with a "x" or "p" variable, optional encoding function)
    and nearest-neighbour / all-to-all two-qubit layers.  Each variable group keeps a running
    index over the whole circuit: features wrap modulo d, each parameter use is fresh
    (layered_encoding_circuit.py:20-76,238-257)."""

    _FIXED = ("h", "x", "y", "z", "s", "sdg", "t", "tdg")
    _ROT = ("rx", "ry", "rz", "p")
    _ENT_FIXED = ("cx", "cy", "cz", "ch", "swap")
    _ENT_ROT = ("cp", "crx", "cry", "crz", "rxx", "ryy", "rzz", "rzx")

    def __init__(self, num_qubits, num_features=None):
        super().__init__(num_qubits, num_features)
        self._ops = []  # (kind, name, variable, encoding fn, ent_strategy)

    def _single(self, name, var=None, encoding=None):
        name = name.lower()
        if var is None:
            if name not in self._FIXED:
                raise ValueError(f"gate {name} needs a variable ('x' or 'p')")
        else:
            if name not in self._ROT:
                raise ValueError(f"gate {name} takes no variable")
            if var not in ("x", "p"):
                raise ValueError("variable must be 'x' or 'p'")
        if encoding not in (None, "arccos", "arctan"):
            raise ValueError("encoding must be None, 'arccos' or 'arctan'")
        self._ops.append(("single", name, var, encoding, None))
        return self

    def H(self): return self._single("h")
    def X(self): return self._single("x")
    def Y(self): return self._single("y")
    def Z(self): return self._single("z")
    def S(self): return self._single("s")
    def T(self): return self._single("t")
    def Rx(self, var, encoding=None): return self._single("rx", var, encoding)
    def Ry(self, var, encoding=None): return self._single("ry", var, encoding)
    def Rz(self, var, encoding=None): return self._single("rz", var, encoding)
    def P(self, var, encoding=None): return self._single("p", var, encoding)

    def _entangle(self, name, var=None, ent_strategy="NN"):
        name = name.lower()
        if ent_strategy not in ("NN", "AA"):
            raise ValueError("ent_strategy must be 'NN' or 'AA'")
        if var is None and name not in self._ENT_FIXED:
            raise ValueError(f"gate {name} needs a variable")
        if var is not None and name not in self._ENT_ROT:
            raise ValueError(f"gate {name} takes no variable")
        self._ops.append(("ent", name, var, None, ent_strategy))
        return self

    def cx_entangling(self, ent_strategy="NN"): return self._entangle("cx", None, ent_strategy)
    def cz_entangling(self, ent_strategy="NN"): return self._entangle("cz", None, ent_strategy)
    def crz_entangling(self, var, ent_strategy="NN"): return self._entangle("crz", var, ent_strategy)
    def crx_entangling(self, var, ent_strategy="NN"): return self._entangle("crx", var, ent_strategy)
    def cry_entangling(self, var, ent_strategy="NN"): return self._entangle("cry", var, ent_strategy)
    def rzz_entangling(self, var, ent_strategy="NN"): return self._entangle("rzz", var, ent_strategy)

    def _pairs(self, strategy):
        n = self.num_qubits
        if strategy == "NN":
            return [(i, i + 1) for i in range(0, n - 1, 2)] + [(i, i + 1) for i in range(1, n - 1, 2)]
        return [(i, j) for i in range(n) for j in range(i + 1, n)]

    @property
    def num_parameters(self) -> int:
        num = 0
        for kind, name, var, _, strat in self._ops:
            if var == "p":
                num += self.num_qubits if kind == "single" else len(self._pairs(strat))
        return num

    @property
    def num_encoding_slots(self):
        num = 0
        for kind, name, var, _, strat in self._ops:
            if var == "x":
                num += self.num_qubits if kind == "single" else len(self._pairs(strat))
        return num if num else np.inf

    def get_program(self, num_features: int) -> Program:
        n, d = self.num_qubits, num_features
        gates, ip, ix = [], 0, 0

        def angle(var, enc):
            nonlocal ip, ix
            if var == "p":
                a = Angle(param=ip)
                ip += 1
            else:
                a = Angle(feature=ix % d, fn=enc or "linear")
                ix += 1
            return a

        for kind, name, var, enc, strat in self._ops:
            if kind == "single":
                for q in range(n):
                    gates.append(GateOp(name, (q,), angle(var, enc) if var else None))
            else:
                for (a, b) in self._pairs(strat):
                    gates.append(GateOp(name, (a, b), angle(var, enc) if var else None))
        return Program(n, gates, self.num_parameters, d)
