"""Gate-list IR of an encoding circuit (host side).

Mirrors what ``PennyLaneCircuit.build_circuit_instructions`` extracts from a Qiskit circuit in the
reference (src/squlearn/util/pennylane/pennylane_circuit.py:249-415): parallel lists of gate
names, wires and angle functions.  Because Qiskit / sympy expressions are not available to the
engine, an angle is kept in the closed form every library circuit uses,

    theta(x, p) = offset + scale * p[param] * f(x[feature]),        f in {1, x, arccos x, arctan x}

(``p[param]`` or ``f(x[feature])`` may be absent), or as an arbitrary host callable that fills a
row of the device angle table.
"""

from dataclasses import dataclass
from typing import Callable, Optional, Sequence, Tuple

import ctypes
import numpy as np

from . import _lib

_FN_CODE = {None: _lib.FN_CONST, "linear": _lib.FN_LINEAR, "arccos": _lib.FN_ARCCOS,
            "arctan": _lib.FN_ARCTAN}
_FN_NUMPY = {"linear": lambda v: v, "arccos": np.arccos, "arctan": np.arctan}


@dataclass(frozen=True)
class Angle:
    """theta = offset + scale * p[param] * fn(x[feature]); missing factors are 1."""

    param: Optional[int] = None
    feature: Optional[int] = None
    fn: Optional[str] = None          # "linear" | "arccos" | "arctan" (None when feature is None)
    scale: float = 1.0
    offset: float = 0.0
    table: Optional[Callable] = None  # general case: callable(x (N,d), p) -> (N,) angles

    def __post_init__(self):
        if self.feature is not None and self.fn is None:
            object.__setattr__(self, "fn", "linear")
        if self.feature is None and self.fn is not None:
            raise ValueError("angle function given without a feature index")
        if self.fn not in _FN_CODE:
            raise ValueError(f"Unknown angle function {self.fn}")

    # -- host evaluation (used for parameter-shift weights and the angle table)
    def feature_factor(self, x: np.ndarray) -> np.ndarray:
        """fn(x[:, feature]) as an (N,) array (ones when the angle has no feature)."""
        if self.feature is None:
            return np.ones(x.shape[0])
        return _FN_NUMPY[self.fn](x[:, self.feature])

    def dtheta_dparam(self, x: np.ndarray, j: int) -> Optional[np.ndarray]:
        """d theta / d p_j per sample, or None when the angle does not depend on p_j."""
        if self.table is not None or self.param is None or self.param != j:
            return None
        return self.scale * self.feature_factor(x)

    def dtheta_dfeature(self, x: np.ndarray, p: Optional[np.ndarray]) -> Optional[np.ndarray]:
        """d theta / d x[feature] per sample, or None when the angle does not depend on a feature
        (theta = offset + scale * p * fn(x): linear -> 1, arccos -> -1/sqrt(1 - x^2), arctan -> 1/(1 + x^2))."""
        if self.table is not None or self.feature is None:
            return None
        pv = 1.0 if self.param is None else float(p[self.param])
        v = x[:, self.feature]
        if self.fn == "linear":
            d = np.ones_like(v)
        elif self.fn == "arccos":
            d = -1.0 / np.sqrt(1.0 - v * v)
        else:
            d = 1.0 / (1.0 + v * v)
        return self.scale * pv * d

    def d2theta(self, x: np.ndarray, p: Optional[np.ndarray], wrt1: str, i1: int, wrt2: str, i2: int):
        """Second derivative of the angle with respect to (parameter / feature i1, parameter / feature i2) per sample,
        or None when it vanishes: theta = offset + scale * p * fn(x) is linear in p, so only d2/dx2 = scale p fn''(x)
        and d2/dp dx = scale fn'(x) survive (fn'': linear 0, arccos -x (1-x^2)^(-3/2), arctan -2x (1+x^2)^(-2))."""
        if self.table is not None or self.feature is None:
            return None
        v = x[:, self.feature]
        if wrt1 == "x" and wrt2 == "x":
            if i1 != self.feature or i2 != self.feature or self.fn == "linear":
                return None
            pv = 1.0 if self.param is None else float(p[self.param])
            d2 = -v / (1.0 - v * v) ** 1.5 if self.fn == "arccos" else -2.0 * v / (1.0 + v * v) ** 2
            return self.scale * pv * d2
        if wrt1 == wrt2:          # param, param
            return None
        ip, ix = (i1, i2) if wrt1 == "param" else (i2, i1)
        if self.param is None or ip != self.param or ix != self.feature:
            return None
        d1 = np.ones_like(v) if self.fn == "linear" else (-1.0 / np.sqrt(1.0 - v * v) if self.fn == "arccos"
                                                        else 1.0 / (1.0 + v * v))
        return self.scale * d1

    def coeffs(self, p: np.ndarray) -> Tuple[float, float]:
        """(c0, c1) of the C ABI: theta = c0 + c1 * f(x[feat])."""
        pv = 1.0 if self.param is None else float(p[self.param])
        if self.feature is None:
            return self.offset + self.scale * pv, 0.0
        return self.offset, self.scale * pv


@dataclass(frozen=True)
class GateOp:
    name: str
    qubits: Tuple[int, ...]
    angle: Optional[Angle] = None
    extra: Tuple[float, float] = (0.0, 0.0)  # (phi, lambda) of the u gate


class Program:
    """An n-qubit gate list plus the sizes of its parameter and feature vectors."""

    def __init__(self, num_qubits: int, gates: Sequence[GateOp], num_parameters: int,
                 num_features: int):
        if num_qubits < 1 or num_qubits > 26:
            raise ValueError("num_qubits must be in 1..26")
        self.num_qubits = int(num_qubits)
        self.num_parameters = int(num_parameters)
        self.num_features = int(num_features)
        self.gates = []
        for g in gates:
            name = g.name.lower()
            if name in ("measure", "reset", "if_else"):
                # mid-circuit measurement / classical control is outside a pure statevector batch
                # engine, as in the Qulacs adapter (qulacs_circuit.py:323-329)
                raise NotImplementedError(f"gate {name} is not supported by the b200 engine")
            if name not in _lib.GATE_KIND:
                raise NotImplementedError(f"gate {name} is not in the gate table")
            nq, has_angle = _lib.GATE_ARITY[name]
            if len(g.qubits) != nq:
                raise ValueError(f"gate {name} takes {nq} qubit(s)")
            if len(set(g.qubits)) != nq or min(g.qubits) < 0 or max(g.qubits) >= num_qubits:
                raise ValueError(f"bad qubits {g.qubits} for gate {name}")
            if has_angle and g.angle is None:
                raise ValueError(f"gate {name} needs an angle")
            if g.angle is not None:
                if g.angle.param is not None and not 0 <= g.angle.param < num_parameters:
                    raise ValueError("angle refers to a parameter outside the parameter vector")
                if g.angle.feature is not None and not 0 <= g.angle.feature < num_features:
                    raise ValueError("angle refers to a feature outside the feature vector")
            self.gates.append(GateOp(name, tuple(int(q) for q in g.qubits), g.angle, g.extra))
        self.table_gates = [i for i, g in enumerate(self.gates)
                            if g.angle is not None and g.angle.table is not None]

    def __len__(self):
        return len(self.gates)

    # ------------------------------------------------------------------ C ABI views
    def coefficient_arrays(self, p: Optional[np.ndarray]):
        p = np.zeros(0) if p is None else np.asarray(p, dtype=np.float64).reshape(-1)
        if len(p) != self.num_parameters:
            raise ValueError(
                f"expected {self.num_parameters} circuit parameters, got {len(p)}")
        c0 = np.zeros(len(self.gates))
        c1 = np.zeros(len(self.gates))
        for i, g in enumerate(self.gates):
            if g.angle is not None and g.angle.table is None:
                c0[i], c1[i] = g.angle.coeffs(p)
        return c0, c1

    def c_gates(self, p: Optional[np.ndarray]):
        """ctypes array of b200sq_gate for the parameter vector ``p``."""
        c0, c1 = self.coefficient_arrays(p)
        arr = (_lib.Gate * max(len(self.gates), 1))()
        slot = 0
        for i, g in enumerate(self.gates):
            cg = arr[i]
            cg.kind = _lib.GATE_KIND[g.name]
            for k in range(3):
                cg.q[k] = g.qubits[k] if k < len(g.qubits) else -1
            cg.fn, cg.feat, cg.slot = _lib.FN_CONST, 0, 0
            if g.angle is not None:
                if g.angle.table is not None:
                    cg.fn, cg.slot = _lib.FN_TABLE, slot
                    slot += 1
                elif g.angle.feature is not None:
                    cg.fn, cg.feat = _FN_CODE[g.angle.fn], g.angle.feature
            cg.c0, cg.c1 = c0[i], c1[i]
            cg.c2, cg.c3 = g.extra
        return arr

    def angle_table(self, x: np.ndarray, p: Optional[np.ndarray]) -> Optional[np.ndarray]:
        """[n_slots][N] float64 host table for the gates with general (callable) angles."""
        if not self.table_gates:
            return None
        rows = [np.broadcast_to(np.asarray(self.gates[i].angle.table(x, p), dtype=np.float64),
                                (x.shape[0],)) for i in self.table_gates]
        return np.ascontiguousarray(np.stack(rows, axis=0))

    # ------------------------------------------------------------------ parameter shift
    def shift_variants(self, x: np.ndarray, wrt: str = "param", p: Optional[np.ndarray] = None):
        """The shifted-circuit batch of src/squlearn/util/optree/optree_derivative.py:130-158.

        ``wrt="param"``: for every occurrence of parameter p_j in a gate angle theta_k (linear in p_j): circuits
        with theta_k +- pi/2 weighted +-1/2 dtheta_k/dp_j (Pauli rotations, p, cp), or the four-term rule for
        controlled rotations (which the reference reaches by transpiling crx/cry/crz to single-qubit rotations
        first, :352-370).  ``wrt="x"``: the same shifted circuits for every gate whose angle depends on a
        feature x_j, weighted with dtheta_k/dx_j (chain rule through the encoding, needs the parameters ``p``).
        Returns (gate_index[V], shift[V], weight[V, N], index[V]) with index = parameter / feature index.
        """
        two_term = {"rx", "ry", "rz", "p", "cp", "rxx", "ryy", "rzz", "rzx"}
        four_term = {"crx", "cry", "crz"}
        cp_ = (np.sqrt(2.0) + 1.0) / (4.0 * np.sqrt(2.0))
        cm_ = (np.sqrt(2.0) - 1.0) / (4.0 * np.sqrt(2.0))
        if wrt not in ("param", "x"):
            raise ValueError("wrt must be 'param' or 'x'")
        gate_idx, shifts, weights, params = [], [], [], []
        for k, g in enumerate(self.gates):
            a = g.angle
            if a is None:
                continue
            if wrt == "param":
                if a.param is None:
                    continue
                if a.table is not None:
                    raise NotImplementedError("parameter shift through a general angle callable")
                fac, index = a.dtheta_dparam(x, a.param), a.param
            else:
                if a.table is not None:
                    raise NotImplementedError("feature derivative through a general angle callable")
                if a.feature is None:
                    continue
                fac, index = a.dtheta_dfeature(x, p), a.feature
            if g.name in two_term:
                rule = [(0.5 * np.pi, 0.5), (-0.5 * np.pi, -0.5)]
            elif g.name in four_term:
                rule = [(0.5 * np.pi, cp_), (-0.5 * np.pi, -cp_), (1.5 * np.pi, -cm_),
                        (-1.5 * np.pi, cm_)]
            else:
                raise NotImplementedError(f"no parameter-shift rule for gate {g.name}")
            for s, w in rule:
                gate_idx.append(k)
                shifts.append(s)
                weights.append(w * fac)
                params.append(index)
        if not gate_idx:
            return (np.zeros(0, dtype=np.int32), np.zeros(0), np.zeros((0, x.shape[0])),
                    np.zeros(0, dtype=np.int64))
        return (np.asarray(gate_idx, dtype=np.int32), np.asarray(shifts, dtype=np.float64),
                np.stack(weights, axis=0), np.asarray(params, dtype=np.int64))


def _shift_rule(name: str):
    """[(shift, coefficient)] of the parameter-shift rule of a gate (optree_derivative.py:130-158, :352-370)."""
    if name in ("rx", "ry", "rz", "p", "cp", "rxx", "ryy", "rzz", "rzx"):
        return [(0.5 * np.pi, 0.5), (-0.5 * np.pi, -0.5)]
    if name in ("crx", "cry", "crz"):
        cp_ = (np.sqrt(2.0) + 1.0) / (4.0 * np.sqrt(2.0))
        cm_ = (np.sqrt(2.0) - 1.0) / (4.0 * np.sqrt(2.0))
        return [(0.5 * np.pi, cp_), (-0.5 * np.pi, -cp_), (1.5 * np.pi, -cm_), (-1.5 * np.pi, cm_)]
    raise NotImplementedError(f"no parameter-shift rule for gate {name}")


def second_order_variants(program: "Program", x: np.ndarray, wrt1: str, wrt2: str, p: Optional[np.ndarray]):
    """The doubly shifted circuits of d2 f / d a_i d b_j (a, b in {"param", "x"}):

        d2f/da_i db_j = sum_k sum_l (dtheta_k/da_i)(dtheta_l/db_j) d2f/dtheta_k dtheta_l  +  sum_k (d2theta_k/da_i db_j) df/dtheta_k

    with d2f/dtheta_k dtheta_l from the shift rule applied to both gates (two shifts per variant; the same gate twice:
    the shifts add) and the second term (curvature of the encoding, e.g. arccos) from single shifts.  This is what the
    reference's nested jacobians evaluate for "dfdxdx", "dfdpdx", "dfdpdp", ... (evaluation_classes.py:105-164).
    Returns (gate1[V], shift1[V], gate2[V], shift2[V], weight[V, N], index1[V], index2[V]); gate2 = -1: single shift."""
    g1, s1, w1, i1 = program.shift_variants(x, wrt=wrt1, p=p)
    g2, s2, w2, i2 = program.shift_variants(x, wrt=wrt2, p=p)
    n = x.shape[0]
    ga, sa, gb, sb, ww, ia, ib = [], [], [], [], [], [], []
    for a in range(len(g1)):
        for b in range(len(g2)):
            ga.append(g1[a]); sa.append(s1[a]); gb.append(g2[b]); sb.append(s2[b])
            ww.append(w1[a] * w2[b]); ia.append(i1[a]); ib.append(i2[b])
    n1 = program.num_parameters if wrt1 == "param" else program.num_features
    n2 = program.num_parameters if wrt2 == "param" else program.num_features
    for k, g in enumerate(program.gates):
        if g.angle is None or g.angle.feature is None:
            continue
        for j1 in range(n1):
            for j2 in range(n2):
                d2 = g.angle.d2theta(x, p, wrt1, j1, wrt2, j2)
                if d2 is None:
                    continue
                for s, c in _shift_rule(g.name):
                    ga.append(k); sa.append(s); gb.append(-1); sb.append(0.0)
                    ww.append(c * d2); ia.append(j1); ib.append(j2)
    if not ga:
        z = np.zeros(0)
        return (z.astype(np.int32), z, z.astype(np.int32), z, np.zeros((0, n)), z.astype(np.int64), z.astype(np.int64))
    return (np.asarray(ga, dtype=np.int32), np.asarray(sa, dtype=np.float64), np.asarray(gb, dtype=np.int32),
            np.asarray(sb, dtype=np.float64), np.stack([np.broadcast_to(w, (n,)) for w in ww], axis=0),
            np.asarray(ia, dtype=np.int64), np.asarray(ib, dtype=np.int64))


def as_uint32_array(values):
    arr = (ctypes.c_uint32 * max(len(values), 1))()
    for i, v in enumerate(values):
        arr[i] = int(v)
    return arr
