"""LowLevelQNN for the b200 backend.

Mirrors the evaluate() contract of the reference's low-level QNNs
(src/squlearn/qnn/lowlevel_qnn/lowlevel_qnn_base.py:15-194; PennyLane implementation
lowlevel_qnn_pennylane.py:190-370; Qulacs implementation lowlevel_qnn_qulacs.py:379-476):

    qnn.evaluate(x, param, param_obs, "f", "dfdp", "dfdop", "fcc", "var", "dvardp", "dvardop")

returns a dict with the requested keys plus "x", "param", "param_op".  Axis order of every array:
[samples (if multi x)] [param sets (if param is 2-D)] [param_obs sets (if 2-D)]
[outputs (if the observable is a list)] [derivative axes]  (lowlevel_qnn_pennylane.py:331-359).

All x are evaluated in one engine call; "dfdp" uses the parameter-shift batch
(optree_derivative.py:130-158), which the reference asserts equal to its autodiff routes
(tests/qnn/lowlevel_qnn/test_lowlevel_qnn.py:36-65).  First-order input derivatives ("dfdx") use the same shifted
batch with chain-rule weights; second-order keys ("dfdxdx", "dfdpdx", "dfdxdp", "dfdpdp", "dfdopdx", "dfdopdxdx",
"laplace", "laplace_dop", ...) are assembled from doubly shifted circuits; third-order keys
raise NotImplementedError.
"""

from typing import Sequence, Union

import numpy as np

from .data_preprocessing import adjust_features, adjust_parameters
from .executor import (
    B200Circuit,
    Executor,
    b200_evaluate,
    b200_gradient,
    b200_hessian,
    b200_operator_gradient,
    b200_term_gradient,
)
from .observables import ObservableBase

_DIRECT = {"f", "fcc", "dfdp", "dfccdp", "dfdop", "dfccdop", "dfdx", "dfccdx",
           # second order: doubly shifted circuits (evaluation_classes.py:105-164); axes follow the key, e.g.
           # "dfdpdx" -> (..., P, d) (the reference reorders its nested jacobians the same way,
           # lowlevel_qnn_pennylane.py:331-338)
           "dfdxdx", "dfdpdp", "dfdpdx", "dfdxdp", "dfccdxdx", "dfccdpdp",
           "dfdopdx", "dfdopdp", "dfdpdop", "dfdopdop", "dfdopdxdx", "dfccdopdx", "dfccdopdop"}
_SECOND = {"dfdxdx": ("x", "x"), "dfdpdp": ("param", "param"), "dfdpdx": ("param", "x"), "dfdxdp": ("x", "param"),
           "dfccdxdx": ("x", "x"), "dfccdpdp": ("param", "param")}
_POST = {
    "var": ("f", "fcc"), "varf": ("f", "fcc"),
    "dvardp": ("f", "dfccdp", "dfdp"), "dvarfdp": ("f", "dfccdp", "dfdp"),
    "dvardop": ("f", "dfccdop", "dfdop"), "dvarfdop": ("f", "dfccdop", "dfdop"),
    "dvardx": ("f", "dfccdx", "dfdx"), "dvarfdx": ("f", "dfccdx", "dfdx"),
    "laplace": ("dfdxdx",), "laplace_dop": ("dfdopdxdx",),
}
# third order in the circuit angles (and the Fisher matrix): not offered
_NOT_YET = {"laplace_dp", "dfdxdxdp", "dfdxdpdx", "dfdpdxdx", "fischer"}


class LowLevelQNN:
    """LowLevelQNN(parameterized_quantum_circuit, observable | [observables], executor, num_features)

    ``parameterized_quantum_circuit`` is one of the builders in :mod:`squlearn_b200.circuits`.
    """

    def __init__(self, parameterized_quantum_circuit,
                 observable: Union[ObservableBase, Sequence[ObservableBase]], executor: Executor,
                 num_features: int, post_processing=None, caching: bool = True, primitive=None):
        if executor.quantum_framework != "b200":
            raise RuntimeError("squlearn_b200.LowLevelQNN needs Executor('b200')")
        self._pqc = parameterized_quantum_circuit
        self._observable = observable
        self._executor = executor
        self._num_features = int(num_features)
        self.caching = caching
        self.result_container = {}
        self.multiple_output = isinstance(observable, (list, tuple))
        self._program = self._pqc.get_program(self._num_features)
        self._circuit = B200Circuit(self._program, observable)
        self._num_parameters_observable = self._circuit.num_parameters_observable

    # ---------------------------------------------------------------- properties
    @property
    def num_qubits(self) -> int:
        return self._pqc.num_qubits

    @property
    def num_features(self) -> int:
        return self._num_features

    @property
    def num_parameters(self) -> int:
        return self._pqc.num_parameters

    @property
    def num_parameters_observable(self) -> int:
        return self._num_parameters_observable

    @property
    def num_operator(self) -> int:
        return len(self._circuit.observables)

    def get_params(self, deep: bool = True) -> dict:
        params = {"num_qubits": self.num_qubits}
        if deep:
            params.update(self._pqc.get_params())
        return params

    # ---------------------------------------------------------------- evaluation
    def __call__(self, x, param, param_obs):
        return self.evaluate(x, param, param_obs, "f")["f"]

    def gradient(self, x, param, param_obs):
        """Returns (dfdp, dfdop) — lowlevel_qnn_base.py:151-174."""
        r = self.evaluate(x, param, param_obs, "dfdp", "dfdop")
        return r["dfdp"], r["dfdop"]

    def _direct(self, key, x_inp, p, pobs):
        """(N, n_obs[, D]) array for one (param, param_obs) combination."""
        ex = self._executor
        sq = key in ("fcc", "dfccdp", "dfccdop", "dfccdx")
        kw = dict(param=p, x=x_inp, param_obs=pobs)
        if key in ("f", "fcc"):
            return ex.b200_execute(b200_evaluate, self._circuit, squared=sq, **kw)
        if key in ("dfdp", "dfccdp"):
            return ex.b200_execute(b200_gradient, self._circuit, squared=sq, **kw)
        if key in ("dfdx", "dfccdx"):
            return ex.b200_execute(b200_gradient, self._circuit, squared=sq, wrt="x", **kw)
        if key in _SECOND:
            return ex.b200_execute(b200_hessian, self._circuit, squared=sq or key.startswith("dfcc"),
                                   wrt=_SECOND[key], **kw)
        if key == "dfdopdxdx":
            return ex.b200_execute(b200_hessian, self._circuit, wrt=("x", "x"), per_term=True, **kw)
        if key in ("dfdopdx", "dfdopdp", "dfdpdop"):
            g = ex.b200_execute(b200_term_gradient, self._circuit, wrt="x" if key == "dfdopdx" else "param", **kw)
            return np.swapaxes(g, -1, -2) if key == "dfdpdop" else g
        if key == "dfdopdop":   # the observables are linear in their parameters (observable_base.py)
            npo = self._num_parameters_observable
            return np.zeros((x_inp.shape[0], self.num_operator, npo, npo))
        if key == "dfdop":
            return ex.b200_execute(b200_operator_gradient, self._circuit, **kw)
        if key == "dfccdop":
            # d<O^2>/d op_j = <O dO/dop_j + dO/dop_j O>: evaluate by central differences of the
            # exactly-squared operator (it is a quadratic polynomial in param_obs, so the
            # symmetric difference with any step is exact)
            out = np.zeros((x_inp.shape[0], self.num_operator, len(pobs)))
            for j in range(len(pobs)):
                e = np.zeros(len(pobs))
                e[j] = 1.0
                fp = ex.b200_execute(b200_evaluate, self._circuit, squared=True, param=p, x=x_inp,
                                     param_obs=pobs + e)
                fm = ex.b200_execute(b200_evaluate, self._circuit, squared=True, param=p, x=x_inp,
                                     param_obs=pobs - e)
                out[:, :, j] = 0.5 * (fp - fm)
            return out
        if key == "dfccdopdx":
            # d/dx of d<O^2>/d op_j (evaluation_classes.py:177-185): <O^2> is quadratic in param_obs, so the symmetric
            # difference of its x-gradient over param_obs is exact -> (N, n_obs, P_op, d)
            out = np.zeros((x_inp.shape[0], self.num_operator, len(pobs), self._num_features))
            for j in range(len(pobs)):
                e = np.zeros(len(pobs))
                e[j] = 1.0
                gp = ex.b200_execute(b200_gradient, self._circuit, squared=True, wrt="x", param=p, x=x_inp,
                                     param_obs=pobs + e)
                gm = ex.b200_execute(b200_gradient, self._circuit, squared=True, wrt="x", param=p, x=x_inp,
                                     param_obs=pobs - e)
                out[:, :, j, :] = 0.5 * (gp - gm)
            return out
        if key == "dfccdopdop":
            # Hessian of the quadratic <O^2> in param_obs (evaluation_classes.py:190-193): second differences with
            # unit steps are exact -> (N, n_obs, P_op, P_op)
            npo = len(pobs)
            f = lambda q: ex.b200_execute(b200_evaluate, self._circuit, squared=True, param=p, x=x_inp, param_obs=q)
            f0 = f(pobs)
            unit = np.eye(npo)
            plus = [f(pobs + unit[j]) for j in range(npo)]
            minus = [f(pobs - unit[j]) for j in range(npo)]
            out = np.zeros((x_inp.shape[0], self.num_operator, npo, npo))
            for j in range(npo):
                out[:, :, j, j] = plus[j] - 2.0 * f0 + minus[j]
                for k in range(j):
                    # f(+j+k) - f(+j) - f(+k) + f(0) is the mixed term of a quadratic exactly
                    out[:, :, j, k] = out[:, :, k, j] = f(pobs + unit[j] + unit[k]) - plus[j] - plus[k] + f0
            return out
        raise ValueError("Unknown input string:", key)

    def evaluate(self, x, param, param_obs, *values) -> dict:
        x_inp, multi_x = adjust_features(x, self._num_features)
        param_inp, multi_param = adjust_parameters(param, self.num_parameters) \
            if self.num_parameters > 0 else (np.zeros((1, 0)), False)
        pobs_inp, multi_pobs = adjust_parameters(param_obs, self._num_parameters_observable) \
            if self._num_parameters_observable > 0 else (np.zeros((1, 0)), False)

        cache_key = None
        if self.caching:
            cache_key = (np.asarray(x_inp).tobytes(), np.asarray(param_inp).tobytes(),
                         np.asarray(pobs_inp).tobytes())
            value_dict = self.result_container.get(cache_key, {})
        else:
            value_dict = {}
        value_dict["x"] = x
        value_dict["param"] = param
        value_dict["param_op"] = param_obs

        todo, post = [], []
        for v in values:
            if not isinstance(v, str):
                raise TypeError("String expected, found type:", type(v))
            if v in _NOT_YET:
                raise NotImplementedError(f"Evaluation not implemented by the b200 backend: {v}")
            if v in _POST:
                post.append(v)
                todo += [k for k in _POST[v]]
            elif v in _DIRECT:
                todo.append(v)
            else:
                raise ValueError("Unknown input string:", v)

        for key in sorted(set(todo)):
            if key in value_dict:
                continue
            blocks = []
            for p in param_inp:
                row = []
                for po in pobs_inp:
                    row.append(self._direct(key, x_inp, p, po))  # (N, n_obs[, D])
                blocks.append(row)
            arr = np.array(blocks)                      # (n_p, n_po, N, n_obs[, D])
            arr = np.moveaxis(arr, 2, 0)                # (N, n_p, n_po, n_obs[, D])
            shape = []
            if multi_x:
                shape.append(arr.shape[0])
            if multi_param:
                shape.append(arr.shape[1])
            if multi_pobs:
                shape.append(arr.shape[2])
            if self.multiple_output:
                shape.append(arr.shape[3])
            shape += list(arr.shape[4:])
            if len(shape) == 0:
                value_dict[key] = arr.reshape(-1)[0]
            else:
                value_dict[key] = arr.reshape(shape)

        for v in post:
            if v in ("var", "varf"):
                value_dict[v] = value_dict["fcc"] - np.square(value_dict["f"])
            elif v in ("dvardp", "dvarfdp"):
                f = value_dict["f"]
                value_dict[v] = value_dict["dfccdp"] - 2.0 * np.multiply(
                    value_dict["dfdp"], np.expand_dims(f, -1))
            elif v in ("dvardop", "dvarfdop"):
                f = value_dict["f"]
                value_dict[v] = value_dict["dfccdop"] - 2.0 * np.multiply(
                    value_dict["dfdop"], np.expand_dims(f, -1))
            elif v in ("dvardx", "dvarfdx"):
                f = value_dict["f"]
                value_dict[v] = value_dict["dfccdx"] - 2.0 * np.multiply(
                    value_dict["dfdx"], np.expand_dims(f, -1))
            elif v in ("laplace", "laplace_dop"):   # sum of the diagonal of the feature Hessian (:292-316)
                h = value_dict[_POST[v][0]]
                value_dict[v] = np.trace(h, axis1=-2, axis2=-1)

        if self.caching:
            self.result_container[cache_key] = value_dict
        return value_dict
