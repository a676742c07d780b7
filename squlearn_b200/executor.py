"""``Executor("b200")`` — the runtime boundary of the new backend.

Mirrors the slice of the reference's Executor that the statevector hot path touches
(src/squlearn/util/executor.py): ``quantum_framework`` (:848), ``is_statevector`` (:2742),
``shots`` / ``get_shots`` / ``set_shots`` / ``reset_shots`` (:2228-2430), ``backend`` (:1106),
``remote`` (:1119), ``execution`` (:1101) and a generic ``b200_execute(fn, circuit, **kwargs)``
dispatcher shaped like ``qulacs_execute`` (:853-906), together with batched execution functions
shaped like src/squlearn/util/qulacs/qulacs_execution.py:11-188:

    b200_evaluate_statevector(circuit, param=..., x=...)          -> (N, 2^n) complex128
    b200_evaluate(circuit, param=..., x=..., param_obs=...)       -> (N, n_obs) float64
    b200_gradient(circuit, param=..., x=..., param_obs=...)       -> (N, n_obs, P)
    b200_operator_gradient(circuit, param=..., x=..., param_obs=...) -> (N, n_obs, P_op)

Unlike the Qulacs functions these take the WHOLE batch ``x`` (N, d) in one call.  Every engine /
CUDA error surfaces as RuntimeError or ValueError — the classes the reference treats as critical
and re-raises immediately instead of retrying ten times (:1054-1099).
"""

import logging
import os
from typing import Callable, List, Optional, Sequence, Union

import numpy as np

from .engine import Engine
from .observables import ObservableBase, pauli_sum_squared
from .program import Program, second_order_variants


class Executor:
    """Executor for the B200 statevector engine.

    Args:
        execution: must be ``"b200"`` (aliases ``"b200sq"``, ``"statevector_b200"``); any other
            string raises ``ValueError("Unknown backend string")`` like the reference (:551-552).
        device: CUDA device index (default: the current device; under torchrun, LOCAL_RANK).
        shots: must be None — this is a statevector engine.
        seed, caching, log_file, max_jobs_retries, ...: accepted for signature compatibility.
    """

    def __init__(self, execution: str = "b200", *, device: Optional[int] = None, shots=None,
                 seed: Optional[int] = None, caching=None, cache_dir: str = "_cache",
                 log_file: str = "", max_jobs_retries: int = 10, wait_restart: int = 1,
                 engine=None, **kwargs):
        if not isinstance(execution, str) or execution.lower() not in (
                "b200", "b200sq", "statevector_b200"):
            raise ValueError("Unknown backend string: " + str(execution))
        self._execution = execution
        self._quantum_framework = "b200"
        self._logger = logging.getLogger("executor")
        if seed is not None and seed >= 0:
            seed = seed + 1  # executor.py:474-482 stores seed + 1
        self._seed = seed
        self._shots = None
        self._inital_num_shots = None
        self.set_shots(shots)
        self._max_jobs_retries = max_jobs_retries
        self._wait_restart = wait_restart
        self._caching = bool(caching) if caching is not None else False
        # ``engine`` lets tests inject a stand-in; the default is the CUDA engine (raises
        # without a GPU — there is no CPU path).
        self._engine = engine if engine is not None else Engine(device)

    # --------------------------------------------------------------- reference surface
    @property
    def execution(self) -> str:
        return self._execution

    @property
    def quantum_framework(self) -> str:
        return self._quantum_framework

    @property
    def backend(self):
        return self._engine

    @property
    def engine(self) -> Engine:
        return self._engine

    @property
    def remote(self) -> bool:
        return False

    @property
    def is_statevector(self) -> bool:
        return True

    @property
    def shots(self):
        return self.get_shots()

    def get_shots(self):
        return self._shots

    def set_shots(self, num_shots: Union[int, None]) -> None:
        if num_shots is not None and num_shots != 0:
            raise ValueError(
                "Executor('b200') is a statevector engine: shots must be None "
                f"(got {num_shots})")
        self._shots = None

    def reset_shots(self) -> None:
        self._shots = self._inital_num_shots

    def b200_execute(self, b200_execution: Callable, b200_circuit: "B200Circuit", **kwargs):
        """Run ``b200_execution(b200_circuit, **kwargs)`` (cf. qulacs_execute, :853-906)."""
        if not callable(b200_execution):
            raise ValueError("Unknown function specified as b200 execution")
        self._logger.info("Execution of b200 circuit")
        result = b200_execution(b200_circuit, executor=self, **kwargs)
        self._logger.info("Execution of b200 successful")
        return result


class B200Circuit:
    """A program plus (optionally) observables, compiled once and reused — the counterpart of
    QulacsCircuit (src/squlearn/util/qulacs/qulacs_circuit.py:22) for this engine."""

    def __init__(self, program: Program,
                 observable: Union[None, ObservableBase, Sequence[ObservableBase]] = None,
                 tile_bits: int = 0):
        self.program = program
        self.multiple_observables = isinstance(observable, (list, tuple))
        if observable is None:
            self.observables: List[ObservableBase] = []
        else:
            self.observables = list(observable) if self.multiple_observables else [observable]
        for o in self.observables:
            if o.num_qubits != program.num_qubits:
                raise ValueError("Number of qubits of the observable and the circuit differ")
        # B200SQ_TILE_BITS overrides the default tile size (tuning / experiments)
        self.tile_bits = tile_bits or int(os.environ.get("B200SQ_TILE_BITS", "0"))
        self._compiled = None

    @property
    def num_qubits(self) -> int:
        return self.program.num_qubits

    @property
    def num_parameters_observable(self) -> int:
        return sum(o.num_parameters for o in self.observables)

    def compiled(self, engine, param):
        if self._compiled is None:
            self._compiled = engine.compile(self.program, param, self.tile_bits)
        else:
            self._compiled.set_params(param)
        return self._compiled

    def observable_terms(self, param_obs, squared: bool = False):
        """Flatten all observables into one term list.

        Returns (strings, per-observable list of (term index, coefficient, param_obs index))."""
        param_obs = np.zeros(0) if param_obs is None else np.asarray(param_obs, dtype=np.float64)
        strings, index, per_obs, off = [], {}, [], 0
        for o in self.observables:
            npo = o.num_parameters
            s, c, idx = o.get_pauli_terms(param_obs[off:off + npo])
            if squared:
                s, c = pauli_sum_squared(s, c)
                idx = -np.ones(len(s), dtype=np.int64)
            entries = []
            for si, ci, ii in zip(s, c, idx):
                if si not in index:
                    index[si] = len(strings)
                    strings.append(si)
                entries.append((index[si], float(ci), int(off + ii) if ii >= 0 else -1))
            per_obs.append(entries)
            off += npo
        return strings, per_obs


def _combine(expv: np.ndarray, per_obs) -> np.ndarray:
    """(N, n_terms) term expectation values -> (N, n_obs) observable expectation values."""
    out = np.zeros((expv.shape[0], len(per_obs)))
    for o, entries in enumerate(per_obs):
        for t, c, _ in entries:
            out[:, o] += c * expv[:, t]
    return out


def b200_evaluate_statevector(circuit: B200Circuit, executor: Executor, param=None, x=None,
                              as_tensor: bool = False, **_):
    """(N, 2^n) complex128 statevectors, little-endian (qulacs_evaluate_statevector, :45-63)."""
    eng = executor.engine
    x = np.zeros((1, 0)) if x is None else x
    psi = eng.simulate(circuit.compiled(eng, param), x)
    return psi if as_tensor else psi.cpu().numpy()


def b200_evaluate(circuit: B200Circuit, executor: Executor, param=None, x=None, param_obs=None,
                  squared: bool = False, **_) -> np.ndarray:
    """(N, n_obs) expectation values of the circuit's observables (qulacs_evaluate, :11-42)."""
    eng = executor.engine
    x = np.zeros((1, 0)) if x is None else x
    strings, per_obs = circuit.observable_terms(param_obs, squared)
    expv = eng.simulate_expvals(circuit.compiled(eng, param), x, strings).cpu().numpy()
    return _combine(expv, per_obs)


def b200_operator_gradient(circuit: B200Circuit, executor: Executor, param=None, x=None,
                           param_obs=None, **_) -> np.ndarray:
    """(N, n_obs, P_op): d<O>/d param_obs = expectation of the term each parameter multiplies
    (qulacs_operator_gradient, :141-185)."""
    eng = executor.engine
    x = np.zeros((1, 0)) if x is None else x
    strings, per_obs = circuit.observable_terms(param_obs)
    expv = eng.simulate_expvals(circuit.compiled(eng, param), x, strings).cpu().numpy()
    out = np.zeros((expv.shape[0], len(per_obs), circuit.num_parameters_observable))
    for o, entries in enumerate(per_obs):
        for t, _, j in entries:
            if j >= 0:
                out[:, o, j] += expv[:, t]
    return out


def b200_gradient(circuit: B200Circuit, executor: Executor, param=None, x=None, param_obs=None,
                  squared: bool = False, variant_slice: Optional[slice] = None, wrt: str = "param", **_) -> np.ndarray:
    """(N, n_obs, P): d<O>/d param by the parameter-shift batch
    (src/squlearn/util/optree/optree_derivative.py:130-158): all shifted circuits of all data
    points are evaluated in one engine call, variant index outermost.  ``wrt="x"`` gives (N, n_obs, d):
    d<O>/d x through the same shifted circuits and the chain rule of the encoding angles (the reference's
    "dfdx", evaluation_classes.py:105-125).

    ``variant_slice`` restricts the evaluation to a slice of the shifted-variant list (used to
    shard the batch over GPUs); the partial sums of all slices add up to the full gradient."""
    eng = executor.engine
    prog = circuit.program
    x2 = np.asarray(np.zeros((1, 0)) if x is None else x, dtype=np.float64)
    if x2.ndim == 1:
        x2 = x2.reshape(-1, max(prog.num_features, 1))
    strings, per_obs = circuit.observable_terms(param_obs, squared)
    gate_idx, shift, weight, pidx = prog.shift_variants(
        x2, wrt=wrt, p=None if param is None else np.asarray(param, dtype=np.float64))
    if variant_slice is not None:
        gate_idx, shift, weight, pidx = (gate_idx[variant_slice], shift[variant_slice],
                                         weight[variant_slice], pidx[variant_slice])
    out = np.zeros((x2.shape[0], len(per_obs), prog.num_parameters if wrt == "param" else prog.num_features))
    if len(gate_idx) == 0:
        return out
    expv = eng.simulate_expvals(circuit.compiled(eng, param), x2, strings,
                                variants=(gate_idx, shift)).cpu().numpy()  # (V, N, n_terms)
    for v in range(len(gate_idx)):
        out[:, :, pidx[v]] += weight[v][:, None] * _combine(expv[v], per_obs)
    return out


def _term_values(expv: np.ndarray, per_obs, n_param_obs: int, per_term: bool) -> np.ndarray:
    """(N, n_terms) term expectation values -> (N, n_obs) observable values, or with ``per_term`` the derivative
    with respect to every observable parameter (N, n_obs, P_op) (the expectation of the term it multiplies)."""
    if not per_term:
        return _combine(expv, per_obs)
    out = np.zeros((expv.shape[0], len(per_obs), n_param_obs))
    for o, entries in enumerate(per_obs):
        for t, _, j in entries:
            if j >= 0:
                out[:, o, j] += expv[:, t]
    return out


def b200_hessian(circuit: B200Circuit, executor: Executor, param=None, x=None, param_obs=None, wrt=("x", "x"),
                 squared: bool = False, per_term: bool = False, chunk: int = 4096, **_) -> np.ndarray:
    """Second derivatives of the expectation values: (N, n_obs, D1, D2) with ``wrt = (a, b)``, a, b in {"param", "x"}
    (axes in that order), from doubly shifted circuits (program.second_order_variants) evaluated ``chunk`` variants
    per engine call.  ``per_term``: (N, n_obs, P_op, D1, D2), the same derivative of d<O>/d param_obs.
    Reference: "dfdxdx", "dfdpdx", "dfdxdp", "dfdpdp", "dfdopdxdx" of evaluation_classes.py:105-164."""
    eng = executor.engine
    prog = circuit.program
    x2 = np.asarray(np.zeros((1, 0)) if x is None else x, dtype=np.float64)
    if x2.ndim == 1:
        x2 = x2.reshape(-1, max(prog.num_features, 1))
    pv = None if param is None else np.asarray(param, dtype=np.float64)
    strings, per_obs = circuit.observable_terms(param_obs, squared)
    ga, sa, gb, sb, weight, ia, ib = second_order_variants(prog, x2, wrt[0], wrt[1], pv)
    dims = tuple(prog.num_parameters if w == "param" else prog.num_features for w in wrt)
    npo = circuit.num_parameters_observable
    shape = (x2.shape[0], len(per_obs)) + ((npo,) if per_term else ()) + dims
    out = np.zeros(shape)
    cprog = circuit.compiled(eng, param)
    for c0 in range(0, len(ga), chunk):
        sl = slice(c0, c0 + chunk)
        expv = eng.simulate_expvals(cprog, x2, strings, variants=(ga[sl], sa[sl], gb[sl], sb[sl])).cpu().numpy()
        for v in range(expv.shape[0]):
            vals = _term_values(expv[v], per_obs, npo, per_term)         # (N, n_obs[, P_op])
            w = weight[c0 + v].reshape((-1,) + (1,) * (vals.ndim - 1))
            out[..., ia[c0 + v], ib[c0 + v]] += w * vals
    return out


def b200_term_gradient(circuit: B200Circuit, executor: Executor, param=None, x=None, param_obs=None,
                       wrt: str = "x", **_) -> np.ndarray:
    """(N, n_obs, P_op, D): derivative with respect to x (or the circuit parameters) of d<O>/d param_obs — the
    reference's "dfdopdx" / "dfdopdp" (evaluation_classes.py:130-160)."""
    eng = executor.engine
    prog = circuit.program
    x2 = np.asarray(np.zeros((1, 0)) if x is None else x, dtype=np.float64)
    if x2.ndim == 1:
        x2 = x2.reshape(-1, max(prog.num_features, 1))
    strings, per_obs = circuit.observable_terms(param_obs)
    gate_idx, shift, weight, pidx = prog.shift_variants(
        x2, wrt=wrt, p=None if param is None else np.asarray(param, dtype=np.float64))
    npo = circuit.num_parameters_observable
    out = np.zeros((x2.shape[0], len(per_obs), npo, prog.num_parameters if wrt == "param" else prog.num_features))
    if len(gate_idx) == 0:
        return out
    expv = eng.simulate_expvals(circuit.compiled(eng, param), x2, strings, variants=(gate_idx, shift)).cpu().numpy()
    for v in range(len(gate_idx)):
        out[..., pidx[v]] += weight[v][:, None, None] * _term_values(expv[v], per_obs, npo, True)
    return out
