"""Quantum reservoir computing on the b200 backend (consumer of LowLevelQNN "f").

Mirrors src/squlearn/qrc/base_qrc.py:30-171 (constructor, fit/predict, operator set-up),
qrc_regressor.py / qrc_classifier.py (classical head selection) and
src/squlearn/qrc/util/random_pauli_list.py:6-38.  The quantum part is one batched engine call
per fit/predict; the classical head is scikit-learn on the host, as in the reference.
"""

from typing import List, Union

import numpy as np

from .data_preprocessing import extract_num_features
from .executor import Executor
from .observables import CustomObservable, ObservableBase, SinglePauli
from .qnn import LowLevelQNN


def random_pauli_list(pauli_string_length: int, num_samples: int, seed: int = 42) -> List[str]:
    total = 4 ** pauli_string_length
    if num_samples > total:
        raise ValueError(f"num_samples={num_samples} exceeds total possible {total}")
    symbols = np.array(["I", "X", "Y", "Z"])
    rng = np.random.default_rng(seed)
    picks = rng.choice(total, size=num_samples, replace=False)
    out = []
    for index in picks:
        chars, t = [], index
        for _ in range(pauli_string_length):
            chars.append(symbols[t % 4])
            t //= 4
        out.append("".join(reversed(chars)))
    return out


class BaseQRC:
    def __init__(self, encoding_circuit, executor: Executor, ml_model: str = "linear",
                 ml_model_options: Union[dict, None] = None,
                 operators: Union[ObservableBase, list, str] = "random_paulis",
                 num_operators: int = None, operator_seed: int = 0, param_ini=None,
                 param_op_ini=None, parameter_seed: Union[int, None] = 0, caching: bool = True):
        self.encoding_circuit = encoding_circuit
        self.executor = executor
        self.ml_model = ml_model
        self.ml_model_options = ml_model_options
        self.num_operators = num_operators
        self.operator_seed = operator_seed
        self.operators = operators
        self.param_ini = param_ini
        self.param_op_ini = param_op_ini
        self.parameter_seed = parameter_seed
        self.caching = caching
        self._qnn = None
        nq = encoding_circuit.num_qubits
        if isinstance(operators, str) and operators == "random_paulis":
            if num_operators is not None and num_operators > 4 ** nq:
                raise ValueError(
                    f"Number of operators ({num_operators}) exceeds the maximum possible "
                    f"({4 ** nq}) for the given number of qubits ({nq}).")
            if num_operators is None:
                self.num_operators = min(100, 4 ** nq)
        self._initialize_observables()
        self._ml_model = self._make_ml_model()

    def _initialize_observables(self):
        nq = self.encoding_circuit.num_qubits
        if isinstance(self.operators, str):
            if self.operators == "random_paulis":
                paulis = random_pauli_list(nq, self.num_operators, seed=self.operator_seed)
                self._operators = [CustomObservable(nq, str(p)) for p in paulis]
            elif self.operators == "single_paulis":
                self._operators = [SinglePauli(nq, i, p) for i in range(nq) for p in "XYZ"]
            else:
                raise ValueError("Invalid string for operators")
        elif isinstance(self.operators, ObservableBase):
            self._operators = [self.operators]
        elif isinstance(self.operators, list):
            self._operators = self.operators
        else:
            raise ValueError("Invalid operators. Must be an ObservableBase object or a list of "
                             "ObservableBase objects or None.")
        self.num_operators = len(self._operators)

    @property
    def used_operators(self):
        return self._operators

    @property
    def qnn(self):
        return self._qnn

    def _make_ml_model(self):
        raise NotImplementedError

    def _features(self, X):
        return self._qnn.evaluate(X, self.param_ini, self.param_op_ini, "f")["f"]

    def fit(self, X, y):
        X = np.asarray(X, dtype=np.float64)
        num_features = extract_num_features(X)
        self._qnn = LowLevelQNN(self.encoding_circuit, self._operators, self.executor,
                                num_features, caching=self.caching)
        if self.param_ini is None or len(self.param_ini) != self._qnn.num_parameters:
            self.param_ini = self.encoding_circuit.generate_initial_parameters(
                seed=self.parameter_seed, num_features=num_features)
        if self.param_op_ini is None or len(self.param_op_ini) != self._qnn.num_parameters_observable:
            self.param_op_ini = np.concatenate(
                [o.generate_initial_parameters(seed=self.parameter_seed) for o in self._operators])
        self._ml_model.fit(self._features(X), y)
        return self

    def predict(self, X):
        if self._qnn is None:
            raise RuntimeError("The model is not fitted.")
        return self._ml_model.predict(self._features(np.asarray(X, dtype=np.float64)))


class QRCRegressor(BaseQRC):
    def _make_ml_model(self):
        opts = self.ml_model_options or {}
        if self.ml_model == "linear":
            from sklearn.linear_model import LinearRegression
            return LinearRegression(**opts)
        if self.ml_model == "mlp":
            from sklearn.neural_network import MLPRegressor
            return MLPRegressor(**opts)
        if self.ml_model == "kernel":
            from sklearn.kernel_ridge import KernelRidge
            return KernelRidge(**opts)
        raise ValueError("Invalid ml_model. Please choose 'mlp', 'linear' or 'kernel'.")


class QRCClassifier(BaseQRC):
    def _make_ml_model(self):
        opts = self.ml_model_options or {}
        if self.ml_model == "linear":
            from sklearn.linear_model import Perceptron
            return Perceptron(**opts)
        if self.ml_model == "mlp":
            from sklearn.neural_network import MLPClassifier
            return MLPClassifier(**opts)
        if self.ml_model == "kernel":
            from sklearn.svm import SVC
            return SVC(**opts)
        raise ValueError("Invalid ml_model. Please choose 'mlp', 'linear' or 'kernel'.")
