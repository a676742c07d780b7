"""QNN losses, the training loop and QNNRegressor on the b200 backend.

The consumers of the parameter-shift batch (BASELINE configs[3]: QNNRegressor, ChebyshevRx(8, 4), batch 1024):

* ``SquaredLoss`` / ``MeanSquaredError`` / ``VarianceLoss`` follow the loss protocol of the reference
  (src/squlearn/qnn/loss/qnn_loss_base.py; squared_loss.py:35-61,92-143; mean_squared_error.py; variance_loss.py):
  ``loss_args_tuple`` / ``gradient_args_tuple`` name the LowLevelQNN.evaluate keys a step needs, ``value`` and
  ``gradient`` reduce the returned dict.  Losses add up (``SquaredLoss() + VarianceLoss(a)``).
* ``train`` is the full-batch loop of src/squlearn/qnn/util/training.py:277-360: the optimiser minimises
  theta = (param, param_op) with ``fun`` = loss value and ``grad`` = loss gradient, each ONE batched engine
  call over all samples (and, for the gradient, all shifted circuits).
* ``QNNRegressor`` has the constructor / fit / partial_fit / predict surface of src/squlearn/qnn/qnnr.py:222-321 and
  base_qnn.py:67-135,306-390 (parameter initialisation rule included).
* ``SLSQP`` / ``LBFGSB`` wrap scipy.optimize.minimize like optimizers/optimizers_wrapper.py:27-130; ``Adam`` is the
  update rule of optimizers/adam.py (bias-corrected effective learning rate, break on the parameter update).
"""

from typing import Callable, Sequence, Union

import numpy as np

from .data_preprocessing import extract_num_features
from .qnn import LowLevelQNN


# ------------------------------------------------------------------------------------------- losses
def _weights(kwargs, like):
    w = kwargs.get("weights")
    return np.ones_like(like) if w is None else np.asarray(w, dtype=np.float64)


def _contract(residual: np.ndarray, derivative: np.ndarray) -> np.ndarray:
    """sum over every sample / output axis of residual * d f / d theta_k (works for one and several outputs)."""
    if derivative.shape[0] == 0:
        return np.array([])
    return np.tensordot(residual, derivative, axes=residual.ndim)


class QNNLossBase:
    """Loss protocol of qnn_loss_base.py: what to evaluate, and how to reduce it."""

    def __init__(self):
        self._opt_param_op = True

    def set_opt_param_op(self, opt_param_op: bool = True):
        self._opt_param_op = opt_param_op

    loss_variance_available = False
    loss_args_tuple: tuple = ()
    variance_args_tuple: tuple = ()

    @property
    def gradient_args_tuple(self) -> tuple:
        raise NotImplementedError

    def value(self, value_dict: dict, **kwargs) -> float:
        raise NotImplementedError

    def variance(self, value_dict: dict, **kwargs) -> float:
        raise NotImplementedError

    def gradient(self, value_dict: dict, **kwargs):
        raise NotImplementedError

    def __add__(self, other):
        return _SumLoss(self, other)


class _SumLoss(QNNLossBase):
    """l1 + l2: union of the evaluation keys, sum of values and gradients."""

    def __init__(self, l1: QNNLossBase, l2: QNNLossBase):
        super().__init__()
        self._parts = (l1, l2)

    def set_opt_param_op(self, opt_param_op: bool = True):
        self._opt_param_op = opt_param_op
        for p in self._parts:
            p.set_opt_param_op(opt_param_op)

    @property
    def loss_variance_available(self):
        return all(p.loss_variance_available for p in self._parts)

    @property
    def loss_args_tuple(self):
        return tuple(dict.fromkeys(self._parts[0].loss_args_tuple + self._parts[1].loss_args_tuple))

    @property
    def variance_args_tuple(self):
        return tuple(dict.fromkeys(self._parts[0].variance_args_tuple + self._parts[1].variance_args_tuple))

    @property
    def gradient_args_tuple(self):
        return tuple(dict.fromkeys(self._parts[0].gradient_args_tuple + self._parts[1].gradient_args_tuple))

    def value(self, value_dict, **kwargs):
        return sum(p.value(value_dict, **kwargs) for p in self._parts)

    def variance(self, value_dict, **kwargs):
        return sum(p.variance(value_dict, **kwargs) for p in self._parts)

    def gradient(self, value_dict, **kwargs):
        g1, g2 = (p.gradient(value_dict, **kwargs) for p in self._parts)
        if self._opt_param_op:
            return g1[0] + g2[0], g1[1] + g2[1]
        return g1 + g2


class SquaredLoss(QNNLossBase):
    """sum_i w_i |f(x_i) - y_i|^2 (squared_loss.py:35-61) and its gradient (:92-143)."""

    loss_variance_available = True
    loss_args_tuple = ("f",)
    variance_args_tuple = ("f", "var")
    _mean = False

    @property
    def gradient_args_tuple(self) -> tuple:
        return ("f", "dfdp", "dfdop") if self._opt_param_op else ("f", "dfdp")

    def _scale(self, kwargs, truth):
        if self._mean:
            if kwargs.get("weights") is not None:
                raise ValueError("Weights are not supported for MeanSquaredError.")
            return 1.0 / len(truth)
        return 1.0

    @staticmethod
    def _truth(kwargs):
        if "ground_truth" not in kwargs:
            raise AttributeError("SquaredLoss requires ground_truth.")
        return np.asarray(kwargs["ground_truth"])

    def value(self, value_dict: dict, **kwargs) -> float:
        truth = self._truth(kwargs)
        scale = self._scale(kwargs, truth)
        return float(scale * np.sum(_weights(kwargs, truth) * np.square(value_dict["f"] - truth)))

    def variance(self, value_dict: dict, **kwargs) -> float:
        truth = self._truth(kwargs)
        scale = self._scale(kwargs, truth)
        return float(4.0 * scale * np.sum(_weights(kwargs, truth) * np.square(value_dict["f"] - truth)
                                          * value_dict["var"]))

    def gradient(self, value_dict: dict, **kwargs):
        truth = self._truth(kwargs)
        residual = 2.0 * self._scale(kwargs, truth) * _weights(kwargs, truth) * (value_dict["f"] - truth)
        d_p = _contract(residual, value_dict["dfdp"])
        if not self._opt_param_op:
            return d_p
        return d_p, _contract(residual, value_dict["dfdop"])


class MeanSquaredError(SquaredLoss):
    """(1/N) sum_i |f(x_i) - y_i|^2 (mean_squared_error.py); weights are rejected like in the reference."""

    _mean = True


class VarianceLoss(QNNLossBase):
    """alpha * sum_i Var_i (variance_loss.py); alpha may be a function of the iteration."""

    loss_variance_available = True
    loss_args_tuple = ("var",)

    def __init__(self, alpha: Union[float, Callable[[int], float]] = 0.005):
        super().__init__()
        self._alpha = alpha

    @property
    def gradient_args_tuple(self) -> tuple:
        return ("var", "dvardp", "dvardop") if self._opt_param_op else ("var", "dvardp")

    def _a(self, kwargs):
        if callable(self._alpha):
            if kwargs.get("iteration") is None:
                raise AttributeError("If alpha is callable, iteration is required.")
            return self._alpha(kwargs["iteration"])
        return self._alpha

    def value(self, value_dict, **kwargs):
        return float(self._a(kwargs) * np.sum(value_dict["var"]))

    def variance(self, value_dict, **kwargs):
        return 0.0

    def gradient(self, value_dict, **kwargs):
        a = self._a(kwargs)
        lead = tuple(range(value_dict["dvardp"].ndim - 1))
        d_p = a * np.sum(value_dict["dvardp"], axis=lead) if value_dict["dvardp"].shape[0] else np.array([])
        if not self._opt_param_op:
            return d_p
        lead = tuple(range(value_dict["dvardop"].ndim - 1))
        d_op = a * np.sum(value_dict["dvardop"], axis=lead) if value_dict["dvardop"].shape[0] else np.array([])
        return d_p, d_op


# ------------------------------------------------------------------------------------------- optimisers
class OptimizerResult:
    x = None
    nit = 0
    fun = 0.0


class _ScipyOptimizer:
    method = ""

    def __init__(self, options: dict = None, callback=None):
        options = dict(options or {})
        self.tol = options.pop("tol", 1e-6)
        self.options = options
        self.callback = callback
        self.iteration = 0

    def minimize(self, fun, x0, grad=None, bounds=None) -> OptimizerResult:
        import scipy.optimize

        def cb(*args):
            self.iteration += 1
            if self.callback is not None:
                self.callback(self.iteration, *args)

        res = scipy.optimize.minimize(fun, jac=grad, x0=np.asarray(x0, dtype=np.float64), method=self.method,
                                      options=self.options, bounds=bounds, tol=self.tol, callback=cb)
        out = OptimizerResult()
        out.x, out.nit, out.fun = res.x, res.nit, res.fun
        return out


class SLSQP(_ScipyOptimizer):
    method = "SLSQP"


class LBFGSB(_ScipyOptimizer):
    method = "L-BFGS-B"


class Adam:
    """ADAM with the reference's conventions (optimizers/adam.py): options lr (0.05), beta_1 (0.9), beta_2 (0.99),
    regularization (1e-8), maxiter (100), tol (1e-6, on the parameter update)."""

    def __init__(self, options: dict = None, callback=None):
        o = dict(options or {})
        self.lr, self.beta_1, self.beta_2 = o.get("lr", 0.05), o.get("beta_1", 0.9), o.get("beta_2", 0.99)
        self.regularization, self.maxiter, self.tol = o.get("regularization", 1e-8), o.get("maxiter", 100), o.get("tol", 1e-6)
        self.callback = callback
        self.iteration = 0
        self.m = self.v = None

    def _lr_eff(self):
        lr = self.lr(self.iteration) if callable(self.lr) else (
            self.lr[self.iteration] if isinstance(self.lr, (list, np.ndarray)) else self.lr)
        t = max(self.iteration, 0) + 1
        return lr * np.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)

    def minimize(self, fun, x0, grad=None, bounds=None) -> OptimizerResult:
        if grad is None:
            raise ValueError("Adam on the b200 backend needs the analytic gradient")
        x = np.asarray(x0, dtype=np.float64).copy()
        while self.iteration < self.maxiter:
            g = np.asarray(grad(x), dtype=np.float64)
            if self.m is None:
                self.m, self.v = np.zeros_like(g), np.zeros_like(g)
            self.m = self.beta_1 * self.m + (1.0 - self.beta_1) * g
            self.v = self.beta_2 * self.v + (1.0 - self.beta_2) * g * g
            x_new = x - self._lr_eff() * self.m / (np.sqrt(self.v) + self.regularization)
            if bounds is not None:
                x_new = np.clip(x_new, np.asarray(bounds)[:, 0], np.asarray(bounds)[:, 1])
            self.iteration += 1
            if self.callback is not None:
                self.callback(self.iteration, x, g, None)
            done = np.linalg.norm(x - x_new) < self.tol
            x = x_new
            if done:
                break
        out = OptimizerResult()
        out.x, out.nit, out.fun = x, self.iteration, fun(x)
        return out


# ------------------------------------------------------------------------------------------- training
def train(qnn, input_values, ground_truth, param_ini, param_op_ini, loss: QNNLossBase, optimizer,
          shot_control=None, weights=None, opt_param_op: bool = True):
    """Full-batch training (training.py:277-360): returns param, or (param, param_op) when the observable
    parameters are optimised too."""
    if weights is not None:
        if not isinstance(weights, np.ndarray):
            raise TypeError(f"Unknown weight format: {type(weights)}")
        if weights.shape != np.shape(ground_truth):
            raise ValueError(f"Shape {weights.shape} of weight values doesn't match shape "
                             f"{np.shape(ground_truth)} of reference values")
    if shot_control is not None:
        raise NotImplementedError("shot control: Executor('b200') is a statevector engine")
    loss.set_opt_param_op(opt_param_op)
    param = np.atleast_1d(np.asarray(param_ini, dtype=np.float64))
    param_op = np.atleast_1d(np.asarray(param_op_ini, dtype=np.float64))
    n_p = len(param)
    theta0 = np.concatenate((param, param_op)) if opt_param_op else param

    def split(theta):
        return (theta[:n_p], theta[n_p:]) if opt_param_op else (theta, param_op)

    def iteration():
        return getattr(optimizer, "iteration", None)

    def fun(theta):
        p, po = split(theta)
        values = qnn.evaluate(input_values, p, po, *loss.loss_args_tuple)
        return loss.value(values, ground_truth=ground_truth, weights=weights, iteration=iteration())

    def grad(theta):
        p, po = split(theta)
        values = qnn.evaluate(input_values, p, po, *loss.gradient_args_tuple)
        g = loss.gradient(values, ground_truth=ground_truth, weights=weights, iteration=iteration(),
                          multiple_output=qnn.multiple_output, opt_param_op=opt_param_op)
        return np.concatenate(g, axis=None) if opt_param_op else np.asarray(g)

    if len(theta0) == 0:
        return (np.array([]), np.array([])) if opt_param_op else np.array([])
    result = optimizer.minimize(fun, theta0, grad, bounds=None)
    theta = getattr(result, "x", result)
    return split(theta) if opt_param_op else theta


class QNNRegressor:
    """QNNRegressor(encoding_circuit, operator, executor, loss, optimizer, param_ini=None, param_op_ini=None, ...)
    — qnnr.py:222-321 on base_qnn.py.  Mini-batch SGD options are accepted for signature compatibility; every
    optimiser here trains on the full batch (the engine evaluates the whole batch in one call)."""

    def __init__(self, encoding_circuit, operator, executor, loss: QNNLossBase, optimizer, param_ini=None,
                 param_op_ini=None, batch_size: int = None, epochs: int = None, shuffle: bool = None,
                 opt_param_op: bool = True, variance: Union[float, Callable] = None, shot_control=None,
                 parameter_seed: Union[int, None] = 0, caching: bool = True, pretrained: bool = False,
                 callback=None, primitive=None, **kwargs):
        if pretrained and (param_ini is None or param_op_ini is None):
            raise ValueError("If pretrained is True, param_ini and param_op_ini must be provided!")
        self.encoding_circuit, self.operator, self.executor = encoding_circuit, operator, executor
        self.loss, self.optimizer, self.variance = loss, optimizer, variance
        self.param_ini, self.param_op_ini = param_ini, param_op_ini
        self.batch_size, self.epochs, self.shuffle = batch_size, epochs, shuffle
        self.opt_param_op, self.parameter_seed, self.caching = opt_param_op, parameter_seed, caching
        self.pretrained, self.callback, self.primitive, self.shot_control = pretrained, callback, primitive, shot_control
        self._param = self._param_op = self._qnn = None
        self._is_fitted = False
        self._is_lowlevel_qnn_initialized = False

    param = property(lambda self: self._param)
    param_op = property(lambda self: self._param_op)
    num_parameters = property(lambda self: self._qnn.num_parameters)
    num_parameters_observable = property(lambda self: self._qnn.num_parameters_observable)

    def _initialize_lowlevel_qnn(self, num_features: int) -> None:
        """base_qnn.py:347-389: build the LowLevelQNN; (re)draw the parameters unless usable initial values exist."""
        if self._is_lowlevel_qnn_initialized:
            return
        self._qnn = LowLevelQNN(self.encoding_circuit, self.operator, self.executor, num_features,
                                caching=self.caching)
        if not self._is_fitted:
            if self.param_ini is None or len(self.param_ini) != self.encoding_circuit.num_parameters:
                self._param = self.encoding_circuit.generate_initial_parameters(
                    seed=self.parameter_seed, num_features=num_features)
            else:
                self._param = np.array(self.param_ini, dtype=np.float64)
            if self.param_op_ini is None or len(self.param_op_ini) != self._qnn.num_parameters_observable:
                ops = self.operator if isinstance(self.operator, (list, tuple)) else [self.operator]
                self._param_op = np.concatenate([o.generate_initial_parameters(seed=self.parameter_seed + i + 1)
                                                 for i, o in enumerate(ops)])
            else:
                self._param_op = np.array(self.param_op_ini, dtype=np.float64)
        self._is_lowlevel_qnn_initialized = True

    @staticmethod
    def _xy(X, y=None):
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X.reshape(-1, 1)
        if y is None:
            return X
        y = np.asarray(y, dtype=np.float64)
        if y.ndim == 2 and y.shape[1] == 1:
            y = y.ravel()
        if len(y) != len(X):
            raise ValueError("Found input variables with inconsistent numbers of samples")
        return X, y

    def fit(self, X, y, weights: np.ndarray = None) -> None:
        """Re-initialise the parameters and fit (base_qnn.py:184-198)."""
        self._is_lowlevel_qnn_initialized = False
        self._is_fitted = False
        self.partial_fit(X, y, weights)

    def partial_fit(self, X, y, weights: np.ndarray = None) -> None:
        X, y = self._xy(X, y)
        self._initialize_lowlevel_qnn(extract_num_features(X))
        loss = self.loss if self.variance is None else self.loss + VarianceLoss(alpha=self.variance)
        out = train(self._qnn, X, y, self._param, self._param_op, loss, self.optimizer, self.shot_control,
                    weights, bool(self.opt_param_op))
        if self.opt_param_op:
            self._param, self._param_op = out
        else:
            self._param = out
        self._is_fitted = True

    def predict(self, X) -> np.ndarray:
        X = self._xy(X)
        if not self._is_fitted and not self.pretrained:
            raise RuntimeError("The model is not fitted.")
        if self._qnn is None:
            if self.pretrained:
                self._param, self._param_op = np.array(self.param_ini), np.array(self.param_op_ini)
                self._is_fitted = True
            self._initialize_lowlevel_qnn(extract_num_features(X))
        return self._qnn.evaluate(X, self._param, self._param_op, "f")["f"]

    def get_params(self, deep: bool = True) -> dict:
        params = {k: getattr(self, k) for k in (
            "encoding_circuit", "operator", "executor", "loss", "optimizer", "param_ini", "param_op_ini",
            "batch_size", "epochs", "shuffle", "opt_param_op", "variance", "parameter_seed", "caching",
            "pretrained", "callback", "primitive", "shot_control")}
        if deep:
            params.update(self.encoding_circuit.get_params())
        return params

    def set_params(self, **params):
        """Encoding-circuit hyper-parameters re-initialise the QNN (and re-draw the parameters when their number
        changes: base_qnn.py:226-304)."""
        own = self.get_params(deep=False)
        enc = {k: v for k, v in params.items() if k not in own}
        for k, v in params.items():
            if k in own:
                setattr(self, k, v)
        if enc:
            self.encoding_circuit.set_params(**enc)
        self._is_fitted = False
        self._is_lowlevel_qnn_initialized = False
        self._qnn = None
        return self
