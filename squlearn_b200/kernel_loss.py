"""Losses and the optimisation loop of trainable quantum kernels with the Gram matrix kept on the GPU
(SURVEY.md §8 row f2).

Mirrors of ``squlearn.kernel.loss.TargetAlignment`` (src/squlearn/kernel/loss/target_alignment.py:42-90),
``NLL`` (loss/negative_log_likelihood.py:43-90) and ``KernelOptimizer``
(lowlevel_kernel/kernel_optimizer.py:14-186).  Every optimiser step re-evaluates the full Gram matrix with new
circuit parameters: ``assign_parameters`` only updates the angle coefficients of the compiled program
(``b200sq_program_set_coeffs`` — no re-planning, and the run-time specialised kernels depend on the gate
structure only), K stays in HBM and the loss is reduced there, so a step costs one Gram evaluation plus a few
O(N^2) device reductions and no 8 N^2-byte device->host copy.
  TargetAlignment:  sum(K * y y^T) = y^T K y,  sum(K * K) = |K|_F^2,  sum(T * T) = (y^T y)^2
  NLL:              Cholesky of K + sigma I on the device (cuSOLVER through torch.linalg: a plain library call)
"""

from functools import partial

import numpy as np


def _device_kernel(kernel, data):
    """The Gram matrix as a float64 torch tensor on whatever device the kernel computes on."""
    from .kernel import device_kernel

    return device_kernel(kernel, data)


class KernelLossBase:
    """kernel_loss_base.py:7-39."""

    def __init__(self) -> None:
        self._quantum_kernel = None

    def set_quantum_kernel(self, quantum_kernel) -> None:
        self._quantum_kernel = quantum_kernel

    def compute(self, parameter_values, data, **kwargs) -> float:
        raise NotImplementedError

    def _check(self, kwargs, what):
        if self._quantum_kernel is None:
            raise ValueError("Quantum kernel is not set, please set the quantum kernel with set_quantum_kernel method")
        labels = kwargs.get("labels", None)
        if labels is None:
            raise ValueError(f"labels must be provided as a keyword argument for {what} computation")
        return labels


class TargetAlignment(KernelLossBase):
    """Negative kernel-target alignment (target_alignment.py:42-90)."""

    def __init__(self, rescale_class_labels=True) -> None:
        super().__init__()
        self._rescale_class_labels = rescale_class_labels

    def compute(self, parameter_values, data, **kwargs) -> float:
        import torch

        labels = self._check(kwargs, "TargetAlignment")
        self._quantum_kernel.assign_parameters(parameter_values)
        if self._rescale_class_labels:
            nplus = np.count_nonzero(np.array(labels) == 1)
            nminus = len(labels) - nplus
            y = np.array([v / nplus if v == 1 else v / nminus for v in labels], dtype=np.float64)
        else:
            y = np.array(labels, dtype=np.float64)
        fused = self._quantum_kernel.evaluate_alignment(data, y) if hasattr(self._quantum_kernel, "evaluate_alignment") \
            else None
        if fused is not None:
            # sum(K * outer(y, y)) and sum(K * K) were reduced in the Gram kernel's epilogue: K is not read again
            sums = fused[1]
            yy = float(np.dot(y, y))
            return float(-(sums[0] / torch.sqrt(sums[1] * yy * yy)).item())
        k = _device_kernel(self._quantum_kernel, data)
        yd = torch.as_tensor(y, device=k.device)
        inner = torch.dot(yd, k @ yd)                              # sum(K * outer(y, y))
        norm = torch.sqrt(torch.sum(k * k) * torch.dot(yd, yd) ** 2)  # sqrt(sum(K*K) * sum(T*T))
        return float(-(inner / norm).item())


class NLL(KernelLossBase):
    """Negative log likelihood of a Gaussian process (negative_log_likelihood.py:43-90)."""

    def __init__(self, sigma=0.0) -> None:
        super().__init__()
        self._sigma = sigma

    def compute(self, parameter_values, data, **kwargs):
        import torch

        labels = np.asarray(self._check(kwargs, "NLL"), dtype=np.float64)
        self._quantum_kernel.assign_parameters(parameter_values)
        k = _device_kernel(self._quantum_kernel, data).clone()
        k.diagonal().add_(self._sigma)
        L = torch.linalg.cholesky(k)
        yd = torch.as_tensor(labels, device=k.device).reshape(len(labels), -1)
        s2 = torch.cholesky_solve(yd, L)
        val = torch.sum(torch.log(torch.diagonal(L))) + 0.5 * torch.sum(yd * s2) + 0.5 * len(data) * np.log(2.0 * np.pi)
        return np.array([float(val.item())])


class KernelOptimizer:
    """Parameter optimisation of a trainable quantum kernel (kernel_optimizer.py:14-186): wraps a kernel, a loss
    and an optimiser with a ``minimize(fun, x0)`` method returning an object with ``.x`` (the reference's
    ``OptimizerBase`` protocol; ``scipy.optimize.minimize`` works through ``ScipyMinimize`` below)."""

    def __init__(self, quantum_kernel, loss=None, optimizer=None, initial_parameters=None):
        self._quantum_kernel = quantum_kernel
        self._loss = loss
        self._optimizer = optimizer
        self._initial_parameters = initial_parameters
        self._optimal_parameters = None
        self._is_fitted = False
        self._loss.set_quantum_kernel(quantum_kernel)

    @property
    def is_fitted(self) -> bool:
        return self._is_fitted

    def run_optimization(self, X, y=None):
        from .data_preprocessing import extract_num_features

        if self._is_fitted:
            return None
        if self._quantum_kernel.num_parameters == 0:
            return None
        self._quantum_kernel._initialize_kernel(num_features=extract_num_features(X))
        if self._initial_parameters is None:
            self._initial_parameters = self._quantum_kernel.parameters
        loss_function = partial(self._loss.compute, data=X, labels=y)
        opt_result = self._optimizer.minimize(fun=loss_function, x0=self._initial_parameters)
        self._optimal_parameters = opt_result.x
        self._quantum_kernel.assign_parameters(self._optimal_parameters)
        self._is_fitted = True
        return opt_result

    def assign_parameters(self, parameters):
        self._is_fitted = False
        self._quantum_kernel.assign_parameters(parameters)
        self._initial_parameters = parameters

    def evaluate(self, x, y=None):
        return self._quantum_kernel.evaluate(x, y)

    def evaluate_device(self, x, y=None):
        from .kernel import device_kernel

        return device_kernel(self._quantum_kernel, x, y)

    def get_optimal_parameters(self):
        return self._optimal_parameters


class ScipyMinimize:
    """Minimal optimiser with the ``minimize(fun, x0)`` protocol, backed by scipy.optimize.minimize."""

    def __init__(self, method="Nelder-Mead", options=None):
        self.method, self.options = method, options or {}

    def minimize(self, fun, x0):
        import scipy.optimize

        return scipy.optimize.minimize(lambda v: float(np.ravel(fun(v))[0]), np.asarray(x0, dtype=np.float64),
                                       method=self.method, options=self.options)
