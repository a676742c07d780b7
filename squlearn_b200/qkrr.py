"""Quantum kernel ridge regression with the Gram matrix kept on the GPU.

Mirror of ``squlearn.kernel.QKRR`` (src/squlearn/kernel/qkrr.py:99-199): ``fit`` evaluates the training
kernel, adds ``alpha * I`` and solves ``(K + alpha I) c = y`` with a Cholesky factorisation
(``scipy.linalg.cholesky`` / ``cho_solve`` there, :158-169); ``predict`` is ``K(X, X_train) @ c``.
With the 4096 x 4096 Gram produced in tens of milliseconds the 128 MiB device->host copy and the host
solve would dominate, so K never leaves HBM: the factorisation runs through cuSOLVER
(``torch.linalg.cholesky_ex`` — a plain library call; this row is a consumer of the hot path, not part of
it) and only ``dual_coeff_`` and the predictions are copied back.
"""

from typing import Optional, Union

import numpy as np

from .data_preprocessing import convert_to_float64, extract_num_features
from .kernel import KernelMatrixBase, device_kernel


class QKRR:
    def __init__(self, quantum_kernel: Optional[Union[KernelMatrixBase, str]] = None,
                 alpha: Union[float, np.ndarray] = 1.0e-6):
        self._quantum_kernel = quantum_kernel
        self._alpha = alpha
        self.X_train = None
        self.k_train = None          # device tensor (K + alpha I), like the reference attribute after fit
        self.k_testtrain = None
        self.dual_coeff_ = None
        self._dual_dev = None

    @property
    def alpha(self):
        return self._alpha

    def _torch(self):
        import torch

        return torch

    def _to_device(self, a):
        """A precomputed kernel matrix as a device tensor (the CUDA device when there is one)."""
        torch = self._torch()
        t = torch.as_tensor(np.asarray(a, dtype=np.float64))
        return t.cuda() if torch.cuda.is_available() else t

    def fit(self, X, y):
        torch = self._torch()
        y = np.asarray(y, dtype=np.float64)
        if isinstance(self._quantum_kernel, str):
            if self._quantum_kernel != "precomputed":
                raise ValueError("Unknown quantum kernel: {}".format(self._quantum_kernel))
            k = self._to_device(X)
        elif hasattr(self._quantum_kernel, "evaluate") or hasattr(self._quantum_kernel, "evaluate_device"):
            # any kernel object: FidelityKernel / ProjectedQuantumKernel keep K in HBM (evaluate_device),
            # wrappers such as KernelOptimizer and host-only kernels go through evaluate (qkrr.py:147-156)
            X = convert_to_float64(X)
            if hasattr(self._quantum_kernel, "_initialize_kernel"):
                self._quantum_kernel._initialize_kernel(num_features=extract_num_features(X))
            if hasattr(self._quantum_kernel, "run_optimization"):
                self._quantum_kernel.run_optimization(X, y)
            k = device_kernel(self._quantum_kernel, X)
        else:
            raise ValueError("Unknown type of quantum kernel: {}".format(type(self._quantum_kernel)))
        self.X_train = X
        n = k.shape[0]
        alpha = torch.as_tensor(np.broadcast_to(np.asarray(self._alpha, dtype=np.float64), (n,)).copy(),
                                device=k.device)
        k = k.clone()
        k.diagonal().add_(alpha)
        self.k_train = k
        # Cholesky decomposition for providing numerical stability (qkrr.py:164-169)
        L, info = torch.linalg.cholesky_ex(k)
        if int(info.item()) != 0:
            print("Increase regularization parameter alpha")
            return self
        yd = torch.as_tensor(y, device=k.device)
        rhs = yd.reshape(n, -1)
        self._dual_dev = torch.cholesky_solve(rhs, L).reshape(yd.shape)
        self.dual_coeff_ = self._dual_dev.cpu().numpy()
        return self

    def predict(self, X) -> np.ndarray:
        if self.k_train is None:
            raise ValueError("The fit() method has to be called beforehand.")
        torch = self._torch()
        if isinstance(self._quantum_kernel, str):
            kt = self._to_device(X).to(self.k_train.device)
        else:
            kt = device_kernel(self._quantum_kernel, convert_to_float64(X), self.X_train)
        self.k_testtrain = kt
        return (kt @ self._dual_dev).cpu().numpy()
