"""Quantum Gaussian process regression with the Gram matrices kept on the GPU.

Mirror of ``squlearn.kernel.QGPR`` (src/squlearn/kernel/qgpr.py:95-283): ``fit`` evaluates the training
kernel (and optionally normalises y), ``predict`` evaluates K_test and K_testtrain, optionally regularises the
FULL Gram matrix [[K_train, K_testtrain^T], [K_testtrain, K_test]] by clipping negative eigenvalues
(``regularize_full_kernel``, lowlevel_kernel/regularization.py:49-70), adds ``sigma * I``, solves with an LU
factorisation and returns mean / std / cov (:258-283).  All matrices stay device tensors (cuSOLVER through
``torch.linalg`` — plain library calls; this row consumes the hot path's output); only the predictions are
copied to the host.

Reference quirk kept: every ``predict`` adds ``sigma * I`` to the stored ``K_train`` in
place (qgpr.py:256), so repeated predictions accumulate the regularisation exactly as the reference does.
"""

import warnings
from typing import Optional, Union

import numpy as np

from .data_preprocessing import convert_to_float64
from .kernel import KernelMatrixBase


def regularize_full_kernel(k_train, k_testtrain, k_test):
    """regularization.py:49-70 on device tensors (the full Gram matrix is symmetric: eigh)."""
    import torch

    n = k_train.shape[0]
    full = torch.cat([torch.cat([k_train, k_testtrain.T], dim=1), torch.cat([k_testtrain, k_test], dim=1)], dim=0)
    evals, evecs = torch.linalg.eigh(0.5 * (full + full.T))
    rec = (evecs * evals.clamp(min=0.0)) @ evecs.T
    return rec[:n, :n].contiguous(), rec[n:, :n].contiguous(), rec[n:, n:].contiguous()


class QGPR:
    def __init__(self, quantum_kernel: Optional[Union[KernelMatrixBase, str]] = None, sigma: float = 1.0e-6,
                 normalize_y: bool = False, full_regularization: bool = True):
        self._quantum_kernel = quantum_kernel
        self.X_train = None
        self.y_train = None
        self._sigma = sigma
        self._normalize_y = normalize_y
        self._full_regularization = full_regularization
        self.K_train = None
        self.K_test = None
        self.K_testtrain = None
        self._LU = None
        self._piv = None
        self._alpha = None

    sigma = property(lambda self: self._sigma)
    normalize_y = property(lambda self: self._normalize_y)
    full_regularization = property(lambda self: self._full_regularization)

    def _kernel(self, x, y=None):
        import torch

        k = self._quantum_kernel.evaluate_device(x, y) if hasattr(self._quantum_kernel, "evaluate_device") \
            else torch.as_tensor(self._quantum_kernel.evaluate(x, y))
        return k.to(torch.float64)

    def fit(self, X, y):
        import torch

        y = np.asarray(y, dtype=np.float64)
        if isinstance(self._quantum_kernel, str):
            if self._quantum_kernel != "precomputed":
                raise ValueError("Unknown quantum kernel: {}".format(self._quantum_kernel))
            self.K_train = torch.as_tensor(np.asarray(X, dtype=np.float64))
            if torch.cuda.is_available():
                self.K_train = self.K_train.cuda()
        elif isinstance(self._quantum_kernel, KernelMatrixBase) or hasattr(self._quantum_kernel, "evaluate_device"):
            X = convert_to_float64(X)
            if self._full_regularization and getattr(self._quantum_kernel, "_regularization", None) is not None:
                warnings.warn("If full_regularization is True, best results are achieved with no additional "
                              "quantum kernel regularization.")
            self.K_train = self._kernel(X).clone()
        else:
            raise ValueError("Unknown type of quantum kernel: {}".format(type(self._quantum_kernel)))
        self.X_train = X
        if self._normalize_y:
            self._y_train_mean = np.mean(y, axis=0)
            std = np.std(y, axis=0)
            self._y_train_std = np.where(std == 0.0, 1.0, std)  # sklearn _handle_zeros_in_scale
            self.y_train = (y - self._y_train_mean) / self._y_train_std
        else:
            self.y_train = y
        return self

    def predict(self, X, return_std=False, return_cov=False):
        import torch

        if self.K_train is None:
            raise ValueError("There is no training data. Please call the fit method first.")
        if return_std and return_cov:
            raise ValueError("Only one of return_std or return_cov can be True. Currently both are True.")
        dev = self.K_train.device
        if isinstance(self._quantum_kernel, str):
            Xd = torch.as_tensor(np.asarray(X, dtype=np.float64), device=dev)
            n_train = self.y_train.shape[0]
            n_test = Xd.shape[0] - n_train
            self.K_test = Xd[-n_test:, -n_test:]
            self.K_testtrain = Xd[n_train:, :n_train]
        else:
            X = convert_to_float64(X)
            self.K_test = self._kernel(X)
            self.K_testtrain = self._kernel(X, self.X_train)
        if self._full_regularization:
            self.K_train, self.K_testtrain, self.K_test = regularize_full_kernel(self.K_train, self.K_testtrain,
                                                                                 self.K_test)
        self.K_train.diagonal().add_(self._sigma)          # in place, like the reference (:256)
        self._LU, self._piv, info = torch.linalg.lu_factor_ex(self.K_train)
        if int(info.item()) != 0:
            self.K_train.diagonal().add_(1e-8)
            self._LU, self._piv = torch.linalg.lu_factor(self.K_train)
        yd = torch.as_tensor(self.y_train, device=dev)
        self._alpha = torch.linalg.lu_solve(self._LU, self._piv, yd.reshape(yd.shape[0], -1)).reshape(yd.shape)
        mean = self.K_testtrain @ self._alpha
        v = torch.linalg.lu_solve(self._LU, self._piv, self.K_testtrain.T.contiguous())
        cov = self.K_test - self.K_testtrain @ v
        mean = mean.cpu().numpy()
        if self._normalize_y:
            mean = self._y_train_std * mean + self._y_train_mean
        if return_std:
            return mean, np.sqrt(np.diag(cov.cpu().numpy()))
        if return_cov:
            return mean, cov.cpu().numpy()
        return mean
