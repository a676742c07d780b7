"""CPU tests of the kernel-method consumers against NumPy / SciPy restatements of the
reference lines.  A stand-in kernel backed by the ORACLE supplies the Gram matrices as CPU torch tensors, so
the device-agnostic torch.linalg code of QGPR runs here exactly as it does on CUDA tensors."""

import numpy as np
import scipy.linalg
import torch

import oracle
from squlearn_b200 import QGPR


class _OracleKernel:
    _regularization = None

    def __init__(self, n, d):
        self.n, self.d = n, d
        self.gates = oracle.chebyshev_pqc(n, 2, d)
        self.p = oracle.chebyshev_pqc_initial_parameters(n, 2, d, seed=0)

    def evaluate_device(self, x, y=None):
        return torch.from_numpy(oracle.fidelity_kernel(self.gates, self.n, x, y, self.p))


def _reference_qgpr(k_train, k_testtrain, k_test, y, sigma, full_regularization):
    """qgpr.py:246-283 + regularization.py:5-27,49-70 restated with NumPy / SciPy."""
    if full_regularization:
        n = k_train.shape[0]
        full = np.block([[k_train, k_testtrain.T], [k_testtrain, k_test]])
        evals, evecs = scipy.linalg.eig(full)
        rec = np.real(evecs @ np.diag(evals.clip(min=0)) @ evecs.T)
        k_train, k_testtrain, k_test = rec[:n, :n], rec[n:, :n], rec[n:, n:]
    k_train = k_train + sigma * np.identity(k_train.shape[0])
    lu = scipy.linalg.lu_factor(k_train)
    alpha = scipy.linalg.lu_solve(lu, y)
    mean = k_testtrain.dot(alpha)
    cov = k_test - k_testtrain @ scipy.linalg.lu_solve(lu, k_testtrain.T)
    return mean, cov


def test_qgpr_matches_reference_restatement():
    n, d = 4, 2
    rng = np.random.default_rng(5)
    X, Xt = rng.uniform(-0.9, 0.9, (25, d)), rng.uniform(-0.9, 0.9, (9, d))
    y = np.cos(2 * X[:, 0]) - X[:, 1]
    kern = _OracleKernel(n, d)
    k_train = oracle.fidelity_kernel(kern.gates, n, X, None, kern.p)
    k_tt = oracle.fidelity_kernel(kern.gates, n, Xt, X, kern.p)
    k_test = oracle.fidelity_kernel(kern.gates, n, Xt, None, kern.p)
    for full_reg in (True, False):
        for normalize in (False, True):
            model = QGPR(kern, sigma=1e-5, normalize_y=normalize, full_regularization=full_reg).fit(X, y)
            mean, cov = model.predict(Xt, return_cov=True)
            yy = (y - y.mean()) / y.std() if normalize else y
            want_mean, want_cov = _reference_qgpr(k_train, k_tt, k_test, yy, 1e-5, full_reg)
            if normalize:
                want_mean = y.std() * want_mean + y.mean()
            assert np.max(np.abs(mean - want_mean)) < 1e-7
            assert np.max(np.abs(cov - want_cov)) < 1e-7


def test_qgpr_precomputed_and_std_and_sigma_accumulation_quirk():
    rng = np.random.default_rng(6)
    a = rng.normal(size=(12, 5))
    full = a @ a.T + 0.1 * np.eye(12)
    n_train = 8
    y = rng.normal(size=n_train)
    model = QGPR("precomputed", sigma=1e-3, full_regularization=False).fit(full[:n_train, :n_train].copy(), y)
    mean, std = model.predict(full, return_std=True)
    want_mean, want_cov = _reference_qgpr(full[:n_train, :n_train], full[n_train:, :n_train], full[n_train:, n_train:],
                                          y, 1e-3, False)
    assert np.max(np.abs(mean - want_mean)) < 1e-9
    assert np.max(np.abs(std - np.sqrt(np.diag(want_cov)))) < 1e-9
    # the reference adds sigma to the stored K_train on every predict (qgpr.py:256): second call uses 2 sigma
    mean2 = model.predict(full)
    want2, _ = _reference_qgpr(full[:n_train, :n_train], full[n_train:, :n_train], full[n_train:, n_train:], y, 2e-3, False)
    assert np.max(np.abs(mean2 - want2)) < 1e-9
