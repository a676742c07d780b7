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


# ------------------------------------------------------------------------------ row f2: trainable kernels
class _TrainableOracleKernel(_OracleKernel):
    """Oracle-backed stand-in with the parameter interface the losses and the optimiser use."""

    def __init__(self, n, d):
        super().__init__(n, d)
        self.parameters = np.array(self.p, dtype=float)
        self.num_parameters = len(self.p)

    def assign_parameters(self, parameters):
        self.parameters = np.array(parameters, dtype=float)
        self.p = self.parameters

    def _initialize_kernel(self, num_features):
        pass

    def evaluate(self, x, y=None):
        return self.evaluate_device(x, y).numpy()


def test_target_alignment_and_nll_match_reference_restatements():
    from squlearn_b200 import NLL, TargetAlignment

    n, d = 3, 2
    rng = np.random.default_rng(8)
    X = rng.uniform(-0.9, 0.9, (14, d))
    labels = np.where(X[:, 0] * X[:, 1] > 0, 1, -1)
    kern = _TrainableOracleKernel(n, d)
    p = rng.uniform(-1, 1, kern.num_parameters)
    k = oracle.fidelity_kernel(kern.gates, n, X, None, p)
    for rescale in (True, False):
        ta = TargetAlignment(rescale_class_labels=rescale)
        ta.set_quantum_kernel(kern)
        got = ta.compute(p, X, labels=labels)
        # target_alignment.py:78-90
        if rescale:
            nplus = np.count_nonzero(labels == 1)
            nminus = len(labels) - nplus
            yy = np.array([v / nplus if v == 1 else v / nminus for v in labels])
        else:
            yy = labels.astype(float)
        T = np.outer(yy, yy)
        want = -np.sum(k * T) / np.sqrt(np.sum(k * k) * np.sum(T * T))
        assert abs(got - want) < 1e-12
    y = np.sin(X[:, 0])
    nll = NLL(sigma=1e-3)
    nll.set_quantum_kernel(kern)
    got = nll.compute(p, X, labels=y)
    # negative_log_likelihood.py:77-88
    km = k + 1e-3 * np.eye(len(X))
    L = scipy.linalg.cholesky(km, lower=True)
    s2 = scipy.linalg.solve_triangular(L.T, scipy.linalg.solve_triangular(L, y, lower=True), lower=False)
    want = np.sum(np.log(np.diagonal(L))) + 0.5 * y.T @ s2 + 0.5 * len(X) * np.log(2.0 * np.pi)
    assert got.shape == (1,) and abs(got[0] - want) < 1e-9


def test_kernel_optimizer_lowers_the_loss_and_assigns_the_optimum():
    from squlearn_b200 import KernelOptimizer, ScipyMinimize, TargetAlignment

    n, d = 2, 1
    rng = np.random.default_rng(9)
    X = rng.uniform(-0.9, 0.9, (12, d))
    labels = np.where(X[:, 0] > 0, 1, -1)
    kern = _TrainableOracleKernel(n, d)
    loss = TargetAlignment()
    opt = KernelOptimizer(kern, loss, ScipyMinimize("Nelder-Mead", {"maxiter": 60}))
    start = loss.compute(kern.parameters, X, labels=labels) if loss._quantum_kernel is not None else None
    res = opt.run_optimization(X, labels)
    assert opt.is_fitted and np.allclose(kern.parameters, res.x) and np.allclose(opt.get_optimal_parameters(), res.x)
    assert res.fun <= start + 1e-12
    assert opt.run_optimization(X, labels) is None          # already fitted (kernel_optimizer.py:66-67)


def test_qsvc_config_c1_matches_sklearn_on_the_oracle_kernel():
    """BASELINE configs[0]: FidelityKernel + QSVC, ChebyshevPQC(4 qubits, 2 layers), 100 synthetic points with
    labels sign(x0 * x1) — through the real FidelityKernel glue on the emulator engine."""
    from sklearn.svm import SVC

    import squlearn_b200 as sq
    from tests.emu_engine import EmuExecutor

    n, d = 4, 2
    X = np.random.default_rng(1).uniform(-0.95, 0.95, (100, d))
    y = np.sign(X[:, 0] * X[:, 1]).astype(int)
    Xt = np.random.default_rng(2).uniform(-0.95, 0.95, (30, d))
    fk = sq.FidelityKernel(sq.ChebyshevPQC(n, 2), EmuExecutor())
    model = sq.QSVC(fk, C=2.0).fit(X, y)
    gates = oracle.chebyshev_pqc(n, 2, d)
    p = oracle.chebyshev_pqc_initial_parameters(n, 2, d, seed=0)
    ref = SVC(kernel="precomputed", C=2.0).fit(oracle.fidelity_kernel(gates, n, X, None, p), y)
    k_test = oracle.fidelity_kernel(gates, n, Xt, X, p)
    assert np.array_equal(model.predict(Xt), ref.predict(k_test))
    assert np.max(np.abs(model.decision_function(Xt) - ref.decision_function(k_test))) < 1e-8
    assert model.get_params()["C"] == 2.0 and model.get_params()["quantum_kernel"] is fk
    pre = sq.QSVC("precomputed").fit(oracle.fidelity_kernel(gates, n, X, None, p), y)
    assert np.array_equal(pre.predict(k_test), sq.QSVC("precomputed", C=1.0).fit(
        oracle.fidelity_kernel(gates, n, X, None, p), y).predict(k_test))
