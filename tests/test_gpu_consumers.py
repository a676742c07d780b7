"""GPU tests (-m gpu) of the consumers of the hot path with the kernel matrices produced by
the CUDA engine and kept in HBM: QSVC at BASELINE configs[0], QGPR, QKRR with every kind of kernel object,
the kernel losses (TargetAlignment with the sums fused into the Gram epilogue, NLL) and KernelOptimizer.  Each
against a NumPy / SciPy / scikit-learn restatement of the cited reference lines on the ORACLE's kernel matrix."""

import numpy as np
import pytest
import scipy.linalg

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sq():
    import squlearn_b200

    return squlearn_b200


@pytest.fixture(scope="module")
def executor(sq):
    return sq.Executor("b200")


def _c1(rng_seed=1, n_points=100):
    X = np.random.default_rng(rng_seed).uniform(-0.95, 0.95, (n_points, 2))
    return X, np.sign(X[:, 0] * X[:, 1]).astype(int)


def _oracle_k(x, y=None, n=4, d=2):
    return oracle.fidelity_kernel(oracle.chebyshev_pqc(n, 2, d), n, x, y,
                                  oracle.chebyshev_pqc_initial_parameters(n, 2, d, seed=0))


def test_qsvc_config_c1_on_the_gpu_kernel(sq, executor):
    """BASELINE configs[0]: FidelityKernel + QSVC, ChebyshevPQC(4 qubits, 2 layers), 100 synthetic points
    (src/squlearn/kernel/qsvc.py:89-147): labels and decision values equal scikit-learn's SVC on the oracle kernel."""
    from sklearn.svm import SVC

    X, y = _c1()
    Xt = np.random.default_rng(2).uniform(-0.95, 0.95, (30, 2))
    fk = sq.FidelityKernel(sq.ChebyshevPQC(4, 2), executor)
    model = sq.QSVC(fk, C=2.0).fit(X, y)
    ref = SVC(kernel="precomputed", C=2.0).fit(_oracle_k(X), y)
    k_test = _oracle_k(Xt, X)
    assert np.array_equal(model.predict(Xt), ref.predict(k_test))
    assert np.max(np.abs(model.decision_function(Xt) - ref.decision_function(k_test))) < 1e-8
    assert model.score(X, y) == ref.score(_oracle_k(X), y)


def test_qgpr_and_full_gram_regularisation_on_the_gpu(sq, executor):
    """qgpr.py:177-283 + regularization.py:49-70 (eigenvalue clipping of the FULL Gram matrix, device eigh) against
    the SciPy restatement."""
    rng = np.random.default_rng(5)
    X, Xt = rng.uniform(-0.9, 0.9, (60, 2)), rng.uniform(-0.9, 0.9, (17, 2))
    y = np.cos(2 * X[:, 0]) - X[:, 1]
    fk = sq.FidelityKernel(sq.ChebyshevPQC(4, 2), executor)
    k_train, k_tt, k_test = _oracle_k(X), _oracle_k(Xt, X), _oracle_k(Xt)
    for full_reg in (True, False):
        model = sq.QGPR(fk, sigma=1e-5, full_regularization=full_reg).fit(X, y)
        mean, cov = model.predict(Xt, return_cov=True)
        kt, ktt, kte = k_train, k_tt, k_test
        if full_reg:
            full = np.block([[kt, ktt.T], [ktt, kte]])
            evals, evecs = scipy.linalg.eig(full)
            rec = np.real(evecs @ np.diag(evals.clip(min=0)) @ evecs.T)
            kt, ktt, kte = rec[:60, :60], rec[60:, :60], rec[60:, 60:]
        lu = scipy.linalg.lu_factor(kt + 1e-5 * np.identity(60))
        want_mean = ktt.dot(scipy.linalg.lu_solve(lu, y))
        want_cov = kte - ktt @ scipy.linalg.lu_solve(lu, ktt.T)
        assert np.max(np.abs(mean - want_mean)) < 1e-6
        assert np.max(np.abs(cov - want_cov)) < 1e-6


def test_qkrr_accepts_every_kernel_object(sq, executor):
    """qkrr.py:99-199: FidelityKernel, ProjectedQuantumKernel, a KernelOptimizer wrapper and "precomputed"."""
    rng = np.random.default_rng(6)
    X, Xt = rng.uniform(-0.9, 0.9, (50, 2)), rng.uniform(-0.9, 0.9, (11, 2))
    y = np.sin(3 * X[:, 0]) + X[:, 1] ** 2

    def ridge(k, kt, alpha):
        L = scipy.linalg.cholesky(k + alpha * np.eye(len(k)), lower=True)
        return kt @ scipy.linalg.cho_solve((L, True), y)

    fk = sq.FidelityKernel(sq.ChebyshevPQC(4, 2), executor)
    got = sq.QKRR(fk, alpha=1e-4).fit(X, y).predict(Xt)
    assert np.max(np.abs(got - ridge(_oracle_k(X), _oracle_k(Xt, X), 1e-4))) < 1e-7
    pqk = sq.ProjectedQuantumKernel(sq.HubregtsenEncodingCircuit(4, 1), executor, gamma=0.7)
    got = sq.QKRR(pqk, alpha=1e-3).fit(X, y).predict(Xt)
    p = oracle.hubregtsen_initial_parameters(4, 1, seed=0)
    gates = oracle.hubregtsen(4, 1, 2)
    f, ft = oracle.projected_kernel_features(gates, 4, X, p), oracle.projected_kernel_features(gates, 4, Xt, p)
    want = ridge(oracle.gaussian_outer_kernel(f, f, 0.7), oracle.gaussian_outer_kernel(ft, f, 0.7), 1e-3)
    assert np.max(np.abs(got - want)) < 1e-7
    wrapped = sq.KernelOptimizer(sq.FidelityKernel(sq.ChebyshevPQC(4, 2), executor), sq.TargetAlignment(),
                                 sq.ScipyMinimize("Nelder-Mead", {"maxiter": 3}))
    wrapped._is_fitted = True                                   # skip the optimisation: same kernel as fk
    wrapped._quantum_kernel._initialize_kernel(2)
    got = sq.QKRR(wrapped, alpha=1e-4).fit(X, y).predict(Xt)
    assert np.max(np.abs(got - ridge(_oracle_k(X), _oracle_k(Xt, X), 1e-4))) < 1e-7
    pre = sq.QKRR("precomputed", alpha=1e-4).fit(_oracle_k(X), y)
    assert np.max(np.abs(pre.predict(_oracle_k(Xt, X)) - ridge(_oracle_k(X), _oracle_k(Xt, X), 1e-4))) < 1e-9


def test_target_alignment_fused_and_nll_on_the_gpu(sq, executor):
    """target_alignment.py:42-90 (sums fused into the INT8 Gram epilogue for a 12-qubit kernel, plain device
    reductions for the 4-qubit one) and negative_log_likelihood.py:43-90."""
    for n, n_points in ((12, 1100), (4, 64)):
        d = 2
        X = np.random.default_rng(8).uniform(-0.9, 0.9, (n_points, d))
        labels = np.where(X[:, 0] * X[:, 1] > 0, 1, -1)
        enc = sq.ChebyshevPQC(n, 2)
        fk = sq.FidelityKernel(enc, executor)
        p = np.random.default_rng(9).uniform(-1, 1, enc.num_parameters)
        k = oracle.fidelity_kernel(oracle.chebyshev_pqc(n, 2, d), n, X, None, p, pairloop=False)
        for rescale in (True, False):
            ta = sq.TargetAlignment(rescale_class_labels=rescale)
            ta.set_quantum_kernel(fk)
            got = ta.compute(p, X, labels=labels)
            if rescale:
                nplus = np.count_nonzero(labels == 1)
                yy = np.array([v / nplus if v == 1 else v / (len(labels) - nplus) for v in labels])
            else:
                yy = labels.astype(float)
            T = np.outer(yy, yy)
            want = -np.sum(k * T) / np.sqrt(np.sum(k * k) * np.sum(T * T))
            assert abs(got - want) < 1e-9, (n, rescale)
        # the fused path returns K as well, and it is the evaluate() matrix
        fused = fk.evaluate_alignment(X, labels.astype(float))
        assert fused is not None and np.max(np.abs(fused[0].cpu().numpy() - k)) < 1e-10
    y = np.sin(X[:, 0])
    nll = sq.NLL(sigma=1e-3)
    nll.set_quantum_kernel(fk)
    got = nll.compute(p, X, labels=y)
    L = scipy.linalg.cholesky(k + 1e-3 * np.eye(len(X)), lower=True)
    s2 = scipy.linalg.cho_solve((L, True), y)
    want = np.sum(np.log(np.diagonal(L))) + 0.5 * y @ s2 + 0.5 * len(X) * np.log(2.0 * np.pi)
    assert abs(got[0] - want) < 1e-7


def test_kernel_optimizer_on_the_gpu(sq, executor):
    """kernel_optimizer.py:14-186: every optimiser step re-evaluates the Gram with new parameters (only the angle
    coefficients are updated: no re-planning)."""
    X = np.random.default_rng(9).uniform(-0.9, 0.9, (80, 1))
    labels = np.where(X[:, 0] > 0, 1, -1)
    fk = sq.FidelityKernel(sq.ChebyshevPQC(3, 1), executor)
    loss = sq.TargetAlignment()
    opt = sq.KernelOptimizer(fk, loss, sq.ScipyMinimize("Nelder-Mead", {"maxiter": 40}))
    fk._initialize_kernel(1)
    start = loss.compute(fk.parameters, X, labels=labels)
    launches0 = executor.engine.launch_count()
    res = opt.run_optimization(X, labels)
    assert opt.is_fitted and np.allclose(fk.parameters, res.x)
    assert res.fun <= start + 1e-12
    assert executor.engine.launch_count() > launches0
    assert opt.run_optimization(X, labels) is None
