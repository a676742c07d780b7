"""The QNN host glue (value / gradient assembly, shapes, the shifted-variant batch for parameters AND for
features) on the CPU: the real package code runs on the emulator engine of tests/emu_engine.py and is checked
against the oracle, finite differences and the reference's golden values."""

import json
import os

import numpy as np
import pytest

import oracle
import squlearn_b200 as sq
from tests.emu_engine import EmuExecutor

GOLDENS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")))


def test_optree_golden_gradient_and_input_derivative():
    """tests/util/optree/test_optree_derivative.py:49-96: grad_p AND d/dx of IZ + ZI."""
    g = GOLDENS["optree_gradient"]
    A = sq.Angle
    prog = sq.Program(2, [
        sq.GateOp("rx", (0,), A(param=0, feature=0)), sq.GateOp("rx", (1,), A(param=1, feature=0)),
        sq.GateOp("ry", (0,), A(param=2)), sq.GateOp("ry", (1,), A(param=3)),
        sq.GateOp("rxx", (0, 1), A(param=0, feature=0))], 4, 1)
    circ = sq.B200Circuit(prog, sq.CustomObservable(2, g["observable"]))
    ex = EmuExecutor()
    kw = dict(param=np.array(g["p"]), x=np.array([[g["x"]]]))
    grad = ex.b200_execute(sq.b200_gradient, circ, **kw)
    assert np.allclose(grad[0, 0], g["reference_grad"])
    dx = ex.b200_execute(sq.b200_gradient, circ, wrt="x", **kw)
    assert dx.shape == (1, 1, 1) and np.allclose(dx[0, 0], g["reference_dx"])


@pytest.mark.parametrize("nonlin", ["arccos", "arctan"])
def test_qnn_values_gradients_and_dfdx_on_the_emulator(nonlin):
    n, d, N = 4, 2, 5
    rng = np.random.default_rng(3)
    X = rng.uniform(-0.8, 0.8, (N, d))
    enc = sq.ChebyshevPQC(n, 1, nonlinearity=nonlin)
    p = rng.uniform(-1, 1, enc.num_parameters)
    pop = rng.uniform(-1, 1, n + 1)
    qnn = sq.LowLevelQNN(enc, sq.SummedPaulis(n), EmuExecutor(), d)
    r = qnn.evaluate(X, p, pop, "f", "dfdp", "dfdop", "dfdx")
    gates = oracle.chebyshev_pqc(n, 1, d, nonlinearity=nonlin)
    w = oracle.qnn_evaluate(gates, n, X, p, ("summed_paulis", {}), pop, ("f", "dfdp", "dfdop"))
    for key in ("f", "dfdp", "dfdop"):
        assert r[key].shape == w[key].shape and np.max(np.abs(r[key] - w[key])) < 1e-11, key
    # dfdx against central differences of the oracle value
    assert r["dfdx"].shape == (N, d)
    h = 1e-6
    for j in range(d):
        e = np.zeros(d)
        e[j] = h
        fp = oracle.qnn_evaluate(gates, n, X + e, p, ("summed_paulis", {}), pop, ("f",))["f"]
        fm = oracle.qnn_evaluate(gates, n, X - e, p, ("summed_paulis", {}), pop, ("f",))["f"]
        assert np.max(np.abs(r["dfdx"][:, j] - (fp - fm) / (2 * h))) < 1e-7


def test_dfdx_shapes_multi_output_and_linear_encoding():
    n, d = 3, 3
    rng = np.random.default_rng(4)
    X = rng.uniform(-1, 1, (4, d))
    enc = sq.HubregtsenEncodingCircuit(n, num_features=d, num_layers=1)   # rz(x) / rx(x): linear encodings
    p = rng.uniform(-1, 1, enc.num_parameters)
    obs = [sq.SinglePauli(n, 0, "Z"), sq.SinglePauli(n, 1, "X")]
    qnn = sq.LowLevelQNN(enc, obs, EmuExecutor(), d)
    r = qnn.evaluate(X, p, [], "f", "dfdx")
    assert r["f"].shape == (4, 2) and r["dfdx"].shape == (4, 2, d)
    h = 1e-6
    for j in range(d):
        e = np.zeros(d)
        e[j] = h
        fd = (qnn.evaluate(X + e, p, [], "f")["f"] - qnn.evaluate(X - e, p, [], "f")["f"]) / (2 * h)
        assert np.max(np.abs(r["dfdx"][:, :, j] - fd)) < 1e-7
    single = qnn.evaluate(X[0], p, [], "dfdx")["dfdx"]
    assert single.shape == (2, d) and np.allclose(single, r["dfdx"][0])
    with pytest.raises(NotImplementedError):
        qnn.evaluate(X, p, [], "dfdxdxdp")


def test_squared_operator_second_derivatives_in_the_operator_parameters():
    """"dfccdopdop" and "dfccdopdx" (evaluation_classes.py:177-193): <O^2> is a quadratic form in the observable
    parameters, so its Hessian is constant in them, symmetric, and reproduces <O^2> and "dfccdop" exactly; the mixed
    derivative is checked against central differences of "dfccdop" in x."""
    n, d, N = 3, 2, 4
    rng = np.random.default_rng(8)
    X = rng.uniform(-0.8, 0.8, (N, d))
    enc = sq.ChebyshevPQC(n, 1)
    p = rng.uniform(-1, 1, enc.num_parameters)
    pop = rng.uniform(-1, 1, n + 1)
    qnn = sq.LowLevelQNN(enc, sq.SummedPaulis(n), EmuExecutor(), d, caching=False)
    r = qnn.evaluate(X, p, pop, "fcc", "dfccdop", "dfccdopdop", "dfccdopdx")
    hess = r["dfccdopdop"]
    assert hess.shape == (N, n + 1, n + 1) and np.allclose(hess, np.swapaxes(hess, -1, -2))
    # quadratic form: <O^2>(c) = 1/2 c^T H c and grad = H c
    assert np.max(np.abs(0.5 * np.einsum("j,njk,k->n", pop, hess, pop) - r["fcc"])) < 1e-11
    assert np.max(np.abs(np.einsum("njk,k->nj", hess, pop) - r["dfccdop"])) < 1e-11
    pop2 = rng.uniform(-1, 1, n + 1)
    assert np.max(np.abs(qnn.evaluate(X, p, pop2, "dfccdopdop")["dfccdopdop"] - hess)) < 1e-11
    mixed = r["dfccdopdx"]
    assert mixed.shape == (N, n + 1, d)
    h = 1e-6
    for j in range(d):
        e = np.zeros(d)
        e[j] = h
        fd = (qnn.evaluate(X + e, p, pop, "dfccdop")["dfccdop"] - qnn.evaluate(X - e, p, pop, "dfccdop")["dfccdop"]) / (2 * h)
        assert np.max(np.abs(mixed[:, :, j] - fd)) < 1e-7
