"""Pin the oracle against the reference's own known-answer tests.

Every expected value comes from tests/golden/reference_goldens.json, which transcribes numbers the
reference asserts against its PennyLane / Qiskit / Qulacs statevector paths.
"""

import numpy as np
import pytest

import oracle
from oracle.statevector import Gate


def _p(k):
    return lambda x, p: p[k]


def test_qnnr_predict_golden(goldens, regression_data):
    g = goldens["qnnr_predict"]
    X, _ = regression_data
    gates = oracle.chebyshev_rx(2, 1, num_features=1)
    out = oracle.qnn_evaluate(gates, 2, X, np.linspace(0.1, 0.4, 4), ("summed_paulis", {}),
                              np.linspace(0.1, 0.3, 3), ("f",))
    assert out["f"].shape == (6,)
    assert np.allclose(out["f"], g["expected"])
    assert np.max(np.abs(out["f"] - np.array(g["expected"]))) < 5e-8


def test_optree_gradient_golden(goldens):
    g = goldens["optree_gradient"]
    gates = [
        Gate("rx", (0,), (lambda x, p: p[0] * x[:, 0],)),
        Gate("rx", (1,), (lambda x, p: p[1] * x[:, 0],)),
        Gate("ry", (0,), (_p(2),)),
        Gate("ry", (1,), (_p(3),)),
        Gate("rxx", (0, 1), (lambda x, p: p[0] * x[:, 0],)),
    ]
    obs = ("custom", {"operator_string": g["observable"]})
    x = np.array([[g["x"]]])
    p = np.array(g["p"])
    for route in ("generator", "parameter_shift"):
        out = oracle.qnn_evaluate(gates, 2, x, p, obs, [], ("dfdp",), gradient=route)
        assert out["dfdp"].shape == (1, 4)
        assert np.allclose(out["dfdp"][0], g["reference_grad"]), route


def test_optree_evaluate_golden(goldens):
    g = goldens["optree_evaluate"]
    c1 = [Gate("h", (0,), ()), Gate("cx", (0, 1), ()), Gate("rz", (0,), (0.5,)),
          Gate("ry", (1,), (0.5,))]
    c2 = [Gate("h", (0,), ()), Gate("cp", (0, 1), (0.5,)), Gate("s", (0,), ()),
          Gate("t", (1,), ()), Gate("h", (0,), ()), Gate("h", (1,), ())]
    x = np.zeros((1, 1))
    for key_op, key_exp in (("z_operator", "z_expected"), ("xy_operator", "xy_expected")):
        vals = []
        for circ in (c1, c2):
            sv = oracle.simulate(circ, 2, x)
            vals.append(sum(c * oracle.pauli_expectation(sv, s)[0]
                            for s, c in zip(g[key_op]["strings"], g[key_op]["coeffs"])))
        assert np.allclose(vals, g[key_exp])


def test_qrc_regressor_golden(goldens, regression_data):
    from sklearn.linear_model import LinearRegression

    g = goldens["qrc_regressor_linear"]
    X, y = regression_data
    gates = oracle.hubregtsen(2, 1, num_features=1)
    p = oracle.hubregtsen_initial_parameters(2, 1, seed=0)
    assert np.allclose(p, np.random.RandomState(0).uniform(-np.pi, np.pi, 2))
    strings = oracle.random_pauli_list(2, 16, seed=0)
    obs = [("custom", {"operator_string": s}) for s in strings]
    feats = oracle.qnn_evaluate(gates, 2, X, p, obs, [], ("f",))["f"]
    assert feats.shape == (6, 16)
    pred = LinearRegression().fit(feats, y).predict(feats)
    assert np.allclose(pred, g["expected"], atol=g["atol"])


def _fqk_single(x, y, p):
    c = np.cos
    return (0.375 * (1 - c(p)) ** 2 * c(x - y) + 0.125 * (1 - c(p)) ** 2 * c(x + y)
            - 0.5 * (1 - c(p)) ** 2 - 1.0 * c(p) - 0.3125 * c(x - y) - 0.1875 * c(x + y)
            - 0.03125 * c(-2 * p + x + y) + 0.125 * c(-p + x + y) + 0.375 * c(p - x + y)
            + 0.375 * c(p + x - y) + 0.125 * c(p + x + y) + 0.03125 * c(2 * p - x + y)
            + 0.03125 * c(2 * p + x - y) - 0.03125 * c(2 * p + x + y) + 1.5)


def test_fqk_closed_forms(goldens):
    g = goldens["fqk_closed_form"]
    s = g["single"]
    gates = oracle.layered_ry_rx(1, 1, first="ry")
    k = oracle.fidelity_kernel(gates, 1, np.array([[s["x"]]]), np.array([[s["y"]]]),
                               np.array([s["p"]]), evaluate_duplicates="all")
    assert abs(k[0, 0] - _fqk_single(s["x"], s["y"], s["p"])) < 1e-12
    m = g["multi"]
    x0, x1 = m["x"]
    y0, y1 = m["y"]
    expect = (0.25 * np.cos(x0 - y0) + 0.25 * np.cos(x1 - y1) + 0.125 * np.cos(x0 - x1 - y0 + y1)
              + 0.125 * np.cos(x0 + x1 - y0 - y1) + 0.25)
    gates = oracle.layered_ry_rx(2, 2, first="rx")
    k = oracle.fidelity_kernel(gates, 2, np.array([m["x"]]), np.array([m["y"]]), np.array(m["p"]),
                               evaluate_duplicates="all")
    assert abs(k[0, 0] - expect) < 1e-12


def test_pqk_closed_forms(goldens):
    g = goldens["pqk_closed_form"]
    s = g["single"]
    gates = oracle.layered_ry_rx(1, 1)
    k = oracle.projected_quantum_kernel(gates, 1, np.array([[s["x"]]]), np.array([[s["y"]]]),
                                        np.array([s["p"]]), gamma=s["gamma"])
    expect = np.exp(-2 * s["gamma"] * (1 - np.cos(s["x"] - s["y"])) * np.cos(s["p"]) ** 2)
    assert abs(k[0, 0] - expect) < 1e-12
    m = g["multi"]
    (x0, x1), (y0, y1), (p0, p1), gam = m["x"], m["y"], m["p"], m["gamma"]
    expect = np.exp(-2.0 * gam * (-np.cos(p0) ** 2 * np.cos(x0 - y0) + np.cos(p0) ** 2
                                  - np.cos(p1) ** 2 * np.cos(x1 - y1) + np.cos(p1) ** 2))
    gates = oracle.layered_ry_rx(2, 2)
    k = oracle.projected_quantum_kernel(gates, 2, np.array([m["x"]]), np.array([m["y"]]),
                                        np.array(m["p"]), gamma=gam)
    assert abs(k[0, 0] - expect) < 1e-12


def test_highdim_golden(goldens):
    g = goldens["highdim_all_zero"]
    gates = oracle.highdim(2, 1, num_layers=1)
    X = np.array(g["X"]).reshape(-1, 1)
    f = oracle.qnn_evaluate(gates, 2, X, [], ("single_pauli", {"qubit": 0, "op_str": "Z"}), [],
                            ("f",))["f"]
    assert np.allclose(f, g["expected"], atol=g["atol"])


def test_executor_smoke_goldens(goldens):
    g = goldens["executor_smoke"]
    gates = [Gate("x", (0,), ()), Gate("x", (1,), ())]
    sv = oracle.simulate(gates, 2, np.zeros((1, 1)))
    assert np.allclose(np.abs(sv[0]) ** 2, g["x_x_state_probs"])
    assert np.isclose(oracle.pauli_expectation(sv, "ZZ")[0], g["zz_expectation_after_xx"])
    gates = [Gate("ry", (0,), (np.pi,)), Gate("ry", (1,), (np.pi,))]
    sv = oracle.simulate(gates, 2, np.zeros((1, 1)))
    assert np.allclose(np.abs(sv[0]) ** 2, [0, 0, 0, 1])


def test_gram_assembly_semantics(goldens):
    g = goldens["gram_assembly"]
    sv = np.array(g["states"], dtype=np.complex128)
    k = oracle.fidelity_gram_pairloop(sv, sv, True, "off_diagonal")
    assert k == pytest.approx(np.array(g["symmetric_off_diagonal"]), abs=1e-12)
    k2 = oracle.fidelity_gram_pairloop(sv, sv[:1], False, "off_diagonal")
    assert k2 == pytest.approx(np.array(g["rectangular_vs_first"]), abs=1e-12)
    # reference quirk: "all" leaves the symmetric diagonal at zero (never computed)
    k3 = oracle.fidelity_gram_pairloop(sv, sv, True, "all")
    assert np.all(np.diag(k3) == 0.0)
    # "none" short-circuits identical states to 1
    same = np.array([[1, 0], [1, 0]], dtype=np.complex128)
    assert np.all(oracle.fidelity_gram_pairloop(same, same, True, "none") == 1.0)


def test_pairloop_equals_zgemm_and_reference_style():
    rng = np.random.default_rng(7)
    n, d, N = 5, 2, 12
    X = rng.uniform(-0.95, 0.95, (N, d))
    p = oracle.chebyshev_pqc_initial_parameters(n, 2, d, seed=0)
    gates = oracle.chebyshev_pqc(n, 2, d)
    a = oracle.simulate(gates, n, X, p)
    b = oracle.simulate_reference_style(gates, n, X, p)
    assert np.max(np.abs(a - b)) < 1e-14
    assert np.allclose(np.sum(np.abs(a) ** 2, axis=1), 1.0, atol=1e-13)
    k1 = oracle.fidelity_gram_pairloop(a, a, True)
    k2 = oracle.fidelity_gram_zgemm(a, a, True)
    assert np.max(np.abs(k1 - k2)) < 1e-14
    # PennyLane ordering is the bit reversal of the little-endian index
    pl = oracle.to_pennylane_order(a, n)
    idx = np.arange(2**n)
    rev = np.array([int(format(i, f"0{n}b")[::-1], 2) for i in idx])
    assert np.array_equal(pl[:, rev], a)


def test_circuit_parameter_counts_and_init():
    # ChebyshevPQC(4, 2): P = 2n + nL + nL = 24 gates of SURVEY config C1
    assert oracle.chebyshev_pqc_num_parameters(4, 2) == 24
    assert len(oracle.chebyshev_pqc(4, 2, 2)) == 24
    assert oracle.chebyshev_pqc_num_parameters(16, 2) == 96
    assert oracle.hubregtsen_num_parameters(10, 1) == 20
    assert len(oracle.hubregtsen(10, 1, 10)) == 40
    assert oracle.chebyshev_rx_num_parameters(8, 4) == 64
    assert len(oracle.chebyshev_rx(8, 4, 2)) == 92
    p = oracle.chebyshev_pqc_initial_parameters(4, 2, 2, seed=0)
    lin = np.linspace(0.01, 4.0, 2)
    assert np.allclose(p[4:8], [lin[0], lin[1], lin[0], lin[1]])
    p = oracle.chebyshev_rx_initial_parameters(8, 4, 2, seed=0)
    lin = np.linspace(0.01, 4.0, 4)
    assert np.allclose(p[16:24], np.tile(lin, 2))


def test_gradient_routes_agree_with_controlled_rotations():
    rng = np.random.default_rng(3)
    n, d, N = 4, 2, 5
    X = rng.uniform(-0.9, 0.9, (N, d))
    p = rng.uniform(-1, 1, oracle.chebyshev_pqc_num_parameters(n, 1))
    gates = oracle.chebyshev_pqc(n, 1, d)  # contains crz -> four-term rule
    obs = [("summed_paulis", {}), ("single_pauli", {"qubit": 1, "op_str": "X"})]
    pobs = rng.uniform(-1, 1, n + 1)
    a = oracle.qnn_evaluate(gates, n, X, p, obs, pobs, ("f", "dfdp", "dfdop", "var"))
    b = oracle.qnn_evaluate(gates, n, X, p, obs, pobs, ("dfdp",), gradient="parameter_shift")
    assert a["dfdp"].shape == (N, 2, len(p))
    assert np.max(np.abs(a["dfdp"] - b["dfdp"])) < 1e-12
    # finite-difference cross-check of dfdp and dfdop
    eps = 1e-6
    for j in (0, 5, len(p) - 1):
        e = np.zeros_like(p)
        e[j] = eps
        fp = oracle.qnn_evaluate(gates, n, X, p + e, obs, pobs, ("f",))["f"]
        fm = oracle.qnn_evaluate(gates, n, X, p - e, obs, pobs, ("f",))["f"]
        assert np.allclose((fp - fm) / (2 * eps), a["dfdp"][:, :, j], atol=1e-8)
    e = np.zeros_like(pobs)
    e[2] = eps
    fp = oracle.qnn_evaluate(gates, n, X, p, obs, pobs + e, ("f",))["f"]
    fm = oracle.qnn_evaluate(gates, n, X, p, obs, pobs - e, ("f",))["f"]
    assert np.allclose((fp - fm) / (2 * eps), a["dfdop"][:, :, 2], atol=1e-8)
    assert np.all(a["var"] > -1e-12)
