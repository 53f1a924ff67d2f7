"""The oracle restatements vs the fixtures recorded from the unmodified reference
(tests/golden/*.npz, produced by oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle.linear_ref import OracleLinear
from oracle.nonlinear_ref import OracleMLP, OracleNonlinear
from oracle import notreks_ref, simulate


def _edges(g, key):
    return tuple(tuple(int(x) for x in e) for e in g[key]) if key in g.files else None


@pytest.mark.parametrize("name,loss", [
    ("linear_l2_d20", "l2"), ("linear_l2_d64", "l2"), ("linear_l2_d7_masks", "l2"),
    ("linear_logistic_d12", "logistic"), ("linear_l2_d100", "l2"),
])
def test_linear_stages_bit_identical(golden, name, loss):
    g = golden(name)
    o = OracleLinear(loss).prepare(g["X"].copy(), float(g["lambda1"]), int(g["checkpoint"]),
                                   _edges(g, "exclude"), _edges(g, "include"))
    assert np.array_equal(o.cov, g["cov"])
    d = o.d
    W = np.zeros((d, d))
    for si, (mu, s, iters, lr) in enumerate(g["stages"]):
        o.trace = []
        W, ok = o.minimize(W.copy(), mu, int(iters), s, lr)
        assert ok == bool(g[f"ok_{si}"])
        assert o.last_iters == int(g[f"iters_{si}"])
        G = g[f"G_first_{si}"]
        for k in range(G.shape[0]):
            assert np.array_equal(o.trace[k]["Gobj"], G[k])
        assert np.array_equal(W, g[f"W_after_{si}"])
        obj, score, h, _ = o.func(W, mu, s)
        assert obj == float(g[f"obj_{si}"]) and score == float(g[f"score_{si}"]) and h == float(g[f"h_{si}"])
    sv, sg = o.score(W)
    hv, hg = o.h(W, 0.9)
    assert sv == float(g["score_val"]) and np.array_equal(sg, g["score_grad"])
    assert hv == float(g["h_val"]) and np.array_equal(hg, g["h_grad"])


def test_linear_full_fit_c1(golden):
    g = golden("fit_c1_seed0")
    o = OracleLinear("l2")
    W = o.fit(g["X"].copy(), lambda1=float(g["lambda1"]))
    assert o.stage_iters == [int(c) for c, _ in g["minimize_calls"]]
    assert np.array_equal(W, g["W_est"])
    assert o.h_final == float(g["h_final"]) and o.score_final == float(g["score_final"])
    acc = simulate.count_accuracy(g["W_true"] != 0, W != 0)
    assert acc["shd"] <= 2


def test_linear_short_fit_c4(golden):
    g = golden("fit_c4_short")
    warm, mx, ck = (int(x) for x in g["fit_kw"])
    o = OracleLinear("l2")
    W = o.fit(g["X"].copy(), lambda1=float(g["lambda1"]), warm_iter=warm, max_iter=mx, checkpoint=ck)
    assert np.array_equal(W, g["W_est"])


@pytest.mark.parametrize("name", ["mlp_d7", "mlp_d40", "mlp_deep_d6", "mlp_deep4_d5", "mlp_lin_d6"])
def test_mlp_value_grad_and_steps(golden, name):
    g = golden(name)
    dims = [int(x) for x in g["dims"]]
    lambda1, lambda2, mu, s, lr = (float(x) for x in g["hyper"])
    params = {k[5:]: g[k] for k in g.files if k.startswith("init.")}
    m = OracleMLP(dims, params)
    X = g["X"]
    np.testing.assert_allclose(m.forward(X), g["X_hat"], rtol=1e-12, atol=1e-13)
    obj, score, h, grads = m.obj_and_grads(X, lambda1, mu, s)
    assert abs(obj - float(g["obj"])) <= 1e-12 * abs(float(g["obj"]))
    assert abs(score - float(g["score"])) <= 1e-12 * abs(float(g["score"]))
    assert abs(h - float(g["h"])) <= 1e-12 * max(abs(float(g["h"])), 1e-3)
    for k, gk in grads.items():
        ref = g["grad." + k].reshape(gk.shape)
        assert np.abs(gk - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1e-30), k
    np.testing.assert_allclose(m.fc1_to_adj(), g["adj"], rtol=1e-13, atol=0)
    steps = int(g["steps"])
    opt = OracleNonlinear(m)
    ok = opt.minimize(X, steps, lr, lambda1, lambda2, mu, s, checkpoint=10 ** 9)
    assert ok == bool(g["ok"])
    for k in params:
        ref = g[f"after{steps}.{k}"]
        assert np.abs(m.p[k] - ref).max() <= 1e-11 * max(np.abs(ref).max(), 1e-30), k


def test_mlp_short_fit(golden):
    g = golden("mlp_fit_d5")
    dims = [int(x) for x in g["dims"]]
    T, warm, mx, ck = (int(x) for x in g["kw"])
    params = {k[5:]: g[k] for k in g.files if k.startswith("init.")}
    m = OracleMLP(dims, params)
    OracleNonlinear(m).fit(g["X"], lambda1=0.02, lambda2=0.005, T=T, warm_iter=warm, max_iter=mx,
                           checkpoint=ck)
    np.testing.assert_allclose(m.fc1_to_adj(), g["W_raw"], rtol=0, atol=1e-9)


def test_notreks_logdet(golden):
    g = golden("notreks_logdet")
    for s in (1.0, 0.8):
        h, G = notreks_ref.logdet_acyc_value_gradA(g["A"], s)
        assert abs(h - float(g[f"h_s{s}"])) <= 1e-13
        np.testing.assert_allclose(G, g[f"G_s{s}"], rtol=1e-12, atol=1e-14)
    d = g["W"].shape[0]
    S = notreks_ref.indicator_from_pairs(g["pairs"], d)
    for version in ("DAG_learning", "exact_trek_graph"):
        pen, gW = notreks_ref.tcc_logdet_value_gradW(g["W"], S, w=0.7, version=version, s=1.0)
        assert abs(pen - float(g[f"tcc_pen_{version}"])) <= 1e-13
        np.testing.assert_allclose(gW, g[f"tcc_grad_{version}"], rtol=1e-11, atol=1e-14)
    assert float(g["noop_val"]) == 0.0 and not g["noop_grad"].any()


def test_known_answers():
    """Analytic KATs: h(0)=0, h(DAG)=0, 2-cycle closed form."""
    o = OracleLinear("l2")
    o.d, o.Id = 5, np.eye(5)
    assert o.h(np.zeros((5, 5)))[0] == 0.0
    W = np.triu(np.arange(25.0).reshape(5, 5) / 10, 1)
    P = np.eye(5)[[3, 0, 4, 1, 2]]
    assert abs(o.h(P @ W @ P.T, 1.0)[0]) < 1e-12
    a, b, s = 0.4, -0.7, 0.9
    o.d, o.Id = 2, np.eye(2)
    h, G = o.h(np.array([[0, a], [b, 0]]), s)
    assert abs(h - (-np.log(s * s - a * a * b * b) + 2 * np.log(s))) < 1e-14
    det = s * s - a * a * b * b
    np.testing.assert_allclose(G, [[0, 2 * a * b * b / det], [2 * b * a * a / det, 0]], rtol=1e-13)


def test_generators():
    X, W = simulate.config_c1(0)
    assert X.shape == (500, 20) and simulate.is_dag(W) and int((W != 0).sum()) == 40
    X, W, lam = simulate.config_c4_problem(5)
    assert X.shape == (1000, 64) and int((W != 0).sum()) == 256 and lam == 0.02
    X, W = simulate.make_linear_problem(50, 4, 10, "SF", "gauss", 0)
    assert simulate.is_dag(W) and 150 <= int((W != 0).sum()) <= 200
    X, W = simulate.config_c2(0, n=200, d=10)
    assert set(np.unique(X)) <= {0.0, 1.0}
    assert not simulate.is_dag(np.array([[0, 1.0], [1.0, 0]]))


@pytest.mark.parametrize("seq", ["inv", "log", "exp", "binom"])
@pytest.mark.parametrize("wn", ["a", "b"])
def test_pst_series_value_grad(golden, seq, wn):
    """Closed-form adjoints of every PST series / aggregation against the reference's autograd."""
    from oracle.notreks_ref import pst_value_grad
    g = golden("pst_series")
    W = g[f"W_{wn}"]
    for agg in ("mean", "sum", "max", "lse"):
        val, grad, H = pst_value_grad(W, g["pairs"], seq=seq, agg=agg)
        ref_v, ref_g = float(g[f"val_{wn}_{seq}_{agg}"]), g[f"grad_{wn}_{seq}_{agg}"]
        assert abs(val - ref_v) <= 1e-11 * max(abs(ref_v), 1e-3), (seq, agg)
        assert np.abs(grad - ref_g).max() <= 1e-10 * max(np.abs(ref_g).max(), 1e-6), (seq, agg)
    assert np.abs(H - g[f"H_{wn}_{seq}"]).max() <= 1e-11 * np.abs(g[f"H_{wn}_{seq}"]).max()
    if seq == "log" and wn == "a":
        val, grad, _ = pst_value_grad(W, g["pairs"], seq="log", agg="mean", K_log=5)
        assert abs(val - float(g["val_a_log_K5"])) <= 1e-12
        assert np.abs(grad - g["grad_a_log_K5"]).max() <= 1e-12


def test_mi_tests_oracle(golden):
    """HSIC / dCor permutation tests restated in numpy against the unmodified reference."""
    from oracle import mi_ref
    g = golden("mi_tests")
    X = g["X"]
    pairs = [tuple(int(v) for v in p) for p in g["pairs"]]
    for test in ("hsic", "dcor"):
        res = mi_ref.pairwise(X, pairs, test=test, num_perm=40, seed=3)
        np.testing.assert_allclose([r[2] for r in res], g[f"{test}_stat"], rtol=1e-12, atol=1e-15)
        assert np.array_equal(np.array([r[3] for r in res]), g[f"{test}_p"])
    assert abs(mi_ref.hsic_stat(X[:, 0], X[:, 1], 0.8, 1.3) - float(g["hsic_01_sig"])) <= 1e-15
    assert abs(mi_ref.dcor_stat(X[:, 0], X[:, 3]) - float(g["dcor_03"])) <= 1e-15
