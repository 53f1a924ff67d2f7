"""SURVEY.md 8f4 on the GPU: HSIC / distance-correlation permutation tests (csrc/mi.cu) against outputs of the
unmodified reference (oracle/make_golden_mi.py -> tests/golden/mi_tests.npz) and against the numpy oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("test", ["hsic", "dcor"])
def test_pairwise_tests_vs_reference(golden, test):
    from midagma_b200 import mi_tests as mi
    g = golden("mi_tests")
    X = g["X"]
    pairs = [tuple(int(v) for v in p) for p in g["pairs"]]
    res = mi.test_pairwise_independence(X, pairs, test=test, num_perm=40, seed=3)
    assert [(r.i, r.j) for r in res] == pairs
    stat = np.array([r.stat for r in res])
    ref = g[f"{test}_stat"]
    assert np.abs(stat - ref).max() <= 1e-11 * max(np.abs(ref).max(), 1e-6), np.abs(stat - ref).max()
    assert np.array_equal(np.array([r.pvalue for r in res]), g[f"{test}_p"])        # same permutations, same counts
    I = mi.get_I_from_full_pairwise_tests(X, alpha=0.05, test=test, num_perm=25, seed=1)
    assert np.array_equal(I, g[f"{test}_I"])
    I2 = mi.get_I_from_full_pairwise_tests(X, alpha=0.2, test=test, num_perm=10, seed=2, bonferroni=False,
                                           undirected=False)
    assert np.array_equal(I2, g[f"{test}_I_dir"])


def test_stat_functions_and_permutation_pvalue(golden):
    from midagma_b200 import mi_tests as mi
    g = golden("mi_tests")
    X = g["X"]
    assert abs(mi.hsic_stat(X[:, 0], X[:, 1]) - float(g["hsic_01"])) <= 1e-13
    assert abs(mi.hsic_stat(X[:, 0], X[:, 1], sigma_x=0.8, sigma_y=1.3) - float(g["hsic_01_sig"])) <= 1e-13
    assert abs(mi.dcor_stat(X[:, 0], X[:, 3]) - float(g["dcor_03"])) <= 1e-12
    assert mi.dcor_stat(X[:, 0], X[:, 5]) == 0.0 and mi.hsic_stat(X[:, 0], X[:, 5]) == 0.0      # constant column
    s, p = mi.permutation_pvalue(mi.hsic_stat, X[:, 0], X[:, 2], num_perm=50, rng=np.random.default_rng(5))
    assert abs(s - g["perm_hsic_02"][0]) <= 1e-13 and p == g["perm_hsic_02"][1]
    s, p = mi.permutation_pvalue(mi.dcor_stat, X[:, 1], X[:, 3], num_perm=50)
    assert abs(s - g["perm_dcor_13"][0]) <= 1e-12 and p == g["perm_dcor_13"][1]
    with pytest.raises(ValueError):
        mi.test_pairwise_independence(X, [(0, 1)], test="nope")
    assert mi.test_pairwise_independence(X, [], test="hsic") == []


def test_larger_n_against_oracle():
    """n = 700 (several row chunks, permutation larger than one block stride) against the numpy restatement."""
    from midagma_b200 import mi_tests as mi
    from oracle import mi_ref
    rng = np.random.default_rng(2)
    n = 700
    x = rng.standard_normal(n)
    X = np.column_stack([x, np.cos(2 * x) + 0.3 * rng.standard_normal(n), rng.standard_normal(n)])
    pairs = [(0, 1), (0, 2), (2, 1)]
    for test in ("hsic", "dcor"):
        got = mi.test_pairwise_independence(X, pairs, test=test, num_perm=12, seed=7)
        ref = mi_ref.pairwise(X, pairs, test=test, num_perm=12, seed=7)
        for r, q in zip(got, ref):
            assert abs(r.stat - q[2]) <= 1e-11 * max(abs(q[2]), 1e-6) and r.pvalue == q[3], (test, r, q)


@pytest.mark.parametrize("test", ["pearson", "spearman"])
def test_analytic_tests_vs_reference(golden, test):
    """Pearson / Spearman: correlations from the fused covariance kernel, scipy's closed-form p-values."""
    from midagma_b200 import mi_tests as mi
    g = golden("mi_tests")
    X = g["X"]
    pairs = [tuple(int(v) for v in p) for p in g["pairs"]]
    res = mi.test_pairwise_independence(X, pairs, test=test)
    stat, p = np.array([r.stat for r in res]), np.array([r.pvalue for r in res])
    np.testing.assert_allclose(stat, g[f"{test}_stat"], rtol=1e-11, atol=1e-14, equal_nan=True)
    np.testing.assert_allclose(p, g[f"{test}_p"], rtol=1e-9, atol=1e-300, equal_nan=True)
    if test == "spearman":
        res = mi.test_pairwise_independence(g["X_ties"], [q for q in pairs if q[1] < 5], test="spearman")
        np.testing.assert_allclose([r.stat for r in res], g["spearman_ties_stat"], rtol=1e-11, atol=1e-14)
        np.testing.assert_allclose([r.pvalue for r in res], g["spearman_ties_p"], rtol=1e-9)
        s, pv = mi.spearman_stat_pvalue(X[:, 0], X[:, 3])
        k = pairs.index((0, 3))
        assert abs(s - g["spearman_stat"][k]) <= 1e-12 and abs(pv - g["spearman_p"][k]) <= 1e-9 * g["spearman_p"][k]
    else:
        s, pv = mi.pearson_stat_pvalue(X[:, 0], X[:, 3])
        k = pairs.index((0, 3))
        assert abs(s - g["pearson_stat"][k]) <= 1e-12 and abs(pv - g["pearson_p"][k]) <= 1e-9 * g["pearson_p"][k]
