"""Generate tests/golden/mi_tests.npz by running the UNMODIFIED reference (build container only):

    python oracle/make_golden_mi.py

Covers SURVEY.md 8f4: src/notreks/mi_tests.py hsic_stat / dcor_stat / test_pairwise_independence /
get_I_from_full_pairwise_tests on a small nonlinear SEM.  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

import numpy as np  # noqa: E402

import notreks.mi_tests as ref  # noqa: E402  (reference)

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    rng = np.random.default_rng(11)
    n = 90
    x0 = rng.standard_normal(n)
    x1 = np.sin(3.0 * x0) + 0.15 * rng.standard_normal(n)
    x2 = rng.standard_normal(n)
    x3 = np.tanh(x0) + (x2 ** 2 - np.mean(x2 ** 2)) + 0.2 * rng.standard_normal(n)
    x4 = rng.standard_normal(n)
    x5 = np.full(n, 0.7)                                       # constant column: median 0 -> sigma^2 = 1, dvar = 0
    X = np.column_stack([x0, x1, x2, x3, x4, x5])
    out = {"X": X}
    d = X.shape[1]
    pairs = [(i, j) for i in range(d) for j in range(i + 1, d)]
    out["pairs"] = np.array(pairs)
    for test in ("hsic", "dcor"):
        res = ref.test_pairwise_independence(X, pairs, test=test, num_perm=40, seed=3)
        out[f"{test}_stat"] = np.array([r.stat for r in res])
        out[f"{test}_p"] = np.array([r.pvalue for r in res])
        out[f"{test}_I"] = ref.get_I_from_full_pairwise_tests(X, alpha=0.05, test=test, num_perm=25, seed=1)
        out[f"{test}_I_dir"] = ref.get_I_from_full_pairwise_tests(X, alpha=0.2, test=test, num_perm=10, seed=2,
                                                                   bonferroni=False, undirected=False)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                            # ConstantInputWarning for the constant column
        for test in ("pearson", "spearman"):
            res = ref.test_pairwise_independence(X, pairs, test=test)
            out[f"{test}_stat"] = np.array([r.stat for r in res])
            out[f"{test}_p"] = np.array([r.pvalue for r in res])
        Xt = np.round(X[:, :5] * 2.0) / 2.0                        # ties in every column
        res = ref.test_pairwise_independence(Xt, [q for q in pairs if q[1] < 5], test="spearman")
        out["X_ties"] = Xt
        out["spearman_ties_stat"] = np.array([r.stat for r in res])
        out["spearman_ties_p"] = np.array([r.pvalue for r in res])
    out["hsic_01"] = np.array(ref.hsic_stat(X[:, 0], X[:, 1]))
    out["hsic_01_sig"] = np.array(ref.hsic_stat(X[:, 0], X[:, 1], sigma_x=0.8, sigma_y=1.3))
    out["dcor_03"] = np.array(ref.dcor_stat(X[:, 0], X[:, 3]))
    out["perm_hsic_02"] = np.array(ref.permutation_pvalue(ref.hsic_stat, X[:, 0], X[:, 2], num_perm=50,
                                                          rng=np.random.default_rng(5)))
    out["perm_dcor_13"] = np.array(ref.permutation_pvalue(ref.dcor_stat, X[:, 1], X[:, 3], num_perm=50))
    np.savez_compressed(os.path.join(GOLD, "mi_tests.npz"), **out)
    for k in ("hsic_stat", "hsic_p", "dcor_stat", "dcor_p"):
        print(k, np.round(out[k], 4))
    print("wrote mi_tests.npz")


if __name__ == "__main__":
    main()
