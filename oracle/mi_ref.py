"""numpy restatement of the fork's pairwise independence tests (reference: src/notreks/mi_tests.py,
CR-delimited line numbers).

TEST INFRASTRUCTURE (see oracle/__init__.py) -- never imported by the product.

HSIC with RBF kernels + median heuristic (:19-64), distance correlation (:67-100), permutation p-value
(:103-135) and the pair loop that shares one RNG stream (:165-203).  Pinned by tests/golden/mi_tests.npz
(outputs of the unmodified reference, oracle/make_golden_mi.py).
"""
from __future__ import annotations

import numpy as np


def _center(K):                                               # :19-27, :67-75
    return K - K.mean(axis=1, keepdims=True) - K.mean(axis=0, keepdims=True) + K.mean()


def _rbf_gram(x, sigma=None):                                 # :30-50
    x = np.asarray(x).reshape(-1, 1)
    D2 = (x - x.T) ** 2
    if sigma is None:
        med = np.median(D2[np.triu_indices(D2.shape[0], k=1)])
        sigma2 = med if med > 0 else 1.0
    else:
        sigma2 = float(sigma) ** 2
        if sigma2 <= 0:
            sigma2 = 1.0
    return np.exp(-D2 / (2.0 * sigma2))


def hsic_stat(x, y, sigma_x=None, sigma_y=None):              # :53-64
    x, y = np.asarray(x).ravel(), np.asarray(y).ravel()
    n = x.shape[0]
    return float((_center(_rbf_gram(x, sigma_x)) * _center(_rbf_gram(y, sigma_y))).sum() / (n * n))


def dcor_stat(x, y):                                          # :78-100
    x, y = np.asarray(x).ravel(), np.asarray(y).ravel()
    n = x.shape[0]
    Ax = _center(np.abs(x[:, None] - x[None, :]))
    Ay = _center(np.abs(y[:, None] - y[None, :]))
    dcov2 = (Ax * Ay).sum() / (n * n)
    dvarx2 = (Ax * Ax).sum() / (n * n)
    dvary2 = (Ay * Ay).sum() / (n * n)
    if dvarx2 <= 0 or dvary2 <= 0:
        return 0.0
    return float(np.sqrt(max(dcov2, 0.0)) / np.sqrt(np.sqrt(dvarx2 * dvary2)))


def permutation_pvalue(stat_fn, x, y, *, num_perm=200, rng=None):      # :103-135
    x, y = np.asarray(x).ravel(), np.asarray(y).ravel()
    if rng is None:
        rng = np.random.default_rng(0)
    stat_obs = float(stat_fn(x, y))
    ge = 0
    for _ in range(num_perm):
        if float(stat_fn(x, y[rng.permutation(x.shape[0])])) >= stat_obs:
            ge += 1
    return stat_obs, float((ge + 1) / (num_perm + 1))


def pairwise(X, pairs, *, test="hsic", num_perm=200, seed=0):          # :165-203 (hsic / dcor branches)
    fn = {"hsic": hsic_stat, "dcor": dcor_stat}[test]
    rng = np.random.default_rng(seed)
    return [(i, j) + permutation_pvalue(fn, X[:, i], X[:, j], num_perm=num_perm, rng=rng) for i, j in pairs]
