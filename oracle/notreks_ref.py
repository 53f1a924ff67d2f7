"""numpy restatement of the fork's log-det constraint variant
(reference: src/notreks/notreks.py, CR-delimited line numbers).

TEST INFRASTRUCTURE (see oracle/__init__.py) -- never imported by the product.

* ``logdet_acyc_value_gradA`` (notreks.py:241-275): ``h(A) = -logabsdet(sI-A) + n log s``,
  ``G_A = solve(sI-A, I)^T``; no Hadamard square, ``eps`` unused.
* the ``cycle_penalty="logdet"`` branch of ``trek_cycle_coupling_value_gradW``
  (notreks.py:319-337 block assembly, :380-413 branch, :278-287 fold-back).

Pinned by tests/golden/notreks_logdet.npz (reference torch outputs).
"""
from __future__ import annotations

import numpy as np


def logdet_acyc_value_gradA(A: np.ndarray, s: float = 1.0):
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise ValueError("A must be square")
    n = A.shape[0]
    M = float(s) * np.eye(n) - A
    _, logabsdet = np.linalg.slogdet(M)
    h = -logabsdet + float(n) * np.log(float(s))
    G_A = np.linalg.solve(M, np.eye(n)).T
    return h, G_A


def indicator_from_pairs(pairs, d: int) -> np.ndarray:
    S = np.zeros((d, d))
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    for i, j in pairs:
        S[i, j] = 1.0
    return S


def tcc_blocks(W: np.ndarray, S: np.ndarray, w: float = 1.0):
    """A = [[W2, w S], [I, W2^T]],  B = [[W2, 0], [I, W2^T]]  (notreks.py:319-337)."""
    d = W.shape[0]
    W2 = W * W
    bot = np.concatenate([np.eye(d), W2.T], axis=1)
    A = np.concatenate([np.concatenate([W2, float(w) * S], axis=1), bot], axis=0)
    B = np.concatenate([np.concatenate([W2, np.zeros_like(S)], axis=1), bot], axis=0)
    return A, B


def tcc_logdet_value_gradW(W: np.ndarray, S: np.ndarray, *, w: float = 1.0,
                           version: str = "DAG_learning", s: float = 1.0):
    d = W.shape[0]
    A, B = tcc_blocks(W, S, w)
    hA, GA = logdet_acyc_value_gradA(A, s)
    gA = 2.0 * W * (GA[:d, :d] + GA[d:, d:].T)               # :278-287, :384
    if version == "DAG_learning":
        return hA, gA
    if version == "exact_trek_graph":
        hB, GB = logdet_acyc_value_gradA(B, s)
        gB = 2.0 * W * (GB[:d, :d] + GB[d:, d:].T)
        return hA - hB, gA - gB
    raise ValueError(f"version {version!r} is not implemented for the logdet penalty")
