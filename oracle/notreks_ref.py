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


# ---------------------------------------------------------------- PST penalty (notreks.py:454-619, 717-736)
def _expm_taylor(M: np.ndarray, squarings: int = 12, degree: int = 18) -> np.ndarray:
    """expm by degree-18 Taylor (Horner) on M / 2^s and s squarings -- the algorithm of the CUDA path; torch's
    ``matrix_exp`` (the reference's, notreks.py:497) agrees to round-off."""
    n = M.shape[0]
    Ms = M * 2.0 ** -squarings
    R = np.eye(n) + Ms / degree
    for j in range(degree - 1, 0, -1):
        R = np.eye(n) + (Ms @ R) / j
    for _ in range(squarings):
        R = R @ R
    return R


def pst_series(W2: np.ndarray, seq: str, *, K_log=None, eps_inv: float = 1e-8):
    """F(W2) and the stored chain needed by the adjoint."""
    d = W2.shape[0]
    if seq == "inv":
        return np.linalg.solve((1.0 + eps_inv) * np.eye(d) - W2, np.eye(d)), None      # :500-507
    if seq == "exp":
        return _expm_taylor(W2), None                                                  # :496-498
    if seq == "log":                                                                   # :509-513, 425-452 (s = 1)
        K = 2 * d if K_log is None else int(K_log)
        P = [W2.copy()]
        for _ in range(K - 1):
            P.append(P[-1] @ W2)
        return np.eye(d) + sum(Pk / (k + 1) for k, Pk in enumerate(P)), P
    if seq == "binom":                                                                 # :515-519, 411-423
        A = np.eye(d) + W2
        P = [A]
        for _ in range(d - 1):
            P.append(P[-1] @ A)
        return P[-1].copy(), P
    raise ValueError("seq must be one of {'exp','log','inv','binom'}")


def pst_value_grad(W: np.ndarray, pairs, *, seq: str = "exp", agg: str = "mean", K_log=None,
                   eps_inv: float = 1e-8):
    """(value, d value / d W) with the closed-form adjoints of midagma_b200/_pst.py (the reference uses autograd)."""
    d = W.shape[0]
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    rows, cols = pairs[:, 0], pairs[:, 1]
    W2 = W * W
    F, P = pst_series(W2, seq, K_log=K_log, eps_inv=eps_inv)
    H = F.T @ F
    vals = H[rows, cols]
    if agg == "mean":
        val, w = vals.mean(), np.full(vals.shape, 1.0 / vals.size)
    elif agg == "sum":
        val, w = vals.sum(), np.ones_like(vals)
    elif agg == "max":
        val = vals.max()
        tie = (vals == val).astype(float)
        w = tie / tie.sum()
    elif agg == "lse":
        mx = vals.max()
        e = np.exp(vals - mx)
        val, w = mx + np.log(e.sum()), e / e.sum()
    else:
        raise ValueError("agg must be one of {'mean','sum','max','lse'}")
    GH = np.zeros((d, d))
    np.add.at(GH, (rows, cols), w)
    Gt = (F @ (GH + GH.T)).T                                  # transposed adjoint of F
    if seq == "inv":
        GT = F @ Gt @ F
    elif seq == "exp":
        E = np.zeros((2 * d, 2 * d))
        E[:d, :d] = W2
        E[d:, d:] = W2
        E[:d, d:] = Gt
        GT = _expm_taylor(E)[:d, d:]
    else:
        K = len(P)
        log = seq == "log"
        lhs = W2 if log else P[0]
        B = Gt / K if log else Gt.copy()
        T = np.zeros((d, d))
        for k in range(K - 1, 0, -1):
            T += B @ P[k - 1]
            B = lhs @ B + (Gt / k if log else 0.0)
        GT = T + B
    return float(val), 2.0 * W * GT.T, H
