"""Seeded numpy generators for the synthetic inputs of BASELINE.json's configs.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference's own generators
live in ``src/dagma/utils.py`` and need ``igraph``, which is not installed
(SURVEY.md section 8c2), so the *distributions* are restated here without it:

* ER-k  (utils.py:50-54): ``s0 = k*d`` undirected edges drawn uniformly, oriented
  by a random node order, nodes relabelled at random (utils.py:39-45, 68).
* SF-k  (utils.py:55-58): Barabasi-Albert preferential attachment with
  ``m = round(s0/d)`` directed edges per new node, then relabelled.
* weights (utils.py:73-96): uniform on [-2,-0.5] U [0.5,2].
* linear SEM (utils.py:124-172): per node, in topological order,
  ``x_j = X[:,pa] @ w + z`` (gauss, unit scale) or ``Bernoulli(sigmoid(X[:,pa] @ w))``.
* MLP SEM (utils.py:199-211): hidden 100, ``x = sigmoid(X W1) W2 + z``.
* SHD / count_accuracy (utils.py:245-310).

Exact igraph RNG streams are irrelevant: parity only needs both
implementations to be fed the same arrays.  Everything uses
``np.random.default_rng(seed)`` and returns C-contiguous float64.
"""
from __future__ import annotations

import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def is_dag(W: np.ndarray) -> bool:
    """Kahn's algorithm on the support of ``W`` (utils.py:13-18 without igraph)."""
    A = (np.asarray(W) != 0)
    d = A.shape[0]
    indeg = A.sum(axis=0).astype(np.int64)
    stack = [i for i in range(d) if indeg[i] == 0]
    seen = 0
    while stack:
        i = stack.pop()
        seen += 1
        for j in np.flatnonzero(A[i]):
            indeg[j] -= 1
            if indeg[j] == 0:
                stack.append(j)
    return seen == d


def topological_order(W: np.ndarray) -> list:
    A = (np.asarray(W) != 0)
    d = A.shape[0]
    indeg = A.sum(axis=0).astype(np.int64)
    ready = [i for i in range(d) if indeg[i] == 0]
    order = []
    while ready:
        i = ready.pop(0)
        order.append(i)
        for j in np.flatnonzero(A[i]):
            indeg[j] -= 1
            if indeg[j] == 0:
                ready.append(j)
    if len(order) != d:
        raise ValueError("W must be a DAG")
    return order


def simulate_dag(d: int, s0: int, graph_type: str, rng: np.random.Generator) -> np.ndarray:
    """Binary adjacency ``B[i, j] = 1`` for an edge i -> j of a random DAG."""
    if graph_type == "ER":
        iu, ju = np.triu_indices(d, k=1)
        pick = rng.choice(iu.size, size=min(int(s0), iu.size), replace=False)
        B = np.zeros((d, d))
        # strictly upper-triangular in a hidden order = acyclic orientation
        B[iu[pick], ju[pick]] = 1.0
    elif graph_type == "SF":
        m = max(int(round(s0 / d)), 1)
        B = np.zeros((d, d))
        deg = np.zeros(d)
        for v in range(1, d):
            k = min(m, v)
            p = deg[:v] + 1.0
            targets = rng.choice(v, size=k, replace=False, p=p / p.sum())
            B[v, targets] = 1.0          # new node points to older nodes
            deg[targets] += 1.0
            deg[v] += k
    elif graph_type == "Fully":
        B = np.triu(np.ones((d, d)), 1)
    else:
        raise ValueError("unknown graph type")
    perm = rng.permutation(d)
    B = B[np.ix_(perm, perm)]
    assert is_dag(B)
    return np.ascontiguousarray(B)


def simulate_parameter(B: np.ndarray, rng: np.random.Generator,
                       w_ranges=((-2.0, -0.5), (0.5, 2.0))) -> np.ndarray:
    W = np.zeros(B.shape)
    which = rng.integers(len(w_ranges), size=B.shape)
    for i, (lo, hi) in enumerate(w_ranges):
        U = rng.uniform(lo, hi, size=B.shape)
        W += B * (which == i) * U
    return W


def simulate_linear_sem(W: np.ndarray, n: int, sem_type: str, rng: np.random.Generator,
                        noise_scale: float = 1.0) -> np.ndarray:
    d = W.shape[0]
    X = np.zeros((n, d))
    for j in topological_order(W):
        pa = np.flatnonzero(W[:, j])
        eta = X[:, pa] @ W[pa, j]
        if sem_type == "gauss":
            X[:, j] = eta + rng.normal(scale=noise_scale, size=n)
        elif sem_type == "exp":
            X[:, j] = eta + rng.exponential(scale=noise_scale, size=n)
        elif sem_type == "gumbel":
            X[:, j] = eta + rng.gumbel(scale=noise_scale, size=n)
        elif sem_type == "uniform":
            X[:, j] = eta + rng.uniform(-noise_scale, noise_scale, size=n)
        elif sem_type == "logistic":
            X[:, j] = rng.binomial(1, _sigmoid(eta)) * 1.0
        else:
            raise ValueError("unknown sem type")
    return np.ascontiguousarray(X)


def simulate_nonlinear_sem(B: np.ndarray, n: int, rng: np.random.Generator,
                           hidden: int = 100) -> np.ndarray:
    """``mlp`` SEM of utils.py:205-211."""
    d = B.shape[0]
    X = np.zeros((n, d))
    for j in topological_order(B):
        pa = np.flatnonzero(B[:, j])
        z = rng.normal(size=n)
        if pa.size == 0:
            X[:, j] = z
            continue
        W1 = rng.uniform(0.5, 2.0, size=(pa.size, hidden))
        W1[rng.random(W1.shape) < 0.5] *= -1
        W2 = rng.uniform(0.5, 2.0, size=hidden)
        W2[rng.random(hidden) < 0.5] *= -1
        X[:, j] = _sigmoid(X[:, pa] @ W1) @ W2 + z
    return np.ascontiguousarray(X)


def count_accuracy(B_true: np.ndarray, B_est: np.ndarray) -> dict:
    """fdr / tpr / fpr / shd / nnz for a 0/1 estimate (utils.py:279-310, DAG branch)."""
    d = B_true.shape[0]
    pred = np.flatnonzero(B_est == 1)
    cond = np.flatnonzero(B_true)
    cond_rev = np.flatnonzero(B_true.T)
    skeleton = np.concatenate([cond, cond_rev])
    true_pos = np.intersect1d(pred, cond, assume_unique=True)
    false_pos = np.setdiff1d(pred, skeleton, assume_unique=True)
    extra = np.setdiff1d(pred, cond, assume_unique=True)
    reverse = np.intersect1d(extra, cond_rev, assume_unique=True)
    pred_size = len(pred)
    cond_neg = 0.5 * d * (d - 1) - len(cond)
    pred_lower = np.flatnonzero(np.tril(B_est + B_est.T))
    cond_lower = np.flatnonzero(np.tril(B_true + B_true.T))
    extra_lower = np.setdiff1d(pred_lower, cond_lower, assume_unique=True)
    missing_lower = np.setdiff1d(cond_lower, pred_lower, assume_unique=True)
    return {
        "fdr": float(len(reverse) + len(false_pos)) / max(pred_size, 1),
        "tpr": float(len(true_pos)) / max(len(cond), 1),
        "fpr": float(len(reverse) + len(false_pos)) / max(cond_neg, 1),
        "shd": len(extra_lower) + len(missing_lower) + len(reverse),
        "nnz": pred_size,
    }


def edge_set_distance(W_a: np.ndarray, W_b: np.ndarray) -> int:
    """Number of directed-edge disagreements between two thresholded estimates."""
    return int(np.count_nonzero((W_a != 0) != (W_b != 0)))


# ---------------------------------------------------------------- named configs
def make_linear_problem(d: int, k: int, n: int, graph: str, sem: str, seed: int):
    """(X, W_true) for an ``{ER,SF}k`` graph with d nodes, n samples."""
    rng = np.random.default_rng(seed)
    B = simulate_dag(d, k * d, graph, rng)
    W = simulate_parameter(B, rng)
    X = simulate_linear_sem(W, n, sem, rng)
    return X, W


def config_c1(seed: int = 0):
    return make_linear_problem(20, 2, 500, "ER", "gauss", seed)


def config_c2(seed: int = 0, n: int = 10_000, d: int = 100):
    return make_linear_problem(d, 2, n, "ER", "logistic", seed)


def config_c3(seed: int = 0, n: int = 2000, d: int = 40):
    rng = np.random.default_rng(seed)
    B = simulate_dag(d, 2 * d, "ER", rng)
    return simulate_nonlinear_sem(B, n, rng), B


def config_c4_problem(p: int, d: int = 64, n: int = 1000):
    """Problem ``p`` of the 4096-problem sweep: seed 1000 + p//4, lambda1 grid p%4."""
    lam = (0.01, 0.02, 0.03, 0.05)[p % 4]
    X, W = make_linear_problem(d, 4, n, "ER", "gauss", 1000 + p // 4)
    return X, W, lam


def config_c5(seed: int = 0, d: int = 2000, n: int = 20_000):
    return make_linear_problem(d, 4, n, "SF", "gauss", seed)
