"""Graph evaluation of the reference's ``dagma.utils`` without igraph, on the GPU and for whole batches
(SURVEY.md 8f4): ``is_dag`` (src/dagma/utils.py:13-18) and ``count_accuracy`` (utils.py:245-310) with the reference's
signatures, return values and ``ValueError``s, plus ``count_accuracy_batch`` / ``is_dag_batch`` for the results of a
``fit_batch`` sweep.  All counting happens in ``dagma_graph_metrics`` (csrc/graph_metrics.cu); there is no CPU path."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _counts(B_true, B_est) -> np.ndarray:
    """int32 [batch, 8]: nnz, condition positive, true positive, false positive, reverse, extra, missing, is_dag."""
    _lib.require_device()
    dev = torch.device("cuda")
    est = torch.as_tensor(np.asarray(B_est) if not isinstance(B_est, torch.Tensor) else B_est, device=dev)
    tru = torch.as_tensor(np.asarray(B_true) if not isinstance(B_true, torch.Tensor) else B_true, device=dev)
    if est.dim() == 2:
        est = est[None]
    batch, d, d2 = est.shape
    shared = tru.dim() == 2
    assert d == d2 and tuple(tru.shape[-2:]) == (d, d) and (shared or tru.shape[0] == batch)
    est8 = est.to(torch.int8).contiguous()
    tru8 = (tru != 0).to(torch.uint8).contiguous()
    out = torch.zeros(batch, 8, dtype=torch.int32, device=dev)
    _lib.check(_lib.load().dagma_graph_metrics(_lib.stream_ptr(), batch, d, est8.data_ptr(), tru8.data_ptr(), int(shared),
                                               out.data_ptr()), "dagma_graph_metrics")
    return out.cpu().numpy()


def is_dag_batch(W) -> np.ndarray:
    """bool [batch]: the support of every ``W[b]`` is acyclic."""
    W = torch.as_tensor(np.asarray(W) if not isinstance(W, torch.Tensor) else W)
    B = (W != 0)
    single = B.dim() == 2
    if single:
        B = B[None]
    z = torch.zeros(B.shape[-2:], dtype=torch.uint8)
    return _counts(z, B.to(torch.int8))[:, 7].astype(bool)


def is_dag(W) -> bool:
    """utils.py:13-18 (``ig.Graph.Weighted_Adjacency(W).is_dag()``): Kahn's algorithm on the support of W."""
    return bool(is_dag_batch(W)[0])


def _check_values(B_est: np.ndarray) -> bool:
    """The reference's input checks (utils.py:269-278); returns whether B_est is a CPDAG (has -1 entries)."""
    cpdag = bool((B_est == -1).any())
    if cpdag:
        if not ((B_est == 0) | (B_est == 1) | (B_est == -1)).all():
            raise ValueError('B_est should take value in {0,1,-1}')
        if ((B_est == -1) & (np.swapaxes(B_est, -1, -2) == -1)).any():
            raise ValueError('undirected edge should only appear once')
    elif not ((B_est == 0) | (B_est == 1)).all():
        raise ValueError('B_est should take value in {0,1}')
    return cpdag


def _ratios(c: np.ndarray, d: int) -> dict:
    nnz, cond, tp, fp, rev, extra, miss = (c[..., k].astype(np.float64) for k in range(7))
    cond_neg = 0.5 * d * (d - 1) - cond
    return {"fdr": (rev + fp) / np.maximum(nnz, 1), "tpr": tp / np.maximum(cond, 1),
            "fpr": (rev + fp) / np.maximum(cond_neg, 1), "shd": (extra + miss + rev).astype(np.int64),
            "nnz": nnz.astype(np.int64)}


def count_accuracy(B_true: np.ndarray, B_est: np.ndarray) -> dict:
    """utils.py:245-310: fdr / tpr / fpr / shd / nnz of one estimate ({0, 1}, or {0, 1, -1} for a CPDAG)."""
    B_est = np.asarray(B_est)
    cpdag = _check_values(B_est)
    c = _counts(np.asarray(B_true), B_est)[0]
    if not cpdag and not c[7]:
        raise ValueError('B_est should be a DAG')
    r = _ratios(c, B_est.shape[-1])
    return {"fdr": float(r["fdr"]), "tpr": float(r["tpr"]), "fpr": float(r["fpr"]), "shd": int(r["shd"]), "nnz": int(r["nnz"])}


def count_accuracy_batch(B_true, B_est) -> dict:
    """``count_accuracy`` for ``B_est [batch, d, d]`` against ``B_true [batch, d, d]`` (or one ``[d, d]`` truth for all):
    arrays of fdr / tpr / fpr / shd / nnz plus ``is_dag`` -- a cyclic estimate is flagged instead of raising."""
    B_est_np = B_est.cpu().numpy() if isinstance(B_est, torch.Tensor) else np.asarray(B_est)
    _check_values(B_est_np)
    c = _counts(B_true, B_est)
    r = _ratios(c, B_est_np.shape[-1])
    r["is_dag"] = c[:, 7].astype(bool)
    return r
