"""The fork's log-det constraint variant on the B200 kernels (reference:
src/notreks/notreks.py, CR-delimited line numbers as in SURVEY.md).

In scope:
* ``logdet_acyc_value_gradA``                                   notreks.py:241-275
* ``trek_cycle_coupling_value_gradW(..., cycle_penalty="logdet")``  notreks.py:291-337, 380-413
* ``trek_value_grad`` no-op path (no / disabled regulariser)    notreks.py:684-689
* the regulariser dataclasses, so objects built for the reference can be passed around.

* PST with ``seq="inv"`` (value and closed-form gradient on the inverse + GEMM kernels)   notreks.py:500-507, 558-619

The other PST series and the spectral TCC penalty are outside the accelerated path
 and raise ``NotImplementedError``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .linear import logdet_inv

TrekRegularizerNames = ["pst", "tcc"]


@dataclass(frozen=True)
class TrekRegularizer:
    """name / mode in {"off", "log", "opt"} / weight / cfg  (notreks.py:21-35)."""
    name: str
    mode: str = "off"
    weight: float = 0.0
    cfg: Dict[str, Any] = field(default_factory=dict)

    def enabled(self) -> bool:
        return self.mode != "off" and self.weight != 0.0


@dataclass(frozen=True)
class PSTRegularizer(TrekRegularizer):
    def __init__(self, *, I, seq="exp", weight: float = 0.0, kwargs: Optional[Dict[str, Any]] = None,
                 mode="opt", name: str = "pst"):
        object.__setattr__(self, "name", name)
        object.__setattr__(self, "mode", mode)
        object.__setattr__(self, "weight", float(weight))
        object.__setattr__(self, "cfg", {"I": I, "seq": seq, "kwargs": {} if kwargs is None else dict(kwargs)})


@dataclass(frozen=True)
class TCCRegularizer(TrekRegularizer):
    def __init__(self, *, I, cycle_penalty="spectral", version="approx_trek_graph", method="eig_troch",
                 weight: float = 1.0, w: float = 1.0, s: float = 1.0, n_iter: int = 10, eps: float = 1e-12,
                 mode="opt", name: str = "tcc"):
        object.__setattr__(self, "cycle_penalty", cycle_penalty)
        object.__setattr__(self, "name", name)
        object.__setattr__(self, "mode", mode)
        object.__setattr__(self, "weight", float(weight))
        object.__setattr__(self, "cfg", {"I": I, "version": version, "w": float(w), "n_iter": int(n_iter),
                                         "eps": float(eps), "s": float(s)})


def _dev64(x: torch.Tensor) -> torch.Tensor:
    return x.detach().to(device="cuda", dtype=torch.float64).contiguous()


def logdet_acyc_value_gradA(A: torch.Tensor, *, s: float = 1.0, eps: float = 1e-12) -> Tuple[torch.Tensor, torch.Tensor]:
    """h(A) = -logabsdet(sI - A) + n log s and dh/dA = (sI - A)^{-T}; no Hadamard square, ``eps`` unused."""
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise ValueError("A must be square")
    out = logdet_inv(_dev64(A)[None], s=float(s), square_input=False, want_inv=False, want_grad=True)
    return out["h"][0].to(A.device, A.dtype), out["grad"][0].to(A.device, A.dtype)


def _indicator_from_pairs(I, d: int) -> torch.Tensor:
    S = torch.zeros((d, d), dtype=torch.float64)
    I_np = np.asarray(I, dtype=np.int64)
    if I_np.size == 0:
        return S
    if I_np.ndim != 2 or I_np.shape[1] != 2:
        raise ValueError("I must be array-like of shape (m,2)")
    S[I_np[:, 0], I_np[:, 1]] = 1.0
    return S


def trek_cycle_coupling_value_gradW(W: torch.Tensor, I, *, w: float = 1.0, cycle_penalty="spectral",
                                    version="approx_trek_graph", method="eig_numpy", n_iter: int = 50,
                                    s: float = 1.0, eps: float = 1e-12) -> Tuple[torch.Tensor, torch.Tensor]:
    """TCC penalty on the 2d x 2d block matrix; only ``cycle_penalty="logdet"`` is accelerated."""
    if W.ndim != 2 or W.shape[0] != W.shape[1]:
        raise ValueError("W must be square")
    if cycle_penalty == "spectral":
        raise NotImplementedError("the spectral TCC penalty is outside the B200 hot path (SURVEY.md 8f3)")
    if cycle_penalty != "logdet":
        raise ValueError("cycle_penalty must be one of {'spectral','logdet'}")
    if version in ("exact_original_graph", "approx_trek_graph"):
        print(version)
        raise ValueError(f"The version '{version}' for the 'logdet' acyclicity constraint is not imlpemented")
    if version not in ("DAG_learning", "exact_trek_graph"):
        raise ValueError("version must be one of {TCCVersion} for logdet")
    _lib.require_device()
    lib = _lib.load()
    d = W.shape[0]
    Wd = _dev64(W)
    Sd = _indicator_from_pairs(I, d).cuda()
    blocks = torch.empty(2, 2 * d, 2 * d, dtype=torch.float64, device="cuda")
    two = version == "exact_trek_graph"
    _lib.check(lib.dagma_tcc_assemble_f64(_lib.stream_ptr(), d, Wd.data_ptr(), Sd.data_ptr(), float(w), 1,
                                          blocks[0].data_ptr()), "dagma_tcc_assemble_f64")
    if two:
        _lib.check(lib.dagma_tcc_assemble_f64(_lib.stream_ptr(), d, Wd.data_ptr(), Sd.data_ptr(), float(w), 0,
                                              blocks[1].data_ptr()), "dagma_tcc_assemble_f64")
    nb = 2 if two else 1
    out = logdet_inv(blocks[:nb].contiguous(), s=float(s), square_input=False, want_inv=False, want_grad=True)
    print(version)                                               # the reference prints the branch name (:387, 392)
    grad = torch.empty(d, d, dtype=torch.float64, device="cuda")
    _lib.check(lib.dagma_tcc_fold_f64(_lib.stream_ptr(), d, Wd.data_ptr(), out["grad"][0].data_ptr(), 1.0, 0,
                                      grad.data_ptr()), "dagma_tcc_fold_f64")
    penalty = out["h"][0]
    if two:
        _lib.check(lib.dagma_tcc_fold_f64(_lib.stream_ptr(), d, Wd.data_ptr(), out["grad"][1].data_ptr(), -1.0, 1,
                                          grad.data_ptr()), "dagma_tcc_fold_f64")
        penalty = out["h"][0] - out["h"][1]
    return penalty.to(W.device, W.dtype), grad.to(W.device, W.dtype)


def pst_inv_value_grad(W, I, *, agg: str = "mean", eps_inv: float = 1e-8, want_grad: bool = True):
    """PST penalty with ``seq="inv"`` (notreks.py:500-507, 558-619) and its gradient w.r.t. W, on device:
    X = ((1 + eps) I - W o W)^{-1} (fused inverse kernel), H = X^T X, pst = agg_{(i,j) in I} H[i,j],
    d pst / d W = 2 W o (X M_s H)^T with M_s the symmetrised, agg-scaled pair mask (closed form of the
    reference's autograd).  Returns (float, ndarray or None)."""
    from ._large import gemm
    _lib.require_device()
    Wd = torch.as_tensor(np.ascontiguousarray(W, dtype=np.float64)).cuda()
    d = Wd.shape[0]
    I_np = np.asarray(I, dtype=np.int64)
    idx = torch.as_tensor(I_np, device="cuda")
    mask = torch.zeros(d, d, dtype=torch.float64, device="cuda")
    mask.index_put_((idx[:, 0], idx[:, 1]), torch.ones(idx.shape[0], dtype=torch.float64, device="cuda"), accumulate=True)
    if agg == "mean":
        mask /= idx.shape[0]
    elif agg != "sum":
        raise NotImplementedError("only agg in {'mean', 'sum'} is accelerated (SURVEY.md 8f3)")
    out = logdet_inv(Wd[None].contiguous(), s=1.0 + float(eps_inv), square_input=True, want_inv=True, want_grad=False)
    if int(out["info"][0].item()) != 0:
        raise _lib.DagmaB200Error("(1 + eps) I - W o W is not an M-matrix: outside the domain of the fused inverse")
    X = out["minv"][0].contiguous()
    H = torch.empty_like(X)
    gemm(X, X, H, trans_a=True)
    val = float((mask * H).sum().item())
    if not want_grad:
        return val, None
    B, GT = torch.empty_like(X), torch.empty_like(X)
    gemm((mask + mask.T).contiguous(), H, B)
    gemm(X, B, GT)
    return val, (2.0 * Wd * GT.T).cpu().numpy()


def trek_value_grad(W: np.ndarray, tr: Optional[TrekRegularizer], *, torch_dtype: torch.dtype = torch.double,
                    device: Optional[torch.device] = None) -> Tuple[float, np.ndarray]:
    """(value, grad) of a trek regulariser (notreks.py:667-736): the no-op branch, and PST ``seq="inv"``."""
    from .linear import _trek_plan
    W_np = np.asarray(W)
    plan = _trek_plan(tr)               # None: disabled / empty I; NotImplementedError outside the accelerated set
    if plan is None:
        return 0.0, np.zeros_like(W_np)
    val, grad = pst_inv_value_grad(W_np, plan["I"], agg=plan["agg"], eps_inv=plan["eps_inv"],
                                   want_grad=(tr.mode == "opt"))
    if tr.mode != "opt":
        return val, np.zeros_like(W_np)
    return val, grad.astype(W_np.dtype, copy=False)
