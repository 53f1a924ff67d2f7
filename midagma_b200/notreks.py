"""The fork's log-det constraint variant on the B200 kernels (reference:
src/notreks/notreks.py, CR-delimited line numbers as in SURVEY.md).

In scope:
* ``logdet_acyc_value_gradA``                                   notreks.py:241-275
* ``trek_cycle_coupling_value_gradW(..., cycle_penalty="logdet")``  notreks.py:291-337, 380-413
* ``trek_value_grad`` no-op path (no / disabled regulariser)    notreks.py:684-689
* the regulariser dataclasses, so objects built for the reference can be passed around.

* PST penalties, every series (``inv``, ``log``, ``exp``, ``binom``) and aggregation (``mean``, ``sum``, ``max``,
  ``lse``): value and closed-form gradient on the inverse + GEMM kernels (``_pst.PstEngine``)   notreks.py:454-619

* the spectral TCC penalty (``perron_eig_with_gradA``, ``trek_cycle_coupling_value_gradW(cycle_penalty="spectral")``,
  a ``TCCRegularizer`` through ``trek_value_grad`` / ``DagmaLinear``) by power iteration on device (``_tcc``)
                                                                                        notreks.py:156-239, 340-378
* ``FORWARD_TCC_CONFIG`` / ``trek_value_grad(..., forward_tcc_config=True)``: the reference's dispatch drops a TCC
  regulariser's ``cycle_penalty`` / ``version`` / ``method`` / ``s`` (notreks.py:699-707, Q14); the flag forwards them.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .linear import logdet_inv

TrekRegularizerNames = ["pst", "tcc"]

# Q14: `trek_value_grad` of the reference calls the TCC penalty with its DEFAULT cycle_penalty
# ("spectral"), version ("approx_trek_graph") and method ("eig_numpy") whatever the regulariser object says
# (notreks.py:699-707).  False keeps that behaviour (drop-in parity); True forwards the regulariser's own settings.
FORWARD_TCC_CONFIG = False


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
        object.__setattr__(self, "method", method)     # never read by the reference ("eig_troch" is its default, Q14)
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


def perron_eig_with_gradA(A: torch.Tensor, *, method="eig_torch", n_iter: int = 50,
                          eps: float = 1e-12) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """(rho, u, v, d rho / dA = u v^T / (u.v + eps)) of a non-negative matrix (notreks.py:156-239).
    ``method="power"``: the reference's n_iter-step power iteration; the eig methods: the same Perron pair by power
    steps iterated to convergence (see ``_tcc``)."""
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise ValueError("A must be square")
    if method not in ("power", "eig_torch", "eig_numpy"):
        raise ValueError("method must be one of {'power','eig_torch','eig_numpy'}")
    from ._tcc import PerronSolver
    Ad = _dev64(A)
    sol = PerronSolver(Ad.shape[0])
    if method == "power":
        sol.power(Ad, n_iter, eps)
    else:
        sol.converged(Ad, eps)
    G = torch.outer(sol.u, sol.v) / (sol.scal[1] + eps)
    to = lambda t: t.clone().to(A.device, A.dtype)  # noqa: E731
    return to(sol.scal[0]), to(sol.u), to(sol.v), to(G)


def trek_cycle_coupling_value_gradW(W: torch.Tensor, I, *, w: float = 1.0, cycle_penalty="spectral",
                                    version="approx_trek_graph", method="eig_numpy", n_iter: int = 50,
                                    s: float = 1.0, eps: float = 1e-12) -> Tuple[torch.Tensor, torch.Tensor]:
    """TCC penalty on the 2d x 2d block matrix (notreks.py:291-413): spectral (Perron root) or log-det."""
    if W.ndim != 2 or W.shape[0] != W.shape[1]:
        raise ValueError("W must be square")
    if cycle_penalty == "spectral":
        from ._tcc import SpectralTcc
        eng = SpectralTcc(W.shape[0], I, w=w, version=version, method=method, n_iter=n_iter, eps=eps)
        grad = eng.compute(_dev64(W))
        return eng.value().to(W.device, W.dtype), grad.clone().to(W.device, W.dtype)
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


def _pst_engine(d: int, I, seq, agg, eps_inv, K_log):
    from ._pst import PstEngine
    return PstEngine(d, I, seq, agg, eps_inv=eps_inv, K_log=K_log)


def pst_mat(W, seq="exp", *, K_log: Optional[int] = None, eps_inv: float = 1e-8, s: float = 1.0) -> torch.Tensor:
    """H = F(W o W)^T F(W o W) for the series ``seq`` (notreks.py:454-527); ``s`` is accepted and ignored as in the
    reference (dead duplicate branch).  No autograd: gradients come from ``pst_value_grad``."""
    seq = str(seq).lower().strip()
    if seq not in {"exp", "log", "inv", "binom"}:
        raise ValueError("seq must be one of {'exp','log','inv','binom'}")
    if not isinstance(W, torch.Tensor):
        W = torch.as_tensor(W, dtype=torch.double, device=torch.device("cpu"))
    if W.ndim != 2 or W.shape[0] != W.shape[1]:
        raise ValueError("W must be a square matrix tensor")
    if eps_inv < 0:
        raise ValueError("eps_inv must be >= 0")
    eng = _pst_engine(W.shape[0], np.zeros((1, 2), dtype=np.int64), seq, "sum", eps_inv, K_log)
    Wd = _dev64(W)
    eng.forward(Wd)
    eng.check_domain(Wd)
    return eng.H.to(W.device, W.dtype)


def get_no_trek_pairs(W, seq="exp", *, K_log: Optional[int] = None, eps_inv: float = 1e-8) -> np.ndarray:
    """Pairs (i, j), i < j, with H[i, j] == 0 (notreks.py:529-554)."""
    H = pst_mat(W, seq, K_log=K_log, eps_inv=eps_inv)
    upper = torch.triu(torch.ones_like(H, dtype=torch.bool), diagonal=1)
    rows, cols = torch.nonzero((H == 0) & upper, as_tuple=True)
    return torch.stack([rows, cols], dim=1).cpu().numpy().astype(np.int64, copy=False)


def pst(W: torch.Tensor, I, seq="exp", *, K_log: Optional[int] = None, eps_inv: float = 1e-8, s: float = 1.,
        agg: str = "mean") -> torch.Tensor:
    """PST penalty value (notreks.py:557-619); ``agg="none"`` returns the vector of pair values."""
    agg = str(agg).lower().strip()
    I_np = np.asarray(I, dtype=np.int64)
    if I_np.size == 0:
        return W.sum() * 0.0
    if agg not in {"mean", "sum", "max", "lse", "none"}:
        raise ValueError("agg must be one of {'mean','sum','max','lse','none'}")
    eng = _pst_engine(W.shape[0], I_np, seq, "sum" if agg == "none" else agg, eps_inv, K_log)
    Wd = _dev64(W)
    val = eng.value(Wd)
    eng.check_domain(Wd)
    if agg == "none":
        return eng.pair_values().to(W.device, W.dtype)
    return val.clone().to(W.device, W.dtype)


def pst_value_grad(W, I, *, seq: str = "exp", agg: str = "mean", eps_inv: float = 1e-8, K_log: Optional[int] = None,
                   want_grad: bool = True):
    """(value, d value / d W) of the PST penalty on device -- closed-form adjoints of every series (see ``_pst``)
    in place of the reference's autograd (notreks.py:717-736).  Returns (float, ndarray or None)."""
    W_np = np.asarray(W, dtype=np.float64)
    eng = _pst_engine(W_np.shape[0], I, seq, agg, eps_inv, K_log)
    return eng.value_grad_host(W_np, want_grad)


def pst_inv_value_grad(W, I, *, agg: str = "mean", eps_inv: float = 1e-8, want_grad: bool = True):
    """PST with ``seq="inv"`` (notreks.py:500-507): X = ((1 + eps) I - W o W)^{-1}, H = X^T X."""
    return pst_value_grad(W, I, seq="inv", agg=agg, eps_inv=eps_inv, want_grad=want_grad)


def _tcc_call_args(tr, forward: bool) -> dict:
    """What ``trek_cycle_coupling_value_gradW`` is called with for a TCC regulariser: the reference passes only I, w,
    n_iter, eps (notreks.py:699-707); with ``forward`` also the regulariser's cycle_penalty / version / method / s."""
    cfg = tr.cfg
    kw = dict(w=cfg.get("w", 1.0), n_iter=cfg.get("n_iter", 10), eps=cfg.get("eps", 1e-12))
    if forward:
        kw["cycle_penalty"] = getattr(tr, "cycle_penalty", "spectral")
        kw["version"] = cfg.get("version", "approx_trek_graph")
        kw["s"] = cfg.get("s", 1.0)
        method = getattr(tr, "method", "eig_numpy")
        kw["method"] = method if method in ("power", "eig_torch", "eig_numpy") else "eig_numpy"   # "eig_troch" typo
    return kw


def trek_value_grad(W: np.ndarray, tr: Optional[TrekRegularizer], *, torch_dtype: torch.dtype = torch.double,
                    device: Optional[torch.device] = None,
                    forward_tcc_config: Optional[bool] = None) -> Tuple[float, np.ndarray]:
    """(value, grad) of a trek regulariser (notreks.py:667-736): the no-op branch, every PST penalty, and TCC."""
    from .linear import _trek_plan
    W_np = np.asarray(W)
    plan = _trek_plan(tr)               # None: disabled / empty I
    if plan is None:
        return 0.0, np.zeros_like(W_np)
    if plan["kind"] == "tcc":
        import contextlib
        import io
        forward = FORWARD_TCC_CONFIG if forward_tcc_config is None else bool(forward_tcc_config)
        kw = _tcc_call_args(tr, forward)
        with contextlib.nullcontext() if kw.get("cycle_penalty") == "logdet" else contextlib.redirect_stdout(io.StringIO()):
            pen, grad = trek_cycle_coupling_value_gradW(torch.as_tensor(np.ascontiguousarray(W_np), dtype=torch.double),
                                                        np.asarray(tr.cfg["I"]), **kw)
        val = float(pen.item())
        if tr.mode != "opt":
            return val, np.zeros_like(W_np)
        return val, grad.numpy().astype(W_np.dtype, copy=False)
    val, grad = pst_value_grad(W_np, plan["I"], seq=plan["seq"], agg=plan["agg"], eps_inv=plan["eps_inv"],
                               K_log=plan["K_log"], want_grad=(tr.mode == "opt"))
    if tr.mode != "opt":
        return val, np.zeros_like(W_np)
    return val, grad.astype(W_np.dtype, copy=False)
