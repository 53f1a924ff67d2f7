"""Pairwise independence tests on the B200 kernels (reference: src/notreks/mi_tests.py, CR-delimited line
numbers) -- the pre-processing step that builds the no-trek pair set ``I``.

Same names and signatures: ``hsic_stat``, ``dcor_stat``, ``permutation_pvalue``, ``test_pairwise_independence``,
``get_I_from_full_pairwise_tests``, ``IndepTestResult``.  HSIC (RBF kernels, median heuristic) and distance
correlation run on the GPU: each variable's centred Gram matrix is built once and every permuted statistic is a
gathered dot product (csrc/mi.cu), instead of the reference's full rebuild per permutation.  The permutations
themselves come from the same ``numpy.random.Generator`` stream as the reference (``rng.permutation(n)`` per
permutation, pairs in order), so p-values are comparable one to one.  ``pearson`` / ``spearman``: all correlations
from one fused centring + covariance launch (on the data or on its tie-averaged ranks), the analytic p-values are
scipy's closed forms evaluated on the host (d^2 scalars).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Optional, Tuple

import numpy as np
import torch

from . import _lib

__all__ = ["IndepTestResult", "hsic_stat", "dcor_stat", "permutation_pvalue", "pearson_stat_pvalue",
           "spearman_stat_pvalue", "test_pairwise_independence", "get_I_from_full_pairwise_tests"]


@dataclass(frozen=True)
class IndepTestResult:
    i: int
    j: int
    stat: float
    pvalue: float


class _GramBank:
    """Centred Gram matrices of the columns ``cols`` of X (device), built once."""

    def __init__(self, X: np.ndarray, cols: List[int], kind: str, sigmas: Optional[List[Optional[float]]] = None):
        _lib.require_device()
        self.lib = _lib.load()
        X = np.ascontiguousarray(X, dtype=np.float64)
        self.n, self.d = X.shape
        self.kind = 0 if kind == "hsic" else 1
        self.cols = list(cols)
        self.slot = {c: k for k, c in enumerate(self.cols)}
        f64 = dict(dtype=torch.float64, device="cuda")
        self.X = torch.from_numpy(X).cuda()
        nv, n = len(self.cols), self.n
        self.G = torch.empty(nv, n, n, **f64)
        sig = torch.ones(nv, **f64)
        if self.kind == 0:
            sig = torch.tensor([self._sigma2(c, None if sigmas is None else sigmas[k])
                                for k, c in enumerate(self.cols)], **f64)
        self.rm = torch.empty(nv * n, **f64)
        self.am = torch.empty(nv, **f64)
        cd = torch.tensor(self.cols, dtype=torch.int32, device="cuda")
        _lib.check(self.lib.dagma_mi_centered_gram_f64(
            _lib.stream_ptr(), n, self.d, nv, cd.data_ptr(), self.X.data_ptr(), sig.data_ptr(), self.kind,
            self.G.data_ptr(), self.rm.data_ptr(), self.am.data_ptr()), "dagma_mi_centered_gram_f64")

    def _sigma2(self, col: int, sigma: Optional[float]) -> float:
        """Median heuristic on the off-diagonal squared distances (mi_tests.py:39-50)."""
        if sigma is not None:
            s2 = float(sigma) ** 2
            return s2 if s2 > 0 else 1.0
        n = self.n
        if n < 2:
            return 1.0
        buf = torch.empty(n * (n - 1) // 2, dtype=torch.float64, device="cuda")
        _lib.check(self.lib.dagma_mi_upper_d2_f64(_lib.stream_ptr(), n, self.d, col, self.X.data_ptr(), buf.data_ptr()),
                   "dagma_mi_upper_d2_f64")
        srt = torch.sort(buf).values
        m = srt.numel()
        med = float(srt[m // 2].item()) if m % 2 else 0.5 * float((srt[m // 2 - 1] + srt[m // 2]).item())
        return med if med > 0 else 1.0

    def dots(self, pairs: List[Tuple[int, int]], perms) -> np.ndarray:
        """[len(pairs), P] values of (1/n^2) sum Kc_i o (P Lc_j P^T).  ``perms``: a [len(pairs), P, n] int array, or a
        ``_PermStream`` that draws the permutations of the next pairs on demand -- host memory then stays at one
        launch's worth (<= 256 MB) however many pairs there are, and the RNG stream is consumed in pair order
        exactly as the reference does (mi_tests.py:124-127, 186-197)."""
        lazy = isinstance(perms, _PermStream)
        n, q, P = self.n, len(pairs), (perms.P if lazy else perms.shape[1])
        out = np.empty((q, P))
        step = max(1, min(q, (256 << 20) // max(1, P * n * 4)))            # <= 256 MB of permutations per launch
        for q0 in range(0, q, step):
            sub = pairs[q0:q0 + step]
            vi = torch.tensor([self.slot[i] for i, _ in sub], dtype=torch.int32, device="cuda")
            vj = torch.tensor([self.slot[j] for _, j in sub], dtype=torch.int32, device="cuda")
            block = perms.take(len(sub)) if lazy else perms[q0:q0 + step]
            pd = torch.from_numpy(np.ascontiguousarray(block, dtype=np.int32)).cuda()
            del block
            res = torch.empty(len(sub) * P, dtype=torch.float64, device="cuda")
            wsb = self.lib.dagma_mi_perm_workspace_bytes(n, len(sub), P)
            ws = torch.empty(wsb // 8 + 1, dtype=torch.float64, device="cuda")
            _lib.check(self.lib.dagma_mi_perm_dots_f64(
                _lib.stream_ptr(), n, self.G.data_ptr(), len(sub), vi.data_ptr(), vj.data_ptr(), pd.data_ptr(), P,
                1.0 / (n * n), ws.data_ptr(), ws.numel() * 8, res.data_ptr()), "dagma_mi_perm_dots_f64")
            out[q0:q0 + len(sub)] = res.cpu().numpy().reshape(len(sub), P)
        return out


def _stats_from_dots(kind: str, dots: np.ndarray, var_i: float, var_j: float) -> np.ndarray:
    if kind == "hsic":
        return dots                                                        # mi_tests.py:64
    if var_i <= 0 or var_j <= 0:                                           # :98-99
        return np.zeros_like(dots)
    return np.sqrt(np.maximum(dots, 0.0)) / np.sqrt(np.sqrt(var_i * var_j))   # :100


def _pair_stats(X: np.ndarray, pairs: List[Tuple[int, int]], kind: str, perms: np.ndarray) -> np.ndarray:
    """Statistic of every pair under every given permutation of the second variable."""
    cols = sorted({c for p in pairs for c in p})
    bank = _GramBank(X, cols, kind)
    dots = bank.dots(pairs, perms)
    if kind == "hsic":
        return dots
    n = X.shape[0]
    ident = np.broadcast_to(np.arange(n, dtype=np.int32), (len(cols), 1, n))
    var = bank.dots([(c, c) for c in cols], np.ascontiguousarray(ident))[:, 0]     # dvar^2 of every variable
    vmap = dict(zip(cols, var))
    return np.stack([_stats_from_dots(kind, dots[q], vmap[i], vmap[j]) for q, (i, j) in enumerate(pairs)])


def _two_col(x, y):
    x = np.asarray(x, dtype=np.float64).ravel()
    y = np.asarray(y, dtype=np.float64).ravel()
    return np.ascontiguousarray(np.stack([x, y], axis=1)), x.shape[0]


def hsic_stat(x: np.ndarray, y: np.ndarray, sigma_x: Optional[float] = None, sigma_y: Optional[float] = None) -> float:
    """Biased HSIC estimator (1/n^2) sum(Kc o Lc) with RBF kernels (mi_tests.py:53-64)."""
    X, n = _two_col(x, y)
    bank = _GramBank(X, [0, 1], "hsic", [sigma_x, sigma_y])
    return float(bank.dots([(0, 1)], np.arange(n, dtype=np.int32)[None, None, :])[0, 0])


def dcor_stat(x: np.ndarray, y: np.ndarray) -> float:
    """Distance correlation (mi_tests.py:78-100)."""
    X, n = _two_col(x, y)
    return float(_pair_stats(X, [(0, 1)], "dcor", np.arange(n, dtype=np.int32)[None, None, :])[0, 0])


def _draw_perms(rng: np.random.Generator, n: int, num_perm: int) -> np.ndarray:
    """identity (the observed statistic) followed by ``num_perm`` draws of ``rng.permutation(n)`` (mi_tests.py:124-127)."""
    out = np.empty((num_perm + 1, n), dtype=np.int32)
    out[0] = np.arange(n)
    for p in range(num_perm):
        out[p + 1] = rng.permutation(n)
    return out


class _PermStream:
    """The permutations of consecutive pairs, drawn block by block from ONE generator (pair order = stream order)."""

    def __init__(self, rng: np.random.Generator, n: int, num_perm: int):
        self.rng, self.n, self.num_perm, self.P = rng, n, num_perm, num_perm + 1

    def take(self, count: int) -> np.ndarray:
        return np.stack([_draw_perms(self.rng, self.n, self.num_perm) for _ in range(count)])


def permutation_pvalue(stat_fn, x: np.ndarray, y: np.ndarray, *, num_perm: int = 200,
                       rng: Optional[np.random.Generator] = None) -> Tuple[float, float]:
    """Permutation p-value (mi_tests.py:103-135); ``stat_fn`` must be this module's ``hsic_stat`` or ``dcor_stat``."""
    if stat_fn is hsic_stat:
        kind = "hsic"
    elif stat_fn is dcor_stat:
        kind = "dcor"
    else:
        raise NotImplementedError("the B200 path accelerates hsic_stat and dcor_stat")
    if rng is None:
        rng = np.random.default_rng(0)
    X, n = _two_col(x, y)
    stats = _pair_stats(X, [(0, 1)], kind, _draw_perms(rng, n, num_perm)[None])[0]
    ge = int((stats[1:] >= stats[0]).sum())
    return float(stats[0]), float((ge + 1) / (num_perm + 1))


def _corr_matrix(Xd: "torch.Tensor") -> np.ndarray:
    """Pearson correlation matrix of the columns of a device matrix via the fused centring + covariance kernel
    (``dagma_center_cov_f64``: X <- X - mean, cov = X^T X / n)."""
    lib = _lib.load()
    n, d = Xd.shape
    Xc = Xd.clone().contiguous()
    cov = torch.empty(d, d, dtype=torch.float64, device="cuda")
    _lib.check(lib.dagma_center_cov_f64(_lib.stream_ptr(), 1, n, d, Xc.data_ptr(), 1, cov.data_ptr()), "dagma_center_cov_f64")
    c = cov.cpu().numpy()
    sd = np.sqrt(np.diag(c))
    with np.errstate(divide="ignore", invalid="ignore"):
        return c / np.outer(sd, sd)


def _average_ranks(Xd: "torch.Tensor") -> "torch.Tensor":
    """Column-wise ranks 1..n with ties averaged (scipy.stats.rankdata, the ranking behind spearmanr)."""
    n, d = Xd.shape
    srt, order = torch.sort(Xd, dim=0, stable=True)
    first = torch.ones(n, d, dtype=torch.bool, device=Xd.device)
    first[1:] = srt[1:] != srt[:-1]
    pos = torch.arange(1, n + 1, dtype=torch.float64, device=Xd.device)[:, None].expand(n, d)
    start = torch.cummax(torch.where(first, pos, torch.zeros_like(pos)), dim=0).values       # first position of the tie group
    last = torch.ones(n, d, dtype=torch.bool, device=Xd.device)
    last[:-1] = srt[1:] != srt[:-1]
    big = torch.full_like(pos, float(n + 1))
    end = torch.flip(torch.cummin(torch.flip(torch.where(last, pos, big), [0]), dim=0).values, [0])
    avg = 0.5 * (start + end)
    ranks = torch.empty_like(avg)
    ranks.scatter_(0, order, avg)
    return ranks


def _analytic_tests(X: np.ndarray, pairs, test: str) -> List[IndepTestResult]:
    """``pearson`` / ``spearman`` (mi_tests.py:137-162): the correlations come from one covariance launch over all
    columns (of the data or of its ranks); the analytic two-sided p-values are scipy's formulas (the reference calls
    ``scipy.stats.pearsonr`` / ``spearmanr``): pearson  p = 2 * Beta(n/2 - 1, n/2 - 1).sf on [-1, 1],
    spearman  t = r sqrt((n - 2) / ((1 + r)(1 - r))),  p = 2 * t_{n-2}.sf(|t|)."""
    from scipy import special, stats
    _lib.require_device()
    Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).cuda()
    n = Xd.shape[0]
    R = _corr_matrix(_average_ranks(Xd) if test == "spearman" else Xd)
    const = (Xd == Xd[0:1]).all(dim=0).cpu().numpy()           # scipy: a constant input has no correlation (nan)
    R[const, :] = np.nan
    R[:, const] = np.nan
    out = []
    for i, j in pairs:
        r = float(R[i, j])
        if not np.isfinite(r):
            if test == "pearson":                                              # constant input: scipy returns (nan, nan)
                out.append(IndepTestResult(i=i, j=j, stat=float("nan"), pvalue=float("nan")))
            else:
                out.append(IndepTestResult(i=i, j=j, stat=0.0, pvalue=1.0))    # mi_tests.py:158-160
            continue
        r = max(-1.0, min(1.0, r))
        if test == "pearson":
            ab = n / 2.0 - 1.0
            p = 2.0 * float(special.betainc(ab, ab, 0.5 * (1.0 - abs(r))))
        else:
            dof = n - 2
            with np.errstate(divide="ignore"):
                t = r * np.sqrt(dof / ((r + 1.0) * (1.0 - r)))
            p = 2.0 * float(stats.t.sf(abs(t), dof))
        out.append(IndepTestResult(i=i, j=j, stat=abs(r), pvalue=min(p, 1.0)))
    return out


def pearson_stat_pvalue(x: np.ndarray, y: np.ndarray) -> Tuple[float, float]:
    """(|r|, p) of the Pearson correlation test (mi_tests.py:137-145)."""
    X, _ = _two_col(x, y)
    r = _analytic_tests(X, [(0, 1)], "pearson")[0]
    return r.stat, r.pvalue


def spearman_stat_pvalue(x: np.ndarray, y: np.ndarray) -> Tuple[float, float]:
    """(|rho|, p) of the Spearman rank correlation test (mi_tests.py:148-162)."""
    X, _ = _two_col(x, y)
    r = _analytic_tests(X, [(0, 1)], "spearman")[0]
    return r.stat, r.pvalue


def test_pairwise_independence(X: np.ndarray, pairs: Iterable[Tuple[int, int]], *, test: str = "hsic",
                               num_perm: int = 200, seed: int = 0) -> List[IndepTestResult]:
    """(stat, p-value) per pair (mi_tests.py:165-203): one RNG stream over the pairs in order, as the reference."""
    X = np.asarray(X)
    pairs = [(int(i), int(j)) for i, j in pairs]
    if test in ("pearson", "spearman"):
        return _analytic_tests(X, pairs, test)
    if test not in ("hsic", "dcor"):
        raise ValueError("test must be one of 'hsic', 'dcor', 'pearson', 'spearman'")
    if not pairs:
        return []
    rng = np.random.default_rng(seed)
    n = X.shape[0]
    stats = _pair_stats(X, pairs, test, _PermStream(rng, n, num_perm))
    out = []
    for q, (i, j) in enumerate(pairs):
        ge = int((stats[q, 1:] >= stats[q, 0]).sum())
        out.append(IndepTestResult(i=i, j=j, stat=float(stats[q, 0]), pvalue=float((ge + 1) / (num_perm + 1))))
    return out


test_pairwise_independence.__test__ = False          # not a pytest test


def get_I_from_full_pairwise_tests(X: np.ndarray, *, alpha: float = 0.05, test: str = "hsic", num_perm: int = 200,
                                   seed: int = 0, bonferroni: bool = True, undirected: bool = True,
                                   exclude_diagonal: bool = True) -> np.ndarray:
    """I = {(i, j): p > alpha_eff} over all pairs (mi_tests.py:206-247)."""
    X = np.asarray(X)
    n, d = X.shape
    if undirected:
        pairs = [(i, j) for i in range(d) for j in range(i + 1, d)]
    else:
        pairs = [(i, j) for i in range(d) for j in range(d) if not (exclude_diagonal and i == j)]
    results = test_pairwise_independence(X, pairs, test=test, num_perm=num_perm, seed=seed)
    m = len(results)
    alpha_eff = (alpha / m) if (bonferroni and m > 0) else alpha
    return np.asarray([(r.i, r.j) for r in results if r.pvalue > alpha_eff], dtype=int)
