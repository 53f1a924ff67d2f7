"""PST trek penalty (reference: src/notreks/notreks.py:454-619, CR-delimited line numbers) for every series of
the reference -- ``inv``, ``log``, ``exp``, ``binom`` -- and every scalar aggregation -- ``mean``, ``sum``,
``max``, ``lse`` -- as fixed launch sequences on the B200 kernels (fused inverse + FP64 DMMA GEMM chains).

    W2 = W o W;  F = series(W2);  H = F^T F;  pst = agg_{(i,j) in I} H[i, j]            (notreks.py:492-527, 599-619)

The reference differentiates with torch autograd; here every adjoint is closed form and is kept TRANSPOSED
(``B = Abar^T``) so that each step is a plain row-major product ``C = A @ B`` of ``dagma_gemm_f64``:

    G_H   = d pst / d H       (pair weights: 1/m, 1, ties-shared one-hot, softmax)
    G_F   = F (G_H + G_H^T),  Gt = G_F^T
    inv   F = X = ((1 + eps) I - W2)^{-1}:                 GT = X Gt X
    log   P_1 = W2, P_{k+1} = P_k W2, F = I + sum_k P_k / k (k <= K = 2d unless K_log is given; ``s`` never
          reaches the series -- dead duplicate branch, notreks.py:509-513 vs 521-525):
              B_K = Gt / K;  B_k = Gt / k + W2 B_{k+1};  T += B_{k+1} P_k;   GT = T + B_1
    binom P_1 = I + W2, P_{k+1} = P_k P_1, F = P_d:          B_d = Gt;  T += B_{k+1} P_k;  B_k = P_1 B_{k+1};  GT = T + B_1
    exp   F = expm(W2) (degree-18 Taylor by Horner on W2 / 2^s, then s squarings):
              GT = top-right block of expm([[W2, Gt], [0, W2]])   (Frechet derivative L(W2, Gt) = L(W2^T, G_F)^T)
    d pst / d W = 2 W o GT^T

``GT`` is exactly what ``dagma_linear_update_ex_f64`` consumes (``extra_t``), so the same object serves the
stand-alone ``trek_value_grad`` and the CUDA-graph-captured inner iteration of ``DagmaLinear`` (no host
synchronisation anywhere in ``grad()``).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._large import gemm

SEQS = ("inv", "log", "exp", "binom")
AGGS = ("mean", "sum", "max", "lse")
EXP_SQUARINGS = 12            # W2 / 2^12: the Taylor remainder is below 1e-20 for ||W2||_1 <= 2048
EXP_DEGREE = 18
MAX_CHAIN_BYTES = 8 << 30     # stored powers of the log / binom chains


class PstEngine:
    """Device buffers + launch sequences of one PST penalty (fixed d, pair set, series and aggregation)."""

    def __init__(self, d: int, I, seq: str = "exp", agg: str = "mean", *, eps_inv: float = 1e-8, K_log=None):
        _lib.require_device()
        self.lib = _lib.load()
        seq, agg = str(seq).lower().strip(), str(agg).lower().strip()
        if seq not in SEQS:
            raise ValueError("seq must be one of {'exp','log','inv','binom'}")
        if agg not in AGGS:
            raise ValueError("agg must be one of {'mean','sum','max','lse','none'}" if agg != "none" else
                             "agg='none' returns a vector: use PstEngine.pair_values()")
        if eps_inv < 0:
            raise ValueError("eps_inv must be >= 0")
        I_np = np.asarray(I, dtype=np.int64)
        if I_np.ndim != 2 or I_np.shape[1] != 2:
            raise ValueError("I must be an array-like of shape (m, 2) with integer indices.")
        self.d, self.seq, self.agg, self.eps_inv = int(d), seq, agg, float(eps_inv)
        self.m = I_np.shape[0]
        f64 = dict(dtype=torch.float64, device="cuda")
        d = self.d
        self.rows = torch.as_tensor(I_np[:, 0], device="cuda")
        self.cols = torch.as_tensor(I_np[:, 1], device="cuda")
        self.W2 = torch.empty(d, d, **f64)
        self.F = torch.empty(d, d, **f64)
        self.H = torch.empty(d, d, **f64)
        self.GH = torch.zeros(d, d, **f64)
        self.GF = torch.empty(d, d, **f64)
        self.Gt = torch.empty(d, d, **f64)
        self.GT = torch.zeros(d, d, **f64)
        self.tmp = torch.empty(d, d, **f64)
        self.val = torch.zeros((), **f64)
        self.info = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.sc = torch.zeros(4, **f64)
        if seq == "inv":
            self.ws = torch.empty(self.lib.dagma_large_workspace_bytes(d) // 8 + 8, **f64)
        elif seq in ("log", "binom"):
            self.K = int(d) if seq == "binom" else (2 * int(d) if K_log is None else int(K_log))
            if self.K < 1:
                raise ValueError("K must be >= 1")
            if self.K * d * d * 8 > MAX_CHAIN_BYTES:
                raise NotImplementedError(f"the {seq} chain would store {self.K} powers of a {d} x {d} matrix")
            self.P = torch.empty(self.K, d, d, **f64)
            self.B = torch.empty(2, d, d, **f64)
            self.T = torch.empty(d, d, **f64)
        else:
            self.E = torch.empty(3, 2 * d, 2 * d, **f64)      # block matrix of the Frechet derivative + Horner ping-pong
            self.e1 = torch.empty(2, d, d, **f64)

    # ------------------------------------------------------------------ series
    def _expm(self, M, R, S):
        """R (and S as scratch) <- expm(M); M is overwritten by M / 2^s.  Returns the buffer holding the result."""
        M.mul_(2.0 ** -EXP_SQUARINGS)
        R.zero_()
        R.diagonal().fill_(1.0)
        R.add_(M, alpha=1.0 / EXP_DEGREE)                   # I + M / 18
        for j in range(EXP_DEGREE - 1, 0, -1):              # R <- I + (M R) / j
            gemm(M, R, S, alpha=1.0 / j)
            S.diagonal().add_(1.0)
            R, S = S, R
        for _ in range(EXP_SQUARINGS):
            gemm(R, R, S)
            R, S = S, R
        return R

    def forward(self, W: torch.Tensor) -> None:
        """F and H = F^T F at W (device d x d)."""
        d = self.d
        torch.mul(W, W, out=self.W2)
        if self.seq == "inv":
            # X = ((1 + eps) I - W o W)^{-1} by the fused inverse kernel (square_input = 1 on W itself)
            _lib.check(self.lib.dagma_logdet_inv_ws_f64(
                _lib.stream_ptr(), d, 1.0 + self.eps_inv, W.data_ptr(), W.stride(0), 1, self.sc.data_ptr(),
                self.sc.data_ptr() + 8, self.F.data_ptr(), None, d, self.sc.data_ptr() + 16, self.info.data_ptr(),
                self.ws.data_ptr(), self.ws.numel() * 8), "dagma_logdet_inv_ws_f64")
        elif self.seq == "log":
            P = self.P
            P[0].copy_(self.W2)
            for k in range(1, self.K):
                gemm(P[k - 1], self.W2, P[k])
            self.F.zero_()
            self.F.diagonal().fill_(1.0)
            for k in range(self.K):
                self.F.add_(P[k], alpha=1.0 / (k + 1))
        elif self.seq == "binom":
            P = self.P
            P[0].copy_(self.W2)
            P[0].diagonal().add_(1.0)
            for k in range(1, self.K):
                gemm(P[k - 1], P[0], P[k])
            self.F.copy_(P[self.K - 1])
        else:
            self.e1[0].copy_(self.W2)
            R = self._expm(self.e1[0], self.e1[1], self.tmp)
            self.F.copy_(R)
        gemm(self.F, self.F, self.H, trans_a=True)

    def pair_values(self) -> torch.Tensor:
        return self.H[self.rows, self.cols]

    def _aggregate(self, want_weights: bool) -> None:
        """val (0-dim device tensor) and, if asked, G_H = d val / d H -- no host synchronisation."""
        vals = self.pair_values()
        if self.agg == "mean":
            self.val.copy_(vals.mean())
            w = torch.full_like(vals, 1.0 / self.m)
        elif self.agg == "sum":
            self.val.copy_(vals.sum())
            w = torch.ones_like(vals)
        elif self.agg == "max":
            mx = vals.max()
            self.val.copy_(mx)
            tie = (vals == mx).to(vals.dtype)               # torch shares the gradient of max() between ties
            w = tie / tie.sum()
        else:
            self.val.copy_(torch.logsumexp(vals, dim=0))
            w = torch.softmax(vals, dim=0)
        if want_weights:
            self.GH.zero_()
            self.GH.index_put_((self.rows, self.cols), w, accumulate=True)

    def value(self, W: torch.Tensor) -> torch.Tensor:
        self.forward(W)
        self._aggregate(False)
        return self.val

    def grad(self, W: torch.Tensor) -> torch.Tensor:
        """value in ``self.val`` and GT with d pst / d W = 2 W o GT^T."""
        d = self.d
        self.forward(W)
        self._aggregate(True)
        torch.add(self.GH, self.GH.t(), out=self.tmp)       # M_s
        gemm(self.F, self.tmp, self.GF)
        self.Gt.copy_(self.GF.t())
        if self.seq == "inv":
            gemm(self.Gt, self.F, self.tmp)
            gemm(self.F, self.tmp, self.GT)
        elif self.seq in ("log", "binom"):
            P, B, T, K = self.P, self.B, self.T, self.K
            log = self.seq == "log"
            cur = 0
            B[cur].copy_(self.Gt)
            if log:
                B[cur].mul_(1.0 / K)
            T.zero_()
            lhs = self.W2 if log else P[0]
            for k in range(K - 1, 0, -1):                   # B[cur] = B_{k+1}
                gemm(B[cur], P[k - 1], T, beta=1.0)         # T += B_{k+1} P_k
                gemm(lhs, B[cur], B[1 - cur])               # B_k = lhs B_{k+1} (+ Gt / k)
                if log:
                    B[1 - cur].add_(self.Gt, alpha=1.0 / k)
                cur = 1 - cur
            torch.add(T, B[cur], out=self.GT)
        else:
            E = self.E
            E[0].zero_()
            E[0][:d, :d].copy_(self.W2)
            E[0][d:, d:].copy_(self.W2)
            E[0][:d, d:].copy_(self.Gt)
            R = self._expm(E[0], E[1], E[2])
            self.GT.copy_(R[:d, d:])
        return self.GT

    # ------------------------------------------------------------------ host conveniences
    def check_domain(self, W: torch.Tensor) -> None:
        """Host-side guards of the eager API (they synchronise, so the captured iteration does not call them)."""
        if self.seq == "inv" and int(self.info.item()) != 0:
            raise _lib.DagmaB200Error("(1 + eps) I - W o W is not an M-matrix: outside the domain of the fused inverse")
        if self.seq == "exp":
            n1 = float((W * W).abs().sum(dim=0).max().item())
            if n1 > 0.5 * 2.0 ** EXP_SQUARINGS:
                raise _lib.DagmaB200Error(f"||W o W||_1 = {n1:.3g} is outside the range of the fixed-scaling expm")

    def value_grad_host(self, W_np: np.ndarray, want_grad: bool = True):
        Wd = torch.as_tensor(np.ascontiguousarray(W_np, dtype=np.float64)).cuda()
        if want_grad:
            GT = self.grad(Wd)
            self.check_domain(Wd)
            return float(self.val.item()), (2.0 * Wd * GT.t()).cpu().numpy()
        self.value(Wd)
        self.check_domain(Wd)
        return float(self.val.item()), None
