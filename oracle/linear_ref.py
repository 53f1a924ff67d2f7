"""numpy restatement of ``DagmaLinear`` (reference: src/dagma/linear.py).

TEST INFRASTRUCTURE (see oracle/__init__.py) -- never imported by the product.

The arithmetic lives in un-vendored third-party routines (pyproject.toml:24-30,
unpinned): ``scipy.linalg.inv`` (getrf+getri), ``numpy.linalg.slogdet``, BLAS
``@``, ``scipy.special.expit``, ``np.logaddexp``.  This file calls the same
routines in the same order and association as the reference, so that with the
container's numpy 2.3 / scipy 1.18 it reproduces the reference bit for bit
(pinned by tests/test_oracle_golden.py against tests/golden/linear_*.npz, which
oracle/make_golden.py produced by running the unmodified reference).

Only the no-regulariser path is restated (``trek_reg is None`` -> value 0,
gradient zeros; notreks.py:684-689), which is the path BASELINE.json names.
"""
from __future__ import annotations

import numpy as np
import numpy.linalg as la
import scipy.linalg as sla
from scipy.special import expit


class OracleLinear:
    """State + methods of DagmaLinear needed by ``minimize``/``fit``.

    ``trace`` (optional list) receives one dict per inner iteration with the
    quantities the parity tests compare (linear.py:226-276).
    """

    def __init__(self, loss_type: str = "l2"):
        assert loss_type in ("l2", "logistic")            # linear.py:52-53
        self.loss_type = loss_type
        self.trace = None
        self.n_adam_calls = 0
        self.stage_iters = []
        self.events = []

    # ------------------------------------------------------------ set-up (fit :406-429)
    def prepare(self, X, lambda1, checkpoint=1000, exclude_edges=None, include_edges=None,
                center_inplace=True):
        self.X, self.lambda1, self.checkpoint = X, lambda1, checkpoint
        self.n, self.d = X.shape
        self.Id = np.eye(self.d)
        if self.loss_type == "l2":                         # :410-411 (in place, Q8)
            if center_inplace:
                self.X -= X.mean(axis=0, keepdims=True)
            else:
                self.X = X - X.mean(axis=0, keepdims=True)
        self.exc_r = self.exc_c = self.inc_r = self.inc_c = None
        if exclude_edges is not None and _is_edge_tuple(exclude_edges):   # :416-420 (Q5)
            self.exc_r, self.exc_c = zip(*exclude_edges)
        if include_edges is not None and _is_edge_tuple(include_edges):   # :422-426
            self.inc_r, self.inc_c = zip(*include_edges)
        self.cov = self.X.T @ self.X / float(self.n)        # :428
        return self

    # ------------------------------------------------------------ score (:70-94)
    def score(self, W):
        if self.loss_type == "l2":
            dif = self.Id - W
            rhs = self.cov @ dif
            loss = 0.5 * np.trace(dif.T @ rhs)
            G = -rhs
        else:
            R = self.X @ W
            loss = 1.0 / self.n * (np.logaddexp(0, R) - self.X * R).sum()
            G = (1.0 / self.n * self.X.T) @ expit(R) - self.cov
        return loss, G

    # ------------------------------------------------------------ h (:97-116)
    def h(self, W, s=1.0):
        M = s * self.Id - W * W
        hval = -la.slogdet(M)[1] + self.d * np.log(s)
        G_h = 2 * W * sla.inv(M).T
        return hval, G_h

    # ------------------------------------------------------------ objective (:118-135)
    def func(self, W, mu, s=1.0):
        score, _ = self.score(W)
        hval, _ = self.h(W, s)
        obj = mu * (score + self.lambda1 * np.abs(W).sum()) + hval
        return obj, score, hval, 0.0

    # ------------------------------------------------------------ Adam (:138-163)
    def adam_update(self, grad, it, b1, b2):
        self.n_adam_calls += 1
        self.opt_m = self.opt_m * b1 + (1 - b1) * grad
        self.opt_v = self.opt_v * b2 + (1 - b2) * (grad ** 2)
        m_hat = self.opt_m / (1 - b1 ** it)
        v_hat = self.opt_v / (1 - b2 ** it)
        return m_hat / (np.sqrt(v_hat) + 1e-8)

    # ------------------------------------------------------------ minimize (:165-333)
    def minimize(self, W, mu, max_iter, s, lr, tol=1e-6, b1=0.99, b2=0.999):
        obj_prev = 1e16
        self.opt_m, self.opt_v = 0, 0                       # :215 (Q4)
        d = self.d
        mask_inc = np.zeros((d, d))
        if self.inc_c is not None:
            mask_inc[self.inc_r, self.inc_c] = -2 * mu * self.lambda1
        mask_exc = np.ones((d, d))
        if self.exc_c is not None:
            mask_exc[self.exc_r, self.exc_c] = 0.0
        self.last_iters = 0
        self.checkpoints = []
        grad = None
        for it in range(1, max_iter + 1):
            M = sla.inv(s * self.Id - W * W) + 1e-16        # :226 (Q2)
            while np.any(M < 0):                            # :230 (Q3)
                if it == 1 or s <= 0.9:
                    self.events.append(("out_of_domain", it, s))
                    return W, False
                W += lr * grad
                lr *= 0.5
                if lr <= 1e-16:
                    return W, True
                W -= lr * grad
                M = sla.inv(s * self.Id - W * W) + 1e-16
                self.events.append(("lr_halved", it, lr))
            if self.loss_type == "l2":
                G_score = -mu * self.cov @ (self.Id - W)    # :244
            elif getattr(self, "logistic_route", "reference") == "reference":
                G_score = mu / self.n * self.X.T @ expit(self.X @ W) - mu * self.cov   # :246
            else:
                # round-off-equivalent route (sum first, scale afterwards) -- NOT what the reference does; used by
                # the tests to measure the reference's own round-off envelope
                G_score = (self.X.T @ expit(self.X @ W)) * (mu / self.n) - mu * self.cov
            sgn = np.sign(W)
            Gobj = G_score + mu * self.lambda1 * sgn + 2 * W * M.T + mask_inc * sgn   # :248 (Q1,Q5)
            grad = self.adam_update(Gobj, it, b1, b2)       # :272
            if self.trace is not None:
                self.trace.append({"iter": it, "lr": lr, "Gobj": Gobj.copy(),
                                   "minM": float(M.min()), "dir": grad.copy(),
                                   "W_before": W.copy()})
            W -= lr * grad                                  # :275
            W *= mask_exc                                   # :276
            self.last_iters = it
            if it % self.checkpoint == 0 or it == max_iter:  # :279 (Q6)
                obj_new, score, hval, _ = self.func(W, mu, s)
                self.checkpoints.append((it, obj_new, score, hval))
                if np.abs((obj_prev - obj_new) / obj_prev) <= tol:
                    break
                obj_prev = obj_new
        self.last_lr = lr
        return W, True

    # ------------------------------------------------------------ fit (:335-462)
    def fit(self, X, lambda1=0.03, w_threshold=0.3, T=5, mu_init=1.0, mu_factor=0.1,
            s=(1.0, 0.9, 0.8, 0.7, 0.6), warm_iter=3e4, max_iter=6e4, lr=0.0003,
            checkpoint=1000, beta_1=0.99, beta_2=0.999, exclude_edges=None,
            include_edges=None, center_inplace=True):
        self.prepare(X, lambda1, checkpoint, exclude_edges, include_edges, center_inplace)
        self.W_est = np.zeros((self.d, self.d))
        mu = mu_init
        if isinstance(s, (list, tuple)):
            s = list(s)                                     # fresh list (Q9)
            if len(s) < T:
                s = s + (T - len(s)) * [s[-1]]
        else:
            s = T * [s]
        self.stage_iters = []
        for i in range(int(T)):
            lr_adam, success = lr, False
            inner = int(max_iter) if i == T - 1 else int(warm_iter)   # :445 (Q10)
            while success is False:
                W_temp, success = self.minimize(self.W_est.copy(), mu, inner, s[i],
                                                lr=lr_adam, b1=beta_1, b2=beta_2)
                if success is False:
                    lr_adam *= 0.5                          # :450
                    s[i] += 0.1                             # :451
            self.stage_iters.append(self.last_iters)
            self.W_est = W_temp
            mu *= mu_factor                                 # :453 (Q7)
        self.h_final, _ = self.h(self.W_est)                # :456 (Q11)
        self.score_final, _ = self.score(self.W_est)
        self.W_raw = self.W_est.copy()
        self.W_est[np.abs(self.W_est) < w_threshold] = 0    # :458
        return self.W_est


def _is_edge_tuple(edges) -> bool:
    return (type(edges) is tuple and type(edges[0]) is tuple
            and bool(np.all(np.array([len(e) for e in edges]) == 2)))


# ---------------------------------------------------------------------------
# Single-step helper used by step-level parity tests: apply ONE inner
# iteration to an explicit state snapshot (W, m, v, iter, lr).
# ---------------------------------------------------------------------------
def l2_step(W, m, v, it, lr, cov, mu, s, lambda1, b1=0.99, b2=0.999,
            mask_inc=None, mask_exc=None):
    """One pass of linear.py:226-276 for the l2 loss; returns dict of results."""
    d = W.shape[0]
    Id = np.eye(d)
    M = sla.inv(s * Id - W * W) + 1e-16
    feasible = not np.any(M < 0)
    G_score = -mu * cov @ (Id - W)
    sgn = np.sign(W)
    Gobj = G_score + mu * lambda1 * sgn + 2 * W * M.T
    if mask_inc is not None:
        Gobj = Gobj + mask_inc * sgn
    m2 = m * b1 + (1 - b1) * Gobj
    v2 = v * b2 + (1 - b2) * (Gobj ** 2)
    direction = (m2 / (1 - b1 ** it)) / (np.sqrt(v2 / (1 - b2 ** it)) + 1e-8)
    W2 = W - lr * direction
    if mask_exc is not None:
        W2 = W2 * mask_exc
    return {"W": W2, "m": m2, "v": v2, "Gobj": Gobj, "Minv": M, "feasible": feasible,
            "dir": direction}
