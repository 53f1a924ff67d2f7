"""numpy restatement of ``DagmaMLP`` / ``DagmaNonlinear`` (reference:
src/dagma/nonlinear.py, src/dagma/locally_connected.py).

TEST INFRASTRUCTURE (see oracle/__init__.py) -- never imported by the product.

The reference differentiates with torch autograd and steps with
``torch.optim.Adam`` (nonlinear.py:208-223; torch is an unpinned dependency,
pyproject.toml:24-30).  Here the backward pass is written in closed form and
Adam is restated from torch's single-tensor implementation
(``exp_avg.lerp_``, ``addcmul_``, ``denom = sqrt(v)/sqrt(bc2) + eps``,
``param -= lr/bc1 * m/denom``; weight decay added to the gradient first).
Pinned by tests/golden/mlp_*.npz (reference autograd + optimizer outputs,
produced by oracle/make_golden.py); agreement is to round-off (~1e-15), not
bit-exact, because the summation order of the reductions differs.

Parameter naming follows the reference ``state_dict``: ``fc1.weight [d*m1, d]``,
``fc1.bias [d*m1]``, ``fc2.0.weight [d, m1, 1]``, ``fc2.0.bias [d, 1]``
(nonlinear.py:36-43); deeper stacks ``dims = [d, m1, m2, ..., 1]`` add ``fc2.l.weight
[d, m_l, m_{l+1}]`` / ``fc2.l.bias [d, m_{l+1}]``; ``dims = [d, 1]`` has no ``fc2`` at all.
"""
from __future__ import annotations

import copy

import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


class OracleMLP:
    def __init__(self, dims, params: dict):
        assert len(dims) >= 2 and dims[-1] == 1              # nonlinear.py:32-33
        self.dims, self.d = list(dims), dims[0]
        self.p = {k: np.array(v, dtype=np.float64) for k, v in params.items()}

    # forward (nonlinear.py:60-65; locally_connected.py:70-74)
    def forward(self, X):
        d, m1 = self.d, self.dims[1]
        z = X @ self.p["fc1.weight"].T + self.p["fc1.bias"]
        x = z.reshape(-1, d, m1)
        l = 0
        while f"fc2.{l}.weight" in self.p:
            x = _sigmoid(x)
            x = np.einsum("njk,jkm->njm", x, self.p[f"fc2.{l}.weight"]) + self.p[f"fc2.{l}.bias"]
            l += 1
        return x[:, :, 0]

    # adjacency A[i, j] = sum_k w[j, k, i]^2 (nonlinear.py:82-84)
    def adj_sq(self):
        d = self.d
        w = self.p["fc1.weight"].reshape(d, -1, d)
        return (w ** 2).sum(axis=1).T

    def h_func(self, s=1.0):                                  # nonlinear.py:85
        A = self.adj_sq()
        return -np.linalg.slogdet(s * np.eye(self.d) - A)[1] + self.d * np.log(s)

    def fc1_l1_reg(self):                                     # nonlinear.py:97
        return np.abs(self.p["fc1.weight"]).sum()

    def fc1_to_adj(self):                                     # nonlinear.py:110-115
        return np.sqrt(self.adj_sq())

    # value and gradient of  mu*(score + lambda1*|fc1|_1) + h   (nonlinear.py:214-222), any stack
    # dims = [d, m1, ..., 1] of LocallyConnected layers (nonlinear.py:39-43, 60-65)
    def obj_and_grads(self, X, lambda1, mu, s):
        d, m1 = self.d, self.dims[1]
        n = X.shape[0]
        L = len(self.dims) - 2
        w1, b1 = self.p["fc1.weight"], self.p["fc1.bias"]
        A = self.adj_sq()
        Mm = s * np.eye(d) - A
        h = -np.linalg.slogdet(Mm)[1] + d * np.log(s)
        Minv = np.linalg.inv(Mm)
        x = (X @ w1.T + b1).reshape(n, d, m1)
        acts = []
        for l in range(L):
            H = _sigmoid(x)
            acts.append(H)
            x = np.einsum("njk,jkm->njm", H, self.p[f"fc2.{l}.weight"]) + self.p[f"fc2.{l}.bias"]
        out = x[:, :, 0]
        res = out - X
        S = (res ** 2).sum()
        score = 0.5 * d * np.log(S / n)                        # nonlinear.py:158
        dx = ((d / S) * res)[:, :, None]
        grads = {}
        for l in reversed(range(L)):
            H, Wl = acts[l], self.p[f"fc2.{l}.weight"]
            grads[f"fc2.{l}.weight"] = mu * np.einsum("njk,njm->jkm", H, dx)
            grads[f"fc2.{l}.bias"] = mu * dx.sum(axis=0)
            dx = np.einsum("njm,jkm->njk", dx, Wl) * H * (1 - H)
        dZ = dx.reshape(n, d * m1)
        gw1_score = dZ.T @ X
        gb1 = dZ.sum(axis=0)
        # dh/dw[j,k,i] = 2 w[j,k,i] * Minv[j,i]
        w3 = w1.reshape(d, m1, d)
        gh = (2.0 * w3 * Minv[:, None, :]).reshape(d * m1, d)
        obj = mu * (score + lambda1 * np.abs(w1).sum()) + h
        grads["fc1.weight"] = mu * (gw1_score + lambda1 * np.sign(w1)) + gh
        grads["fc1.bias"] = mu * gb1
        return obj, score, h, grads


class OracleNonlinear:
    """``DagmaNonlinear.minimize`` / ``fit`` (nonlinear.py:161-331)."""

    def __init__(self, model: OracleMLP):
        self.model = model
        self.n_iters = 0
        self.trace = None

    def minimize(self, X, max_iter, lr, lambda1, lambda2, mu, s, lr_decay=False,
                 tol=1e-6, checkpoint=1000):
        p = self.model.p
        m = {k: np.zeros_like(v) for k, v in p.items()}       # optimizer re-created (Q13)
        v2 = {k: np.zeros_like(v) for k, v in p.items()}
        b1, b2, eps = 0.99, 0.999, 1e-8
        wd = mu * lambda2
        obj_prev = 1e16
        for i in range(max_iter):
            obj, score, h, g = self.model.obj_and_grads(X, lambda1, mu, s)
            if h < 0:                                          # nonlinear.py:215-217
                return False
            step = i + 1
            bc1 = 1 - b1 ** step
            bc2 = 1 - b2 ** step
            for k in p:
                gk = g[k] + wd * p[k]
                m[k] = m[k] + (gk - m[k]) * (1 - b1)
                v2[k] = v2[k] * b2 + (1 - b2) * gk * gk
                denom = np.sqrt(v2[k]) / np.sqrt(bc2) + eps
                p[k] = p[k] - (lr / bc1) * (m[k] / denom)
            self.n_iters += 1
            if self.trace is not None:
                self.trace.append({"i": i, "obj": obj, "score": score, "h": h})
            if lr_decay and (i + 1) % 1000 == 0:
                lr *= 0.8                                      # ExponentialLR(gamma=.8)
            if i % checkpoint == 0 or i == max_iter - 1:       # :226 (Q6)
                if np.abs((obj_prev - obj) / obj_prev) <= tol:
                    break
                obj_prev = obj
        return True

    def fit(self, X, lambda1=0.02, lambda2=0.005, T=4, mu_init=0.1, mu_factor=0.1, s=1.0,
            warm_iter=5e4, max_iter=8e4, lr=0.0002, w_threshold=0.3, checkpoint=1000):
        mu = mu_init
        s = list(s) if isinstance(s, (list, tuple)) else T * [s]
        if len(s) < T:
            s = s + (T - len(s)) * [s[-1]]
        for i in range(int(T)):
            success, s_cur = False, s[i]
            inner = int(max_iter) if i == T - 1 else int(warm_iter)
            saved = copy.deepcopy(self.model.p)
            lr_decay = False
            while success is False:
                success = self.minimize(X, inner, lr, lambda1, lambda2, mu, s_cur, lr_decay,
                                        checkpoint=checkpoint)
                if success is False:
                    self.model.p = copy.deepcopy(saved)
                    lr *= 0.5
                    lr_decay = True
                    if lr < 1e-10:
                        break
                    s_cur = 1
            mu *= mu_factor
        W = self.model.fc1_to_adj()
        W[np.abs(W) < w_threshold] = 0
        return W
