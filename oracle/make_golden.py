"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

The reference is imported from /root/reference/src (read-only tree, so byte-code
writing is disabled).  Inputs come from oracle/simulate.py with fixed seeds and
are stored in the fixture next to the reference's outputs, so the parity tests
never have to regenerate them.  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import simulate  # noqa: E402

from dagma.linear import DagmaLinear  # noqa: E402  (reference)
from dagma.nonlinear import DagmaMLP, DagmaNonlinear  # noqa: E402  (reference)
import notreks.notreks as ref_nt  # noqa: E402  (reference)

GOLD = os.path.join(ROOT, "tests", "golden")


class _Bar:
    def update(self, *_a, **_k):
        pass


class RecordingLinear(DagmaLinear):
    """Reference class with a tap on ``_adam_update`` (linear.py:272) -- records the
    raw gradient handed to Adam; nothing else is changed."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.rec_G = []
        self.n_calls = 0
        self.rec_limit = 0

    def _adam_update(self, grad, iter, beta_1, beta_2):
        self.n_calls += 1
        if len(self.rec_G) < self.rec_limit:
            self.rec_G.append(grad.copy())
        return super()._adam_update(grad, iter, beta_1, beta_2)


def _prep(model, X, lambda1, checkpoint, exclude=None, include=None):
    """The variable set-up of fit() (linear.py:406-429) without the path loop."""
    model.X, model.lambda1, model.checkpoint = X, lambda1, checkpoint
    model.n, model.d = X.shape
    model.Id = np.eye(model.d)
    if model.loss_type == "l2":
        model.X -= X.mean(axis=0, keepdims=True)
    model.exc_r = model.exc_c = model.inc_r = model.inc_c = None
    if exclude is not None:
        model.exc_r, model.exc_c = zip(*exclude)
    if include is not None:
        model.inc_r, model.inc_c = zip(*include)
    model.cov = X.T @ X / float(model.n)


def linear_stages(name, loss, d, n, k, seed, lambda1, stages, checkpoint, rec=8,
                  exclude=None, include=None, graph="ER"):
    sem = "gauss" if loss == "l2" else "logistic"
    X, W_true = simulate.make_linear_problem(d, k, n, graph, sem, seed)
    X_in = X.copy()
    model = RecordingLinear(loss)
    _prep(model, X, lambda1, checkpoint, exclude, include)
    W = np.zeros((d, d))
    out = {"X": X_in, "W_true": W_true, "lambda1": lambda1, "checkpoint": checkpoint,
           "stages": np.array(stages, dtype=np.float64)}
    if exclude is not None:
        out["exclude"] = np.array(exclude)
    if include is not None:
        out["include"] = np.array(include)
    for si, (mu, s, iters, lr) in enumerate(stages):
        model.rec_G, model.rec_limit = [], rec
        before = model.n_calls
        W, ok = model.minimize(W.copy(), mu, int(iters), s, lr=lr, pbar=_Bar())
        out[f"W_after_{si}"] = W.copy()
        out[f"ok_{si}"] = ok
        out[f"iters_{si}"] = model.n_calls - before
        out[f"G_first_{si}"] = np.array(model.rec_G)
        obj, score, h, _ = model._func(W, mu, s)
        out[f"obj_{si}"], out[f"score_{si}"], out[f"h_{si}"] = obj, score, h
    # value/gradient KATs at the last W
    out["score_val"], out["score_grad"] = model._score(W)
    out["h_val"], out["h_grad"] = model._h(W, 0.9)
    out["cov"] = model.cov
    np.savez_compressed(os.path.join(GOLD, name), **out)
    print(name, {k: out[k] for k in out if k.startswith("iters_")})


def linear_full_fit(name, d, n, k, seed, lambda1, graph="ER", **fit_kw):
    X, W_true = simulate.make_linear_problem(d, k, n, graph, "gauss", seed)
    X_in = X.copy()
    model = RecordingLinear("l2")
    counts = []
    orig = model.minimize

    def tapped(*a, **kw):
        before = model.n_calls
        r = orig(*a, **kw)
        counts.append((model.n_calls - before, bool(r[1])))
        return r

    model.minimize = tapped
    W_est = model.fit(X, lambda1=lambda1, s=[1.0, .9, .8, .7, .6], **fit_kw)
    np.savez_compressed(os.path.join(GOLD, name), X=X_in, W_true=W_true, lambda1=lambda1,
                        W_est=W_est, h_final=model.h_final, score_final=model.score_final,
                        minimize_calls=np.array(counts, dtype=np.int64),
                        fit_kw=np.array(sorted(fit_kw.items()), dtype=object) if False else
                        np.array([fit_kw.get("warm_iter", 3e4), fit_kw.get("max_iter", 6e4),
                                  fit_kw.get("checkpoint", 1000)]))
    print(name, counts, "nnz", int((W_est != 0).sum()), "true", int((W_true != 0).sum()))


def mlp_case(name, d, m1, n, seed, steps, lambda1=0.02, lambda2=0.005, mu=0.1, s=1.0, lr=2e-4,
             init_scale=0.1, dims=None):
    Xn, B = simulate.config_c3(seed=seed, n=n, d=d)
    torch.manual_seed(seed)
    dims = list(dims) if dims is not None else [d, m1, 1]
    m1 = dims[1]
    model = DagmaMLP(dims=dims, bias=True, dtype=torch.double)
    # fc1 is zero-initialised in the reference (nonlinear.py:37-38): perturb it so the
    # gradient KAT is non-trivial, then also record a run from the true zero init.
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        model.fc1.weight.copy_(init_scale * torch.randn(d * m1, d, generator=g, dtype=torch.double))
        model.fc1.bias.copy_(init_scale * torch.randn(d * m1, generator=g, dtype=torch.double))
    X = torch.from_numpy(Xn)
    out = {"X": Xn, "B": B, "dims": np.array(dims), "hyper": np.array([lambda1, lambda2, mu, s, lr])}
    for k_, v in model.state_dict().items():
        out["init." + k_] = v.detach().numpy().copy()
    # one autograd evaluation (nonlinear.py:213-222)
    eq = DagmaNonlinear(model)
    h_val = model.h_func(s)
    X_hat = model(X)
    score = eq.log_mse_loss(X_hat, X)
    obj = mu * (score + lambda1 * model.fc1_l1_reg()) + h_val
    obj.backward()
    out["obj"], out["score"], out["h"] = obj.item(), score.item(), h_val.item()
    out["X_hat"] = X_hat.detach().numpy()
    for k_, p in model.named_parameters():
        out["grad." + k_] = p.grad.detach().numpy().copy()
    out["adj"] = model.fc1_to_adj()
    # `steps` iterations of the reference minimize from this init
    for p in model.parameters():
        p.grad = None
    eq.X = X
    eq.checkpoint = 10 ** 9
    ok = eq.minimize(steps, lr, lambda1, lambda2, mu, s, pbar=_Bar())
    out["ok"] = ok
    for k_, v in model.state_dict().items():
        out[f"after{steps}." + k_] = v.detach().numpy().copy()
    out["steps"] = steps
    np.savez_compressed(os.path.join(GOLD, name), **out)
    print(name, "obj", out["obj"], "ok", ok)


def mlp_deep_cases():
    """Deeper LocallyConnected stacks and the degenerate dims = [d, 1] (nonlinear.py:39-43)."""
    mlp_case("mlp_deep_d6", 6, 5, 60, 3, steps=20, dims=[6, 5, 3, 1])
    mlp_case("mlp_deep4_d5", 5, 4, 40, 4, steps=15, dims=[5, 4, 3, 2, 1])
    mlp_case("mlp_lin_d6", 6, 1, 60, 5, steps=20, dims=[6, 1])


def mlp_fit_case(name, d, m1, n, seed, **kw):
    Xn, B = simulate.config_c3(seed=seed, n=n, d=d)
    torch.manual_seed(seed)
    model = DagmaMLP(dims=[d, m1, 1], bias=True, dtype=torch.double)
    out = {"X": Xn, "B": B, "dims": np.array([d, m1, 1])}
    for k_, v in model.state_dict().items():
        out["init." + k_] = v.detach().numpy().copy()
    eq = DagmaNonlinear(model)
    W = eq.fit(Xn, **kw)
    out["W_est"] = W
    out["W_raw"] = model.fc1_to_adj()
    out["kw"] = np.array([kw["T"], kw["warm_iter"], kw["max_iter"], kw["checkpoint"]])
    for k_, v in model.state_dict().items():
        out["final." + k_] = v.detach().numpy().copy()
    np.savez_compressed(os.path.join(GOLD, name), **out)
    print(name, "nnz", int((W != 0).sum()))


def notreks_case(name, d, seed):
    rng = np.random.default_rng(seed)
    out = {}
    # (i) raw logdet on a nonnegative matrix with spectral radius < s
    n = 2 * d
    A = rng.uniform(0, 1, size=(n, n)) * (rng.random((n, n)) < 0.2)
    A *= 0.5 / max(np.abs(np.linalg.eigvals(A)).max(), 1e-12)
    for s in (1.0, 0.8):
        h, G = ref_nt.logdet_acyc_value_gradA(torch.from_numpy(A), s=s)
        out[f"h_s{s}"], out[f"G_s{s}"] = h.item(), G.numpy()
    out["A"] = A
    # (ii) TCC-logdet on W with a pair set I
    W = rng.normal(size=(d, d)) * (rng.random((d, d)) < 0.3) * 0.4
    np.fill_diagonal(W, 0.0)
    pairs = np.array([(0, 1), (2, 3), (1, 4)])
    out["W"], out["pairs"] = W, pairs
    import contextlib
    import io
    for version in ("DAG_learning", "exact_trek_graph"):
        with contextlib.redirect_stdout(io.StringIO()):
            pen, gW = ref_nt.trek_cycle_coupling_value_gradW(
                torch.from_numpy(W), pairs, w=0.7, cycle_penalty="logdet", version=version, s=1.0)
        out[f"tcc_pen_{version}"], out[f"tcc_grad_{version}"] = pen.item(), gW.numpy()
    # (iii) the no-op fast path (notreks.py:684-689)
    v, g = ref_nt.trek_value_grad(W, None, torch_dtype=torch.double, device=torch.device("cpu"))
    out["noop_val"], out["noop_grad"] = v, g
    np.savez_compressed(os.path.join(GOLD, name), **out)
    print(name, out["h_s1.0"], out["tcc_pen_DAG_learning"])


if __name__ == "__main__" and "--deep-only" in sys.argv:
    mlp_deep_cases()
    sys.exit(0)

if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    # chained short stages (step-level / short-horizon parity)
    linear_stages("linear_l2_d20", "l2", 20, 500, 2, 0, 0.02,
                  [(1.0, 1.0, 400, 3e-4), (0.1, 0.9, 400, 3e-4), (0.01, 0.8, 300, 3e-4)], 100)
    linear_stages("linear_l2_d64", "l2", 64, 400, 4, 1000, 0.02,
                  [(1.0, 1.0, 300, 3e-4), (0.1, 0.9, 200, 3e-4)], 100, rec=3)
    linear_stages("linear_l2_d7_masks", "l2", 7, 200, 2, 3, 0.03,
                  [(1.0, 1.0, 300, 3e-4), (0.1, 0.9, 300, 3e-4)], 50,
                  exclude=((0, 1), (2, 5), (6, 3)), include=((1, 2), (4, 0)))
    linear_stages("linear_logistic_d12", "logistic", 12, 800, 2, 5, 0.02,
                  [(1.0, 1.0, 300, 3e-4), (0.1, 0.9, 300, 3e-4)], 100)
    linear_stages("linear_l2_d100", "l2", 100, 300, 2, 7, 0.02,
                  [(1.0, 1.0, 60, 3e-4), (0.1, 0.9, 40, 3e-4)], 20, rec=2)
    # full default fit, C1 (about 5 s) and a reduced-schedule C4 instance
    linear_full_fit("fit_c1_seed0", 20, 500, 2, 0, 0.02)
    linear_full_fit("fit_c4_short", 64, 1000, 4, 1000, 0.02, warm_iter=2000, max_iter=3000,
                    checkpoint=500)
    mlp_case("mlp_d7", 7, 5, 50, 0, steps=25)
    mlp_case("mlp_d40", 40, 10, 300, 1, steps=10, init_scale=0.02)
    mlp_fit_case("mlp_fit_d5", 5, 4, 200, 2, T=2, warm_iter=300, max_iter=400, checkpoint=100,
                 lambda1=0.02, lambda2=0.005)
    notreks_case("notreks_logdet", 6, 0)
    mlp_deep_cases()
