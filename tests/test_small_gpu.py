"""GPU parity of the on-chip (d <= 64 / inverse d <= 128) path against the oracle and the
golden fixtures recorded from the reference.  Tolerances follow SURVEY.md 7.4 / north_star:
per-call h / gradient 1e-9 max-norm relative; short-horizon W (mu >= 0.1 stages) 1e-6
absolute (measured ~1e-10); full fits: identical thresholded edge set."""
import numpy as np
import pytest
import torch

from oracle import simulate
from oracle.linear_ref import OracleLinear

pytestmark = pytest.mark.gpu


def _relmax(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("d", [1, 2, 6, 16, 17, 20, 33, 40, 64, 65, 100, 128])
@pytest.mark.parametrize("square", [True, False])
def test_logdet_inv_vs_numpy(d, square):
    from midagma_b200.linear import logdet_inv
    rng = np.random.default_rng(d)
    batch = 5
    A = rng.normal(size=(batch, d, d)) * (rng.random((batch, d, d)) < 0.3)
    if not square:
        A = np.abs(A)
    B = A * A if square else A
    for b in range(batch):   # spectral radius 0.6 < s
        B_b = B[b]
        rho = max(np.abs(np.linalg.eigvals(B_b)).max(), 1e-12)
        A[b] *= np.sqrt(0.6 / rho) if square else 0.6 / rho
    s = 0.9
    out = logdet_inv(torch.from_numpy(A).cuda(), s=s, square_input=square, want_inv=True, want_grad=True)
    for b in range(batch):
        M = s * np.eye(d) - (A[b] * A[b] if square else A[b])
        Minv = np.linalg.inv(M)
        lad = np.linalg.slogdet(M)[1]
        assert abs(out["logabsdet"][b].item() - lad) <= 1e-11 * max(1.0, abs(lad))
        assert abs(out["h"][b].item() - (-lad + d * np.log(s))) <= 1e-11 * max(1.0, abs(lad))
        assert _relmax(out["minv"][b].cpu().numpy(), Minv) <= 1e-11
        G = 2 * A[b] * Minv.T if square else Minv.T
        assert _relmax(out["grad"][b].cpu().numpy(), G) <= 1e-11
        assert int(out["info"][b].item()) == 0
        assert abs(out["min_entry"][b].item() - Minv.min()) <= 1e-12 * np.abs(Minv).max()


def test_logdet_known_answers():
    from midagma_b200.linear import logdet_inv
    d = 6
    W = np.triu(np.arange(36.0).reshape(6, 6) / 20, 1)                 # a DAG: h = 0
    P = np.eye(d)[[3, 0, 4, 1, 5, 2]]
    A = np.stack([np.zeros((d, d)), P @ W @ P.T])
    out = logdet_inv(torch.from_numpy(A).cuda(), s=1.0)
    assert abs(out["h"][0].item()) == 0.0 and abs(out["h"][1].item()) < 1e-13
    a, b, s = 0.4, -0.7, 0.9                                            # 2-cycle closed form
    out = logdet_inv(torch.tensor([[[0, a], [b, 0]]], dtype=torch.float64).cuda(), s=s)
    det = s * s - a * a * b * b
    assert abs(out["h"][0].item() - (-np.log(det) + 2 * np.log(s))) < 1e-14
    np.testing.assert_allclose(out["grad"][0].cpu().numpy(),
                               [[0, 2 * a * b * b / det], [2 * b * a * a / det, 0]], rtol=1e-13)
    # outside the M-matrix domain: flagged
    out = logdet_inv(torch.tensor([[[0, 1.2], [1.1, 0]]], dtype=torch.float64).cuda(), s=1.0)
    assert int(out["info"][0].item()) != 0


def _edges(g, key):
    return tuple(tuple(int(x) for x in e) for e in g[key]) if key in g.files else None


@pytest.mark.parametrize("name", ["linear_l2_d20", "linear_l2_d64", "linear_l2_d7_masks"])
def test_minimize_stages_vs_reference(golden, name):
    """Chained DagmaLinear.minimize stages: the drop-in class against the reference trace."""
    from midagma_b200 import DagmaLinear
    g = golden(name)
    X = g["X"].copy()
    model = DagmaLinear("l2")
    # set-up exactly as fit() would, then drive minimize by hand like the fixture did
    model.fit(X, lambda1=float(g["lambda1"]), T=1, warm_iter=0, max_iter=0, checkpoint=int(g["checkpoint"]),
              exclude_edges=_edges(g, "exclude"), include_edges=_edges(g, "include"))
    assert _relmax(model.cov, g["cov"]) <= 1e-13
    d = model.d
    W = np.zeros((d, d))
    for si, (mu, s, iters, lr) in enumerate(g["stages"]):
        W, ok = model.minimize(W.copy(), mu, int(iters), s, lr=lr)
        assert ok == bool(g[f"ok_{si}"])
        assert model.last_iters == int(g[f"iters_{si}"])
        ref = g[f"W_after_{si}"]
        err = np.abs(W - ref).max()
        print(name, "stage", si, "max|dW| =", err)
        assert err <= 1e-6, (si, err)
        if mu >= 0.1:
            assert err <= 1e-8, (si, err)
        st, it, obj, score, h, _ = model.checkpoint_log[-1]
        assert abs(obj - float(g[f"obj_{si}"])) <= 1e-8 * abs(float(g[f"obj_{si}"]))
        assert abs(score - float(g[f"score_{si}"])) <= 1e-8 * abs(float(g[f"score_{si}"]))
        assert abs(h - float(g[f"h_{si}"])) <= 1e-8 * max(abs(float(g[f"h_{si}"])), 1e-6)
    if "exclude" in g.files:
        for (i, j) in g["exclude"]:
            assert W[i, j] == 0.0


def test_single_steps_vs_oracle():
    """K <= 100 steps from W = 0: every W within 1e-9 of the oracle."""
    from midagma_b200 import minimize_batch
    X, _ = simulate.make_linear_problem(48, 3, 400, "ER", "gauss", 11)
    o = OracleLinear("l2").prepare(X, 0.03, checkpoint=10 ** 9)
    for K in (1, 2, 7, 100):
        W_ref, _ = o.minimize(np.zeros((48, 48)), 1.0, K, 1.0, 3e-4)
        W, ok, st = minimize_batch(np.zeros((1, 48, 48)), o.cov[None], 0.03, 1.0, K, 1.0, 3e-4, checkpoint=10 ** 9)
        assert ok[0] and int(st[0, 0, 0]) == K
        assert np.abs(W[0] - W_ref).max() <= 1e-9 * max(np.abs(W_ref).max(), 1e-300), K


def test_full_fit_c1_edge_set(golden):
    from midagma_b200 import DagmaLinear
    g = golden("fit_c1_seed0")
    X = g["X"].copy()
    model = DagmaLinear("l2")
    W = model.fit(X, lambda1=float(g["lambda1"]), s=[1.0, .9, .8, .7, .6])
    ref = g["W_est"]
    assert simulate.edge_set_distance(W, ref) == 0
    ref_iters = [int(c) for c, _ in g["minimize_calls"]]
    print("stage iters", model.stage_iters, "reference", ref_iters, "max|dW|", np.abs(W - ref).max())
    for a, b in zip(model.stage_iters, ref_iters):
        assert abs(a - b) <= 1000          # knife-edge tol test may flip one checkpoint
    assert np.abs(W - ref).max() <= 5e-3   # reference's own round-off envelope is ~1e-3
    assert abs(model.h_final - float(g["h_final"])) <= 1e-6
    # centring side effect on the caller's array (Q8)
    assert np.abs(X.mean(axis=0)).max() < 1e-12


def test_short_fit_c4_vs_reference(golden):
    from midagma_b200 import DagmaLinear, fit_batch
    g = golden("fit_c4_short")
    warm, mx, ck = (int(x) for x in g["fit_kw"])
    X = g["X"].copy()
    W = DagmaLinear("l2").fit(X.copy(), lambda1=float(g["lambda1"]), warm_iter=warm, max_iter=mx, checkpoint=ck,
                              s=[1.0, .9, .8, .7, .6])
    ref = g["W_est"]
    diff = simulate.edge_set_distance(W, ref)
    print("C4 short fit: edge diff", diff, "max|dW|", np.abs(W - ref).max())
    assert diff == 0
    # batched entry point gives the same answer for every replica and for a lambda grid
    Xb = np.stack([X, X, X])
    Wb, info = fit_batch(Xb, lambda1=np.array([0.02, 0.02, 0.05]), warm_iter=warm, max_iter=mx, checkpoint=ck,
                         return_info=True)
    assert np.array_equal(Wb[0], Wb[1]) and np.array_equal(Wb[0], W)
    assert (Wb[2] != 0).sum() <= (Wb[0] != 0).sum()
    assert info["status"].tolist() == [0, 0, 0]


def test_infeasible_start_returns_false():
    """minimize from a W outside the M-matrix domain: (W, False) at iteration 1 (linear.py:231-233)."""
    from midagma_b200 import minimize_batch
    d = 8
    X, _ = simulate.make_linear_problem(d, 2, 100, "ER", "gauss", 0)
    cov = (X.T @ X / 100)[None]
    W0 = np.zeros((1, d, d))
    W0[0, 0, 1], W0[0, 1, 0] = 1.5, 1.5
    W, ok, st = minimize_batch(W0.copy(), cov, 0.02, 1.0, 50, 1.0, 3e-4)
    assert not ok[0] and int(st[0, 0, 0]) == 0 and np.array_equal(W, W0)


def test_backtracking_matches_oracle():
    """Large lr at s = 1.0 drives W out of the domain -> lr halving path (linear.py:235-241)."""
    from midagma_b200 import minimize_batch
    d = 10
    X, _ = simulate.make_linear_problem(d, 2, 300, "ER", "gauss", 4)
    o = OracleLinear("l2").prepare(X, 0.0, checkpoint=50)
    W_ref, ok_ref = o.minimize(np.zeros((d, d)), 1.0, 40, 1.0, 1.0)
    n_halved = sum(1 for e in o.events if e[0] == "lr_halved")
    W, ok, st = minimize_batch(np.zeros((1, d, d)), o.cov[None], 0.0, 1.0, 40, 1.0, 1.0, checkpoint=50)
    print("halvings oracle", n_halved, "gpu", st[0, 0, 7], "lr", o.last_lr if ok_ref else None, st[0, 0, 1])
    assert n_halved > 0, "test input must exercise back-tracking"
    assert bool(ok[0]) == ok_ref
    assert int(st[0, 0, 7]) == n_halved and int(st[0, 0, 0]) == o.last_iters
    assert np.abs(W[0] - W_ref).max() <= 1e-6


def test_fit_retry_path():
    """Stage failure at s <= 0.9 -> retry with lr/2, s + 0.1 inside the kernel (linear.py:446-451)."""
    from midagma_b200 import DagmaLinear
    d = 10
    X, _ = simulate.make_linear_problem(d, 2, 300, "ER", "gauss", 4)
    o = OracleLinear("l2")
    kw = dict(lambda1=0.01, T=2, warm_iter=300, max_iter=300, lr=0.01, checkpoint=100)
    W_ref = o.fit(X.copy(), s=[1.0, 0.3], **kw)
    fails = [e for e in o.events if e[0] == "out_of_domain"]
    assert fails, "test input must exercise the retry path"
    s_list = [1.0, 0.3]
    m = DagmaLinear("l2")
    W = m.fit(X.copy(), s=s_list, **kw)
    print("retries oracle", len(fails), "gpu", m.stage_stats[:, 6], "s", s_list)
    assert int(m.stage_stats[:, 6].sum()) == len(fails)
    assert abs(s_list[1] - (0.3 + 0.1 * m.stage_stats[1, 6])) < 1e-12      # caller's list mutated (Q9)
    assert simulate.edge_set_distance(W, W_ref) == 0


def test_minimize_batch_host_pipeline(monkeypatch):
    """Pinned host buffers of a batch of several waves: the copies are hidden behind the kernel chunk by chunk
    (linear._minimize_batch_host_pipelined) -- bit for bit the result of copy / launch / copy."""
    import torch
    from midagma_b200 import minimize_batch
    from midagma_b200.linear import _resident_ctas, _host_pipeline_ok
    d, batch = 40, 4 * _resident_ctas() + 37
    rng = np.random.default_rng(0)
    Xs = rng.normal(size=(16, 120, d))
    cov16 = np.einsum("bni,bnj->bij", Xs, Xs) / 120
    cov = torch.from_numpy(np.ascontiguousarray(np.tile(cov16, (batch // 16 + 1, 1, 1))[:batch])).pin_memory()
    lam = np.linspace(0.01, 0.05, batch)
    outs = []
    for mode in ("1", "0"):
        monkeypatch.setenv("DAGMA_HOST_PIPELINE", mode)
        W = torch.zeros(batch, d, d, dtype=torch.float64).pin_memory()
        assert _host_pipeline_ok(W, cov) == (mode == "1")
        Wr, ok, st = minimize_batch(W, cov, lam, 1.0, 60, 1.0, 3e-4, checkpoint=20)
        assert Wr is W and ok.all() and st.shape == (batch, 1, 8) and (st[:, 0, 0] == 60).all()
        outs.append(W.clone())
    assert torch.equal(outs[0], outs[1]) and float(outs[0].abs().max()) > 0.0
