"""The persistent one-kernel inner iteration of DagmaLinear for d <= 128 (csrc/lin_iter.cu) against the launch
sequence it replaces (DAGMA_LIN_FUSED=0: fused inverse -> score GEMMs -> update kernel, graph-replayed) and against
the numpy oracle of src/dagma/linear.py:165-333."""
import numpy as np
import pytest

from oracle import simulate
from oracle.linear_ref import OracleLinear

pytestmark = pytest.mark.gpu

STAGES = ((1.0, 1.0, 130), (0.1, 0.9, 90))        # (mu, s, iterations): two chained minimize calls, checkpoints inside


def _problem(loss, d, n, seed):
    sem = "logistic" if loss == "logistic" else "gauss"
    X, _ = simulate.make_linear_problem(d, 2, n, "ER", sem, seed)
    return X


def _run(monkeypatch, fused, loss, X, lam, masks, lr=3e-4, stages=STAGES, checkpoint=50):
    from midagma_b200 import DagmaLinear
    monkeypatch.setenv("DAGMA_LIN_FUSED", "1" if fused else "0")
    m = DagmaLinear(loss)
    m.fit(X.copy(), lambda1=lam, T=1, warm_iter=0, max_iter=0, checkpoint=checkpoint, **masks)
    assert m._large_engine().one_kernel == fused
    d = X.shape[1]
    W = np.zeros((d, d))
    out = []
    for mu, s, iters in stages:
        W, ok = m.minimize(W.copy(), mu, iters, s, lr=lr)
        out.append((W.copy(), ok, m.last_iters, [row[2] for row in m.checkpoint_log]))
    return out, m


@pytest.mark.parametrize("loss,d,n", [("l2", 65, 200), ("l2", 72, 300), ("l2", 100, 400), ("l2", 127, 300), ("l2", 128, 500),
                                      ("logistic", 5, 40), ("logistic", 12, 300), ("logistic", 33, 1000),
                                      ("logistic", 64, 2000), ("logistic", 70, 900), ("logistic", 88, 1500), ("logistic", 100, 3000),
                                      ("logistic", 101, 777), ("logistic", 120, 2500),
                                      ("logistic", 128, 10584)])
def test_fused_iteration_matches_launch_sequence(monkeypatch, loss, d, n):
    X = _problem(loss, d, n, 100 + d)
    a, _ = _run(monkeypatch, True, loss, X, 0.02, {})
    b, _ = _run(monkeypatch, False, loss, X, 0.02, {})
    for (Wa, oka, ita, obja), (Wb, okb, itb, objb) in zip(a, b):
        assert oka == okb and ita == itb
        err = np.abs(Wa - Wb).max()
        print(loss, d, n, "max|dW| fused vs sequence =", err)
        assert err <= 1e-11
        assert np.allclose(obja, objb, rtol=1e-11, atol=0)


@pytest.mark.parametrize("loss,d,n", [("l2", 80, 300), ("logistic", 40, 600)])
def test_fused_iteration_with_masks(monkeypatch, loss, d, n):
    X = _problem(loss, d, n, 7)
    masks = dict(exclude_edges=((0, 1), (5, 2), (d - 1, 0)), include_edges=((3, 4), (1, d - 2)))
    a, _ = _run(monkeypatch, True, loss, X, 0.05, masks)
    b, _ = _run(monkeypatch, False, loss, X, 0.05, masks)
    for (Wa, *_), (Wb, *_) in zip(a, b):
        assert np.abs(Wa - Wb).max() <= 1e-11
        assert Wa[0, 1] == 0.0 and Wa[5, 2] == 0.0 and Wa[d - 1, 0] == 0.0


def test_fused_iteration_vs_oracle_l2(monkeypatch):
    d = 96
    X = _problem("l2", d, 500, 3)
    out, m = _run(monkeypatch, True, "l2", X, 0.03, {})
    o = OracleLinear("l2").prepare(X.copy(), 0.03, checkpoint=50)
    W = np.zeros((d, d))
    for (mu, s, iters), (Wg, okg, itg, _) in zip(STAGES, out):
        W, ok = o.minimize(W.copy(), mu, iters, s, 3e-4)
        assert ok == okg and o.last_iters == itg
        err = np.abs(Wg - W).max()
        print("l2 d=96 vs oracle max|dW| =", err)
        assert err <= 1e-9


def test_fused_iteration_vs_oracle_logistic(monkeypatch):
    d = 30
    X = _problem("logistic", d, 800, 5)
    W0 = np.random.default_rng(1).normal(size=(d, d)) * 0.05       # a generic start (no exact-tie entries)
    np.fill_diagonal(W0, 0.0)
    from midagma_b200 import DagmaLinear
    monkeypatch.setenv("DAGMA_LIN_FUSED", "1")
    m = DagmaLinear("logistic")
    m.fit(X.copy(), lambda1=0.02, T=1, warm_iter=0, max_iter=0, checkpoint=40)
    assert m._large_engine().one_kernel
    Wg, okg = m.minimize(W0.copy(), 1.0, 100, 1.0, lr=3e-4)
    o = OracleLinear("logistic").prepare(X.copy(), 0.02, checkpoint=40)
    Wr, okr = o.minimize(W0.copy(), 1.0, 100, 1.0, 3e-4)
    assert okg == okr and m.last_iters == o.last_iters
    err = np.abs(Wg - Wr).max()
    print("logistic d=30 vs oracle max|dW| =", err)
    assert err <= 1e-9


@pytest.mark.parametrize("iters,checkpoint", [(30, 20), (200, 150)])
def test_fused_iteration_backtracking_and_failure(monkeypatch, iters, checkpoint):
    """lr = 1: the inverse turns infeasible inside a launch, the kernel latches `halted`, the host halves lr exactly as
    linear.py:230-241 and relaunches; infeasible start: untouched W, ok = False.  (200, 150): the checkpoint interval is
    longer than the launch sequence's probe interval of the latch, LargeLinearEngine._replay_probed.)"""
    d = 70
    X, _ = simulate.make_linear_problem(d, 2, 300, "ER", "gauss", 4)
    a, ma = _run(monkeypatch, True, "l2", X, 0.0, {}, lr=1.0, stages=((1.0, 1.0, iters),), checkpoint=checkpoint)
    b, mb = _run(monkeypatch, False, "l2", X, 0.0, {}, lr=1.0, stages=((1.0, 1.0, iters),), checkpoint=checkpoint)
    assert a[0][1] == b[0][1] and a[0][2] == b[0][2]
    assert ma._large.last_lr == mb._large.last_lr and ma._large.last_lr < 1.0
    assert np.abs(a[0][0] - b[0][0]).max() <= 1e-9
    W0 = np.zeros((d, d))
    W0[0, 1] = W0[1, 0] = 1.5
    W1, ok = ma.minimize(W0.copy(), 1.0, 10, 1.0, lr=3e-4)
    assert not ok and np.array_equal(W1, W0)


def test_fused_iteration_limits():
    """Shapes outside the kernel's reach are reported as such (the engine then keeps the launch sequence)."""
    import torch
    from midagma_b200 import _lib
    lib = _lib.load()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    assert lib.dagma_linear_iter_supported(0, 0, 128) == 1 and lib.dagma_linear_iter_supported(0, 0, 129) == 0
    assert lib.dagma_linear_iter_supported(1, 72 * (sms - 1), 101) == 1          # resident rows: any d
    assert lib.dagma_linear_iter_supported(1, 72 * (sms - 1) + 1, 100) == 0      # (streamed rows: opt-in, see below)
    assert lib.dagma_linear_iter_workspace_doubles(1, 1000, 10) >= 100


def test_mid_d_batch_runs_problems_side_by_side(monkeypatch):
    """fit_batch at 64 < d <= 128: several problems at once (one persistent kernel, one stream and one host thread
    each) give bit for bit what the same problems give one after the other."""
    import time
    from midagma_b200 import fit_batch
    from midagma_b200.linear import _batch_lanes
    d, batch = 80, 10
    Xs = np.stack([simulate.make_linear_problem(d, 2, 300, "ER", "gauss", 60 + p)[0] for p in range(batch)])
    lam = np.linspace(0.01, 0.05, batch)
    kw = dict(T=2, warm_iter=200, max_iter=300, checkpoint=100, s=(1.0, .9), return_info=True)
    monkeypatch.delenv("DAGMA_BATCH_LANES", raising=False)
    assert _batch_lanes(d, batch) > 1
    t0 = time.perf_counter()
    _, par = fit_batch(Xs, lam, **kw)
    t1 = time.perf_counter()
    monkeypatch.setenv("DAGMA_BATCH_LANES", "1")
    _, seq = fit_batch(Xs, lam, **kw)
    t2 = time.perf_counter()
    print(f"mid-d batch: {_batch_lanes(d, batch)} lane(s) {t2 - t1:.2f} s, side by side {t1 - t0:.2f} s")
    assert np.array_equal(par["W_raw"], seq["W_raw"])
    assert par["stage_iters"].tolist() == seq["stage_iters"].tolist()
    assert np.array_equal(par["h_final"], seq["h_final"])


def test_streamed_rows_variant():
    """The opt-in variant that streams the rows of X through shared memory when they do not fit (DAGMA_LIN_STREAM=1,
    measured slower than the launch sequence and therefore off) stays correct.  The switch is read once per process,
    hence the subprocess."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import numpy as np\n"
        "from oracle import simulate\n"
        "from midagma_b200 import DagmaLinear\n"
        "import os\n"
        "for d, n in ((100, 20000), (128, 15001), (64, 30011), (10, 12000)):\n"
        "    X = simulate.make_linear_problem(d, 2, n, 'ER', 'logistic', 100 + d)[0]\n"
        "    out = []\n"
        "    for fused in ('1', '0'):\n"
        "        os.environ['DAGMA_LIN_FUSED'] = fused\n"
        "        m = DagmaLinear('logistic')\n"
        "        m.fit(X.copy(), lambda1=0.02, T=1, warm_iter=0, max_iter=0, checkpoint=50)\n"
        "        assert m._large_engine().one_kernel == (fused == '1')\n"
        "        W, ok = m.minimize(np.zeros((d, d)), 1.0, 130, 1.0, lr=3e-4)\n"
        "        out.append(W)\n"
        "    err = np.abs(out[0] - out[1]).max()\n"
        "    print(d, n, err)\n"
        "    assert err <= 1e-11\n")
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DAGMA_LIN_STREAM="1", PYTHONPATH=root), cwd=root,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_minimize_batch_beyond_onchip_size():
    """minimize_batch accepts d > 64: each problem equals DagmaLinear.minimize on the same covariance."""
    from midagma_b200 import DagmaLinear, minimize_batch
    from midagma_b200.linear import center_cov
    import torch
    d, batch = 72, 5
    Xs = np.stack([simulate.make_linear_problem(d, 2, 300, "ER", "gauss", 80 + p)[0] for p in range(batch)])
    lam = np.linspace(0.01, 0.03, batch)
    cov = center_cov(torch.from_numpy(Xs).cuda(), center=True).cpu().numpy()
    W0 = np.zeros((batch, d, d))
    W, ok, stats = minimize_batch(W0, cov, lam, 1.0, 120, 1.0, 3e-4, checkpoint=50)
    assert W is W0 and ok.all() and stats[:, 0, 0].tolist() == [120.0] * batch
    for b in range(batch):
        m = DagmaLinear("l2")
        m.fit(Xs[b].copy(), lambda1=float(lam[b]), T=1, warm_iter=0, max_iter=0, checkpoint=50)
        Wb, okb = m.minimize(np.zeros((d, d)), 1.0, 120, 1.0, lr=3e-4)
        assert okb and np.abs(Wb - W[b]).max() <= 1e-12


@pytest.mark.parametrize("d", [1, 8, 20, 48, 64])
def test_fused_iteration_l2_small_d(monkeypatch, d):
    """The l2 branch of the persistent kernel at d <= 64 (tensor-core sweep in the inverse CTA) -- reachable through the
    C ABI, while DagmaLinear itself uses the on-chip fit kernel there: forced here, against the launch sequence and
    the oracle."""
    from midagma_b200 import DagmaLinear
    monkeypatch.setattr(DagmaLinear, "_small_ok", lambda self: False)
    X = _problem("l2", d, 150, 40 + d) if d > 1 else np.random.default_rng(0).normal(size=(150, 1))
    a, _ = _run(monkeypatch, True, "l2", X, 0.02, {})
    b, _ = _run(monkeypatch, False, "l2", X, 0.02, {})
    o = OracleLinear("l2").prepare(X.copy(), 0.02, checkpoint=50)
    W = np.zeros((d, d))
    for (mu, s, iters), (Wa, oka, ita, _), (Wb, okb, itb, _) in zip(STAGES, a, b):
        W, ok = o.minimize(W.copy(), mu, iters, s, 3e-4)
        assert oka == okb == ok and ita == itb == o.last_iters
        assert np.abs(Wa - Wb).max() <= 1e-11 and np.abs(Wa - W).max() <= 1e-9
