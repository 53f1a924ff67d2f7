"""GPU parity of the multi-CTA path (DMMA GEMM, blocked inverse, logistic score, fused update)
against numpy / the oracle / the reference fixtures."""
import numpy as np
import pytest
import torch

from oracle import simulate
from oracle.linear_ref import OracleLinear

pytestmark = pytest.mark.gpu


def _relmax(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("M,N,K,trans", [(128, 128, 64, False), (257, 130, 33, False), (100, 100, 10000, True),
                                          (10000, 100, 100, False), (2000, 2000, 64, False), (65, 7, 5, True),
                                          (1, 1, 1, False), (300, 300, 300, False)])
def test_gemm_vs_numpy(M, N, K, trans):
    from midagma_b200._large import gemm
    rng = np.random.default_rng(M * 7 + N)
    A = rng.normal(size=(K, M) if trans else (M, K))
    B = rng.normal(size=(K, N))
    C0 = rng.normal(size=(M, N))
    ref = 0.7 * ((A.T if trans else A) @ B) - 1.3 * C0
    C = torch.from_numpy(C0.copy()).cuda()
    ws = torch.empty(148 * M * N + 8, dtype=torch.float64, device="cuda") if M * N <= 20000 else None
    gemm(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), C, trans_a=trans, alpha=0.7, beta=-1.3, ws=ws)
    assert _relmax(C.cpu().numpy(), ref) <= 1e-13
    # sigmoid epilogue
    C2 = torch.empty(M, N, dtype=torch.float64, device="cuda")
    gemm(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), C2, trans_a=trans, alpha=0.01, epilogue=1, ws=ws)
    ref2 = 1 / (1 + np.exp(-0.01 * ((A.T if trans else A) @ B)))
    assert np.abs(C2.cpu().numpy() - ref2).max() <= 1e-14


@pytest.mark.parametrize("d", [129, 192, 200, 256, 257, 500, 513, 1000])
@pytest.mark.parametrize("square", [True, False])
def test_blocked_logdet_inv_vs_numpy(d, square):
    from midagma_b200.linear import logdet_inv
    rng = np.random.default_rng(d)
    A = rng.normal(size=(d, d)) * (rng.random((d, d)) < 0.05)
    if not square:
        A = np.abs(A)
    B = A * A if square else A
    rho = max(np.abs(np.linalg.eigvals(B)).max(), 1e-12)
    A *= np.sqrt(0.7 / rho) if square else 0.7 / rho
    s = 0.9
    out = logdet_inv(torch.from_numpy(A[None]).cuda(), s=s, square_input=square, want_inv=True, want_grad=True)
    M = s * np.eye(d) - (A * A if square else A)
    Minv = np.linalg.inv(M)
    lad = np.linalg.slogdet(M)[1]
    assert abs(out["logabsdet"][0].item() - lad) <= 1e-10 * max(1.0, abs(lad))
    assert _relmax(out["minv"][0].cpu().numpy(), Minv) <= 1e-10
    G = 2 * A * Minv.T if square else Minv.T
    assert _relmax(out["grad"][0].cpu().numpy(), G) <= 1e-10
    assert int(out["info"][0].item()) == 0
    # a matrix outside the M-matrix domain is flagged
    A2 = A.copy()
    A2[0, 1] = A2[1, 0] = 3.0
    out = logdet_inv(torch.from_numpy(A2[None]).cuda(), s=s, square_input=square)
    assert int(out["info"][0].item()) != 0


@pytest.mark.parametrize("d", [130, 258, 300, 520, 1000, 1234])
def test_inverse_with_rider_gemm(d):
    """dagma_logdet_inv_gemm_ws_f64: the inverse and an independent d x d GEMM in one call (for even d > 256 one
    dependency-driven kernel); both results against numpy."""
    from midagma_b200 import _lib
    lib = _lib.load()
    _lib.require_device()
    rng = np.random.default_rng(d)
    W = rng.normal(size=(d, d)) * (rng.random((d, d)) < 0.05)
    rho = max(np.abs(np.linalg.eigvals(W * W)).max(), 1e-12)
    W *= np.sqrt(0.6 / rho)
    cov = rng.normal(size=(d, d))
    s = 0.8
    f64 = dict(dtype=torch.float64, device="cuda")
    Wd, covd = torch.from_numpy(W).cuda(), torch.from_numpy(cov).cuda()
    minv, T = torch.empty(d, d, **f64), torch.full((d, d), float("nan"), **f64)
    sc = torch.zeros(4, **f64)
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    ws = torch.empty(lib.dagma_large_workspace_bytes(d) // 8 + 8, **f64)
    for _ in range(2):      # twice: the counters of the workspace are re-armed by every call
        _lib.check(lib.dagma_logdet_inv_gemm_ws_f64(
            _lib.stream_ptr(), d, s, Wd.data_ptr(), d, 1, sc.data_ptr(), sc.data_ptr() + 8, minv.data_ptr(), None, d,
            sc.data_ptr() + 16, info.data_ptr(), ws.data_ptr(), ws.numel() * 8, covd.data_ptr(), Wd.data_ptr(),
            T.data_ptr()), "dagma_logdet_inv_gemm_ws_f64")
    torch.cuda.synchronize()
    M = s * np.eye(d) - W * W
    assert int(info.item()) == 0
    assert _relmax(minv.cpu().numpy(), np.linalg.inv(M)) <= 1e-10
    lad = np.linalg.slogdet(M)[1]
    assert abs(sc[0].item() - lad) <= 1e-10 * max(1.0, abs(lad))
    assert abs(sc[1].item() - (-lad + d * np.log(s))) <= 1e-9 * max(1.0, abs(lad))
    assert _relmax(T.cpu().numpy(), cov @ W) <= 1e-13


def test_flow_kernel_mode():
    """The opt-in single-kernel ("flow") variant of the large-d inverse, with and without the rider GEMM
    (DAGMA_LOOKAHEAD=2 is read once per process, hence the subprocess)."""
    import os, subprocess, sys
    env = dict(os.environ, DAGMA_LOOKAHEAD="2")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.join(root, "tests", "test_large_gpu.py"),
                        "-k", "rider_gemm or (blocked_logdet and (500 or 1000))"], env=env, cwd=root,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("env", [{"DAGMA_FIRST_BLOCK": "1"}, {"DAGMA_TMA": "0"}, {"DAGMA_TMA": "3"}, {"DAGMA_CS_HEAD": "1"},
                                 {"DAGMA_CS_HEAD": "1", "DAGMA_FIRST_BLOCK": "1", "DAGMA_TMA": "0"}],
                         ids=["first-block-step", "cp-async-only", "tma-gemm-everywhere", "cs-strip-at-step-head",
                              "cs-head-first-block-cp-async"])
def test_optional_variants(env):
    """The A/B switches of the multi-CTA path stay correct: the first pivot block as a step of the persistent kernel,
    cp.async slabs everywhere, and the TMA GEMM for every product that qualifies (alpha / beta / sigmoid epilogues
    through the 128 x 128 kernel).  The switches are read once per process, hence the subprocess."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.join(root, "tests", "test_large_gpu.py"),
                        "-k", "gemm_vs_numpy or rider_gemm or (blocked_logdet and (500 or 513 or 1000))"],
                       env=dict(os.environ, **env), cwd=root, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def _edges(g, key):
    return tuple(tuple(int(x) for x in e) for e in g[key]) if key in g.files else None


@pytest.mark.parametrize("name,loss", [("linear_l2_d100", "l2"), ("linear_logistic_d12", "logistic")])
def test_minimize_stages_large_path(golden, name, loss):
    from midagma_b200 import DagmaLinear
    g = golden(name)
    X = g["X"].copy()
    model = DagmaLinear(loss)
    model.fit(X, lambda1=float(g["lambda1"]), T=1, warm_iter=0, max_iter=0, checkpoint=int(g["checkpoint"]))
    assert _relmax(model.cov, g["cov"]) <= 1e-13
    d = model.d
    W = np.zeros((d, d))
    for si, (mu, s, iters, lr) in enumerate(g["stages"]):
        W, ok = model.minimize(W.copy(), mu, int(iters), s, lr=lr)
        assert ok == bool(g[f"ok_{si}"]) and model.last_iters == int(g[f"iters_{si}"])
        ref = g[f"W_after_{si}"]
        err = np.abs(W - ref).max()
        print(name, "stage", si, "max|dW| =", err)
        assert err <= 1e-8, (si, err)
        st, it, obj, score, h, _ = model.checkpoint_log[-1]
        assert abs(obj - float(g[f"obj_{si}"])) <= 1e-9 * abs(float(g[f"obj_{si}"]))
    # value / gradient KATs of the reference at the last W (per-call parity, 1e-9 relative)
    sv, sg = model._score(W)
    assert abs(sv - float(g["score_val"])) <= 1e-9 * abs(float(g["score_val"]))
    assert _relmax(sg, g["score_grad"]) <= 1e-9
    hv, hg = model._h(W, 0.9)
    assert abs(hv - float(g["h_val"])) <= 1e-9 * max(abs(float(g["h_val"])), 1e-3)
    assert _relmax(hg, g["h_grad"]) <= 1e-9
    obj, score, h, trek = model._func(W, 0.1, 0.9)
    assert trek == 0.0 and abs(h - hv) <= 1e-12


def test_adam_update_api():
    from midagma_b200 import DagmaLinear
    rng = np.random.default_rng(0)
    o = OracleLinear("l2")
    o.opt_m, o.opt_v = 0, 0
    m = DagmaLinear("l2")
    m.opt_m, m.opt_v = 0, 0
    for it in range(1, 6):
        g = rng.normal(size=(9, 9))
        ref = o.adam_update(g, it, 0.99, 0.999)
        out = m._adam_update(g, it, 0.99, 0.999)
        assert _relmax(out, ref) <= 1e-14


def test_large_path_backtracking_and_failure():
    from midagma_b200 import DagmaLinear
    d = 70                                                     # > 64: multi-CTA path, on-chip inverse
    X, _ = simulate.make_linear_problem(d, 2, 300, "ER", "gauss", 4)
    o = OracleLinear("l2").prepare(X.copy(), 0.0, checkpoint=20)
    W_ref, ok_ref = o.minimize(np.zeros((d, d)), 1.0, 30, 1.0, 1.0)
    n_halved = sum(1 for e in o.events if e[0] == "lr_halved")
    assert n_halved > 0
    m = DagmaLinear("l2")
    m.fit(X.copy(), lambda1=0.0, T=1, warm_iter=0, max_iter=0, checkpoint=20)
    W, ok = m.minimize(np.zeros((d, d)), 1.0, 30, 1.0, lr=1.0)
    assert ok == ok_ref and m.last_iters == o.last_iters
    assert abs(m._large.last_lr - o.last_lr) == 0.0
    assert np.abs(W - W_ref).max() <= 1e-6
    # infeasible start
    W0 = np.zeros((d, d))
    W0[0, 1] = W0[1, 0] = 1.5
    W1, ok = m.minimize(W0.copy(), 1.0, 10, 1.0, lr=3e-4)
    assert not ok and np.array_equal(W1, W0)


def test_fit_d200_blocked_short_schedule():
    """Reduced-schedule full fit through the blocked inverse (d > 128): same graph as the oracle."""
    from midagma_b200 import DagmaLinear
    d = 200
    X, W_true = simulate.make_linear_problem(d, 2, 1000, "ER", "gauss", 3)
    kw = dict(lambda1=0.02, T=3, warm_iter=400, max_iter=600, checkpoint=200)
    W_ref = OracleLinear("l2").fit(X.copy(), s=[1.0, .9, .8], **kw)
    m = DagmaLinear("l2")
    W = m.fit(X.copy(), s=[1.0, .9, .8], **kw)
    print("d=200 edge diff", simulate.edge_set_distance(W, W_ref), "max|dW|", np.abs(W - W_ref).max(), m.stage_iters)
    assert simulate.edge_set_distance(W, W_ref) == 0
    assert np.abs(W - W_ref).max() <= 1e-6


def test_fit_batch_beyond_onchip_size():
    """fit_batch / the batched entry accept d > 64 (the reference's fit takes any d, linear.py:335): problems run one
    after the other on the multi-CTA engine and equal DagmaLinear.fit on the same data."""
    from midagma_b200 import DagmaLinear, fit_batch
    d = 72
    kw = dict(T=2, warm_iter=150, max_iter=200, checkpoint=50)
    Xs, Ws = [], []
    for p, lam in ((0, 0.02), (1, 0.04)):
        X, _ = simulate.make_linear_problem(d, 2, 400, "ER", "gauss", 30 + p)
        Xs.append(X)
        m = DagmaLinear("l2")
        m.fit(X.copy(), lambda1=lam, s=[1.0, .9], **kw)
        Ws.append(m.W_raw)
    exc = ((0, 1), (5, 9))
    W_est, info = fit_batch(np.stack(Xs), np.array([0.02, 0.04]), s=(1.0, .9), return_info=True, **kw)
    assert info["stage_iters"].tolist() == [[150, 200], [150, 200]]
    for b in range(2):
        assert np.abs(info["W_raw"][b] - Ws[b]).max() <= 1e-12
    _, info_x = fit_batch(np.stack(Xs), np.array([0.02, 0.04]), s=(1.0, .9), return_info=True, exclude_edges=exc, **kw)
    assert all(info_x["W_raw"][b][i, j] == 0.0 for b in range(2) for (i, j) in exc)
