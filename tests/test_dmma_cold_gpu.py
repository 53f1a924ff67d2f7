"""Cold branches of the tensor-core fit kernel (32 < d <= 64, csrc/small_fit_dmma.cu) against the oracle:
back-tracking (linear.py:235-241, Adam direction rebuilt from the moments in tensor memory), stage retry
(linear.py:446-451: read_W / start_attempt), infeasible start (:231-233) and the include / exclude bit masks
(:217-222, 248, 276).  The d <= 32 twins of these tests (tests/test_small_gpu.py) run on the scalar kernel."""
import numpy as np
import pytest

from oracle import simulate
from oracle.linear_ref import OracleLinear

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d", [48, 64])
@pytest.mark.parametrize("lr", [1.0, 0.5])
def test_dmma_backtracking_matches_oracle(d, lr):
    from midagma_b200 import minimize_batch
    X, _ = simulate.make_linear_problem(d, 2, 300, "ER", "gauss", 4)
    o = OracleLinear("l2").prepare(X, 0.0, checkpoint=50)
    W_ref, ok_ref = o.minimize(np.zeros((d, d)), 1.0, 40, 1.0, lr)
    n_halved = sum(1 for e in o.events if e[0] == "lr_halved")
    assert n_halved > 0, "test input must exercise back-tracking"
    W, ok, st = minimize_batch(np.zeros((1, d, d)), o.cov[None], 0.0, 1.0, 40, 1.0, lr, checkpoint=50)
    print("d", d, "halvings oracle", n_halved, "gpu", st[0, 0, 7], "lr", o.last_lr, st[0, 0, 1],
          "max|dW|", np.abs(W[0] - W_ref).max())
    assert bool(ok[0]) == ok_ref
    assert int(st[0, 0, 7]) == n_halved and int(st[0, 0, 0]) == o.last_iters
    assert st[0, 0, 1] == o.last_lr
    assert np.abs(W[0] - W_ref).max() <= 1e-6


@pytest.mark.parametrize("d", [48, 64])
@pytest.mark.parametrize("s2,lr", [(0.3, 0.01), (0.9, 0.1)])
def test_dmma_fit_retry_path(d, s2, lr):
    """(0.3, 0.01): the stage fails at iteration 1 twice (s 0.3 -> 0.4 -> 0.5).  (0.9, 0.1): back-tracking in the
    s = 1 stage, then a failure at iteration 2 of the s = 0.9 stage -> retry with lr / 2 and s = 1.0."""
    from midagma_b200 import DagmaLinear
    X, _ = simulate.make_linear_problem(d, 2, 300, "ER", "gauss", 4)
    kw = dict(lambda1=0.01, T=2, warm_iter=200, max_iter=200, lr=lr, checkpoint=100)
    o = OracleLinear("l2")
    W_ref = o.fit(X.copy(), s=[1.0, s2], **kw)
    fails = [e for e in o.events if e[0] == "out_of_domain"]
    halved = [e for e in o.events if e[0] == "lr_halved"]
    assert fails, "test input must exercise the retry path"
    s_list = [1.0, s2]
    m = DagmaLinear("l2")
    W = m.fit(X.copy(), s=s_list, **kw)
    print("d", d, "retries oracle", len(fails), "gpu", m.stage_stats[:, 6], "halvings", len(halved),
          m.stage_stats[:, 7], "s", s_list, "max|dW|", np.abs(m.W_raw - o.W_raw).max())
    assert int(m.stage_stats[:, 6].sum()) == len(fails)
    assert abs(s_list[1] - (s2 + 0.1 * len(fails))) < 1e-12                 # caller's list mutated (Q9)
    assert m.stage_iters == o.stage_iters
    assert simulate.edge_set_distance(W, W_ref) == 0
    assert np.abs(m.W_raw - o.W_raw).max() <= 1e-5


@pytest.mark.parametrize("d", [48, 64])
def test_dmma_infeasible_start(d):
    from midagma_b200 import minimize_batch
    X, _ = simulate.make_linear_problem(d, 2, 100, "ER", "gauss", 0)
    cov = (X.T @ X / 100)[None]
    W0 = np.zeros((1, d, d))
    W0[0, 0, d - 1], W0[0, d - 1, 0] = 1.5, 1.5
    W, ok, st = minimize_batch(W0.copy(), cov, 0.02, 1.0, 50, 1.0, 3e-4)
    assert not ok[0] and int(st[0, 0, 0]) == 0 and np.array_equal(W, W0)
    # mid-batch: the neighbours of the infeasible problem are unaffected
    Wb = np.concatenate([np.zeros((1, d, d)), W0, np.zeros((1, d, d))])
    W, ok, st = minimize_batch(Wb.copy(), np.repeat(cov, 3, axis=0), 0.02, 1.0, 50, 1.0, 3e-4)
    assert ok.tolist() == [True, False, True] and np.array_equal(W[0], W[2]) and np.array_equal(W[1], W0[0])


@pytest.mark.parametrize("d", [48, 64])
def test_dmma_masks_match_oracle(d):
    """include / exclude edges through fit (drop-in class) and through the batched entry points."""
    from midagma_b200 import DagmaLinear, fit_batch
    X, W_true = simulate.make_linear_problem(d, 2, 300, "ER", "gauss", 9)
    rng = np.random.default_rng(d)
    true_edges = np.argwhere(W_true != 0)
    exc = tuple((int(i), int(j)) for i, j in true_edges[rng.choice(len(true_edges), 5, replace=False)])
    exc = exc + ((0, d - 1), (d - 1, 0), (d // 2, d // 2 + 1))
    inc = tuple((int(i), int(j)) for i, j in true_edges[rng.choice(len(true_edges), 6, replace=False)]
                if (int(i), int(j)) not in exc) + ((1, d - 2),)
    kw = dict(lambda1=0.03, T=2, warm_iter=400, max_iter=300, checkpoint=100)
    o = OracleLinear("l2")
    W_ref = o.fit(X.copy(), s=[1.0, 0.9], exclude_edges=exc, include_edges=inc, **kw)
    m = DagmaLinear("l2")
    W = m.fit(X.copy(), s=[1.0, 0.9], exclude_edges=exc, include_edges=inc, **kw)
    err = np.abs(m.W_raw - o.W_raw).max()
    print("d", d, "masks: max|dW|", err, "stage iters", m.stage_iters, o.stage_iters)
    assert m.stage_iters == o.stage_iters
    assert err <= 1e-8
    for (i, j) in exc:
        assert m.W_raw[i, j] == 0.0
    assert simulate.edge_set_distance(W, W_ref) == 0
    # the same masks through fit_batch (shared by the batch); raw W: the short schedule keeps |W| under the threshold
    Wb, info = fit_batch(np.stack([X, X]), s=(1.0, 0.9), exclude_edges=exc, include_edges=inc, return_info=True, **kw)
    assert np.array_equal(info["W_raw"][0], m.W_raw) and np.array_equal(info["W_raw"][1], m.W_raw)
    assert np.array_equal(Wb[0], W)
    # without masks the answer differs (the masks are not silently ignored)
    _, info0 = fit_batch(X[None], s=(1.0, 0.9), return_info=True, **kw)
    assert not np.array_equal(info0["W_raw"][0], m.W_raw)
    assert any(info0["W_raw"][0][i, j] != 0.0 for (i, j) in exc)


def test_dmma_two_ctas_per_sm_are_resident():
    """The launch geometry of the 32 < d <= 64 kernel assumes two CTAs per SM (registers, shared memory and tensor
    memory are sized for exactly that) although the occupancy API answers 1: the kernel counts for itself."""
    import ctypes as C
    import torch
    from midagma_b200 import _lib, minimize_batch
    sms = _lib.require_device()
    d, nprob = 64, 2 * sms
    rng = np.random.default_rng(0)
    X = rng.normal(size=(nprob, 100, d))
    cov = np.einsum("bni,bnj->bij", X, X) / 100
    minimize_batch(np.zeros((nprob, d, d)), cov, 0.02, 1.0, 300, 1.0, 3e-4, tol=0.0)
    got = C.c_int(-1)
    _lib.check(_lib.load().dagma_fit_dmma_residency(C.byref(got)), "dagma_fit_dmma_residency")
    print("CTAs of the fit kernel per SM:", got.value)
    assert got.value == 2
