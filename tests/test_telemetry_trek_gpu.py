"""SURVEY.md 8f2 / 8f3 on the GPU: the 25-key ``minimize.checkpoint`` telemetry rows and the PST ``seq="inv"`` trek
regulariser (modes "opt" and "log"), against rows / trajectories recorded from the unmodified reference
(oracle/make_golden_trek.py -> tests/golden/linear_trek_events.npz)."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _reg(name, pairs):
    from midagma_b200.notreks import PSTRegularizer
    return {
        "plain": None,
        "pst_opt": PSTRegularizer(I=pairs, seq="inv", weight=0.7, mode="opt"),
        "pst_log": PSTRegularizer(I=pairs, seq="inv", weight=0.7, mode="log"),
        "pst_opt_sum": PSTRegularizer(I=pairs, seq="inv", weight=0.05, mode="opt", kwargs={"agg": "sum"}),
    }[name]


@pytest.mark.parametrize("agg", ["mean", "sum"])
def test_pst_inv_value_grad(golden, agg):
    from midagma_b200.notreks import PSTRegularizer, trek_value_grad
    g = golden("linear_trek_events")
    reg = PSTRegularizer(I=g["pairs"], seq="inv", weight=1.0, mode="opt", kwargs={"agg": agg})
    val, grad = trek_value_grad(g["pst_W"], reg)
    ref_v, ref_g = float(g[f"pst_val_{agg}"]), g[f"pst_grad_{agg}"]
    assert abs(val - ref_v) <= 1e-11 * max(1.0, abs(ref_v))
    assert np.abs(grad - ref_g).max() <= 1e-10 * max(1.0, np.abs(ref_g).max())
    # mode "log": value only
    v2, g2 = trek_value_grad(g["pst_W"], PSTRegularizer(I=g["pairs"], seq="inv", weight=1.0, mode="log", kwargs={"agg": agg}))
    assert abs(v2 - ref_v) <= 1e-11 * max(1.0, abs(ref_v)) and not g2.any()
    # disabled / empty pair list: the reference's no-op branch
    v3, g3 = trek_value_grad(g["pst_W"], PSTRegularizer(I=np.zeros((0, 2), dtype=np.int64), seq="inv", weight=1.0))
    assert v3 == 0.0 and not g3.any()
    from midagma_b200.notreks import TCCRegularizer
    v4, g4 = trek_value_grad(g["pst_W"], TCCRegularizer(I=g["pairs"], weight=1.0, mode="log"))   # spectral TCC, value only
    assert np.isfinite(v4) and not g4.any()


@pytest.mark.parametrize("case", ["plain", "pst_opt", "pst_log", "pst_opt_sum"])
def test_checkpoint_events_and_trajectory(golden, case):
    from midagma_b200 import DagmaLinear
    from midagma_b200.logger import LogConfig
    g = golden("linear_trek_events")
    meta = json.loads(str(g["meta_json"]))[case]
    keys = [str(k) for k in g["numeric_keys"]]
    rows = []
    cfg = LogConfig(enabled=True, store_jsonl=False, store_csv=False, keep_in_memory=True, callback=rows.append)
    m = DagmaLinear("l2", trek_reg=_reg(case, g["pairs"]), log_cfg=cfg)
    X = g[f"{case}_X"].copy()
    d = X.shape[1]
    m.fit(X, lambda1=0.02, T=1, warm_iter=0, max_iter=0, checkpoint=100,
          include_edges=((0, 3), (2, 5)) if case == "plain" else None)
    W = np.zeros((d, d))
    for si, (mu, iters, s, lr) in enumerate(g["stages"]):
        W, ok = m.minimize(W, float(mu), int(iters), float(s), float(lr))
        assert ok == bool(g[f"{case}_ok"][si])
        assert np.abs(W - g[f"{case}_W"][si]).max() <= 1e-9
    ref = g[f"{case}_events"]
    assert len(rows) == meta["n_events"] == ref.shape[0]
    assert sorted(rows[0].keys()) == meta["keys"]                     # the same 25 (+ "event") columns
    assert rows[0]["event"] == meta["event"] and rows[0]["reg_dag_name"] == meta["reg_dag_name"]
    assert rows[0]["reg_trek_name"] == meta["reg_trek_name"] and rows[0]["trek_mode"] == meta["trek_mode"]
    assert rows[0]["reg_dag_cfg"] == meta["reg_dag_cfg"]
    assert sorted(rows[0]["reg_trek_cfg"].keys()) == meta["reg_trek_cfg_keys"]
    got = np.array([[float(r[k]) for k in keys] for r in rows])
    err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-6)
    worst = {k: float(err[:, i].max()) for i, k in enumerate(keys)}
    assert err.max() <= 1e-7, worst
    # the same rows through the logger's own loader
    cols = m._slog.load(event="minimize.checkpoint")
    assert list(cols["iter"]) == [int(x) for x in ref[:, 0]]


def test_telemetry_stays_on_the_onchip_kernel():
    """With telemetry on, a d <= 64 fit still runs as ONE launch of the on-chip kernel (the telemetry rows are produced
    inside it at the checkpoint iterations): bit-identical W with and without logging, rows at every checkpoint."""
    from midagma_b200 import DagmaLinear
    from midagma_b200.logger import LogConfig
    from oracle import simulate
    X, _ = simulate.config_c1(3)
    ma = DagmaLinear("l2")
    a = ma.fit(X.copy(), lambda1=0.02, T=2, warm_iter=2000, max_iter=3000, s=[1.0, 0.9])
    rows = []
    cfg = LogConfig(enabled=True, store_jsonl=False, callback=rows.append)
    mb = DagmaLinear("l2", log_cfg=cfg)
    b = mb.fit(X.copy(), lambda1=0.02, T=2, warm_iter=2000, max_iter=3000, s=[1.0, 0.9])
    assert mb._large is None                                   # the multi-CTA engine was never built
    assert np.array_equal(ma.W_raw, mb.W_raw) and np.array_equal(a, b)
    assert [r["iter"] for r in rows] == [int(x[1]) for x in mb.checkpoint_log] and len(rows) >= 4
    assert rows[0]["iter"] == 1000 and rows[0]["mu"] == 1.0 and rows[-1]["mu"] == 0.1
    assert rows[0]["reg_dag_cfg"] == {"s": 1.0} and rows[-1]["reg_dag_cfg"] == {"s": 0.9}
    assert all(r["elapsed_sec"] > 0 for r in rows) and rows[1]["elapsed_sec"] > rows[0]["elapsed_sec"]


@pytest.mark.parametrize("d", [12, 24, 40, 64])
def test_onchip_telemetry_rows_vs_oracle(d):
    """The kernel's telemetry rows (scalar kernel d <= 32, tensor-core kernel above) against the same quantities
    computed from the oracle's iteration trace (src/dagma/linear.py:262-273, 307-324), with include / exclude edges."""
    from midagma_b200 import DagmaLinear
    from midagma_b200.logger import LogConfig
    from oracle import simulate
    from oracle.linear_ref import OracleLinear
    import scipy.linalg as sla
    X, W_true = simulate.make_linear_problem(d, 2, 300, "ER", "gauss", 40 + d)
    edges = np.argwhere(W_true != 0)
    inc = ((int(edges[0][0]), int(edges[0][1])), (int(edges[1][0]), int(edges[1][1])))
    exc = ((int(edges[2][0]), int(edges[2][1])), (0, d - 1))
    mu, s, lr, ck, iters, lam = 0.5, 0.9, 3e-4, 60, 150, 0.03
    o = OracleLinear("l2").prepare(X.copy(), lam, checkpoint=ck, exclude_edges=exc, include_edges=inc)
    o.trace = []
    W_ref, _ = o.minimize(np.zeros((d, d)), mu, iters, s, lr)
    rows = []
    m = DagmaLinear("l2", log_cfg=LogConfig(enabled=True, store_jsonl=False, callback=rows.append))
    m.fit(X.copy(), lambda1=lam, T=1, warm_iter=0, max_iter=0, checkpoint=ck, exclude_edges=exc, include_edges=inc)
    W, ok = m.minimize(np.zeros((d, d)), mu, iters, s, lr)
    assert ok and np.abs(W - W_ref).max() <= 1e-9
    assert [r["iter"] for r in rows] == [60, 120, 150]
    mask_inc = np.zeros((d, d))
    mask_inc[tuple(zip(*inc))] = -2 * mu * lam
    mask_exc = np.ones((d, d))
    mask_exc[tuple(zip(*exc))] = 0.0
    for r in rows:
        t = o.trace[r["iter"] - 1]
        Wb, Gobj, dirn = t["W_before"], t["Gobj"], t["dir"]
        M = sla.inv(s * np.eye(d) - Wb * Wb) + 1e-16
        Wn = (Wb - t["lr"] * dirn) * mask_exc
        nz = np.abs(Wn[Wn != 0])
        want = {"grad_raw_norm": np.linalg.norm(Gobj), "grad_score_norm": np.linalg.norm(-mu * o.cov @ (np.eye(d) - Wb)),
                "grad_dag_norm": np.linalg.norm(2 * Wb * M.T), "grad_l1_norm": np.linalg.norm(mu * lam * np.sign(Wb)),
                "grad_inc_norm": np.linalg.norm(mask_inc * np.sign(Wb)), "grad_step_norm": np.linalg.norm(dirn),
                "step_norm": t["lr"] * np.linalg.norm(dirn), "w_norm": np.linalg.norm(Wn), "w_abs_sum": np.abs(Wn).sum(),
                "max_abs_w": np.abs(Wn).max(), "min_abs_w_nonzero": nz.min() if nz.size else 0.0, "lr": t["lr"], "mu": mu}
        for k, v in want.items():
            assert abs(r[k] - v) <= 1e-9 * max(abs(v), 1e-12), (d, r["iter"], k, r[k], v)


@pytest.mark.parametrize("seq", ["inv", "log", "exp", "binom"])
@pytest.mark.parametrize("wn", ["a", "b"])
def test_pst_series_value_grad(golden, seq, wn):
    """Every PST series x aggregation on the GEMM kernels against the reference's autograd
    (oracle/make_golden_trek.py --pst-series -> tests/golden/pst_series.npz)."""
    import torch
    from midagma_b200.notreks import PSTRegularizer, trek_value_grad, pst_mat, pst, get_no_trek_pairs
    g = golden("pst_series")
    W, pairs = g[f"W_{wn}"], g["pairs"]
    for agg in ("mean", "sum", "max", "lse"):
        reg = PSTRegularizer(I=pairs, seq=seq, weight=1.0, mode="opt", kwargs={"agg": agg})
        val, grad = trek_value_grad(W, reg)
        ref_v, ref_g = float(g[f"val_{wn}_{seq}_{agg}"]), g[f"grad_{wn}_{seq}_{agg}"]
        assert abs(val - ref_v) <= 1e-9 * max(abs(ref_v), 1e-3), (seq, agg, val, ref_v)
        assert np.abs(grad - ref_g).max() <= 1e-9 * max(np.abs(ref_g).max(), 1e-6), (seq, agg)
        v2, g2 = trek_value_grad(W, PSTRegularizer(I=pairs, seq=seq, weight=1.0, mode="log", kwargs={"agg": agg}))
        assert abs(v2 - ref_v) <= 1e-9 * max(abs(ref_v), 1e-3) and not g2.any()
    H_ref = g[f"H_{wn}_{seq}"]
    H = pst_mat(torch.from_numpy(W), seq).numpy()
    assert np.abs(H - H_ref).max() <= 1e-10 * np.abs(H_ref).max()
    vec = pst(torch.from_numpy(W), pairs, seq, agg="none").numpy()
    assert np.abs(vec - g[f"none_{wn}_{seq}"]).max() <= 1e-10 * max(np.abs(H_ref).max(), 1e-6)
    if seq == "log" and wn == "a":
        reg = PSTRegularizer(I=pairs, seq="log", weight=1.0, mode="opt", kwargs={"agg": "mean", "K_log": 5})
        val, grad = trek_value_grad(W, reg)
        assert abs(val - float(g["val_a_log_K5"])) <= 1e-11 and np.abs(grad - g["grad_a_log_K5"]).max() <= 1e-11
    if seq == "exp":                                                   # pairs with no trek: H == 0 exactly
        Wd = np.triu(W, 1) * (np.abs(np.triu(W, 1)) > 0.25)
        ref_pairs = {(i, j) for i in range(W.shape[0]) for j in range(i + 1, W.shape[0])}
        got = {tuple(p) for p in get_no_trek_pairs(torch.from_numpy(Wd), "exp")}
        assert got <= ref_pairs


@pytest.mark.parametrize("case", ["exp_mean", "log_lse", "binom_max"])
def test_pst_series_trajectory(golden, case):
    """``minimize`` in mode "opt" with the exp / log / binom series: W after each stage and the logged trek value."""
    from midagma_b200 import DagmaLinear
    from midagma_b200.logger import LogConfig
    from midagma_b200.notreks import PSTRegularizer
    g = golden("pst_series")
    pairs = g["pairs"]
    reg = {"exp_mean": PSTRegularizer(I=pairs, seq="exp", weight=0.7, mode="opt"),
           "log_lse": PSTRegularizer(I=pairs, seq="log", weight=0.3, mode="opt", kwargs={"agg": "lse"}),
           "binom_max": PSTRegularizer(I=pairs, seq="binom", weight=0.5, mode="opt", kwargs={"agg": "max"})}[case]
    rows = []
    cfg = LogConfig(enabled=True, store_jsonl=False, store_csv=False, keep_in_memory=True, callback=rows.append)
    m = DagmaLinear("l2", trek_reg=reg, log_cfg=cfg)
    X = g[f"{case}_X"].copy()
    d = X.shape[1]
    m.fit(X, lambda1=0.02, T=1, warm_iter=0, max_iter=0, checkpoint=100)
    W = np.zeros((d, d))
    for si, (mu, iters, s, lr) in enumerate(g["stages"]):
        W, ok = m.minimize(W, float(mu), int(iters), float(s), float(lr))
        assert ok == bool(g[f"{case}_ok"][si])
        assert np.abs(W - g[f"{case}_W"][si]).max() <= 1e-9, case
    ref = g[f"{case}_trek_vals"]
    got = np.array([float(r["reg_trek_value"]) for r in rows])
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-8 * max(np.abs(ref).max(), 1e-3)
