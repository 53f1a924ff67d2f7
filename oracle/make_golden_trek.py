"""Generate tests/golden/linear_trek_events.npz by running the UNMODIFIED reference (build container only):

    python oracle/make_golden_trek.py

Covers SURVEY.md 8f2 / 8f3: the 25-key ``minimize.checkpoint`` telemetry events (src/dagma/linear.py:279-326,
src/logger.py) and the PST ``seq="inv"`` trek regulariser in mode="opt" / "log" (src/notreks/notreks.py:454-619,
667-736).  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import json
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

import numpy as np  # noqa: E402

from oracle import simulate  # noqa: E402
from oracle.make_golden import _prep, _Bar  # noqa: E402

from dagma.linear import DagmaLinear  # noqa: E402  (reference)
from logger import LogConfig  # noqa: E402  (reference)
import notreks.notreks as ref_nt  # noqa: E402  (reference)

GOLD = os.path.join(ROOT, "tests", "golden")
NUMERIC_KEYS = ["iter", "stage", "obj_total", "score_datafit", "reg_dag_value", "reg_trek_value", "trek_weight", "mu",
                "lr", "w_norm", "w_abs_sum", "max_abs_w", "min_abs_w_nonzero", "grad_raw_norm", "grad_step_norm",
                "step_norm", "grad_score_norm", "grad_dag_norm", "grad_l1_norm", "grad_inc_norm", "grad_trek_norm"]


def run(d, n, seed, reg, stages, include=None):
    rng = np.random.default_rng(seed)
    B = simulate.simulate_dag(d, 2 * d, "ER", rng)
    W_true = simulate.simulate_parameter(B, rng=rng)
    X = simulate.simulate_linear_sem(W_true, n, "gauss", rng=rng)
    X = X - X.mean(axis=0, keepdims=True)
    cfg = LogConfig(enabled=True, store_jsonl=False, store_csv=False, keep_in_memory=True)
    m = DagmaLinear("l2", trek_reg=reg, log_cfg=cfg)
    _prep(m, X.copy(), 0.02, 100, include=include)
    W = np.zeros((d, d))
    Ws, oks = [], []
    for mu, iters, s, lr in stages:
        W, ok = m.minimize(W, mu, iters, s, lr, pbar=_Bar())
        Ws.append(W.copy())
        oks.append(bool(ok))
    rows = list(m._slog._rows)
    return X, np.stack(Ws), oks, rows


def main():
    d, n = 12, 300
    rng = np.random.default_rng(7)
    pairs = np.array([(i, j) for i in range(d) for j in range(i + 1, d) if rng.random() < 0.25], dtype=np.int64)
    stages = [(1.0, 300, 1.0, 3e-4), (0.1, 300, 0.9, 3e-4)]
    out = {}
    cases = {
        "plain": None,
        "pst_opt": ref_nt.PSTRegularizer(I=pairs, seq="inv", weight=0.7, mode="opt"),
        "pst_log": ref_nt.PSTRegularizer(I=pairs, seq="inv", weight=0.7, mode="log"),
        "pst_opt_sum": ref_nt.PSTRegularizer(I=pairs, seq="inv", weight=0.05, mode="opt", kwargs={"agg": "sum"}),
    }
    meta = {}
    for name, reg in cases.items():
        X, Ws, oks, rows = run(d, n, 11, reg, stages, include=((0, 3), (2, 5)) if name == "plain" else None)
        out[f"{name}_X"] = X
        out[f"{name}_W"] = Ws
        out[f"{name}_ok"] = np.array(oks)
        out[f"{name}_events"] = np.array([[float(r[k]) for k in NUMERIC_KEYS] for r in rows])
        meta[name] = {"keys": sorted(rows[0].keys()), "reg_trek_name": rows[0]["reg_trek_name"],
                      "trek_mode": rows[0]["trek_mode"], "reg_dag_name": rows[0]["reg_dag_name"],
                      "reg_dag_cfg": rows[0]["reg_dag_cfg"], "reg_trek_cfg_keys": sorted(rows[0]["reg_trek_cfg"].keys()),
                      "event": rows[0]["event"], "n_events": len(rows)}
        print(name, "events", len(rows), "ok", oks, "max|W|", float(np.abs(Ws[-1]).max()))
    # value / gradient of the PST-inv penalty at a fixed W (both aggregations)
    Wp = np.random.default_rng(3).uniform(-0.4, 0.4, size=(d, d)) * (np.random.default_rng(4).random((d, d)) < 0.3)
    out["pst_W"] = Wp
    for agg in ("mean", "sum"):
        reg = ref_nt.PSTRegularizer(I=pairs, seq="inv", weight=1.0, mode="opt", kwargs={"agg": agg})
        v, g = ref_nt.trek_value_grad(Wp, reg)
        out[f"pst_val_{agg}"] = np.array(v)
        out[f"pst_grad_{agg}"] = g
    out["pairs"] = pairs
    out["stages"] = np.array(stages)
    out["numeric_keys"] = np.array(NUMERIC_KEYS)
    out["meta_json"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLD, "linear_trek_events.npz"), **out)
    print("wrote linear_trek_events.npz")


def pst_series():
    """tests/golden/pst_series.npz: value and autograd gradient of the PST penalty for every series x aggregation
    of the reference at two fixed matrices (notreks.py:454-619, 717-736), the matrix H of ``pst_mat``, the pair
    values of ``agg="none"``, and short ``minimize`` trajectories with the exp / log / binom series in mode "opt"."""
    d, n = 12, 300
    rng = np.random.default_rng(7)
    pairs = np.array([(i, j) for i in range(d) for j in range(i + 1, d) if rng.random() < 0.25], dtype=np.int64)
    out = {"pairs": pairs}
    Ws = {"a": np.random.default_rng(3).uniform(-0.4, 0.4, size=(d, d)) * (np.random.default_rng(4).random((d, d)) < 0.3),
          "b": np.random.default_rng(5).uniform(-1.2, 1.2, size=(d, d)) * (np.random.default_rng(6).random((d, d)) < 0.2)}
    np.fill_diagonal(Ws["b"], 0.0)
    Ws["b"] = np.triu(Ws["b"], 1) + 0.05 * np.tril(Ws["b"], -1)      # near-DAG with larger weights
    import torch
    for wn, Wp in Ws.items():
        out[f"W_{wn}"] = Wp
        for seq in ("inv", "log", "exp", "binom"):
            H = ref_nt.pst_mat(torch.from_numpy(Wp), seq)
            out[f"H_{wn}_{seq}"] = H.numpy()
            out[f"none_{wn}_{seq}"] = ref_nt.pst(torch.from_numpy(Wp), pairs, seq, agg="none").numpy()
            for agg in ("mean", "sum", "max", "lse"):
                reg = ref_nt.PSTRegularizer(I=pairs, seq=seq, weight=1.0, mode="opt", kwargs={"agg": agg})
                v, g = ref_nt.trek_value_grad(Wp, reg)
                out[f"val_{wn}_{seq}_{agg}"] = np.array(v)
                out[f"grad_{wn}_{seq}_{agg}"] = g
    reg = ref_nt.PSTRegularizer(I=pairs, seq="log", weight=1.0, mode="opt", kwargs={"agg": "mean", "K_log": 5})
    v, g = ref_nt.trek_value_grad(Ws["a"], reg)
    out["val_a_log_K5"], out["grad_a_log_K5"] = np.array(v), g
    stages = [(1.0, 200, 1.0, 3e-4), (0.1, 200, 0.9, 3e-4)]
    for name, reg in {
        "exp_mean": ref_nt.PSTRegularizer(I=pairs, seq="exp", weight=0.7, mode="opt"),
        "log_lse": ref_nt.PSTRegularizer(I=pairs, seq="log", weight=0.3, mode="opt", kwargs={"agg": "lse"}),
        "binom_max": ref_nt.PSTRegularizer(I=pairs, seq="binom", weight=0.5, mode="opt", kwargs={"agg": "max"}),
    }.items():
        X, Wt, oks, rows = run(d, n, 11, reg, stages)
        out[f"{name}_X"], out[f"{name}_W"], out[f"{name}_ok"] = X, Wt, np.array(oks)
        out[f"{name}_trek_vals"] = np.array([float(r["reg_trek_value"]) for r in rows])
        print(name, "ok", oks, "max|W|", float(np.abs(Wt[-1]).max()), "trek", out[f"{name}_trek_vals"][-1])
    out["stages"] = np.array(stages)
    np.savez_compressed(os.path.join(GOLD, "pst_series.npz"), **out)
    print("wrote pst_series.npz")


def tcc_spectral():
    """tests/golden/tcc_spectral.npz: the spectral trek-cycle-coupling penalty of the reference
    (notreks.py:156-239 perron_eig_with_gradA, :340-378) -- every version x (power, eig_numpy) at a fixed W with a
    simple Perron root, the Perron pairs themselves, `trek_value_grad` with a TCCRegularizer (the dispatch that ignores
    the regulariser's cycle_penalty / version / s, :699-707, Q14) and a short `minimize` trajectory in mode "opt"."""
    import contextlib
    import io
    import torch
    d, n = 10, 300
    rng = np.random.default_rng(21)
    pairs = np.array([(i, j) for i in range(d) for j in range(i + 1, d) if rng.random() < 0.3], dtype=np.int64)
    W = rng.uniform(-0.8, 0.8, size=(d, d)) * (rng.random((d, d)) < 0.45)
    np.fill_diagonal(W, 0.0)
    out = {"pairs": pairs, "W": W}
    Wt = torch.from_numpy(W)
    versions = ("DAG_learning", "exact_trek_graph", "exact_original_graph", "approx_trek_graph")
    errors = {}
    for method, n_iter in (("power", 50), ("power", 7), ("eig_numpy", 50), ("eig_torch", 50)):
        for version in versions:
            key = f"{method}{n_iter if method == 'power' else ''}_{version}"
            try:
                pen, g = ref_nt.trek_cycle_coupling_value_gradW(Wt, pairs, w=0.7, cycle_penalty="spectral",
                                                                version=version, method=method, n_iter=n_iter)
                out[f"pen_{key}"], out[f"grad_{key}"] = np.array(pen.item()), g.numpy()
            except Exception as e:                              # noqa: BLE001
                errors[key] = type(e).__name__
    # the Perron pairs (2d x 2d block matrix built as the reference does, :319-337)
    W2 = Wt * Wt
    S = torch.zeros(d, d, dtype=torch.double)
    S[pairs[:, 0], pairs[:, 1]] = 1.0
    A = torch.cat([torch.cat([W2, 0.7 * S], dim=1), torch.cat([torch.eye(d, dtype=torch.double), W2.T], dim=1)], dim=0)
    out["A"] = A.numpy()
    for method, n_iter in (("power", 50), ("power", 7), ("eig_numpy", 50)):
        rho, u, v, G = ref_nt.perron_eig_with_gradA(A, method=method, n_iter=n_iter)
        tag = f"{method}{n_iter if method == 'power' else ''}"
        out[f"rho_{tag}"], out[f"u_{tag}"], out[f"v_{tag}"], out[f"G_{tag}"] = np.array(rho.item()), u.numpy(), v.numpy(), G.numpy()
    # dispatch through trek_value_grad: cycle_penalty / version / s of the regulariser are ignored (Q14)
    for mode in ("opt", "log"):
        reg = ref_nt.TCCRegularizer(I=pairs, cycle_penalty="logdet", version="DAG_learning", weight=0.3, w=0.7, s=0.8,
                                    n_iter=10, mode=mode)
        v, g = ref_nt.trek_value_grad(W, reg)
        out[f"tvg_val_{mode}"], out[f"tvg_grad_{mode}"] = np.array(v), g
    # short minimize trajectory, TCC in mode "opt"
    stages = [(1.0, 150, 1.0, 3e-4), (0.1, 150, 0.9, 3e-4)]
    reg = ref_nt.TCCRegularizer(I=pairs, weight=0.5, w=0.7, n_iter=10, mode="opt")
    with contextlib.redirect_stdout(io.StringIO()):
        X, Ws, oks, rows = run(d, n, 13, reg, stages)
    out["fit_X"], out["fit_W"], out["fit_ok"] = X, Ws, np.array(oks)
    out["fit_trek_vals"] = np.array([float(r["reg_trek_value"]) for r in rows])
    out["stages"] = np.array(stages)
    out["errors_json"] = np.array(json.dumps(errors))
    np.savez_compressed(os.path.join(GOLD, "tcc_spectral.npz"), **out)
    print("wrote tcc_spectral.npz; reference errors:", errors, "fit ok", oks, "trek", out["fit_trek_vals"])
    print({k: float(out[k]) for k in out if k.startswith("pen_")})


if __name__ == "__main__":
    if "--tcc-spectral" in sys.argv:
        tcc_spectral()
        sys.exit(0)
    if "--pst-series" in sys.argv:
        pst_series()
    else:
        main()
        pst_series()
