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


if __name__ == "__main__":
    if "--pst-series" in sys.argv:
        pst_series()
    else:
        main()
        pst_series()
