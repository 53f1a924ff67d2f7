"""The fork's spectral trek-cycle-coupling penalty on device (midagma_b200/_tcc.py, csrc/spectral.cu) against values,
Perron pairs, gradients and an optimisation trajectory recorded from the unmodified reference
(oracle/make_golden_trek.py --tcc-spectral -> tests/golden/tcc_spectral.npz; reference:
src/notreks/notreks.py:156-239, 291-378, 699-707)."""
import json

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _relmax(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("tag,method,n_iter", [("power50", "power", 50), ("power7", "power", 7),
                                               ("eig_numpy", "eig_numpy", 50)])
def test_perron_pair_vs_reference(golden, tag, method, n_iter):
    from midagma_b200.notreks import perron_eig_with_gradA
    g = golden("tcc_spectral")
    rho, u, v, G = perron_eig_with_gradA(torch.from_numpy(g["A"]), method=method, n_iter=n_iter)
    tol = 1e-12 if method == "power" else 1e-10
    assert abs(rho.item() - float(g[f"rho_{tag}"])) <= tol * abs(float(g[f"rho_{tag}"]))
    assert _relmax(u.numpy(), g[f"u_{tag}"]) <= tol and _relmax(v.numpy(), g[f"v_{tag}"]) <= tol
    assert _relmax(G.numpy(), g[f"G_{tag}"]) <= tol


@pytest.mark.parametrize("tag,method,n_iter", [("power50", "power", 50), ("power7", "power", 7),
                                               ("eig_numpy", "eig_numpy", 50), ("eig_torch", "eig_torch", 50)])
@pytest.mark.parametrize("version", ["DAG_learning", "exact_trek_graph", "approx_trek_graph", "exact_original_graph"])
def test_tcc_spectral_value_grad_vs_reference(golden, tag, method, n_iter, version):
    from midagma_b200.notreks import trek_cycle_coupling_value_gradW
    g = golden("tcc_spectral")
    errors = json.loads(str(g["errors_json"]))
    W, pairs = torch.from_numpy(g["W"]), g["pairs"]
    call = lambda: trek_cycle_coupling_value_gradW(W, pairs, w=0.7, cycle_penalty="spectral", version=version,  # noqa: E731
                                                   method=method, n_iter=n_iter)
    key = f"{tag}_{version}"
    if key in errors:                                   # exact_original_graph: the reference raises RuntimeError
        with pytest.raises(RuntimeError):
            call()
        return
    if version == "exact_trek_graph" and method != "power":
        # defective baseline matrix: numpy and torch eigenvectors of the REFERENCE disagree (1e-3 on the gradient)
        a, b = g["grad_eig_numpy_exact_trek_graph"], g["grad_eig_torch_exact_trek_graph"]
        assert _relmax(a, b) > 1e-6
        with pytest.raises(NotImplementedError):
            call()
        return
    pen, grad = call()
    tol = 1e-11 if method == "power" else 1e-9
    assert abs(pen.item() - float(g[f"pen_{key}"])) <= tol * abs(float(g[f"pen_{key}"]))
    assert _relmax(grad.numpy(), g[f"grad_{key}"]) <= tol


def test_trek_value_grad_tcc_dispatch(golden):
    """A TCCRegularizer through trek_value_grad: the reference ignores its cycle_penalty / version / s (Q14) and runs
    spectral / approx_trek_graph / eig_numpy; the forward flag honours the regulariser."""
    from midagma_b200 import notreks
    g = golden("tcc_spectral")
    W, pairs = g["W"], g["pairs"]
    for mode in ("opt", "log"):
        reg = notreks.TCCRegularizer(I=pairs, cycle_penalty="logdet", version="DAG_learning", weight=0.3, w=0.7, s=0.8,
                                     n_iter=10, mode=mode)
        v, gr = notreks.trek_value_grad(W, reg)
        assert abs(v - float(g[f"tvg_val_{mode}"])) <= 1e-9 * abs(float(g[f"tvg_val_{mode}"]))
        assert np.abs(gr - g[f"tvg_grad_{mode}"]).max() <= 1e-9 * max(np.abs(g[f"tvg_grad_{mode}"]).max(), 1e-300) + 0.0
    # forwarded configuration: the log-det DAG_learning penalty of the same block matrix (row a11)
    reg = notreks.TCCRegularizer(I=pairs, cycle_penalty="logdet", version="DAG_learning", weight=0.3, w=0.7, s=3.0,
                                 mode="opt")
    v_fwd, g_fwd = notreks.trek_value_grad(W, reg, forward_tcc_config=True)
    pen, grad = notreks.trek_cycle_coupling_value_gradW(torch.from_numpy(W), pairs, w=0.7, cycle_penalty="logdet",
                                                        version="DAG_learning", s=3.0)
    assert abs(v_fwd - pen.item()) <= 1e-12 * abs(pen.item()) and np.array_equal(g_fwd, grad.numpy())
    assert abs(v_fwd - float(g["tvg_val_opt"])) > 1e-3          # not the spectral default
    # disabled / empty pair set: the no-op branch
    off = notreks.TCCRegularizer(I=pairs, weight=0.0)
    v0, g0 = notreks.trek_value_grad(W, off)
    assert v0 == 0.0 and g0.shape == W.shape and not g0.any()


def test_minimize_with_tcc_regulariser_vs_reference(golden):
    """DagmaLinear.minimize with a TCC regulariser in mode "opt" (linear.py:251-258): trajectory of the reference."""
    from midagma_b200 import DagmaLinear, notreks
    from midagma_b200.logger import LogConfig
    g = golden("tcc_spectral")
    rows = []
    reg = notreks.TCCRegularizer(I=g["pairs"], weight=0.5, w=0.7, n_iter=10, mode="opt")
    m = DagmaLinear("l2", trek_reg=reg,
                    log_cfg=LogConfig(enabled=True, store_jsonl=False, store_csv=False, callback=rows.append))
    X = g["fit_X"].copy()
    m.fit(X, lambda1=0.02, T=1, warm_iter=0, max_iter=0, checkpoint=100)
    d = m.d
    W = np.zeros((d, d))
    for si, (mu, iters, s, lr) in enumerate(g["stages"]):
        W, ok = m.minimize(W, mu, int(iters), s, lr)
        assert ok == bool(g["fit_ok"][si])
        err = np.abs(W - g["fit_W"][si]).max()
        print("TCC minimize stage", si, "max|dW|", err)
        assert err <= 1e-8
    vals = np.array([r["reg_trek_value"] for r in rows if r["event"] == "minimize.checkpoint"])
    assert vals.shape == g["fit_trek_vals"].shape
    assert np.abs(vals - g["fit_trek_vals"]).max() <= 1e-8 * np.abs(g["fit_trek_vals"]).max()
