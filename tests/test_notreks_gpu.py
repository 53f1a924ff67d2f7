"""GPU parity of the fork's log-det constraint variant against the reference fixture."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_logdet_acyc_and_tcc_vs_reference(golden, capsys):
    from midagma_b200 import notreks
    g = golden("notreks_logdet")
    for s in (1.0, 0.8):
        h, G = notreks.logdet_acyc_value_gradA(torch.from_numpy(g["A"]), s=s)
        assert abs(h.item() - float(g[f"h_s{s}"])) <= 1e-12
        np.testing.assert_allclose(G.numpy(), g[f"G_s{s}"], rtol=1e-11, atol=1e-14)
    h0, G0 = notreks.logdet_acyc_value_gradA(torch.zeros(5, 5, dtype=torch.float64))
    assert h0.item() == 0.0 and torch.equal(G0, torch.eye(5, dtype=torch.float64))     # SURVEY section 4 KAT
    for version in ("DAG_learning", "exact_trek_graph"):
        pen, gW = notreks.trek_cycle_coupling_value_gradW(torch.from_numpy(g["W"]), g["pairs"], w=0.7,
                                                           cycle_penalty="logdet", version=version, s=1.0)
        assert abs(pen.item() - float(g[f"tcc_pen_{version}"])) <= 1e-12
        np.testing.assert_allclose(gW.numpy(), g[f"tcc_grad_{version}"], rtol=1e-10, atol=1e-14)
        assert version in capsys.readouterr().out
    with pytest.raises(ValueError):
        notreks.trek_cycle_coupling_value_gradW(torch.from_numpy(g["W"]), g["pairs"], cycle_penalty="logdet",
                                                version="approx_trek_graph")
    with pytest.raises(ValueError):
        notreks.logdet_acyc_value_gradA(torch.zeros(3, 4, dtype=torch.float64))


def test_trek_value_grad_noop():
    from midagma_b200 import notreks
    W = np.arange(9.0).reshape(3, 3)
    v, gr = notreks.trek_value_grad(W, None)
    assert v == 0.0 and gr.shape == W.shape and not gr.any()
    off = notreks.TCCRegularizer(I=[(0, 1)], mode="off")
    v, gr = notreks.trek_value_grad(W, off)
    assert v == 0.0 and not gr.any()
    # an enabled PST regulariser (default series "exp") is computed on the GPU: check it against the numpy oracle
    from oracle.notreks_ref import pst_value_grad
    Ws = 0.1 * W
    v, gr = notreks.trek_value_grad(Ws, notreks.PSTRegularizer(I=[(0, 1)], weight=0.1))
    rv, rg, _ = pst_value_grad(Ws, [(0, 1)], seq="exp", agg="mean")
    assert abs(v - rv) <= 1e-12 * abs(rv) and np.abs(gr - rg).max() <= 1e-12 * np.abs(rg).max()
    # TCC: dispatched to the spectral penalty as in the reference (parity: tests/test_tcc_spectral_gpu.py)
    v, gr = notreks.trek_value_grad(W, notreks.TCCRegularizer(I=np.array([(0, 1)]), weight=0.1))
    assert np.isfinite(v) and gr.shape == W.shape
