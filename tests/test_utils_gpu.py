"""Batched graph evaluation (csrc/graph_metrics.cu, midagma_b200/utils.py) against the numpy restatement of the
reference's utils.count_accuracy / is_dag (oracle/simulate.py; src/dagma/utils.py:13-18, 245-310)."""
import numpy as np
import pytest

from oracle import simulate

pytestmark = pytest.mark.gpu


def _random_dag(d, p, rng):
    B = np.triu((rng.random((d, d)) < p).astype(np.int64), 1)
    perm = rng.permutation(d)
    return B[np.ix_(perm, perm)]


@pytest.mark.parametrize("d", [1, 2, 7, 20, 64, 100, 257])
def test_count_accuracy_and_is_dag_vs_oracle(d):
    from midagma_b200 import utils
    rng = np.random.default_rng(d)
    for trial in range(6):
        B_true = _random_dag(d, 0.2, rng)
        B_est = _random_dag(d, 0.25, rng)
        # share some edges with the truth, reverse a few
        keep = rng.random((d, d)) < 0.5
        B_est = np.where(keep, B_true, B_est)
        if not simulate.is_dag(B_est):
            B_est = np.triu(B_est | B_est.T, 1)
        assert utils.is_dag(B_est) == simulate.is_dag(B_est) is True
        ref = simulate.count_accuracy(B_true, B_est)
        out = utils.count_accuracy(B_true, B_est)
        assert out == ref, (trial, out, ref)
    # cyclic estimates
    C = np.zeros((max(d, 2), max(d, 2)), dtype=np.int64)
    C[0, 1] = C[1, 0] = 1
    assert not utils.is_dag(C) and not simulate.is_dag(C)
    with pytest.raises(ValueError, match="DAG"):
        utils.count_accuracy(np.zeros_like(C), C)
    with pytest.raises(ValueError, match="value in"):
        utils.count_accuracy(np.zeros_like(C), 2 * C)


def test_count_accuracy_batch_and_weights():
    from midagma_b200 import utils
    rng = np.random.default_rng(3)
    d, batch = 64, 37
    truth = np.stack([_random_dag(d, 0.1, rng) for _ in range(batch)])
    W = np.stack([_random_dag(d, 0.12, rng) * rng.normal(size=(d, d)) for _ in range(batch)])
    W[5, 3, 9] = W[5, 9, 3] = 0.7                                       # one cyclic estimate
    out = utils.count_accuracy_batch(truth, (W != 0).astype(np.int8))
    for b in range(batch):
        assert out["is_dag"][b] == simulate.is_dag(W[b])
        ref = simulate.count_accuracy(truth[b], (W[b] != 0).astype(np.int64))
        for k in ("fdr", "tpr", "fpr", "shd", "nnz"):
            assert out[k][b] == ref[k], (b, k)
    assert not out["is_dag"][5] and list(utils.is_dag_batch(W)) == list(out["is_dag"])
    # one truth shared by the whole batch
    out1 = utils.count_accuracy_batch(truth[0], (W != 0).astype(np.int8))
    assert out1["shd"][0] == out["shd"][0]
    assert out1["shd"][1] == simulate.count_accuracy(truth[0], (W[1] != 0).astype(np.int64))["shd"]


def test_count_accuracy_cpdag():
    """-1 marks an undirected edge (utils.py:269-273, 284-292): counted favourably against the skeleton."""
    from midagma_b200 import utils
    B_true = np.zeros((5, 5), dtype=np.int64)
    B_true[0, 1] = B_true[1, 2] = B_true[3, 4] = 1
    B_est = np.zeros((5, 5), dtype=np.int64)
    B_est[1, 0] = -1          # undirected, on the skeleton: true positive
    B_est[2, 1] = 1           # reversed
    B_est[0, 4] = -1          # undirected, off the skeleton: false positive
    out = utils.count_accuracy(B_true, B_est)
    # by hand: pred_size 3, true_pos 1, reverse 1, false_pos 1; lower: pred {(1,0),(2,1),(4,0)}, cond {(1,0),(2,1),(4,3)}
    assert out == {"fdr": 2 / 3, "tpr": 1 / 3, "fpr": 2 / (0.5 * 5 * 4 - 3), "shd": 1 + 1 + 1, "nnz": 3}
    B_bad = B_est.copy()
    B_bad[0, 1] = -1
    with pytest.raises(ValueError, match="only appear once"):
        utils.count_accuracy(B_true, B_bad)
