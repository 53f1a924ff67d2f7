"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: problem / row sharding, the gather
that re-orders results, and the sum-all-reduce protocol of the row-sharded scores.  The compute
stand-in is the numpy oracle -- the point here is the plumbing, not the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from midagma_b200 import parallel


def test_shard_helpers():
    for n, ws in ((4096, 8), (10, 3), (7, 8), (10000, 4)):
        seen = np.concatenate([parallel.problem_shard(n, r, ws) for r in range(ws)])
        assert sorted(seen.tolist()) == list(range(n))
        rows = [parallel.row_shard(n, r, ws) for r in range(ws)]
        assert rows[0].start == 0 and rows[-1].stop == n
        assert all(a.stop == b.start for a, b in zip(rows, rows[1:]))
        sizes = [r.stop - r.start for r in rows]
        assert max(sizes) - min(sizes) <= 1
    assert parallel.world() == (0, 1)
    t = torch.arange(6.0).reshape(3, 2)
    assert parallel.gather_problems(t, 3) is t and torch.equal(parallel.allreduce_sum_(t.clone()), t)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from oracle import simulate
    from oracle.linear_ref import OracleLinear
    from scipy.special import expit
    # --- (e1) independent problems: round-robin split, gather in problem order
    n_prob, d = 5, 6
    Xs = np.stack([simulate.make_linear_problem(d, 1, 80, "ER", "gauss", 10 + p)[0] for p in range(n_prob)])
    lams = np.array([0.01, 0.02, 0.03, 0.05, 0.02])
    kw = dict(T=2, warm_iter=60, max_iter=80, checkpoint=20)

    def solver(Xl, lam, **k):
        return np.stack([OracleLinear("l2").fit(x.copy(), lambda1=float(l), w_threshold=0.0, **k)
                         for x, l in zip(Xl, lam)]) if len(Xl) else np.zeros((0, d, d))

    W_all = parallel.fit_batch_sharded(Xs, lams, solver=solver, **kw)
    ref = solver(Xs, lams, **kw)
    assert W_all.shape == (n_prob, d, d) and np.array_equal(W_all, ref)
    # fewer problems than ranks: the rank with an empty share still joins the gather (no hang, no solver call)
    W_one = parallel.fit_batch_sharded(Xs[:1], lams[:1], solver=solver, **kw)
    assert W_one.shape == (1, d, d) and np.array_equal(W_one[0], ref[0])
    # --- (e2) row-sharded logistic score: partial X^T sigmoid(XW) summed == full gradient
    X, _ = simulate.make_linear_problem(8, 2, 101, "ER", "logistic", 3)
    W = np.random.default_rng(0).normal(size=(8, 8)) * 0.1
    rows = parallel.row_shard(X.shape[0], rank, ws)
    Xl = X[rows]
    part = torch.from_numpy(Xl.T @ expit(Xl @ W))
    cov = torch.from_numpy(Xl.T @ Xl)
    n_tot = torch.tensor([float(Xl.shape[0])])
    for t in (part, cov, n_tot):
        parallel.allreduce_sum_(t)
    n = int(n_tot.item())
    G = part.numpy() / n - cov.numpy() / n
    o = OracleLinear("logistic").prepare(X.copy(), 0.02)
    assert n == X.shape[0] and np.abs(G - o.score(W)[1]).max() < 1e-12
    # gather of a ragged last shard
    loc = torch.full((len(parallel.problem_shard(7, rank, ws)), 2), float(rank))
    full = parallel.gather_problems(loc, 7)
    assert full[:, 0].tolist() == [float(p % ws) for p in range(7)]
    # --- peer-memory exchange of the persistent kernels: the decision is collective; without NCCL every rank gets None
    # (the engines then keep the launch sequence + all-reduce), also when only one rank does not want it
    from midagma_b200._peer import PeerExchange
    assert PeerExchange.create(dist.group.WORLD, 4096, "cpu", True) is None
    assert PeerExchange.create(dist.group.WORLD, 4096, "cpu", rank == 0) is None
    assert PeerExchange._agree(dist.group.WORLD, "cpu", True) is True
    assert PeerExchange._agree(dist.group.WORLD, "cpu", rank == 1) is False
    dist.barrier()
    open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_gloo(tmp_path):
    ws = 2
    mp.spawn(_worker, args=(ws, _free_port(), str(tmp_path)), nprocs=ws, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(ws))
