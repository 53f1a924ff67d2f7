"""torchrun check (N >= 2 GPUs of one box) of the row-sharded persistent DagmaLinear iteration whose cross-GPU sum runs
inside the kernel over NVLink peer memory (csrc/lin_iter.cu, csrc/peer.cu):
  parity   sharded (peer) == sharded (launch sequence + NCCL all-reduce) == one GPU, to reduction-order round-off; the
           replicas on the ranks are bit-identical;
  timing   us per inner iteration at d = 100 for growing n: peer kernel / NCCL sequence / one GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from midagma_b200 import DagmaLinear, parallel
from oracle import simulate


def fit_rows(X, peer, **kw):
    os.environ["DAGMA_LIN_PEER"] = "1" if peer else "0"
    m = DagmaLinear("logistic").shard_rows()
    rows = parallel.row_shard(X.shape[0], rank, ws)
    W = m.fit(X[rows].copy(), w_threshold=0.0, **kw)
    used = m._large is not None and m._large._peer is not None
    m.close()
    return W, used, m.stage_iters


# ---- parity
for d, n in ((20, 4001), (100, 3000), (128, 2049)):
    X, _ = simulate.make_linear_problem(d, 2, n, "ER", "logistic", 1)
    kw = dict(lambda1=0.02, T=2, warm_iter=150, max_iter=200, checkpoint=50)
    W_one = DagmaLinear("logistic").fit(X.copy(), w_threshold=0.0, **kw)
    W_peer, used, it_p = fit_rows(X, True, **kw)
    W_nccl, used0, it_n = fit_rows(X, False, **kw)
    assert used and not used0, (used, used0)
    e1, e2 = np.abs(W_peer - W_one).max(), np.abs(W_peer - W_nccl).max()
    t = torch.from_numpy(W_peer).cuda()
    allw = [torch.empty_like(t) for _ in range(ws)]
    dist.all_gather(allw, t)
    same = all(torch.equal(allw[0], a) for a in allw)
    if rank == 0:
        print(f"parity d={d} n={n}: |dW| peer vs one GPU {e1:.2e}, peer vs NCCL sequence {e2:.2e}, replicas identical: {same}, "
              f"stage iterations {it_p} / {it_n}", flush=True)
    # peer vs NCCL sequence: the same decomposition, different summation order of the rank sums (round-off).  Against ONE
    # GPU the logistic fit from W = 0 can differ by the l1 chatter of the exact-tie entries: any
    # two summation orders do -- reported, and bounded like the NCCL sequence's own distance to one GPU
    e3 = np.abs(W_nccl - W_one).max()
    assert e2 < 1e-9 and same and it_p == it_n and e1 <= max(1e-9, 4 * e3), (e1, e2, e3)


# ---- timing at d = 100
def time_minimize(X, mode, iters):
    rows = parallel.row_shard(X.shape[0], rank, ws) if mode != "one" else slice(None)
    os.environ["DAGMA_LIN_PEER"] = "1" if mode == "peer" else "0"
    m = DagmaLinear("logistic")
    if mode != "one":
        m.shard_rows()
    m.fit(X[rows].copy(), lambda1=0.02, T=1, warm_iter=0, max_iter=0, checkpoint=10 ** 9)
    W = np.zeros((X.shape[1], X.shape[1]))
    m.minimize(W, 1.0, 100, 1.0, lr=3e-4, tol=0.0)
    W = np.zeros((X.shape[1], X.shape[1]))
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    m.minimize(W, 1.0, iters, 1.0, lr=3e-4, tol=0.0)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    kind = "peer kernel" if (m._large._peer is not None) else ("one kernel" if m._large.one_kernel else "launch sequence")
    m.close()
    return t.item() / iters * 1e6, kind, W


X0, _ = simulate.config_c2(0)
for mult in (1, 2, 4, 8):
    X = np.tile(X0, (mult, 1))
    n = X.shape[0]
    iters = 2000 if mult <= 2 else 500
    t_p, k_p, W_p = time_minimize(X, "peer", iters)
    t_n, k_n, W_n = time_minimize(X, "nccl", iters)
    t_1, k_1, W_1 = time_minimize(X, "one", iters)
    if rank == 0:
        print(f"d=100 n={n} on {ws} GPUs: sharded ({k_p}) {t_p:.1f} us/iter | sharded NCCL ({k_n}) {t_n:.1f} | one GPU ({k_1}) "
              f"{t_1:.1f} | |dW| vs one GPU {np.abs(W_p - W_1).max():.1e}", flush=True)
# ---- the MLP (DagmaNonlinear, dims [d, m1, 1]): the same exchange inside csrc/mlp_iter.cu
from midagma_b200.nonlinear import DagmaMLP, DagmaNonlinear


def mlp_fit(Xn, mode, dims, **kw):
    os.environ["DAGMA_MLP_PEER"] = "1" if mode == "peer" else "0"
    torch.manual_seed(0)
    model = DagmaMLP(dims)
    eq = DagmaNonlinear(model)
    Xl = Xn
    if mode != "one":
        eq.group, eq.n_total = dist.group.WORLD, Xn.shape[0]
        Xl = Xn[parallel.row_shard(Xn.shape[0], rank, ws)]
    A = eq.fit(Xl, w_threshold=0.0, **kw)
    used = eq._engine is not None and eq._engine._peer is not None
    eq.close()
    return A, used


for d, m1, n in ((12, 4, 1001), (40, 10, 2000)):
    Xn, _ = simulate.config_c3(seed=0, n=n, d=d)
    kw = dict(lambda1=0.02, lambda2=0.005, T=2, warm_iter=120, max_iter=150, checkpoint=50)
    A_one, _ = mlp_fit(Xn, "one", [d, m1, 1], **kw)
    A_peer, used = mlp_fit(Xn, "peer", [d, m1, 1], **kw)
    A_nccl, used0 = mlp_fit(Xn, "nccl", [d, m1, 1], **kw)
    assert used and not used0, (used, used0)
    e1, e2 = np.abs(A_peer - A_one).max(), np.abs(A_peer - A_nccl).max()
    t = torch.from_numpy(np.ascontiguousarray(A_peer)).cuda()
    allw = [torch.empty_like(t) for _ in range(ws)]
    dist.all_gather(allw, t)
    same = all(torch.equal(allw[0], a) for a in allw)
    if rank == 0:
        print(f"MLP parity [{d}, {m1}, 1] n={n}: |dA| peer vs one GPU {e1:.2e}, peer vs NCCL sequence {e2:.2e}, replicas identical: {same}",
              flush=True)
    assert e1 < 1e-9 and e2 < 1e-9 and same


def mlp_time(Xn, mode, iters):
    os.environ["DAGMA_MLP_PEER"] = "1" if mode == "peer" else "0"
    d = Xn.shape[1]
    torch.manual_seed(0)
    model = DagmaMLP([d, 10, 1])
    eq = DagmaNonlinear(model)
    Xl = Xn
    if mode != "one":
        eq.group, eq.n_total = dist.group.WORLD, Xn.shape[0]
        Xl = Xn[parallel.row_shard(Xn.shape[0], rank, ws)]
    eq.X = torch.from_numpy(np.ascontiguousarray(Xl)).cuda()
    if mode == "one":
        eq.n_total = Xn.shape[0]
    eq.checkpoint = 10 ** 9
    eq.minimize(100, 2e-4, 0.02, 0.005, 0.1, 1.0, tol=0.0)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    eq.minimize(iters, 2e-4, 0.02, 0.005, 0.1, 1.0, tol=0.0)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    eng = eq._engine
    kind = "peer kernel" if eng._peer is not None else ("one kernel" if eng.one_kernel else "launch sequence")
    eq.close()
    return t.item() / iters * 1e6, kind


X3, _ = simulate.config_c3(0)
for mult in (1, 4, 16):
    Xn = np.tile(X3, (mult, 1))
    t_p, k_p = mlp_time(Xn, "peer", 1000)
    t_n, k_n = mlp_time(Xn, "nccl", 1000)
    t_1, k_1 = mlp_time(Xn, "one", 1000)
    if rank == 0:
        print(f"MLP [40, 10, 1] n={Xn.shape[0]} on {ws} GPUs: sharded ({k_p}) {t_p:.1f} us/iter | sharded NCCL ({k_n}) {t_n:.1f} | "
              f"one GPU ({k_1}) {t_1:.1f}", flush=True)
dist.barrier()
import gc
gc.collect()
torch.cuda.synchronize()
dist.destroy_process_group()
