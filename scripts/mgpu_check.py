"""torchrun check of the multi-GPU paths (run with N >= 2 GPUs):
  (e1) fit_batch_sharded == fit_batch on one GPU;  (e2) row-sharded logistic DagmaLinear and
  row-sharded DagmaNonlinear == their single-GPU results (to reduction-order round-off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from midagma_b200 import DagmaLinear, fit_batch, parallel
from midagma_b200.nonlinear import DagmaMLP, DagmaNonlinear
from oracle import simulate

# (e1)
n_prob = 9
Xs = np.stack([simulate.make_linear_problem(16, 2, 200, "ER", "gauss", 50 + p)[0] for p in range(n_prob)])
lams = np.linspace(0.01, 0.05, n_prob)
kw = dict(T=3, warm_iter=300, max_iter=400, checkpoint=100)
W_sh = parallel.fit_batch_sharded(Xs, lams, **kw)
W_1 = fit_batch(Xs, lams, **kw)
assert np.array_equal(W_sh, W_1), "sharded batch differs"
# (e2) logistic rows
X, _ = simulate.make_linear_problem(20, 2, 4001, "ER", "logistic", 1)
kw = dict(lambda1=0.02, T=2, warm_iter=150, max_iter=200, checkpoint=50)
W_full = DagmaLinear("logistic").fit(X.copy(), w_threshold=0.0, **kw)
rows = parallel.row_shard(X.shape[0], rank, ws)
W_rows = DagmaLinear("logistic").shard_rows().fit(X[rows].copy(), w_threshold=0.0, **kw)
err = np.abs(W_rows - W_full).max()
assert err < 1e-9, err
# (e2) MLP rows
Xn, _ = simulate.config_c3(seed=0, n=1001, d=12)
torch.manual_seed(0)
m1 = DagmaMLP([12, 4, 1]); m2 = DagmaMLP([12, 4, 1]); m2.load_state_dict(m1.state_dict())
kw = dict(lambda1=0.02, lambda2=0.005, T=2, warm_iter=120, max_iter=150, checkpoint=50)
A_full = DagmaNonlinear(m1).fit(Xn, w_threshold=0.0, **kw)
eq = DagmaNonlinear(m2); eq.group = dist.group.WORLD; eq.n_total = Xn.shape[0]
rows = parallel.row_shard(Xn.shape[0], rank, ws)
A_rows = eq.fit(Xn[rows], w_threshold=0.0, **kw)
err2 = np.abs(A_rows - A_full).max()
assert err2 < 1e-9, err2
dist.barrier()
if rank == 0:
    print(f"mgpu_check ok on {ws} GPUs: batch identical, logistic |dW|={err:.2e}, mlp |dW|={err2:.2e}", flush=True)
# the engines hold CUDA graphs with captured NCCL all-reduces: release them before the communicator goes away
# (destroy_process_group with such graphs still alive does not return)
del eq, m1, m2
import gc
gc.collect()
torch.cuda.synchronize()
dist.destroy_process_group()
