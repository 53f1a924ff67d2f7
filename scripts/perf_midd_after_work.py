"""fit_batch at d = 128 (8 lanes) in a process that has already done other GPU work (a C2 minimize on the launch
sequence with graph capture, side streams, a d = 300 blocked inverse): the lanes must not depend on the state of the
framework's stream pool."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from midagma_b200 import DagmaLinear, fit_batch
from midagma_b200.linear import _batch_lanes
rng = np.random.default_rng(0)
Xs = rng.normal(size=(16, 512, 128))
kw = dict(T=1, warm_iter=3000, max_iter=3000, checkpoint=1000, s=(1.0,), return_info=True)


def run(tag):
    fit_batch(Xs[:2], 0.02, **dict(kw, warm_iter=100, max_iter=100))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _, info = fit_batch(Xs, 0.02, **kw)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    print(f"{tag}: lanes {_batch_lanes(128, 16)}, {info['total_iters'] / t:,.0f} problem-iterations/s", flush=True)


run("fresh process")
os.environ["DAGMA_LIN_FUSED"] = "0"
X = (rng.random((10000, 100)) < 0.5) * 1.0
m = DagmaLinear("logistic"); m.fit(X, lambda1=0.02, T=1, warm_iter=0, max_iter=0)
m.minimize(np.zeros((100, 100)), 1.0, 300, 1.0, lr=3e-4, tol=0.0)
Xl = rng.normal(size=(1200, 300))
m2 = DagmaLinear("l2"); m2.fit(Xl, lambda1=0.02, T=2, warm_iter=200, max_iter=200, s=[1.0, .9])
for _ in range(37):
    torch.cuda.Stream()
os.environ["DAGMA_LIN_FUSED"] = "1"
run("after other work")
