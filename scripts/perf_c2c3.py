"""Per-iteration time of C2 (logistic d=100 n=10000) and C3 (MLP d=40 m1=10 n=2000): wall clock over many
graph-replayed iterations (the host only synchronises at the checkpoints, every 1000 iterations)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from midagma_b200 import DagmaLinear
from midagma_b200.nonlinear import DagmaMLP, DagmaNonlinear

rng = np.random.default_rng(0)
# ---- C2
d, n = 100, 10000
X = (rng.random((n, d)) < 0.5) * 1.0
m = DagmaLinear("logistic")
m.fit(X, lambda1=0.02, T=1, warm_iter=0, max_iter=0)
W = np.zeros((d, d))
m.minimize(W, 1.0, 200, 1.0, lr=3e-4); torch.cuda.synchronize()
t0 = time.perf_counter(); m.minimize(W, 1.0, 4000, 1.0, lr=3e-4, tol=0.0); torch.cuda.synchronize()
t = time.perf_counter() - t0
print(f"C2 logistic d=100 n=10000: {t/4000*1e6:.1f} us/iter ({(4*n*d*d+2*d**3)/(t/4000)/1e12:.2f} TF/s)")
# ---- C3
d, m1, n = 40, 10, 2000
X = rng.normal(size=(n, d))
torch.manual_seed(0)
model = DagmaMLP(dims=[d, m1, 1], bias=True)
nl = DagmaNonlinear(model)
t0 = time.perf_counter()
nl.fit(X, lambda1=0.02, lambda2=0.005, T=1, warm_iter=3000, max_iter=3000, checkpoint=1000)
torch.cuda.synchronize(); t = time.perf_counter() - t0
t0 = time.perf_counter()
nl.fit(X, lambda1=0.02, lambda2=0.005, T=1, warm_iter=6000, max_iter=6000, checkpoint=1000)
torch.cuda.synchronize(); t2 = time.perf_counter() - t0
print(f"C3 MLP d=40 m1=10 n=2000: {(t2 - t)/3000*1e6:.1f} us/iter (difference of a 6000- and a 3000-iteration fit)")
