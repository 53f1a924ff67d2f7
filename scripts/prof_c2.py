"""ncu target: one launch of the persistent DagmaLinear iteration kernel at C2 (logistic d=100 n=10000, 1000 iterations)
or at a mid-d l2 problem.  Usage: prof_c2.py [logistic|l2] [d] [n] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from midagma_b200 import DagmaLinear
loss = sys.argv[1] if len(sys.argv) > 1 else "logistic"
d = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
rng = np.random.default_rng(0)
X = (rng.random((n, d)) < 0.5) * 1.0 if loss == "logistic" else rng.normal(size=(n, d))
m = DagmaLinear(loss)
m.fit(X, lambda1=0.02, T=1, warm_iter=0, max_iter=0, checkpoint=iters)
W = np.zeros((d, d))
m.minimize(W, 1.0, iters, 1.0, lr=3e-4, tol=0.0)
torch.cuda.synchronize()
print("done", loss, d, n, iters, "iterations", m.last_iters)
