"""Tiny driver for ncu: a few inner iterations of the multi-CTA path at C5 size (l2, d = 2000)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from midagma_b200 import DagmaLinear
d = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rng = np.random.default_rng(0)
X = rng.normal(size=(4 * d, d))
m = DagmaLinear("l2")
m.fit(X, lambda1=0.02, T=1, warm_iter=0, max_iter=0)
W = np.zeros((d, d))
W, ok = m.minimize(W, 1.0, 6, 1.0, lr=3e-4)
print("ok", ok, float(np.abs(W).max()))
