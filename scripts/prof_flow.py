"""Tiny driver for ncu: a few fused inverse + cov@W calls (the dependency-driven kernel) at d from argv (default 2000)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from midagma_b200 import DagmaLinear
d = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rng = np.random.default_rng(0)
X = rng.normal(size=(4 * d, d))
m = DagmaLinear("l2")
m.fit(X, lambda1=0.02, T=1, warm_iter=0, max_iter=0)
eng = m._large_engine()
eng.W.copy_(torch.from_numpy(rng.uniform(-0.01, 0.01, size=(d, d))))
for _ in range(4):
    eng._inverse_and_score(1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng._inverse_and_score(1.0); e1.record(); torch.cuda.synchronize()
print(f"d={d}: inverse + cov@W {e0.elapsed_time(e1):.3f} ms")
