"""Tiny driver for ncu: a few d x d blocked inverses through the C ABI (d from argv, default 2000)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from midagma_b200.linear import logdet_inv
d = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rng = np.random.default_rng(0)
W = rng.uniform(-0.02, 0.02, size=(1, d, d))
A = torch.from_numpy(W).cuda()
for _ in range(3):
    out = logdet_inv(A, s=1.0, square_input=True, want_inv=True, want_grad=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = logdet_inv(A, s=1.0, square_input=True, want_inv=True, want_grad=False); e1.record()
torch.cuda.synchronize()
print(f"d={d}: inverse {e0.elapsed_time(e1):.3f} ms  h={float(out["h"][0]):.6e} info={int(out["info"][0])}")
