"""Tiny driver for ncu: one launch of the on-chip fit kernel (d, problems, iters from argv)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from midagma_b200 import _lib
from midagma_b200.linear import _run_small

d = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nprob = int(sys.argv[2]) if len(sys.argv) > 2 else 148
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
_lib.require_device()
rng = np.random.default_rng(0)
X = rng.normal(size=(nprob, 200, d))
cov = torch.from_numpy(np.einsum("bni,bnj->bij", X, X) / 200).cuda()
lam = torch.full((nprob,), 0.02, dtype=torch.float64, device="cuda")
for rep in range(2):
    W = torch.zeros(nprob, d, d, dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = _run_small(cov, W, lam, [1.0], [1.0], [iters], lr=3e-4, tol=0.0, beta1=.99, beta2=.999,
                     checkpoint=1000, retry=False, want_final=False)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    print(f"d={d} {nprob}x{iters}: {t*1e3:.2f} ms, {t/iters/max(nprob/148,1)*1e6:.2f} us/iter/SM")
