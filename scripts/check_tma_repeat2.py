"""Repeatability of the TMA GEMM schedules and of the blocked inverse (d = 2000) with TMA-fed update tiles."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from midagma_b200 import _lib
from midagma_b200.linear import logdet_inv
lib = _lib.load(); _lib.require_device()
q = torch.zeros(4, dtype=torch.int32, device="cuda")
def tma(a, b, c, mode):
    M, K = a.shape; N = b.shape[1]
    _lib.check(lib.dagma_bench_tma_gemm(_lib.stream_ptr(), M, N, K, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0),
                                        c.data_ptr(), c.stride(0), 1.0, 0.0, mode, q.data_ptr()), "tma_gemm")
torch.manual_seed(1)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 100
d = 2000
a = torch.randn(d, d, dtype=torch.float64, device="cuda"); b = torch.randn(d, d, dtype=torch.float64, device="cuda")
ref = a @ b
for mode in [int(x) for x in os.environ.get("MODES", "0,1,2,4,6").split(",")]:
    nbad = 0; worst = 0.0
    for r in range(R):
        c = torch.randn(d, d, dtype=torch.float64, device="cuda")
        tma(a, b, c, mode)
        e = (c - ref).abs().max().item()
        worst = max(worst, e)
        nbad += e > 1e-9
    print(f"GEMM d={d} mode={mode}: {nbad} of {R} runs wrong (max err {worst:.2e})", flush=True)
if os.environ.get("NO_INV"): sys.exit(0)
rng = np.random.default_rng(0)
A = rng.normal(size=(d, d)) * (rng.random((d, d)) < 0.01)
A *= np.sqrt(0.7 / max(np.abs(np.linalg.eigvals(A * A)).max(), 1e-12))
Ad = torch.from_numpy(A[None]).cuda()
first = None; nbad = 0
for r in range(R // 2):
    out = logdet_inv(Ad, s=0.9, square_input=True, want_inv=True, want_grad=False)["minv"][0].clone()
    if first is None: first = out
    elif not bool((out == first).all().item()): nbad += 1
Minv = np.linalg.inv(0.9 * np.eye(d) - A * A)
print(f"inverse d={d} DAGMA_TMA={os.environ.get('DAGMA_TMA', 'default')}: {nbad} of {R // 2 - 1} runs differ from run 0; "
      f"rel err vs numpy {np.abs(first.cpu().numpy() - Minv).max() / np.abs(Minv).max():.2e}", flush=True)
