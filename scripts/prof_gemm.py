"""Tiny driver for ncu: rank-k update and square DGEMM launches of the library kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from midagma_b200._large import gemm
d = 2000; k = int(sys.argv[1]) if len(sys.argv) > 1 else 256
a = torch.randn(d, k, dtype=torch.float64, device="cuda"); b = torch.randn(k, d, dtype=torch.float64, device="cuda")
c = torch.zeros(d, d, dtype=torch.float64, device="cuda")
s = torch.randn(d, d, dtype=torch.float64, device="cuda"); o = torch.empty_like(s)
for _ in range(3):
    gemm(a, b, c, beta=1.0)
    gemm(s, s, o)
torch.cuda.synchronize()
print("ok")
