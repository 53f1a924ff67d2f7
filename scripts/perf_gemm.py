"""DGEMM timing: library kernel (DAGMA_GEMM_TILE selects the tile) versus cuBLAS at the shapes the path uses."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from midagma_b200._large import gemm

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best

print("tile variant", os.environ.get("DAGMA_GEMM_TILE", "0"))
for n in (2000, 2048, 4096):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); c = torch.empty_like(a)
    t = timed(lambda: gemm(a, a, c)); t2 = timed(lambda: torch.matmul(a, a, out=c))
    err = (c - a @ a).abs().max().item()
    print(f"{n}^3: mine {2*n**3/t/1e12:.2f} TF/s  cuBLAS {2*n**3/t2/1e12:.2f} TF/s  (check {err:.1e})")
for k in (64, 128, 256):
    d = 2000
    a = torch.randn(d, k, dtype=torch.float64, device="cuda"); b = torch.randn(k, d, dtype=torch.float64, device="cuda")
    c = torch.zeros(d, d, dtype=torch.float64, device="cuda")
    t = timed(lambda: gemm(a, b, c, beta=1.0))
    c.zero_(); gemm(a, b, c, beta=1.0); err = (c - a @ b).abs().max().item()
    print(f"rank-{k} update d={d}: {t*1e6:.1f} us, {2*d*d*k/t/1e12:.2f} TF/s (check {err:.1e})")
a = torch.randn(2000, 256, dtype=torch.float64, device="cuda"); q = torch.randn(256, 256, dtype=torch.float64, device="cuda")
c = torch.zeros(2000, 256, dtype=torch.float64, device="cuda")
t = timed(lambda: gemm(a, q, c)); print(f"CS = C Q (2000x256x256): {t*1e6:.1f} us, {2*2000*256*256/t/1e12:.2f} TF/s")
