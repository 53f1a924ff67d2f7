"""Small shapes of the TMA GEMM (all schedules, edge tiles, K tails) and one blocked inverse with TMA-fed update tiles,
checked against torch / numpy: a quick stand-alone correctness run (compute-sanitizer is closed on the GPU pool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from midagma_b200 import _lib
from midagma_b200.linear import logdet_inv
lib = _lib.load(); _lib.require_device()
q = torch.zeros(4, dtype=torch.int32, device="cuda")
torch.manual_seed(0)
for (M, N, K) in [(1300, 1290, 520), (130, 70, 50)]:
    a = torch.randn(M, K, dtype=torch.float64, device="cuda"); b = torch.randn(K, N, dtype=torch.float64, device="cuda")
    ref = a @ b
    for mode in range(8):
        c = torch.zeros(M, N, dtype=torch.float64, device="cuda")
        _lib.check(lib.dagma_bench_tma_gemm(_lib.stream_ptr(), M, N, K, a.data_ptr(), K, b.data_ptr(), N, c.data_ptr(), N, 1.0, 0.0, mode, q.data_ptr()), "tma")
        torch.cuda.synchronize()
        assert (c - ref).abs().max().item() < 1e-9, (M, N, K, mode)
d = 520
rng = np.random.default_rng(0)
A = rng.normal(size=(d, d)) * (rng.random((d, d)) < 0.05) * 0.2
out = logdet_inv(torch.from_numpy(A[None]).cuda(), s=1.0, square_input=True, want_inv=True, want_grad=False)
err = np.abs(out["minv"][0].cpu().numpy() - np.linalg.inv(np.eye(d) - A * A)).max()
assert err < 1e-9, err
print("sanitize_tma ok", err)
