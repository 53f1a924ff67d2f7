"""Timing experiment: the same 64 x 64 tiles of C = A @ B (d = 2000) by the stand-alone GEMM kernel (128-thread CTAs,
four per SM) and by the engine pairs of the persistent inverse kernels (256-thread CTAs, two per SM, named barriers),
one tile per engine or pulled from an atomic queue."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from midagma_b200 import _lib
from midagma_b200._large import gemm
lib = _lib.load(); _lib.require_device()
d = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
a = torch.randn(d, d, dtype=torch.float64, device="cuda"); b = torch.randn(d, d, dtype=torch.float64, device="cuda")
c0 = torch.empty_like(a); c1 = torch.empty_like(a); q = torch.zeros(4, dtype=torch.int32, device="cuda")
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps
t = timed(lambda: gemm(a, b, c0)); print(f"stand-alone GEMM        : {t*1e3:.3f} ms = {2*d**3/t/1e12:.2f} TF/s")
for pers in (0, 1):
    fn = lambda: _lib.check(lib.dagma_bench_engine_gemm(_lib.stream_ptr(), d, a.data_ptr(), b.data_ptr(), c1.data_ptr(), pers, q.data_ptr()))
    t = timed(fn); print(f"engine pairs, persistent={pers}: {t*1e3:.3f} ms = {2*d**3/t/1e12:.2f} TF/s   max|diff| {(c0-c1).abs().max().item():.2e}")
