"""TMA-fed GEMM (gemm_tma.cuh) against the cp.async GEMM and torch: correctness on edge shapes, then timing of
C = A @ B at d = 2000 / 4096 and of the rank-256 update C += A @ B (the outer step of the blocked inverse)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from midagma_b200 import _lib
from midagma_b200._large import gemm
lib = _lib.load(); _lib.require_device()
q = torch.zeros(4, dtype=torch.int32, device="cuda")

def tma(a, b, c, alpha=1.0, beta=0.0, mode=0):
    M, K = a.shape; N = b.shape[1]
    _lib.check(lib.dagma_bench_tma_gemm(_lib.stream_ptr(), M, N, K, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0),
                                        c.data_ptr(), c.stride(0), alpha, beta, mode, q.data_ptr()), "tma_gemm")

def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps

torch.manual_seed(0)
if not os.environ.get('SKIP_CHECK'):
  # mode bits: 1 = atomic tile queue, 2 = balanced k-slab ranges (stream-K, beta = 0 only), 4 = 128 x 128 tiles
  for (M, N, K, beta, modes) in [(2000, 2000, 2000, 0.0, range(8)), (2000, 2000, 256, 1.0, (0, 1, 4, 5)),
                                 (1990, 1234, 208, 1.0, (0, 1, 4, 5)), (1990, 1234, 208, 0.0, range(8)), (130, 70, 50, 0.0, range(8)),
                                 (64, 64, 16, 0.0, (0, 4)), (4096, 4096, 4096, 0.0, (1, 6)), (3000, 2500, 1111 * 2, 0.0, (2, 6))]:
      a = torch.randn(M, K, dtype=torch.float64, device="cuda"); b = torch.randn(K, N, dtype=torch.float64, device="cuda")
      c0 = torch.randn(M, N, dtype=torch.float64, device="cuda")
      ref = a @ b + beta * c0
      for mode in modes:
          c = c0.clone()
          tma(a, b, c, 1.0, beta, mode)
          torch.cuda.synchronize()
          err = (c - ref).abs().max().item()
          c2 = c0.clone(); tma(a, b, c2, 1.0, beta, mode); torch.cuda.synchronize()
          print(f"check M={M} N={N} K={K} beta={beta} mode={mode}: max|diff| {err:.2e}  (|ref| {ref.abs().max().item():.1f})  "
                f"repeatable={bool((c == c2).all().item())}", flush=True)

for d in (2000, 4096):
    a = torch.randn(d, d, dtype=torch.float64, device="cuda"); b = torch.randn(d, d, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    os.environ["DAGMA_TMA"] = "1"
    t = timed(lambda: gemm(a, b, c)); print(f"d={d} dagma_gemm_f64 (default policy): {t*1e3:.3f} ms = {2*d**3/t/1e12:.2f} TF/s")
    for mode in range(8):
        t = timed(lambda: tma(a, b, c, 1.0, 0.0, mode)); print(f"d={d} TMA GEMM mode {mode}    : {t*1e3:.3f} ms = {2*d**3/t/1e12:.2f} TF/s")
    t = timed(lambda: torch.matmul(a, b, out=c)); print(f"d={d} cuBLAS              : {t*1e3:.3f} ms = {2*d**3/t/1e12:.2f} TF/s")
d, k = 2000, 256
a = torch.randn(d, k, dtype=torch.float64, device="cuda"); b = torch.randn(k, d, dtype=torch.float64, device="cuda")
c = torch.zeros(d, d, dtype=torch.float64, device="cuda")
t = timed(lambda: gemm(a, b, c, alpha=1.0, beta=1.0)); print(f"rank-256 update dagma_gemm_f64: {t*1e6:.1f} us = {2*d*d*k/t/1e12:.2f} TF/s")
for mode in (0, 1, 4, 5):
    t = timed(lambda: tma(a, b, c, 1.0, 1.0, mode)); print(f"rank-256 update TMA mode {mode}: {t*1e6:.1f} us = {2*d*d*k/t/1e12:.2f} TF/s")
