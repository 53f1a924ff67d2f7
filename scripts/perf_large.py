"""Ad-hoc timing of the multi-CTA path: per-iteration time at C5 (l2 d=2000) and C2 (logistic d=100 n=10000)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from midagma_b200 import DagmaLinear
from midagma_b200._large import gemm

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best

for n in (2048, 4096):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); c = torch.empty_like(a)
    t = timed(lambda: gemm(a, a, c)); print(f"my DGEMM {n}^3: {2*n**3/t/1e12:.2f} TFLOP/s")
    t = timed(lambda: torch.matmul(a, a, out=c)); print(f"cuBLAS  {n}^3: {2*n**3/t/1e12:.2f} TFLOP/s")
a = torch.randn(2000, 64, dtype=torch.float64, device="cuda"); b = torch.randn(64, 2000, dtype=torch.float64, device="cuda")
c = torch.zeros(2000, 2000, dtype=torch.float64, device="cuda")
t = timed(lambda: gemm(a, b, c, beta=1.0)); print(f"rank-64 update d=2000: {t*1e6:.1f} us, {2*2000*2000*64/t/1e12:.2f} TFLOP/s")

which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("c5", "both"):
    d, n = 2000, 20000
    rng = np.random.default_rng(0)
    X = rng.normal(size=(n, d))
    m = DagmaLinear("l2")
    t0 = time.time(); m.fit(X, lambda1=0.02, T=1, warm_iter=0, max_iter=0); print("setup+cov s", time.time() - t0)
    eng = m._large_engine()
    W = np.zeros((d, d))
    t0 = time.time(); m.minimize(W, 1.0, 50, 1.0, lr=3e-4); torch.cuda.synchronize(); t1 = time.time() - t0
    t0 = time.time(); m.minimize(W, 1.0, 250, 1.0, lr=3e-4); torch.cuda.synchronize(); t2 = time.time() - t0
    per = (t2 - t1) / 200
    print(f"C5 d=2000: {per*1e3:.3f} ms/iter -> {4*d**3/per/1e12:.2f} TFLOP/s (4d^3)")
    t = timed(lambda: eng._inverse(1.0)); print(f"  inverse d=2000: {t*1e3:.3f} ms ({2*d**3/t/1e12:.2f} TF)")
    t = timed(lambda: eng._score_T()); print(f"  cov@W: {t*1e3:.3f} ms ({2*d**3/t/1e12:.2f} TF)")
    t = timed(lambda: eng._update()); print(f"  update: {t*1e3:.3f} ms")
if which in ("c2", "both"):
    d, n = 100, 10000
    rng = np.random.default_rng(0)
    X = (rng.random((n, d)) < 0.5) * 1.0
    m = DagmaLinear("logistic")
    m.fit(X, lambda1=0.02, T=1, warm_iter=0, max_iter=0)
    eng = m._large_engine()
    W = np.zeros((d, d))
    m.minimize(W, 1.0, 100, 1.0, lr=3e-4); torch.cuda.synchronize()
    t0 = time.time(); m.minimize(W, 1.0, 2100, 1.0, lr=3e-4); torch.cuda.synchronize(); t2 = time.time() - t0
    print(f"C2 logistic d=100 n=10000: {t2/2100*1e6:.1f} us/iter")
    t = timed(lambda: eng._inverse(1.0)); print(f"  inverse d=100: {t*1e6:.1f} us")
    t = timed(lambda: eng._score_T()); print(f"  score GEMMs: {t*1e6:.1f} us")
    t = timed(lambda: eng._update()); print(f"  update: {t*1e6:.1f} us")
