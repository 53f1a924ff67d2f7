"""us per inner iteration of the logistic DagmaLinear at d = 100 for growing n (C2's rows tiled): the persistent kernel
(rows resident up to 72 x 147, streamed beyond) against the launch sequence (DAGMA_LIN_FUSED=0)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time
sys.path.insert(0, %(root)r)
import numpy as np, torch
from midagma_b200 import DagmaLinear
from oracle import simulate
X0, _ = simulate.config_c2(0)
for mult in (1, 2, 4, 8, 16):
    X = np.tile(X0, (mult, 1))
    n, d = X.shape
    m = DagmaLinear("logistic")
    m.fit(X.copy(), lambda1=0.02, T=1, warm_iter=0, max_iter=0, checkpoint=10 ** 9)
    W = np.zeros((d, d))
    m.minimize(W, 1.0, 100, 1.0, lr=3e-4, tol=0.0)
    W[...] = 0.0
    iters = 2000 if mult <= 4 else 500
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.minimize(W, 1.0, iters, 1.0, lr=3e-4, tol=0.0)
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / iters
    print(f"n={n:7d}: {t * 1e6:7.1f} us/iter  ({(4.0 * n * d * d + 2.0 * d ** 3) / t / 1e12:5.2f} TF/s)  "
          f"{'one kernel' if m._large.one_kernel else 'launch sequence'}  checksum {np.abs(W).sum():.12e}", flush=True)
'''
for fused in ("1", "0"):
    print("== DAGMA_LIN_FUSED=" + fused, flush=True)
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=dict(os.environ, DAGMA_LIN_FUSED=fused),
                       capture_output=True, text=True, timeout=900)
    print(r.stdout.strip() or r.stderr[-2000:], flush=True)
