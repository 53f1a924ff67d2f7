"""Per-iteration time of the multi-CTA path at C5 (l2, d = 2000): whole iteration, inverse only and
score GEMM only, each replayed as a CUDA graph and timed with CUDA events (no host launch overhead)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from midagma_b200 import DagmaLinear

d = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n = 4 * d
rng = np.random.default_rng(0)
X = rng.normal(size=(n, d))
m = DagmaLinear("l2")
m.fit(X, lambda1=0.02, T=1, warm_iter=0, max_iter=0)
eng = m._large_engine()
W = np.zeros((d, d))
m.minimize(W, 1.0, 30, 1.0, lr=3e-4)          # builds the iteration graph, leaves a non-trivial W
torch.cuda.synchronize()

def graph_of(fn):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g

def timed(g, reps=20):
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps

t_it = timed(eng._graph)
t_inv = timed(graph_of(lambda: eng._inverse(1.0)))
t_gemm = timed(graph_of(lambda: eng._score_T()))
t_upd = timed(graph_of(lambda: eng._update()))
t_fused = timed(graph_of(lambda: eng._inverse_and_score(1.0)))
print(f"d={d}: iteration {t_it*1e3:.3f} ms = {4*d**3/t_it/1e12:.2f} TF/s (4d^3) | inverse {t_inv*1e3:.3f} ms = "
      f"{2*d**3/t_inv/1e12:.2f} TF/s | cov@W {t_gemm*1e3:.3f} ms = {2*d**3/t_gemm/1e12:.2f} TF/s | update {t_upd*1e3:.3f} ms | "
      f"inverse+cov@W fused {t_fused*1e3:.3f} ms = {4*d**3/t_fused/1e12:.2f} TF/s")
