"""Where the wall clock of a reduced-schedule C5 fit goes beyond the kernels: cProfile of DagmaLinear.fit (d = 2000)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from midagma_b200 import DagmaLinear
from oracle import simulate
X, _ = simulate.config_c5(0)
m = DagmaLinear("l2")
m.fit(X.copy(), lambda1=0.02, T=1, warm_iter=20, max_iter=20)          # warm-up: library, allocations
torch.cuda.synchronize()
m = DagmaLinear("l2")
Xh = X.copy()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
m.fit(Xh, lambda1=0.02, warm_iter=500, max_iter=500)
torch.cuda.synchronize()
pr.disable()
print(f"fit wall {time.perf_counter() - t0:.3f} s, iterations {sum(m.stage_iters)}")
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
