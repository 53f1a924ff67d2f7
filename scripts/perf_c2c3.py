"""Per-iteration time of C2 (logistic d=100 n=10000; wall clock over 4000 graph-replayed iterations, the host only
synchronises at the checkpoints) and C3 (MLP d=40 m1=10 n=2000; CUDA events over 4000 replays of the iteration graph)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from midagma_b200 import DagmaLinear
from midagma_b200.nonlinear import DagmaMLP, DagmaNonlinear

rng = np.random.default_rng(0)
# ---- C2
d, n = 100, 10000
X = (rng.random((n, d)) < 0.5) * 1.0
m = DagmaLinear("logistic")
m.fit(X, lambda1=0.02, T=1, warm_iter=0, max_iter=0)
W = np.zeros((d, d))
m.minimize(W, 1.0, 200, 1.0, lr=3e-4); torch.cuda.synchronize()
t0 = time.perf_counter(); m.minimize(W, 1.0, 4000, 1.0, lr=3e-4, tol=0.0); torch.cuda.synchronize()
t = time.perf_counter() - t0
print(f"C2 logistic d=100 n=10000: {t/4000*1e6:.1f} us/iter ({(4*n*d*d+2*d**3)/(t/4000)/1e12:.2f} TF/s)")
# ---- C3: the per-iteration CUDA graph of DagmaNonlinear.minimize, replayed and timed with CUDA events
from midagma_b200 import nonlinear as nlmod
d, m1, n = 40, 10, 2000
X = rng.normal(size=(n, d))
torch.manual_seed(0)
model = DagmaMLP(dims=[d, m1, 1], bias=True)
eng = nlmod._MlpEngine(model, torch.from_numpy(X).cuda())
sh = eng.state_host
sh.zero_()
for f, val in ((nlmod.F_MU, 0.1), (nlmod.F_S, 1.0), (nlmod.F_LR, 2e-4), (nlmod.F_LAM1, 0.02), (nlmod.F_LAM2, 0.005),
               (nlmod.F_B1, 0.99), (nlmod.F_B2, 0.999), (nlmod.F_GAMMA, 1.0)):
    sh[f] = float(val)
eng.state.copy_(sh)
eng.replay(1.0, 500); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.replay(1.0, 4000); e1.record(); torch.cuda.synchronize()
print(f"C3 MLP d=40 m1=10 n=2000: {e0.elapsed_time(e1)/4000*1e3:.1f} us/iter (4000 graph replays, CUDA events)")
