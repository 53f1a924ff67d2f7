"""ncu target: one launch of the persistent MLP iteration kernel at C3 ([40, 10, 1], n = 2000, 300 iterations)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from midagma_b200 import nonlinear as nlmod
from midagma_b200.nonlinear import DagmaMLP
from oracle import simulate
X, _ = simulate.config_c3(0)
torch.manual_seed(0)
model = DagmaMLP(dims=[40, 10, 1], bias=True)
eng = nlmod._MlpEngine(model, torch.from_numpy(X).cuda())
sh = eng.state_host
sh.zero_()
for f, val in ((nlmod.F_MU, 0.1), (nlmod.F_S, 1.0), (nlmod.F_LR, 2e-4), (nlmod.F_LAM1, 0.02), (nlmod.F_LAM2, 0.005),
               (nlmod.F_B1, 0.99), (nlmod.F_B2, 0.999), (nlmod.F_GAMMA, 1.0)):
    sh[f] = float(val)
eng.state.copy_(sh)
eng.replay(1.0, int(sys.argv[1]) if len(sys.argv) > 1 else 300)
torch.cuda.synchronize()
print("done", eng.pull()[1])
