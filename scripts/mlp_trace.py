"""Debug: %globaltimer stamps of one iteration of the one-kernel MLP iteration (needs a -DDAGMA_MLP_TRACE build:
python scripts/build_variant.py mtrace -DDAGMA_MLP_TRACE; DAGMA_B200_LIB=build/variants/libdagma_mtrace.so)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from midagma_b200 import _lib
from midagma_b200 import nonlinear as nlmod
from midagma_b200.nonlinear import DagmaMLP
d, m1, n = 40, 10, 2000
rng = np.random.default_rng(0)
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
eng.replay(1.0, 200); torch.cuda.synchronize()
lib = C.CDLL(_lib.LIB_PATH)
buf = (C.c_ulonglong * 16)()
lib.dagma_debug_mlp_trace.argtypes = [C.c_void_p]
print("rc", lib.dagma_debug_mlp_trace(buf))
t = np.array(buf[:], dtype=np.uint64).astype(np.int64)
names = ["start", "staged", "forward", "res+dZ", "gW1", "row written", "h: A built", "h: sweep", "h: done", "past barrier 1",
         "S", "Adam", "past barrier 2", "last arrival at barrier 1", "  (its CTA)", "first arrival at barrier 1"]
for k, nm in enumerate(names):
    print(f"{nm:28s} {t[k]:7d}" if k == 14 else f"{nm:28s} {(t[k] - t[0]) / 1e3:7.2f} us")
