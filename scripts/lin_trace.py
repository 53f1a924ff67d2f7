"""Debug: %globaltimer stamps of the last iteration of the one-kernel DagmaLinear iteration (csrc/lin_iter.cu; needs a
-DDAGMA_LIN_TRACE build: python scripts/build_variant.py ltrace -DDAGMA_LIN_TRACE;
DAGMA_B200_LIB=build/variants/libdagma_ltrace.so).  Usage: lin_trace.py [logistic|l2] [d] [n]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from midagma_b200 import DagmaLinear, _lib
loss = sys.argv[1] if len(sys.argv) > 1 else "logistic"
d = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
rng = np.random.default_rng(0)
X = (rng.random((n, d)) < 0.5) * 1.0 if loss == "logistic" else rng.normal(size=(n, d))
m = DagmaLinear(loss)
m.fit(X, lambda1=0.02, T=1, warm_iter=0, max_iter=0)
assert m._large_engine().one_kernel
W = np.zeros((d, d))
m.minimize(W, 1.0, 300, 1.0, lr=3e-4, tol=0.0)
torch.cuda.synchronize()
lib = C.CDLL(_lib.LIB_PATH)
buf = (C.c_ulonglong * 16)()
lib.dagma_debug_lin_trace.argtypes = [C.c_void_p]
print("rc", lib.dagma_debug_lin_trace(buf), loss, "d", d, "n", n)
t = np.array(buf[:], dtype=np.uint64).astype(np.int64)
names = ["start", "W staged", "Z done", "R written", "partial written", "past workers' barrier", "T reduced",
         "past barrier 1", "step taken", "past barrier 2", "inv: start", "inv: M built", "inv: inverted", "inv: outputs",
         "last arrival (workers)", "first arrival (workers)"]
for k, nm in enumerate(names):
    print(f"{nm:24s} {(t[k] - t[0]) / 1e3:7.2f} us" if t[k] else f"{nm:24s}       -")
