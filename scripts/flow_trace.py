"""Debug: %globaltimer stamps of the dependency-driven inverse kernel (needs a -DDAGMA_OUTER_TRACE build:
python scripts/build_variant.py otrace -DDAGMA_OUTER_TRACE; DAGMA_B200_LIB=build/variants/libdagma_otrace.so)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from midagma_b200 import DagmaLinear, _lib
d = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rng = np.random.default_rng(0)
X = rng.normal(size=(4 * d, d))
m = DagmaLinear("l2")
m.fit(X, lambda1=0.02, T=1, warm_iter=0, max_iter=0)
eng = m._large_engine()
W = np.zeros((d, d))
m.minimize(W, 1.0, 30, 1.0, lr=3e-4)
torch.cuda.synchronize()
fn = eng._inverse_and_score if (len(sys.argv) > 2 and sys.argv[2] == "rider") else eng._inverse
for _ in range(3): fn(1.0)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * 640)()
lib.dagma_debug_outer_trace.argtypes = [C.c_void_p]
print("rc", lib.dagma_debug_outer_trace(buf))
t = np.array(buf[:], dtype=np.uint64).astype(np.int64).reshape(16, 40)
nob = (d + 255) // 256
z = t[0][2]
f = lambda v: (v - z) / 1e3 if 0 < v < (1 << 62) else float('nan')
print("us since the chain team started block 0's inversion")
for k in range(nob):
    r = t[k]
    print(f"block {k}: deps {f(r[0]):7.1f} CS {f(r[1]):7.1f} inv-start {f(r[2]):7.1f} Q done {f(r[3]):7.1f} | step {k} updates: first {f(r[12]):7.1f} last {f(r[8]):7.1f}")
    for kb in range(4):
        b = 16 + 5 * kb
        if k == 1 and r[b] > 0:
            print(f"      tile step {kb}: loads {f(r[b]):5.1f} sweep {f(r[b+1]):5.1f} CS {f(r[b+2]):5.1f} Rload {f(r[b+3]):5.1f} written {f(r[b+4]):5.1f}")
