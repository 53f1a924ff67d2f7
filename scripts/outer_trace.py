"""Debug: %globaltimer stamps of the fused outer steps of the two-level inverse (needs a -DDAGMA_OUTER_TRACE build:
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
for _ in range(3): eng._inverse(1.0)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * 640)()
lib.dagma_debug_outer_trace.argtypes = [C.c_void_p]
print("rc", lib.dagma_debug_outer_trace(buf))
t = np.array(buf[:], dtype=np.uint64).astype(np.int64).reshape(16, 40)
nob = (d + 255) // 256
t00 = t[0][0]
print("all times in us relative to the first CTA of the step; 'abs' = start of the step relative to step 0")
for ob in range(nob):
    r = t[ob]; z = r[0]
    f = lambda q: (r[q] - z) / 1e3 if r[q] > 0 and r[q] < (1 << 62) else float('nan')
    print(f"step {ob}: abs {(z - t00)/1e3:7.1f} | srv: tile_upd {f(1):5.1f} barrier {f(2):5.1f} steps " +
          " ".join(f"{f(3+k):5.1f}" for k in range(4)) + f" all_done {f(7):5.1f} | upd first {f(12):5.1f} last {f(8):5.1f} | "
          f"CS' first {f(9):5.1f} last {f(10):5.1f} | R' last {f(13):5.1f} | exit {f(11):5.1f}")
    for kb in range(4):
        b = 16 + 5 * kb
        if r[b] > 0:
            print(f"      tile step {kb}: loads {f(b):5.1f} sweep {f(b+1):5.1f} CS {f(b+2):5.1f} Rload {f(b+3):5.1f} written {f(b+4):5.1f}")
