"""Debug: cycle stamps of the serial chain of the DMMA sweep (needs a -DDAGMA_SWEEP_TRACE build)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from midagma_b200 import _lib
from midagma_b200.linear import _run_small
lib = _lib.load(); _lib.require_device()
nprob = int(sys.argv[1]) if len(sys.argv) > 1 else 148
d = 64
rng = np.random.default_rng(0)
X = rng.normal(size=(nprob, 200, d))
cov = torch.from_numpy(np.einsum("bni,bnj->bij", X, X) / 200).cuda()
lam = torch.full((nprob,), 0.02, dtype=torch.float64, device="cuda")
W = torch.zeros(nprob, d, d, dtype=torch.float64, device="cuda")
_run_small(cov, W, lam, [1.0], [1.0], [100], lr=3e-4, tol=0.0, beta1=.99, beta2=.999, checkpoint=1000, retry=False, want_final=False)
torch.cuda.synchronize()
buf = (C.c_longlong * 512)()
lib.dagma_debug_sweep_trace.argtypes = [C.c_void_p]
print("rc", lib.dagma_debug_sweep_trace(buf))
t = np.array(buf[:], dtype=np.int64).reshape(64, 8)
print("step: diag[wait_enter->wait_exit | ->tile_dmma | ->inverse | ->publish+arrive]   other[wait | work]   step period (diag exit-to-exit)")
ph = t[16]
print(f"phases of iteration 51 (clk): build M {ph[1]-ph[0]} | sweep {ph[2]-ph[1]} | score GEMM {ph[3]-ph[2]} | feasibility "
      f"{ph[4]-ph[3]} | Adam pass {ph[5]-ph[4]} | whole iteration {ph[5]-ph[0]}")
for b in range(7):
    r = t[b]
    nxt = t[b + 1][1] if b + 1 < 8 else 0
    print(f"{b:2d}: {r[1]-r[0]:5d} {r[2]-r[1]:5d} {r[3]-r[2]:5d} {r[4]-r[3]:5d}   other: {r[6]-r[5]:5d} {r[7]-r[6]:5d}   period {nxt - r[1] if nxt else 0:6d}")
