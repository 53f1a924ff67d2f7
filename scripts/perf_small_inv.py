"""On-chip fused slogdet + inverse (dagma_logdet_inv_f64) for d <= 64: latency of one problem and throughput of a
batch, tensor-core sweep (default for 32 < d <= 64) against the scalar rank-1 sweep (DAGMA_SMALL_INV_DMMA=0)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from midagma_b200.linear import logdet_inv
def timed(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps
rng = np.random.default_rng(0)
for d in [int(x) for x in os.environ.get("DS", "40,48,64").split(",")]:
    for batch in (1, 4096):
        W = rng.uniform(-0.1, 0.1, size=(batch, d, d))
        A = torch.from_numpy(W).cuda()
        out = logdet_inv(A, s=1.0, square_input=True, want_inv=True, want_grad=True)
        M = np.eye(d) - W[0] * W[0]
        err = np.abs(out["minv"][0].cpu().numpy() - np.linalg.inv(M)).max()
        t = timed(lambda: logdet_inv(A, s=1.0, square_input=True, want_inv=True, want_grad=True))
        print(f"d={d} batch={batch}: {t*1e6:9.1f} us per call ({2*d**3*batch/t/1e12:6.2f} TF/s)  err {err:.1e}  "
              f"DAGMA_SMALL_INV_DMMA={os.environ.get('DAGMA_SMALL_INV_DMMA', '1')}", flush=True)
