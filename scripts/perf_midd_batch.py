"""Throughput of fit_batch at 64 < d <= 128 (l2): problems side by side (one persistent kernel per problem,
csrc/lin_iter.cu) against one after the other.  Usage: perf_midd_batch.py [d] [batch] [iters] [lanes,lanes,...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from midagma_b200 import fit_batch
from midagma_b200.linear import _batch_lanes
d = int(sys.argv[1]) if len(sys.argv) > 1 else 100
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
rng = np.random.default_rng(0)
Xs = rng.normal(size=(batch, 4 * d, d))
kw = dict(T=1, warm_iter=iters, max_iter=iters, checkpoint=1000, s=(1.0,), tol=0.0, return_info=True)
fit_batch(Xs[:2], 0.02, **dict(kw, warm_iter=100, max_iter=100))          # warm-up (module load, kernel attributes)
for lanes in (sys.argv[4].split(",") if len(sys.argv) > 4 else ("auto", "1")):
    if lanes == "auto":
        os.environ.pop("DAGMA_BATCH_LANES", None)
    else:
        os.environ["DAGMA_BATCH_LANES"] = lanes
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, info = fit_batch(Xs, 0.02, **kw)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    print(f"d={d} batch={batch} lanes={_batch_lanes(d, batch)}: {t:.2f} s, {info['total_iters'] / t:,.0f} problem-iterations/s "
          f"({t / info['total_iters'] * _batch_lanes(d, batch) * 1e6:.1f} us per iteration and lane)")
