#!/bin/bash
mkdir -p gpurun_out
echo "env at start: CUDA_DEVICE_MAX_CONNECTIONS=${CUDA_DEVICE_MAX_CONNECTIONS:-unset}" > gpurun_out/c28.log
timeout 200 python scripts/perf_midd_batch.py 128 16 3000 8 >> gpurun_out/c28.log 2>&1
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 200 python scripts/perf_midd_batch.py 128 16 3000 8 >> gpurun_out/c28.log 2>&1
timeout 200 python - >> gpurun_out/c28.log 2>&1 <<'P'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
x = torch.zeros(10, device="cuda")          # CUDA context BEFORE the package is imported
import midagma_b200
print("context first, env now", os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"))
from midagma_b200 import fit_batch
rng = np.random.default_rng(0)
Xs = rng.normal(size=(16, 512, 128))
kw = dict(T=1, warm_iter=3000, max_iter=3000, checkpoint=1000, s=(1.0,), return_info=True)
fit_batch(Xs[:2], 0.02, **dict(kw, warm_iter=100, max_iter=100))
torch.cuda.synchronize(); t0 = time.perf_counter()
_, info = fit_batch(Xs, 0.02, **kw)
torch.cuda.synchronize(); t = time.perf_counter() - t0
print(f"context first: {info['total_iters'] / t:,.0f} problem-iterations/s")
P
cat gpurun_out/c28.log
