"""DMMA / DFMA issue rate versus resident warps per SM (one CTA per SM, 8 independent accumulators per warp)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from midagma_b200 import _lib
lib = _lib.load(); sms = _lib.require_device()
sink = torch.zeros(8, dtype=torch.float64, device="cuda")
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best
for threads in (32, 64, 128, 256, 512, 1024):
    iters = 20000
    t = timed(lambda: lib.dagma_bench_fp64_dmma(_lib.stream_ptr(), sms, threads, iters, sink.data_ptr()))
    tf = sms * (threads // 32) * iters * 8 * 512 / t / 1e12
    clk_per_dmma_per_smsp = t * 1.965e9 / (iters * 8 * max(threads // 128, 1))
    t2 = timed(lambda: lib.dagma_bench_fp64_fma(_lib.stream_ptr(), sms, threads, iters, sink.data_ptr()))
    tf2 = sms * threads * iters * 16 * 2 / t2 / 1e12
    print(f"warps/SM={threads//32:2d}: DMMA {tf:6.2f} TF/s ({clk_per_dmma_per_smsp:5.1f} clk per DMMA per SMSP-warp-slot)  DFMA {tf2:6.2f} TF/s")

for mode in (0, 1):
    for threads in (128, 256, 512):
        iters = 10000
        t = timed(lambda: lib.dagma_bench_fp64_dmma_tiles(_lib.stream_ptr(), sms, threads, iters, mode, sink.data_ptr()))
        print(f"tiles mode={mode} warps/SM={threads//32:2d}: {sms * (threads // 32) * iters * 16 * 512 / t / 1e12:6.2f} TF/s")
