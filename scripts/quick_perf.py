"""Ad-hoc GPU timing: FP64 yardsticks and the batched on-chip iteration rate."""
import ctypes as C
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from midagma_b200 import _lib
from midagma_b200.linear import _run_small

lib = _lib.load()
sms = _lib.require_device()
print("SMs", sms, torch.cuda.get_device_name())
sink = torch.zeros(8, dtype=torch.float64, device="cuda")


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best

for threads in (256, 512, 1024):
    ctas = sms * (2048 // threads)
    iters = 20000
    t = timed(lambda: lib.dagma_bench_fp64_fma(_lib.stream_ptr(), ctas, threads, iters, sink.data_ptr()))
    print(f"DFMA  threads={threads}: {ctas*threads*iters*16*2/t/1e12:.2f} TFLOP/s")
    t = timed(lambda: lib.dagma_bench_fp64_dmma(_lib.stream_ptr(), ctas, threads, iters // 8, sink.data_ptr()))
    print(f"DMMA  threads={threads}: {ctas*(threads//32)*(iters//8)*8*512/t/1e12:.2f} TFLOP/s")

a = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
t = timed(lambda: torch.matmul(a, a))
print(f"cuBLAS DGEMM 4096^3: {2*4096**3/t/1e12:.2f} TFLOP/s")
del a

for d, nprob, iters in ((64, sms * 4, 2000), (20, sms * 8, 2000), (32, sms * 4, 2000), (48, sms * 4, 2000)):
    rng = np.random.default_rng(0)
    X = rng.normal(size=(nprob, 200, d))
    cov = torch.from_numpy(np.einsum("bni,bnj->bij", X, X) / 200).cuda()
    lam = torch.full((nprob,), 0.02, dtype=torch.float64, device="cuda")
    def run():
        W = torch.zeros(nprob, d, d, dtype=torch.float64, device="cuda")
        return _run_small(cov, W, lam, [1.0], [1.0], [iters], lr=3e-4, tol=0.0, beta1=.99, beta2=.999,
                          checkpoint=1000, retry=False, want_final=False)
    t = timed(run, reps=2)
    res = run(); torch.cuda.synchronize()
    done = res.stage_stats[:, 0, 0].sum().item()
    rate = done / t
    print(f"d={d}: {nprob} problems x {iters} iters in {t*1e3:.1f} ms -> {rate/1e6:.2f} M it/s, "
          f"{rate*4*d**3/1e12:.2f} TFLOP/s (4d^3), {t/iters/(nprob/sms)*1e6:.2f} us/iter/SM")
