"""Repeatability of the balanced (stream-K) TMA GEMM: the same product 30 times into fresh buffers; reports which
64 x 64 / 128 x 128 tiles ever differ between runs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from midagma_b200 import _lib
lib = _lib.load(); _lib.require_device()
q = torch.zeros(4, dtype=torch.int32, device="cuda")
def tma(a, b, c, mode):
    M, K = a.shape; N = b.shape[1]
    _lib.check(lib.dagma_bench_tma_gemm(_lib.stream_ptr(), M, N, K, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0),
                                        c.data_ptr(), c.stride(0), 1.0, 0.0, mode, q.data_ptr()), "tma_gemm")
torch.manual_seed(1)
for d in (2000, 3000):
    a = torch.randn(d, d, dtype=torch.float64, device="cuda"); b = torch.randn(d, d, dtype=torch.float64, device="cuda")
    for mode in (2,):
        outs = []
        for r in range(60):
            c = torch.randn(d, d, dtype=torch.float64, device="cuda")
            tma(a, b, c, mode)
            outs.append(c)
        torch.cuda.synchronize()
        bad = [r for r in range(1, 60) if not bool((outs[r] == outs[0]).all().item())]
        msg = ""
        if bad:
            diff = (outs[bad[0]] != outs[0])
            t = 64 if mode == 2 else 128
            rows = sorted(set((diff.nonzero()[:, 0] // t).tolist())); cols = sorted(set((diff.nonzero()[:, 1] // t).tolist()))
            nz = diff.nonzero()
            ref = a @ b
            e_bad = (outs[bad[0]] - ref)[diff]; e_ok = (outs[0] - ref)[diff]
            print("   in-tile rows", sorted(set((nz[:, 0] % t).tolist())), "cols", sorted(set((nz[:, 1] % t).tolist())))
            print("   err of bad run vs torch: max %.3e  | err of run 0 vs torch at the same elements: max %.3e" % (e_bad.abs().max().item(), e_ok.abs().max().item()))
            print("   bad values / reference:", (outs[bad[0]][diff][:6]).tolist(), ref[diff][:6].tolist())
            for rb in bad[:3]:
                dd = (outs[rb] != outs[0]).nonzero()
                print("   run", rb, "tile", (dd[0, 0] // t).item(), (dd[0, 1] // t).item(), "count", dd.shape[0])
            msg = f" first bad run {bad[0]}: {int(diff.sum())} elements, tile rows {rows[:8]} cols {cols[:8]} max {(outs[bad[0]]-outs[0]).abs().max().item():.2e}"
        print(f"d={d} mode={mode}: {len(bad)} of 59 runs differ from run 0{msg}", flush=True)
