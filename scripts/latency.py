"""Print the dependent-issue latencies measured by dagma_bench_latency (cycles per op, one warp)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from midagma_b200 import _lib
lib = _lib.load(); _lib.require_device()
out = torch.zeros(16, dtype=torch.float64, device="cuda")
for _ in range(2):
    _lib.check(lib.dagma_bench_latency(_lib.stream_ptr(), out.data_ptr())); torch.cuda.synchronize()
names = ["DFMA", "DMMA (acc chain)", "DMMA (A-operand chain)", "SHFL.64", "MUFU.RCP64H+DFMA", "LDS chase",
         "STS+syncwarp+LDS+syncwarp", "DADD", "DMUL"]
for n, v in zip(names, out.tolist()): print(f"{n:28s} {v:7.1f} clk")
print(f"rcp.approx.ftz.f64 max rel err {out[9].item():.3e}   rsqrt.approx.ftz.f64 max rel err {out[10].item():.3e}")
_lib.check(lib.dagma_bench_stage(_lib.stream_ptr(), out.data_ptr())); torch.cuda.synchronize()
print(f"8x8 pivot-block inversion (stage_pivot_block + hand-over): {out[0].item():.0f} clk; with 10 DMMAs interleaved: "
      f"{out[1].item():.0f} clk; |P - inv(inv(P))| = {out[2].item():.2e}")
