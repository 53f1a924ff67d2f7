"""A/B of builds of the DMMA fit kernel (run on the GPU box): for every library given on the command line (or
build/variants/libdagma_fit_*.so) -- parity of 100 steps against the oracle at d = 48 / 64, then the rate of the
bench step (4096 problems x 1000 iterations at d = 64), each in its own process (the library is chosen at import)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time
sys.path.insert(0, %(root)r)
import numpy as np, torch
from midagma_b200 import _lib, minimize_batch
from midagma_b200.linear import _run_small
from oracle import simulate
from oracle.linear_ref import OracleLinear
errs = []
for d in (48, 64):
    X, _ = simulate.make_linear_problem(d, 3, 400, "ER", "gauss", 11)
    o = OracleLinear("l2").prepare(X, 0.03, checkpoint=10 ** 9)
    W_ref, _ = o.minimize(np.zeros((d, d)), 1.0, 100, 1.0, 3e-4)
    W, ok, st = minimize_batch(np.zeros((1, d, d)), o.cov[None], 0.03, 1.0, 100, 1.0, 3e-4, checkpoint=10 ** 9)
    errs.append(float(np.abs(W[0] - W_ref).max() / np.abs(W_ref).max()))
nprob, d = 4096, 64
rng = np.random.default_rng(0)
Xs = rng.normal(size=(256, 300, d))
cov = torch.from_numpy(np.einsum("bni,bnj->bij", Xs, Xs) / 300).cuda().repeat(16, 1, 1).contiguous()
lam = torch.full((nprob,), 0.02, dtype=torch.float64, device="cuda")
W = torch.zeros(nprob, d, d, dtype=torch.float64, device="cuda")
def step():
    return _run_small(cov, W, lam, [1.0], [1.0], [1000], lr=3e-4, tol=0.0, beta1=.99, beta2=.999, checkpoint=1000,
                      retry=False, want_final=False)
step(); torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = step(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) * 1e-3)
assert int(r.stage_stats[:, 0, 0].sum().item()) == nprob * 1000
print(f"{os.path.basename(os.environ.get('DAGMA_B200_LIB', 'default')):40s} parity(100 steps, d=48/64) {errs[0]:.1e} {errs[1]:.1e}   "
      f"{nprob * 1000 / best / 1e6:7.3f} M it/s  ({nprob * 1000 / best * 4 * d ** 3 / 1e12:.2f} TFLOP/s)", flush=True)
'''
libs = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "build", "variants", "libdagma_fit_*.so")))
for lib in [None] + libs:
    env = dict(os.environ)
    if lib:
        env["DAGMA_B200_LIB"] = lib
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout.strip() or ("FAILED " + str(lib) + "\n" + r.stderr[-1500:]), flush=True)
