"""Evidence run (GPU box): the blocked d = 2000 inverse on iterates of a REAL SF4 path-following run (BASELINE config 5)
against LAPACK on the host -- the "late-stage snapshots" of SURVEY.md 7.4 (4).  A reduced schedule (`iters` inner
iterations per stage, default 4000; the reference needs ~31 h for the default one) is driven stage by stage; at the end
of every stage the current W is inverted at this stage's s and at the next stage's s (the matrix the next stage starts
from) on the GPU and by numpy.linalg.inv, and the table below goes to profiles/.

    python scripts/c5_path_snapshots.py [iters_per_stage] > profiles/c5_path_snapshots_r2.txt
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import simulate
from midagma_b200 import DagmaLinear
from midagma_b200.linear import logdet_inv

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
X, W_true = simulate.config_c5(0)
d = X.shape[1]
m = DagmaLinear("l2")
m.fit(X, lambda1=0.02, T=1, warm_iter=0, max_iter=0, checkpoint=1000)
S = [1.0, .9, .8, .7, .6]
mu, W = 1.0, np.zeros((d, d))
Id = np.eye(d)
print(f"# C5 path snapshots: SF4 d={d} n={X.shape[0]}, lambda1=0.02, {iters} iterations per stage (reduced schedule), "
      f"{torch.cuda.get_device_name()}")
print("# stage  s_eval  iters  ok   max|W|   cond_1(M)   max(M^-1)  min(M^-1) lapack / gpu   feasible lapack / gpu(info)   "
      "max-norm rel |gpu - lapack|   residual |M X - I| gpu / lapack   gpu inverse ms")
for t, s in enumerate(S):
    t0 = time.time()
    W, ok = m.minimize(W, mu, iters, s, lr=3e-4)
    wall = time.time() - t0
    for s_eval in ([s] if t + 1 == len(S) else [s, S[t + 1]]):
        Wd = torch.from_numpy(np.ascontiguousarray(W[None])).cuda()
        out = logdet_inv(Wd, s=s_eval, square_input=True, want_inv=True, want_grad=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = logdet_inv(Wd, s=s_eval, square_input=True, want_inv=True, want_grad=False)
        e1.record()
        torch.cuda.synchronize()
        Xg = out["minv"][0].cpu().numpy()
        M = s_eval * Id - W * W
        Xl = np.linalg.inv(M)
        cond1 = np.abs(M).sum(0).max() * np.abs(Xl).sum(0).max()
        rel = np.abs(Xg - Xl).max() / np.abs(Xl).max()
        rg, rl = np.abs(M @ Xg - Id).max(), np.abs(M @ Xl - Id).max()
        print(f"{t}  {s_eval:.1f}  {m.last_iters:6d}  {int(ok)}  {np.abs(W).max():.3f}  {cond1:.2e}  {Xl.max():.2e}  "
              f"{Xl.min():+.2e} / {out['min_entry'][0].item():+.2e}   {int(not np.any(Xl + 1e-16 < 0))} / "
              f"{int(out['info'][0].item() == 0)}({int(out['info'][0].item())})   {rel:.2e}   {rg:.2e} / {rl:.2e}   "
              f"{e0.elapsed_time(e1):.3f}   [stage wall {wall:.1f} s]", flush=True)
    mu *= 0.1
Wt = W.copy()
Wt[np.abs(Wt) < 0.3] = 0
acc = simulate.count_accuracy(W_true != 0, Wt != 0) if simulate.is_dag(Wt) else {"shd": "not a DAG yet"}
print("# thresholded estimate after the reduced schedule:", int((Wt != 0).sum()), "edges, true", int((W_true != 0).sum()), acc)
