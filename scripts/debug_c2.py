"""Debug: where does the logistic d=100 n=10000 iteration leave the oracle?  (run on the GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import simulate
from oracle.linear_ref import OracleLinear
from midagma_b200 import DagmaLinear
from midagma_b200._large import gemm
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 100
X, _ = simulate.config_c2(0, n=n, d=d)
o = OracleLinear("logistic").prepare(X.copy(), 0.02, checkpoint=10 ** 9)
m = DagmaLinear("logistic")
m.fit(X.copy(), lambda1=0.02, T=1, warm_iter=0, max_iter=0, checkpoint=10 ** 9)
print("env", {k: v for k, v in os.environ.items() if k.startswith("DAGMA")}, "n", n, "d", d)
print("cov err", np.abs(m.cov - o.cov).max())
W0 = np.zeros((d, d))
sv, sg = m._score(W0)
ov, og = o.score(W0)
print("score at 0: val", sv, ov, "grad err", np.abs(sg - og).max(), "max|G|", np.abs(og).max())
rng = np.random.default_rng(0)
Wr = rng.normal(size=(d, d)) * 0.05
sv, sg = m._score(Wr)
ov, og = o.score(Wr)
e = np.abs(sg - og)
print("score at random W: val", sv, ov, "grad err", e.max(), "at", np.unravel_index(e.argmax(), e.shape))
for K in (1, 2, 3, 5, 10, 50, 150):
    Wo, _ = o.minimize(np.zeros((d, d)), 1.0, K, 1.0, 3e-4)
    Wg, ok = m.minimize(np.zeros((d, d)), 1.0, K, 1.0, lr=3e-4)
    e = np.abs(Wg - Wo)
    i, j = np.unravel_index(e.argmax(), e.shape)
    print(f"K={K:4d}: max|dW| {e.max():.3e} at ({i},{j}) Wo {Wo[i, j]:+.6e} Wg {Wg[i, j]:+.6e}  count>1e-9: {(e > 1e-9).sum()}  ok {ok}")
