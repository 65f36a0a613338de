"""Quick numerics check of the fused step on the GPU box (a few sizes, cholesky vs numpy and LL fused vs chain)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200._lib import check, lib, ptr
from cugp_b200.loaders import synthetic_sine

L = lib()
rng = np.random.default_rng(5)
for n in [92, 128, 300, 1500, 3000]:
    M = rng.standard_normal((n, n + 8))
    A = M @ M.T / n + np.eye(n)
    out = np.empty((n, n))
    check(L.cugp_cholesky(ptr(A), ptr(out), n))
    L0 = np.linalg.cholesky(A)
    print(f"n={n} rel fro err {np.linalg.norm(out - L0) / np.linalg.norm(L0):.3e}", flush=True)
X, y = synthetic_sine(1500, 10)
for fused in (1, 0):
    L.cugp_set_tuning(b"fused_step", fused)
    g = cg.Covsum(1500, 10)
    g.set_data(X, y)
    g.set_loghyperparam([3.762111, -1.152105, -0.384461])
    print("fused", fused, "LL", repr(g.loglik_resident()), "grad", g.grad_resident())
    g.close()
L.cugp_set_tuning(b"fused_step", 1)
