"""A/B of one tuning key on the GPU box: LL+gradient time (several n), 2-expert BCM evaluation and prediction.
usage: r2_ab.py <key> <v0,v1,...> [n list]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200._lib import lib
from cugp_b200.loaders import synthetic_sine

L = lib()
key = sys.argv[1].encode()
vals = [int(v) for v in sys.argv[2].split(",")]
sizes = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1500, 2048, 4096]
TH_B = [3.762111, -1.152105, -0.384461]
d = np.load("tests/golden/data_si24000.npz")
Xt, _ = synthetic_sine(10000, 10, seed=7)
for v in vals:
    assert L.cugp_set_tuning(key, v) == 0, L.cugp_last_error()
    line = f"{key.decode()}={v}:"
    for n in sizes:
        X, y = synthetic_sine(n, 10)
        g = cg.Covsum(n, 10)
        g.set_data(X, y)
        ts, tc = [], []
        for r in range(7):
            g.set_loghyperparam([TH_B[0] + 1e-7 * r, TH_B[1], TH_B[2]])
            t = time.perf_counter()
            g.loglik_resident()
            g.grad_resident()
            ts.append(time.perf_counter() - t)
        for r in range(3):
            g.set_loghyperparam([TH_B[0] + 1e-7 * (r + 20), TH_B[1], TH_B[2]])
            tc.append(g.factorize_resident()[1])
        line += f"  n={n}: LL+grad {1e3 * min(ts[2:]):.3f} ms (chol {min(tc):.3f})"
        g.close()
    b = cg.BCM(d["X"][:3000], d["y"][:3000], K=2, rank=0, world=1)
    te, tp = [], []
    for r in range(7):
        b.set_BCM_log_hyperparam([2.0 + 1e-7 * r, 2.0, 2.0])
        t = time.perf_counter()
        b.loglik_and_gradient()
        te.append(time.perf_counter() - t)
        t = time.perf_counter()
        b.compute_BCM_test_means_and_var(Xt)
        tp.append(time.perf_counter() - t)
    b.close()
    line += f"  BCM 2x1500: eval {1e3 * min(te[2:]):.3f} ms, predict(10000, incl. new factor) {1e3 * min(tp[2:]):.3f} ms"
    print(line, flush=True)
