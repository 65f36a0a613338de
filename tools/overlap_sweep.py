"""Inverse overlapped with the factorisation (tuning overlap_inv_max_n / overlap_inv_cap): LL+gradient time per
evaluation against the sequential path, single GPs of several sizes and the C4-shaped batch."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200._lib import lib
from cugp_b200.loaders import synthetic_sine

L = lib()
TH_B = [3.762111, -1.152105, -0.384461]
modes = [("seq", 0, 64), ("cap96", 1 << 20, 96), ("cap148", 1 << 20, 148)]


def bench(make, evaluate, label):
    base = None
    for name, mx, cap in modes:
        L.cugp_set_tuning(b"overlap_inv_max_n", mx)
        L.cugp_set_tuning(b"overlap_inv_cap", cap)
        obj = make()
        ts, last = [], None
        for r in range(7):
            t0 = time.perf_counter()
            last = evaluate(obj, [TH_B[0] + 1e-7 * r, TH_B[1], TH_B[2]])
            ts.append(time.perf_counter() - t0)
        obj.close()
        ms = 1e3 * sorted(ts[2:])[2]
        if base is None:
            base = (ms, last)
        err = float(np.max(np.abs(last[1] - base[1][1]) / np.maximum(np.abs(base[1][1]), np.abs(base[1][1]).max())))
        print(f"{label} {name:7s}: LL+grad {ms:8.3f} ms  x{base[0]/ms:5.2f}  grad relerr vs seq {err:.1e}  LL {last[0]:.9f}", flush=True)


sizes = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [3000, 4096, 6000, 8192]
for n in sizes:
    X, y = synthetic_sine(n, 10)

    def make():
        g = cg.Covsum(n, 10)
        g.set_data(X, y)
        return g

    def ev(g, th):
        g.set_loghyperparam(th)
        return g.loglik_resident(), g.grad_resident().copy()
    bench(make, ev, f"n={n:6d}      ")
for K, nn in ((16, 1500), (4, 3000)):
    X, y = synthetic_sine(K * nn, 10)

    def make():
        return cg.BCM(X, y, K=K, rank=0, world=1)

    def ev(b, th):
        b.set_BCM_log_hyperparam(th)
        ll, g = b.loglik_and_gradient()
        return ll, g.copy()
    bench(make, ev, f"BCM {K:2d} x {nn:5d}")
L.cugp_set_tuning(b"overlap_inv_max_n", 0)
