"""Covariance build (K1, lower triangle) time and HBM rate at a few sizes; cov_fast A/B."""
import sys

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200._lib import lib
from cugp_b200.loaders import synthetic_sine

for n in [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["10000", "40000"])]:
    X, y = synthetic_sine(n, 10)
    for fast in (0, 1):
        lib().cugp_set_tuning(b"cov_fast", fast)
        g = cg.Covsum(n, 10)
        g.set_data(X, y)
        best = 1e9
        for r in range(4):
            g.set_loghyperparam([3.762111 + 1e-7 * r, -1.152105, -0.384461])
            best = min(best, g.factorize_resident()[0])
        print(f"n={n} cov_fast={fast}: covariance {best:.3f} ms = {4.0*n*(n+1)/best/1e6:.0f} GB/s", flush=True)
        g.close()
lib().cugp_set_tuning(b"cov_fast", 1)
