"""Timing of the phases on device-resident data: covariance, Cholesky (+ fused forward substitution), backward sweep."""
import sys

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200.loaders import synthetic_sine

TH_B = [3.762111, -1.152105, -0.384461]
for n in [int(a) for a in sys.argv[1].split(",")]:
    X, y = synthetic_sine(n, 10)
    g = cg.Covsum(n, 10)
    g.set_data(X, y)
    best = None
    for r in range(3):
        g.set_loghyperparam([TH_B[0] + 1e-7 * r, TH_B[1], TH_B[2]])
        ms_cov, ms_chol = g.factorize_resident()
        ms_solve = g.solve_resident()
        best = ms_solve if best is None else min(best, ms_solve)
    tri = 4.0 * n * (n + 1)
    print(f"n={n}: cov {ms_cov:.3f} ms ({tri/ms_cov/1e6:.0f} GB/s)  chol {ms_chol:.2f} ms ({n**3/3/ms_chol/1e9:.2f} TF)  "
          f"backward sweep {best:.3f} ms ({tri/best/1e6:.0f} GB/s)  LL={g.loglik_resident():.9f}", flush=True)
    g.close()
