"""Covariance build + Cholesky (+ solves) once or twice on synthetic data: the command for ncu captures."""
import sys

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200.loaders import synthetic_sine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
TH_B = [3.762111, -1.152105, -0.384461]
X, y = synthetic_sine(n, 10)
g = cg.Covsum(n, 10)
g.set_data(X, y)
for r in range(reps):
    g.set_loghyperparam([TH_B[0] + 1e-6 * r, TH_B[1], TH_B[2]])
    ms_cov, ms_chol = g.factorize_resident()
    print(f"n={n} cov {ms_cov:.3f} ms chol {ms_chol:.3f} ms = {n**3/3/ms_chol/1e9:.2f} TF, LL={g.loglik_resident():.9f}", flush=True)
