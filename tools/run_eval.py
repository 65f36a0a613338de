"""One (LL, gradient) evaluation + a small prediction on synthetic data: the command ncu launch lists are taken from."""
import sys

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200.loaders import synthetic_sine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
TH_B = [3.762111, -1.152105, -0.384461]
X, y = synthetic_sine(n + 256, 10)
g = cg.Covsum(n, 10)
g.set_data(X[:n], y[:n])
for r in range(reps):
    g.set_loghyperparam([TH_B[0] + 1e-6 * r, TH_B[1], TH_B[2]])
    print("LL", g.loglik_resident(), "grad", g.grad_resident())
mu, var = g.compute_test_means_and_variances(X[:n], y[:n], X[n:])
print("pred", mu[:2], var[:2])
