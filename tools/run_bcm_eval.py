"""One (LL, gradient) evaluation + prediction of a BCM ensemble on synthetic data: the command ncu launch lists of
the C4-shaped path are taken from.  usage: run_bcm_eval.py [experts] [rows_per_expert] [test_points] [reps]"""
import sys
import time

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200.loaders import synthetic_sine

K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
m = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
TH_C = [2.0, 2.0, 2.0]
X, y = synthetic_sine(K * n + m, 10)
b = cg.BCM(X[:K * n], y[:K * n], K=K, rank=0, world=1)
for r in range(reps):
    b.set_BCM_log_hyperparam([TH_C[0] + 1e-6 * r, TH_C[1], TH_C[2]])
    t0 = time.perf_counter()
    ll, g = b.loglik_and_gradient()
    t1 = time.perf_counter()
    mu, var = b.compute_BCM_test_means_and_var(X[K * n:])
    t2 = time.perf_counter()
    print(f"K={K} n={n} m={m}: eval {1e3*(t1-t0):.3f} ms  predict {1e3*(t2-t1):.3f} ms  LL={ll:.6f}", flush=True)
