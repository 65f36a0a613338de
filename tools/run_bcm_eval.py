"""One (LL, gradient) evaluation + prediction of a BCM ensemble on synthetic data: the command ncu launch lists of
the C4-shaped path are taken from.  usage: run_bcm_eval.py [experts] [rows_per_expert] [test_points] [reps] [world]
world > 1 times rank 0's share only (experts 0, world, 2*world, ...; no exchange): the per-GPU work of a W-GPU run."""
import sys
import time

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200.loaders import synthetic_sine

K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
m = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
world = int(sys.argv[5]) if len(sys.argv) > 5 else 1
TH_C = [2.0, 2.0, 2.0]
X, y = synthetic_sine(K * n + m, 10)
b = cg.BCM(X[:K * n], y[:K * n], K=K, rank=0, world=1) if world == 1 else None
if world > 1:
    import numpy as np
    from cugp_b200.bcm import _CudaLocal

    class _Rank0:
        def __init__(self):
            self.l = _CudaLocal(np.ascontiguousarray(X[:K * n]), np.ascontiguousarray(y[:K * n]), K, 0, world)

        def set_BCM_log_hyperparam(self, th):
            self.l.set_theta(np.array(th, dtype=float))

        def loglik_and_gradient(self):
            o = self.l.loglik_grad(True)
            return o[0], o[1:]

        def compute_BCM_test_means_and_var(self, Xt):
            PQ = self.l.moments(np.ascontiguousarray(Xt))
            return PQ[1] / PQ[0], 1.0 / PQ[0]
    b = _Rank0()
for r in range(reps):
    b.set_BCM_log_hyperparam([TH_C[0] + 1e-6 * r, TH_C[1], TH_C[2]])
    t0 = time.perf_counter()
    ll, g = b.loglik_and_gradient()
    t1 = time.perf_counter()
    mu, var = b.compute_BCM_test_means_and_var(X[K * n:])
    t2 = time.perf_counter()
    print(f"K={K} n={n} m={m} world={world}: eval {1e3*(t1-t0):.3f} ms  predict {1e3*(t2-t1):.3f} ms  LL={ll:.6f}", flush=True)
