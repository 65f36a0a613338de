"""Cholesky / LL+grad timing over sizes with a tuning key A/B (device events for the factorisation)."""
import sys
import time

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200._lib import lib
from cugp_b200.loaders import synthetic_sine

L = lib()
key = sys.argv[1].encode()
vals = [int(v) for v in sys.argv[2].split(",")]
sizes = [int(v) for v in sys.argv[3].split(",")]
grad = len(sys.argv) > 4 and sys.argv[4] == "grad"
TH_B = [3.762111, -1.152105, -0.384461]
for n in sizes:
    X, y = synthetic_sine(n, 10)
    g = cg.Covsum(n, 10)
    g.set_data(X, y)
    for v in vals:
        assert L.cugp_set_tuning(key, v) == 0
        tc, ts = [], []
        for r in range(4 if n <= 20000 else 2):
            g.set_loghyperparam([TH_B[0] + 1e-7 * r, TH_B[1], TH_B[2]])
            tc.append(g.factorize_resident()[1])
        ll = g.loglik_resident()
        line = f"n={n:6d} {key.decode()}={v}: chol {min(tc):9.3f} ms = {n**3/3/min(tc)/1e9:6.2f} TF  LL={ll:.9f}"
        if grad:
            for r in range(4):
                g.set_loghyperparam([TH_B[0] + 1e-7 * (r + 10), TH_B[1], TH_B[2]])
                t = time.perf_counter()
                g.loglik_resident()
                g.grad_resident()
                ts.append(time.perf_counter() - t)
            line += f"  LL+grad {1e3 * min(ts[1:]):.3f} ms"
        print(line, flush=True)
    g.close()
