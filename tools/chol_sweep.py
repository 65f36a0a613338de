"""Cholesky timing sweep on the GPU box: outer block width x diagonal-kernel variant x n (device-resident)."""
import ctypes as C
import sys

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200._lib import lib
from cugp_b200.loaders import synthetic_sine

TH_B = [3.762111, -1.152105, -0.384461]
sizes = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1500, 4096, 10000, 20000]
nbs = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [128, 256, 512, 1024]
variants = [int(a) for a in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1]
lookahead = [int(a) for a in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1]
L = lib()
for n in sizes:
    X, y = synthetic_sine(n, 10)
    g = cg.Covsum(n, 10)
    g.set_data(X, y)
    for dv in variants:
        for nb, la in [(nb, la) for nb in nbs for la in lookahead]:
            if nb > n:
                continue
            L.cugp_set_tuning(b"potrf_nb", nb)
            L.cugp_set_tuning(b"lookahead", la)
            best = None
            for r in range(3 if n <= 20000 else 2):
                g.set_loghyperparam([TH_B[0] + 1e-7 * r, TH_B[1], TH_B[2]])
                ms_cov, ms_chol = g.factorize_resident()
                best = ms_chol if best is None else min(best, ms_chol)
            ll = g.loglik_resident()
            print(f"n={n:6d} diag={dv} nb={nb:5d} la={la}: chol {best:9.3f} ms = {n**3/3/best/1e9:6.2f} TF  cov {ms_cov:7.3f} ms  LL={ll:.9f}", flush=True)
    g.close()
L.cugp_set_tuning(b"potrf_nb", 0)
