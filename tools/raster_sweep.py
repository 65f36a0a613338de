"""Raster strip width sweep of the DMMA GEMM (probe shape + Cholesky), and graph replay on/off timings of one
(LL, gradient) evaluation at small and medium n."""
import ctypes as C
import sys
import time

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200._lib import lib
from cugp_b200.loaders import synthetic_sine

L = lib()
TH_B = [3.762111, -1.152105, -0.384461]
what = sys.argv[1] if len(sys.argv) > 1 else "raster,graph"
if "raster" in what:
    for rw in (8, 16, 32):
        L.cugp_set_tuning(b"gemm_raster", rw)
        t = C.c_double()
        L.cugp_probe_gemm(32768, 32768, 1024, 5, C.byref(t))
        print(f"raster={rw} gemm 32768^2x1024: {t.value:.2f} TF", flush=True)
    n = 40000
    X, y = synthetic_sine(n, 10)
    g = cg.Covsum(n, 10)
    g.set_data(X, y)
    for rw in (8, 16, 32):
        L.cugp_set_tuning(b"gemm_raster", rw)
        best = None
        for r in range(3):
            g.set_loghyperparam([TH_B[0] + 1e-7 * r, TH_B[1], TH_B[2]])
            ms_cov, ms_chol = g.factorize_resident()
            best = ms_chol if best is None else min(best, ms_chol)
        print(f"n={n} raster={rw}: chol {best:.3f} ms = {n**3/3/best/1e9:.2f} TF LL={g.loglik_resident():.9f}", flush=True)
    g.close()
    L.cugp_set_tuning(b"gemm_raster", 0)
if "graph" in what:
    for n in (1500, 4096, 10000):
        X, y = synthetic_sine(n, 10)
        g = cg.Covsum(n, 10)
        g.set_data(X, y)
        for mode, mx in (("direct", 0), ("graph", 16384)):
            L.cugp_set_tuning(b"graph_max_n", mx)
            ts, tl = [], []
            for r in range(8):
                g.set_loghyperparam([TH_B[0] + 1e-7 * r, TH_B[1], TH_B[2]])
                t0 = time.perf_counter()
                ll = g.loglik_resident()
                t1 = time.perf_counter()
                gr = g.grad_resident()
                ts.append(time.perf_counter() - t0)
                tl.append(t1 - t0)
            print(f"n={n} {mode}: LL+grad {1e3*sorted(ts[3:])[2]:.3f} ms (LL alone {1e3*sorted(tl[3:])[2]:.3f} ms)  LL={ll:.9f}", flush=True)
        g.close()
