"""Tiles-per-CTA sweep of the DMMA GEMM: probe shapes, then the whole Cholesky at a few sizes."""
import ctypes as C
import sys

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200._lib import lib
from cugp_b200.loaders import synthetic_sine

L = lib()
tpcs = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 2, 4, 8]
sizes = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [10000, 40000]
for tpc in tpcs:
    L.cugp_set_tuning(b"gemm_tpc", tpc)
    for M, N, K in [(8192, 8192, 1024), (32768, 32768, 1024)]:
        t = C.c_double()
        rc = L.cugp_probe_gemm(M, N, K, 5, C.byref(t))
        print(f"tpc={tpc} gemm {M}x{N}x{K}: {t.value:.2f} TF (rc={rc})", flush=True)
TH_B = [3.762111, -1.152105, -0.384461]
for n in sizes:
    X, y = synthetic_sine(n, 10)
    g = cg.Covsum(n, 10)
    g.set_data(X, y)
    for tpc in [0] + tpcs:
        L.cugp_set_tuning(b"gemm_tpc", tpc)
        best = None
        for r in range(3 if n <= 40000 else 1):
            g.set_loghyperparam([TH_B[0] + 1e-7 * r, TH_B[1], TH_B[2]])
            ms_cov, ms_chol = g.factorize_resident()
            best = ms_chol if best is None else min(best, ms_chol)
        print(f"n={n:6d} tpc={tpc}: chol {best:9.3f} ms = {n**3/3/best/1e9:6.2f} TF  LL={g.loglik_resident():.9f}", flush=True)
    g.close()
L.cugp_set_tuning(b"gemm_tpc", 0)
