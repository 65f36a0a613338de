"""Round-2 GPU check of the fused Cholesky block step (cholstep.cu): numerics against numpy / the round-1 launch chain,
then timings.  Writes everything to stdout; run under gpurun."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200._lib import check, lib, ptr
from cugp_b200.loaders import synthetic_sine

L = lib()
TH_B = [3.762111, -1.152105, -0.384461]
rng = np.random.default_rng(5)


def chol(A):
    n = A.shape[0]
    out = np.empty((n, n))
    check(L.cugp_cholesky(ptr(A), ptr(out), n))
    return out


print("== cugp_cholesky vs numpy (fused step)")
for n in [64, 92, 128, 129, 200, 256, 300, 500, 1000, 1500, 2048, 3000]:
    M = rng.standard_normal((n, n + 8))
    A = M @ M.T / n + np.eye(n)
    L0 = np.linalg.cholesky(A)
    for fused in (1, 0):
        L.cugp_set_tuning(b"fused_step", fused)
        try:
            L1 = chol(A)
            err = np.linalg.norm(L1 - L0) / np.linalg.norm(L0)
            print(f"n={n:5d} fused={fused} rel fro err {err:.3e} max abs {np.abs(L1 - L0).max():.3e} nan={int(np.isnan(L1).sum())}", flush=True)
            if fused and not err < 1e-12:
                bad = np.argwhere(~(np.abs(L1 - L0) < 1e-9))
                print("   first bad entries:", bad[:6].tolist(), " bad count", len(bad), " rows", sorted(set(bad[:, 0] // 32))[:12],
                      "cols", sorted(set(bad[:, 1] // 32))[:12])
        except Exception as e:
            print(f"n={n} fused={fused} FAILED {e!r}", flush=True)
L.cugp_set_tuning(b"fused_step", 1)

print("== LL / gradient / alpha, fused vs chain")
for n in [300, 1500, 2048, 4096]:
    X, y = synthetic_sine(n, 10)
    res = {}
    for fused in (1, 0):
        L.cugp_set_tuning(b"fused_step", fused)
        g = cg.Covsum(n, 10)
        g.set_data(X, y)
        g.set_loghyperparam(TH_B)
        ll = g.loglik_resident()
        gr = g.grad_resident()
        g.set_loghyperparam([TH_B[0] + 1e-9, TH_B[1], TH_B[2]])
        ll2 = g.loglik_resident()
        res[fused] = (ll, gr, ll2)
        g.close()
    d = abs(res[1][0] - res[0][0]) / abs(res[0][0])
    dg = np.abs(res[1][1] - res[0][1]).max() / np.abs(res[0][1]).max()
    print(f"n={n}: LL fused {res[1][0]:.9f} chain {res[0][0]:.9f} rel {d:.2e} | grad rel {dg:.2e} | 2nd eval {res[1][2]:.9f}", flush=True)
L.cugp_set_tuning(b"fused_step", 1)

print("== timings (device events inside the library)")
for n in [1500, 2048, 3000, 4096]:
    X, y = synthetic_sine(n, 10)
    g = cg.Covsum(n, 10)
    g.set_data(X, y)
    for fused in (0, 1):
        L.cugp_set_tuning(b"fused_step", fused)
        for la in (1, 0):
            L.cugp_set_tuning(b"lookahead", la)
            best = 1e9
            for r in range(5):
                g.set_loghyperparam([TH_B[0] + 1e-7 * r, TH_B[1], TH_B[2]])
                ms_cov, ms_chol = g.factorize_resident()
                best = min(best, ms_chol)
            # LL + gradient wall time
            ts = []
            for r in range(6):
                g.set_loghyperparam([TH_B[0] + 1e-7 * (r + 10), TH_B[1], TH_B[2]])
                t = time.perf_counter()
                g.loglik_resident()
                g.grad_resident()
                ts.append(time.perf_counter() - t)
            print(f"n={n:5d} fused={fused} la={la}: chol {best:8.3f} ms = {n**3/3/best/1e9:6.2f} TF   LL+grad {1e3*min(ts[2:]):7.3f} ms", flush=True)
    L.cugp_set_tuning(b"lookahead", 1)
    L.cugp_set_tuning(b"fused_step", 1)
    g.close()

print("== BCM 2 x 1500 and 16 x 1500 LL+grad")
d = np.load("tests/golden/data_si24000.npz")
for K, rows in ((2, 3000), (16, 24000)):
    for fused in (0, 1):
        L.cugp_set_tuning(b"fused_step", fused)
        b = cg.BCM(d["X"][:rows], d["y"][:rows], K=K, rank=0, world=1)
        ts = []
        for r in range(6):
            b.set_BCM_log_hyperparam([2.0 + 1e-7 * r, 2.0, 2.0])
            t = time.perf_counter()
            ll, gr = b.loglik_and_gradient()
            ts.append(time.perf_counter() - t)
        print(f"BCM {K} x 1500 fused={fused}: LL+grad {1e3*min(ts[2:]):.3f} ms  LL={ll:.9f} g={gr}", flush=True)
        b.close()
L.cugp_set_tuning(b"fused_step", 1)
