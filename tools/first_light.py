"""GPU bring-up script: exercises the C ABI piece by piece, printing what it sees (run under gpurun)."""
import ctypes as C
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200._lib import lib, ptr
from cugp_b200.loaders import synthetic_sine
from oracle import oracle

PORT = oracle.port()
TH_B = [3.762111, -1.152105, -0.384461]


def step(name, fn):
    t = time.time()
    try:
        out = fn()
        print(f"[ok]   {name}: {out}  ({time.time() - t:.2f}s)", flush=True)
    except Exception:
        print(f"[FAIL] {name}", flush=True)
        traceback.print_exc()


def probes():
    a, b = C.c_double(), C.c_double()
    rc = lib().cugp_probe_fp64_peak(300.0, C.byref(a), C.byref(b))
    g = C.c_double()
    lib().cugp_probe_copy(1 << 30, 10, C.byref(g))
    return f"rc={rc} dmma={a.value:.2f} TF dfma={b.value:.2f} TF copy={g.value:.0f} GB/s"


def gemm_probe():
    out = []
    for M, N, K in ((8192, 8192, 128), (8192, 8192, 256), (8192, 8192, 1024), (16384, 16384, 128)):
        t = C.c_double()
        rc = lib().cugp_probe_gemm(M, N, K, 5, C.byref(t))
        out.append(f"{M}x{N}x{K}: {t.value:.2f} TF (rc={rc})")
    return "; ".join(out)


def small_gemm():
    rng = np.random.default_rng(0)
    A, B, C0 = rng.standard_normal((200, 170)), rng.standard_normal((150, 170)), rng.standard_normal((200, 150))
    res = []
    for cfg in (0, 1, 2):
        out = C0.copy()
        rc = lib().cugp_debug_gemm(ptr(A), ptr(B), ptr(out), 200, 150, 170, -1.0, 1.0, 1, 1, 0, cfg, None)
        res.append((rc, float(np.abs(out - (C0 - A @ B.T)).max())))
    return res


def chol(n):
    rng = np.random.default_rng(n)
    X = rng.uniform(-3, 3, (n, 4))
    K = PORT.K_train(X, [0.7, 0.3, -1.0])
    L = cg.get_cholesky(K)
    Lr = np.linalg.cholesky(K)
    return float(np.linalg.norm(L - Lr) / np.linalg.norm(Lr))


def covsum(n):
    X, y = synthetic_sine(n + 8, 10)
    Xt, X, y = X[n:], X[:n], y[:n]
    g = cg.Covsum(n, 10)
    g.set_loghyperparam(TH_B)
    ll = g.compute_loglikelihood(X, y)
    gr = g.compute_gradient_loghyperparam(X, y)
    mu, var = g.compute_test_means_and_variances(X, y, Xt)
    if n <= 600:
        return ll, PORT.loglik(X, y, TH_B), gr, PORT.grad(X, y, TH_B), mu[:2], var[:2], PORT.predict(X, y, TH_B, Xt[:2])
    return ll, gr, mu[:2], var[:2]


def timing(n):
    X, y = synthetic_sine(n, 10)
    g = cg.Covsum(n, 10)
    g.set_loghyperparam(TH_B)
    g.set_data(X, y)
    g.factorize_resident()
    best = None
    for _ in range(3):
        a, b = g.factorize_resident()
        best = (a, b) if best is None or b < best[1] else best
    tf = n ** 3 / 3 / (best[1] * 1e-3) / 1e12
    t0 = time.time()
    g.set_loghyperparam([TH_B[0] + 1e-9, TH_B[1], TH_B[2]])
    ll = g.loglik_resident()
    t1 = time.time()
    gr = g.grad_resident()
    t2 = time.time()
    return f"n={n}: cov {best[0]:.3f} ms, chol {best[1]:.3f} ms = {tf:.2f} TFLOP/s; LL {1e3 * (t1 - t0):.1f} ms, +grad {1e3 * (t2 - t1):.1f} ms, ll={ll:.6f}"


if __name__ == "__main__":
    print(lib().cugp_version())
    step("small gemm", small_gemm)
    step("probes", probes)
    step("gemm probe", gemm_probe)
    for n in (64, 128, 200, 1000):
        step(f"cholesky n={n}", lambda n=n: chol(n))
    for n in (100, 500, 2000):
        step(f"covsum n={n}", lambda n=n: covsum(n))
    for n in (1024, 4096, 10000, 20000):
        step(f"timing n={n}", lambda n=n: timing(n))
