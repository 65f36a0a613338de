"""ctypes doorway onto the CPU checker.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module (see ``oracle/cugp_oracle.h``).  It exposes two libraries with one
interface:

* ``port()``      -> ``oracle/liboracle.so``       the C restatement (``oracle/cugp_oracle.c``)
* ``reference()`` -> ``oracle/_ref/libcugp_ref.so`` the UNMODIFIED reference objects (built by
  ``make -C oracle ref`` where ``/root/reference`` exists; ``None`` when the file is absent)

All arrays are C-contiguous float64; theta = (log ell, log sigma_f, log sigma_n).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class CpuGP:
    """One CPU implementation (restatement or reference) behind the flat-array interface."""

    def __init__(self, path: str, prefix: str, kind: str):
        self.kind = kind  # "port" | "reference"
        self.path = path
        self._l = C.CDLL(path)
        self._pre = prefix
        d, i, v = C.c_double, C.c_int, None
        sig = {
            "K_train": (v, [_dp, i, i, _dp, _dp]),
            "k_test": (v, [_dp, i, i, _dp, _dp, _dp]),
            "cholesky": (v, [_dp, i, _dp]),
            "chol_and_det": (v, [_dp, _dp, i, _dp, _dp]),
            "kinv_y": (v, [_dp, _dp, i, _dp]),
            "k_inverse": (v, [_dp, i, _dp]),
            "loglik": (d, [_dp, _dp, i, i, _dp]),
            "grad": (v, [_dp, _dp, i, i, _dp, _dp]),
            "predict": (v, [_dp, _dp, i, i, _dp, _dp, i, _dp, _dp]),
            "nlpp": (d, [_dp, _dp, _dp, i]),
            "bcm_loglik": (d, [_dp, _dp, i, i, i, _dp]),
            "bcm_grad": (v, [_dp, _dp, i, i, i, _dp, _dp]),
            "bcm_predict": (v, [_dp, _dp, i, i, i, _dp, _dp, i, _dp, _dp]),
            "cg_solve": (i, [_dp, _dp, i, i, i, _dp, _dp, i]),
            "rprop_solve": (i, [_dp, _dp, i, i, _dp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(self._l, prefix + name)
            fn.restype = res
            fn.argtypes = args
            setattr(self, "_" + name, fn)

    # -- Covsum -------------------------------------------------------------------------------
    def K_train(self, X, theta):
        X, th = _f64(X), _f64(theta)
        n, d = X.shape
        K = np.empty((n, n))
        self._K_train(_p(X), n, d, _p(th), _p(K))
        return K

    def k_test(self, X, theta, xt):
        X, th, xt = _f64(X), _f64(theta), _f64(xt)
        n, d = X.shape
        out = np.empty(n)
        self._k_test(_p(X), n, d, _p(th), _p(xt), _p(out))
        return out

    def cholesky(self, A):
        A = _f64(A)
        n = A.shape[0]
        L = np.empty((n, n))
        self._cholesky(_p(A), n, _p(L))
        return L

    def chol_and_det(self, K, y):
        K, y = _f64(K), _f64(y)
        q, ld = C.c_double(), C.c_double()
        self._chol_and_det(_p(K), _p(y), K.shape[0], C.byref(q), C.byref(ld))
        return q.value, ld.value

    def kinv_y(self, K, y):
        K, y = _f64(K), _f64(y)
        a = np.empty(K.shape[0])
        self._kinv_y(_p(K), _p(y), K.shape[0], _p(a))
        return a

    def k_inverse(self, K):
        K = _f64(K)
        out = np.empty_like(K)
        self._k_inverse(_p(K), K.shape[0], _p(out))
        return out

    def loglik(self, X, y, theta):
        X, y, th = _f64(X), _f64(y), _f64(theta)
        return float(self._loglik(_p(X), _p(y), X.shape[0], X.shape[1], _p(th)))

    def grad(self, X, y, theta):
        X, y, th = _f64(X), _f64(y), _f64(theta)
        g = np.empty(3)
        self._grad(_p(X), _p(y), X.shape[0], X.shape[1], _p(th), _p(g))
        return g

    def predict(self, X, y, theta, Xt):
        X, y, th, Xt = _f64(X), _f64(y), _f64(theta), _f64(Xt)
        m = Xt.shape[0]
        mean, var = np.empty(m), np.empty(m)
        self._predict(_p(X), _p(y), X.shape[0], X.shape[1], _p(th), _p(Xt), m, _p(mean), _p(var))
        return mean, var

    def nlpp(self, actual, mean, var):
        a, mu, v = _f64(actual), _f64(mean), _f64(var)
        return float(self._nlpp(_p(a), _p(mu), _p(v), a.shape[0]))

    # -- BCM ----------------------------------------------------------------------------------
    def bcm_loglik(self, X, y, K, theta):
        X, y, th = _f64(X), _f64(y), _f64(theta)
        return float(self._bcm_loglik(_p(X), _p(y), X.shape[0], X.shape[1], int(K), _p(th)))

    def bcm_grad(self, X, y, K, theta):
        X, y, th = _f64(X), _f64(y), _f64(theta)
        g = np.empty(3)
        self._bcm_grad(_p(X), _p(y), X.shape[0], X.shape[1], int(K), _p(th), _p(g))
        return g

    def bcm_predict(self, X, y, K, theta, Xt):
        X, y, th, Xt = _f64(X), _f64(y), _f64(theta), _f64(Xt)
        m = Xt.shape[0]
        mean, var = np.empty(m), np.empty(m)
        self._bcm_predict(_p(X), _p(y), X.shape[0], X.shape[1], int(K), _p(th), _p(Xt), m, _p(mean), _p(var))
        return mean, var

    # -- optimisers ---------------------------------------------------------------------------
    def cg_solve(self, X, y, theta, K=0, trace_cap=256):
        """Returns (theta*, n_evals, f_trace).  n_evals is -1 and the trace empty for the reference."""
        X, y = _f64(X), _f64(y)
        th = _f64(theta).copy()
        tr = np.full(trace_cap, np.nan)
        ne = self._cg_solve(_p(X), _p(y), X.shape[0], X.shape[1], int(K), _p(th), _p(tr), trace_cap)
        return th, ne, tr[: max(ne, 0)]

    def rprop_solve(self, X, y, theta):
        X, y = _f64(X), _f64(y)
        th = _f64(theta).copy()
        self._rprop_solve(_p(X), _p(y), X.shape[0], X.shape[1], _p(th))
        return th


_PORT = None
_REF = None


def build(ref: bool = True) -> None:
    """Compile the checker (``make -C oracle``; plus ``make ref`` where /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    if ref and os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", _HERE, "-s", "ref"], check=True)
        subprocess.run(["make", "-C", _HERE, "-s", "refgpu"], check=True)


def port() -> CpuGP:
    global _PORT
    if _PORT is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _PORT = CpuGP(path, "oracle_", "port")
    return _PORT


def reference_gpu_path():
    """The reference's own cuSOLVER / cuBLAS GPU variant (cuda_bettersinglenode_ver2/cuda_gp.cu, unchanged, sm_100), or
    None when it has not been built.  bench.py's `library_baseline` leg is its only user."""
    path = os.path.join(_HERE, "_ref", "libcugp_refgpu.so")
    return path if os.path.exists(path) else None


def reference():
    """The unmodified reference, or None when oracle/_ref/libcugp_ref.so has not been built."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libcugp_ref.so")
        if not os.path.exists(path):
            return None
        _REF = CpuGP(path, "ref_", "reference")
    return _REF
