"""Host-side mirror of the reference's ``class Covsum`` (cpp_serial_gp/covkernel.h:3-38) and of the
``matrixops`` free functions (common/matrixops.h:5-25) over the C ABI.  Method names, argument meaning,
sign conventions and error behaviour are the reference's; arrays are NumPy float64 instead of ``double**``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, f64, lib, ptr


class Covsum:
    """SE + noise exact GP with theta = (log ell, log sigma_f, log sigma_n)."""

    def __init__(self, n: int, d: int):  # Covsum::Covsum(int n, int d), covkernel.cpp:13
        self.inputdatasize, self.numdim = int(n), int(d)
        self._h = C.c_void_p()
        check(lib().cugp_covsum_create(self.inputdatasize, self.numdim, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().cugp_covsum_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _X(self, X):
        X = f64(X)
        if X.shape != (self.inputdatasize, self.numdim):
            raise ValueError(f"X must be {self.inputdatasize} x {self.numdim}, got {X.shape}")
        return X

    def _y(self, y):
        y = f64(y)
        if y.shape != (self.inputdatasize,):
            raise ValueError(f"y must have {self.inputdatasize} entries, got {y.shape}")
        return y

    # -- hyper-parameters (covkernel.cpp:266-274, 308-312) ----------------------------------------------
    def set_loghyperparam(self, theta):
        th = f64(theta)
        assert th.shape == (3,)
        check(lib().cugp_covsum_set_loghyper(self._h, ptr(th)))

    set_loghyper_eigen = set_loghyperparam

    def get_loghyperparam(self):
        th = np.empty(3)
        check(lib().cugp_covsum_get_loghyper(self._h, ptr(th)))
        return th

    def get_param_dim(self):  # covkernel.cpp:640-642 returns numdim
        return self.numdim

    # -- covariance (covkernel.cpp:64-116) ----------------------------------------------------------------
    def compute_K_train(self, X):
        X = self._X(X)
        K = np.empty((self.inputdatasize, self.inputdatasize))
        check(lib().cugp_covsum_K_train(self._h, ptr(X), ptr(K)))
        return K

    def compute_k_test(self, X, xtest):
        X, xt = self._X(X), f64(xtest)
        out = np.empty(self.inputdatasize)
        check(lib().cugp_covsum_k_test(self._h, ptr(X), ptr(xt), ptr(out)))
        return out

    # -- log-likelihood and gradient (covkernel.cpp:118-129, 162-263) ----------------------------------------
    def compute_loglikelihood(self, X, y):
        X, y = self._X(X), self._y(y)
        ll = C.c_double()
        check(lib().cugp_covsum_loglik(self._h, ptr(X), ptr(y), C.byref(ll)))
        return ll.value

    def compute_gradient_loghyperparam(self, X, y):
        """d(-LL)/dtheta -- the reference's sign (covkernel.cpp:221, 259-261)."""
        X, y = self._X(X), self._y(y)
        g = np.empty(3)
        check(lib().cugp_covsum_grad(self._h, ptr(X), ptr(y), ptr(g)))
        return g

    # -- prediction (covkernel.cpp:277-306, 629-638) ------------------------------------------------------------
    def compute_test_means_and_variances(self, X, y, Xtest):
        X, y, Xt = self._X(X), self._y(y), f64(Xtest).reshape(-1, self.numdim)
        m = Xt.shape[0]
        mean, var = np.empty(m), np.empty(m)
        if m:
            check(lib().cugp_covsum_predict(self._h, ptr(X), ptr(y), ptr(Xt), m, ptr(mean), ptr(var)))
        return mean, var

    @staticmethod
    def get_negative_log_predprob(actual, predmean, predvar):
        a, mu, v = f64(actual), f64(predmean), f64(predvar)
        out = C.c_double()
        check(lib().cugp_nlpp(ptr(a), ptr(mu), ptr(v), a.shape[0], C.byref(out)))
        return out.value

    # -- optimisers (covkernel.cpp:320-627) ------------------------------------------------------------------
    def cg_solve(self, X, y, verbose: bool = False, trace_cap: int = 256):
        """Returns the trace of f = -LL at every trial point; theta of the object is updated."""
        X, y = self._X(X), self._y(y)
        tr = np.full(trace_cap, np.nan)
        ne = C.c_int()
        check(lib().cugp_covsum_cg_solve(self._h, ptr(X), ptr(y), ptr(tr), trace_cap, C.byref(ne)))
        tr = tr[: min(ne.value, trace_cap)]
        if verbose:
            for f in tr:
                print(f)
        return tr

    def rprop_solve(self, X, y, verbose: bool = False):
        X, y = self._X(X), self._y(y)
        check(lib().cugp_covsum_rprop_solve(self._h, ptr(X), ptr(y)))

    # -- device-resident evaluation (upload once, evaluate many thetas) ---------------------------------------
    def set_data(self, X, y):
        X, y = self._X(X), self._y(y)
        check(lib().cugp_covsum_set_data(self._h, ptr(X), ptr(y)))

    def loglik_resident(self):
        ll = C.c_double()
        check(lib().cugp_covsum_loglik_resident(self._h, C.byref(ll)))
        return ll.value

    def grad_resident(self):
        g = np.empty(3)
        check(lib().cugp_covsum_grad_resident(self._h, ptr(g)))
        return g

    def scalars_resident(self):
        """(y'K^-1 y, logdet K, LL)."""
        s = np.empty(3)
        check(lib().cugp_covsum_scalars_resident(self._h, ptr(s)))
        return s

    def alpha_resident(self):
        a = np.empty(self.inputdatasize)
        check(lib().cugp_covsum_alpha_resident(self._h, ptr(a)))
        return a

    def residual_resident(self):
        """r = (K + sn2 I) alpha - y, K rebuilt on the fly: ||r|| / ||y|| checks build + Cholesky + solves at any n."""
        r = np.empty(self.inputdatasize)
        check(lib().cugp_covsum_residual_resident(self._h, ptr(r)))
        return r

    def factorize_resident(self):
        """Covariance build + Cholesky only; returns (ms_cov, ms_chol) measured with CUDA events."""
        a, b = C.c_float(), C.c_float()
        check(lib().cugp_covsum_factorize_resident(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value


    def solve_resident(self):
        """Device ms of the two triangular sweeps + log-det + LL on the cached factor."""
        a = C.c_float()
        check(lib().cugp_covsum_solve_resident(self._h, C.byref(a)))
        return a.value

    def profile(self, enable: bool = True):
        check(lib().cugp_covsum_profile(self._h, int(enable)))

    def profile_read(self):
        """(summed SYRK launch ms, their algorithmic flops, launch count) since profile(True)."""
        ms, fl, cnt = C.c_double(), C.c_double(), C.c_long()
        check(lib().cugp_covsum_profile_read(self._h, C.byref(ms), C.byref(fl), C.byref(cnt)))
        return ms.value, fl.value, cnt.value


# ---- matrixops (common/matrixops.h:5-25) ---------------------------------------------------------------------
def get_cholesky(A):
    """matrixops.cpp:68-108: dense L with zeroed upper triangle; NaN (no exception) on a negative pivot."""
    A = f64(A)
    n = A.shape[0]
    L = np.empty((n, n))
    check(lib().cugp_cholesky(ptr(A), ptr(L), n))
    return L


def compute_chol_and_det(K, y):
    """matrixops.cpp:232-234: (y' K^-1 y, log det K)."""
    K, y = f64(K), f64(y)
    q, ld = C.c_double(), C.c_double()
    check(lib().cugp_chol_and_det(ptr(K), ptr(y), K.shape[0], C.byref(q), C.byref(ld)))
    return q.value, ld.value


def vector_Kinvy_using_cholesky(K, y):
    K, y = f64(K), f64(y)
    a = np.empty(K.shape[0])
    check(lib().cugp_kinv_y(ptr(K), ptr(y), ptr(a), K.shape[0]))
    return a


def compute_K_inverse(K):
    K = f64(K)
    out = np.empty_like(K)
    check(lib().cugp_k_inverse(ptr(K), ptr(out), K.shape[0]))
    return out
