"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the committed golden vectors.

Tolerances are BASELINE.json's: log-likelihood and gradient 1e-9 relative, predictive mean/variance 1e-8
relative (conftest.assert_*); element-wise covariance to a few ulp (only exp() differs from glibc's)."""
import ctypes as C

import numpy as np
import pytest

import cugp_b200 as cg
from cugp_b200._lib import lib, ptr
from oracle import oracle
from tests.conftest import (assert_grad, assert_ll, assert_pred, case_inputs, load_data, load_golden)

pytestmark = pytest.mark.gpu
GOLD = load_golden()
PORT = oracle.port()
TH_B = [3.762111, -1.152105, -0.384461]


# ---------------------------------------------------------------------------------------------- kernel level
def _gemm(A, B, Cm, alpha, beta, a_kc, b_kc, flags, config, css=False):
    M, N = Cm.shape
    K = A.shape[1] if a_kc else A.shape[0]
    out = np.ascontiguousarray(Cm.copy())
    tm = -(-M // (128 if config == 0 else 64))
    S = np.zeros((tm, N)) if css else None
    rc = lib().cugp_debug_gemm(ptr(np.ascontiguousarray(A)), ptr(np.ascontiguousarray(B)), ptr(out), M, N, K,
                               alpha, beta, int(a_kc), int(b_kc), flags, config, ptr(S) if css else None)
    assert rc == 0, lib().cugp_last_error()
    return out, S


@pytest.mark.parametrize("config", [0, 1, 2])
@pytest.mark.parametrize("shape", [(200, 150, 170), (129, 77, 45), (64, 128, 128), (1, 1, 1), (300, 257, 19)])
def test_gemm_layouts(config, shape):
    M, N, K = shape
    rng = np.random.default_rng(M * 1000 + N)
    A, B, C0 = rng.standard_normal((M, K)), rng.standard_normal((N, K)), rng.standard_normal((M, N))
    ref = -1.0 * A @ B.T + 1.0 * C0
    for a_kc, b_kc in ((1, 1), (1, 0), (0, 0), (0, 1)):
        Ain = A if a_kc else A.T
        Bin = B if b_kc else B.T
        out, _ = _gemm(Ain, Bin, C0, -1.0, 1.0, a_kc, b_kc, 0, config)
        assert np.allclose(out, ref, rtol=1e-13, atol=1e-12), (a_kc, b_kc, np.abs(out - ref).max())
    out, _ = _gemm(A, B, np.full((M, N), np.nan), 2.0, 0.0, 1, 1, 0, config)  # beta == 0 must not read C
    assert np.allclose(out, 2.0 * A @ B.T, rtol=1e-13, atol=1e-12)


@pytest.mark.parametrize("config", [0, 2])
def test_gemm_triangular_ranges(config):
    """The k-range flags must be exact for triangular operands (TRTRI / LAUUM / predictive variance)."""
    rng = np.random.default_rng(3)
    n, N = 333, 210
    T = np.tril(rng.standard_normal((n, n)))
    Bf = rng.standard_normal((n, N))
    # khi_ti: A lower triangular [M][K], B [K][N] row-contig           (T21 = -T22 * tmp)
    out, _ = _gemm(T, Bf, np.zeros((n, N)), -1.0, 0.0, 1, 0, 1 << 3, config)
    assert np.allclose(out, -T @ Bf, rtol=1e-13, atol=1e-12)
    # klo_tj: B lower triangular stored [K][N], A full [M][K]            (tmp = L21 * T11)
    Af = rng.standard_normal((N, n))
    out, _ = _gemm(Af, T, np.zeros((N, n)), 1.0, 0.0, 1, 0, 1 << 2, config)
    assert np.allclose(out, Af @ T, rtol=1e-13, atol=1e-12)
    # LAUUM: C(lower) = T^T T with A = B = T stored [K][M]; lower tiles, k >= ti*BM
    out, _ = _gemm(T, T, np.zeros((n, n)), 1.0, 0.0, 0, 0, 1 | (1 << 1), config)
    assert np.allclose(np.tril(out), np.tril(T.T @ T), rtol=1e-13, atol=1e-12)
    # SYRK: lower tiles of C - P P^T
    P = rng.standard_normal((n, 128))
    C0 = rng.standard_normal((n, n))
    out, _ = _gemm(P, P, C0, -1.0, 1.0, 1, 1, 1, config)
    assert np.allclose(np.tril(out), np.tril(C0 - P @ P.T), rtol=1e-13, atol=1e-12)
    # lower trapezoid (panel update of the two-level Cholesky): M > N, tiles with ti >= tj only
    Pn = P[:N + 7]
    C1 = rng.standard_normal((n, N + 7))
    out, _ = _gemm(P, Pn, C1, -1.0, 1.0, 1, 1, 1, config)
    ref = C1 - P @ Pn.T
    bs = 128 if config == 0 else 64
    mask = (np.arange(n)[:, None] // bs) >= (np.arange(N + 7)[None, :] // bs)
    assert np.allclose(out[mask], ref[mask], rtol=1e-13, atol=1e-12)
    assert np.array_equal(out[~mask], C1[~mask])          # tiles above the diagonal are not touched
    # predictive variance epilogue: column sums of squares of T Kstar^T, per row tile
    Ks = rng.standard_normal((N, n))
    _, S = _gemm(T, Ks, np.zeros((n, N)), 1.0, 0.0, 1, 1, 1 << 3, config, css=True)
    assert np.allclose(S.sum(0), ((T @ Ks.T) ** 2).sum(0), rtol=1e-12)


# ---------------------------------------------------------------------------------------------- matrixops
def _spd(n, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-3, 3, (n, 4))
    return PORT.K_train(X, [0.7, 0.3, -1.0]), rng.standard_normal(n)


@pytest.mark.parametrize("n", [1, 2, 127, 128, 129, 300, 1000])
def test_matrixops_vs_oracle(n):
    K, y = _spd(n, n)
    L, Lo = cg.get_cholesky(K), PORT.cholesky(K)
    assert np.all(np.triu(L, 1) == 0.0)
    assert np.linalg.norm(L - Lo) <= 1e-12 * np.linalg.norm(Lo)
    q, ld = cg.compute_chol_and_det(K, y)
    qo, ldo = PORT.chol_and_det(K, y)
    assert abs(q - qo) <= 1e-10 * abs(qo) and abs(ld - ldo) <= 1e-11 * max(1.0, abs(ldo))
    a, ao = cg.vector_Kinvy_using_cholesky(K, y), PORT.kinv_y(K, y)
    assert np.linalg.norm(a - ao) <= 1e-10 * np.linalg.norm(ao)
    if n <= 300:
        Ki, Kio = cg.compute_K_inverse(K), PORT.k_inverse(K)
        assert np.linalg.norm(Ki - Kio) <= 1e-10 * np.linalg.norm(Kio)
        assert np.array_equal(Ki, Ki.T)


@pytest.mark.parametrize("n,nb", [(700, 256), (1000, 512), (1300, 384), (520, 1024)])
def test_two_level_cholesky(n, nb):
    """The outer block width only changes the order of the trailing updates: L must still match the oracle."""
    K, y = _spd(n, n + 1)
    Lo = PORT.cholesky(K)
    try:
        assert lib().cugp_set_tuning(b"potrf_nb", nb) == 0
        L = cg.get_cholesky(K)
        q, ld = cg.compute_chol_and_det(K, y)
    finally:
        lib().cugp_set_tuning(b"potrf_nb", 0)
    assert np.all(np.triu(L, 1) == 0.0)
    assert np.linalg.norm(L - Lo) <= 1e-12 * np.linalg.norm(Lo)
    qo, ldo = PORT.chol_and_det(K, y)
    assert abs(q - qo) <= 1e-10 * abs(qo) and abs(ld - ldo) <= 1e-11 * max(1.0, abs(ldo))
    assert lib().cugp_set_tuning(b"potrf_nb", 100) != 0 and lib().cugp_set_tuning(b"nope", 1) != 0


@pytest.mark.parametrize("n,nb", [(1000, 128), (1500, 256), (2100, 512)])
def test_lookahead_is_bitwise_neutral(n, nb):
    """Panel look-ahead only reorders independent launches across two streams: every C tile still receives the same
    updates in the same order, so L is bit-identical with and without it."""
    K, _ = _spd(n, 3 * n)
    try:
        lib().cugp_set_tuning(b"potrf_nb", nb)
        lib().cugp_set_tuning(b"lookahead", 0)
        L0 = cg.get_cholesky(K)
        lib().cugp_set_tuning(b"lookahead", 1)
        L1 = cg.get_cholesky(K)
    finally:
        lib().cugp_set_tuning(b"potrf_nb", 0)
        lib().cugp_set_tuning(b"lookahead", 1)
    assert np.array_equal(L0, L1)
    Lo = PORT.cholesky(K)
    assert np.linalg.norm(L1 - Lo) <= 1e-12 * np.linalg.norm(Lo)


@pytest.mark.parametrize("n", [700, 3300, 4200])
def test_backward_sweep_lookahead_is_bitwise_neutral(n):
    """alpha = L^-T z on the blocked sweep (no T at hand): with look-ahead the far part of every 1024-row panel update
    runs on the main stream while the next panel's block steps run on the chain stream; every entry of the work vector
    still receives its updates in the same order, so alpha is bit-identical -- and equals K^-1 y."""
    from cugp_b200.loaders import synthetic_sine
    X, y = synthetic_sine(n, 10)
    th = [3.762111, -1.152105, -0.384461]
    g = cg.Covsum(n, 10)
    g.set_data(X, y)
    out = {}
    try:
        for cl in (0, 1):        # panel chain: one launch per 128-row block / one thread-block-cluster launch per panel
            for la in (0, 1):
                lib().cugp_set_tuning(b"bwd_cluster", cl)
                lib().cugp_set_tuning(b"lookahead", la)
                g.set_loghyperparam([th[0] + 1e-9, th[1], th[2]])   # new theta: refactorise, no cached alpha
                g.set_loghyperparam(th)
                out[cl, la] = g.alpha_resident().copy()
    finally:
        lib().cugp_set_tuning(b"lookahead", 1)
        lib().cugp_set_tuning(b"bwd_cluster", 1)
    assert np.array_equal(out[0, 0], out[0, 1]) and np.array_equal(out[1, 0], out[1, 1])
    # the cluster kernel sums the 128x128 mat-vec in a different (fixed) order than the per-block kernel
    assert np.linalg.norm(out[1, 1] - out[0, 1]) <= 1e-11 * np.linalg.norm(out[0, 1])
    K = g.compute_K_train(X)
    for a in (out[0, 1], out[1, 1]):
        assert np.linalg.norm(K @ a - y) <= 1e-9 * np.linalg.norm(y)
    g.close()


@pytest.mark.parametrize("n,B", [(700, 1), (1500, 1), (600, 3)])
def test_graph_replay_is_bitwise_neutral(n, B):
    """The factorisation and the inverse chain (TRTRI, alpha, LAUUM) are replayed as CUDA graphs from the second
    evaluation on: same kernels, same arguments, same order per stream -- LL, gradient and predictions must not change
    by a bit against direct launches, across several thetas (the graphs are theta independent) and for a batch."""
    from cugp_b200.loaders import synthetic_sine
    X, y = synthetic_sine(n * B + 16, 10)
    Xt = X[n * B:]
    thetas = [[3.762111, -1.152105, -0.384461], [2.0, 2.0, 2.0], [0.882908, 0.098703, -2.971479], [3.0, 0.5, -1.0]]
    res = {}
    try:
        lib().cugp_set_tuning(b"overlap_inv_max_n", 0)     # graph replay in isolation: the inverse stays sequential
        for mode, max_n in (("direct", 0), ("graph", 16384)):
            lib().cugp_set_tuning(b"graph_max_n", max_n)
            out = []
            if B == 1:
                g = cg.Covsum(n, 10)
                g.set_data(X[:n], y[:n])
                for th in thetas:
                    g.set_loghyperparam(th)
                    out.append((g.loglik_resident(), g.grad_resident().copy()))
                mu, var = g.compute_test_means_and_variances(X[:n], y[:n], Xt)
                out.append((mu.copy(), var.copy()))
                g.close()
            else:
                b = cg.BCM(X[:n * B], y[:n * B], K=B, rank=0, world=1)
                for th in thetas:
                    b.set_BCM_log_hyperparam(th)
                    ll, gr = b.loglik_and_gradient()
                    out.append((ll, gr.copy()))
                mu, var = b.compute_BCM_test_means_and_var(Xt)
                out.append((mu.copy(), var.copy()))
                b.close()
            res[mode] = out
    finally:
        lib().cugp_set_tuning(b"graph_max_n", 2048)
        lib().cugp_set_tuning(b"overlap_inv_max_n", OVERLAP_DEFAULT)
    for a, b_ in zip(res["direct"], res["graph"]):
        assert np.array_equal(np.asarray(a[0]), np.asarray(b_[0])) and np.array_equal(a[1], b_[1])
    ref = PORT.loglik(X[:n], y[:n], thetas[1]) if B == 1 else PORT.bcm_loglik(X[:n * B], y[:n * B], B, thetas[1])
    assert_ll(res["graph"][1][0], ref)


@pytest.mark.parametrize("n,B", [(2300, 1), (3000, 2)])
def test_overlapped_inverse_matches_sequential(n, B):
    """T = L^-1 computed in row groups on a third stream while the factorisation advances (tuning overlap_inv_max_n)
    against the sequential recursive doubling: same K^-1 up to rounding, so gradient and predictions agree to 1e-11;
    repeated over thetas (every evaluation after the first takes the overlapped path) and capped / uncapped."""
    from cugp_b200.loaders import synthetic_sine
    X, y = synthetic_sine(n * B + 16, 10)
    Xt = X[n * B:]
    thetas = [[3.762111, -1.152105, -0.384461], [2.0, 2.0, 2.0], [3.0, 0.5, -1.0]]
    res = {}
    try:
        lib().cugp_set_tuning(b"idrows_max_n", 0)   # (below 3500 the identity rows would take over: not the path under test)
        for mode, max_n, cap in (("seq", 0, 64), ("ovl64", 8192, 64), ("ovl8", 8192, 8)):
            lib().cugp_set_tuning(b"overlap_inv_max_n", max_n)
            lib().cugp_set_tuning(b"overlap_inv_cap", cap)
            out = []
            if B == 1:
                g = cg.Covsum(n, 10)
                g.set_data(X[:n], y[:n])
                for th in thetas + thetas[:1]:
                    g.set_loghyperparam(th)
                    out.append((g.loglik_resident(), g.grad_resident().copy()))
                mu, var = g.compute_test_means_and_variances(X[:n], y[:n], Xt)
                out.append((mu.copy(), var.copy()))
                g.close()
            else:
                b = cg.BCM(X[:n * B], y[:n * B], K=B, rank=0, world=1)
                for th in thetas + thetas[:1]:
                    b.set_BCM_log_hyperparam(th)
                    ll, gr = b.loglik_and_gradient()
                    out.append((ll, gr.copy()))
                mu, var = b.compute_BCM_test_means_and_var(Xt)
                out.append((mu.copy(), var.copy()))
                b.close()
            res[mode] = out
    finally:
        lib().cugp_set_tuning(b"overlap_inv_max_n", OVERLAP_DEFAULT)
        lib().cugp_set_tuning(b"overlap_inv_cap", 148)
        lib().cugp_set_tuning(b"idrows_max_n", 3500)
    for mode in ("ovl64", "ovl8"):
        for a, b_ in zip(res["seq"], res[mode]):
            assert np.array_equal(np.asarray(a[0]), np.asarray(b_[0])) or np.allclose(a[0], b_[0], rtol=1e-11, atol=0)
            assert_grad(b_[1], a[1], 1e-10) if np.asarray(a[1]).shape == (3,) else np.testing.assert_allclose(b_[1], a[1], rtol=1e-10)
    # first and last evaluation use the same theta: the overlapped path must reproduce itself
    assert res["ovl64"][0][0] == res["ovl64"][3][0]
    assert_grad(res["ovl64"][3][1], res["ovl64"][0][1], 1e-11)


def test_non_pd_is_nan_not_an_error():
    """sqrt of a negative pivot gives NaN that propagates; status stays OK (matrixops.cpp:77)."""
    A = np.array([[1.0, 2.0], [2.0, 1.0]])
    L = cg.get_cholesky(A)
    assert np.isnan(L[1, 1])
    q, ld = cg.compute_chol_and_det(A, np.ones(2))
    assert np.isnan(q) or np.isnan(ld)
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (200, 3))
    X[150] = X[3]                                   # duplicate point + vanishing noise => singular K
    c = cg.Covsum(200, 3)
    c.set_loghyperparam([5.0, 0.0, -400.0])
    ll = c.compute_loglikelihood(X, np.sin(X[:, 0]))
    assert not np.isfinite(ll) or ll == ll            # no exception is the contract; NaN/Inf allowed


# ---------------------------------------------------------------------------------------------- covariance
@pytest.mark.parametrize("fast", [0, 1])
@pytest.mark.parametrize("n,d", [(1, 1), (63, 3), (64, 10), (65, 7), (300, 10), (1024, 10)])
def test_K_train_and_k_test(n, d, fast):
    """fast = 0: the reference's operation order (separately rounded sub / mul / add, a true division): the exponent is
    bit-identical and only exp() differs from glibc's by a couple of ulp.  fast = 1 (the default): FMA accumulation and a
    multiply by -0.5/ell^2 -- the exponent moves by a few ulp (ten fused steps), i.e. K by a few |arg| ulp, arg = |xi-xj|^2 / (2 ell^2)."""
    rng = np.random.default_rng(n + d)
    X = rng.uniform(-20, 20, (n, d)) if d == 10 else rng.uniform(-3, 3, (n, d))
    try:
        assert lib().cugp_set_tuning(b"cov_fast", fast) == 0
        for th in ([0.5, 0.5, 0.5], TH_B):
            c = cg.Covsum(n, d)
            c.set_loghyperparam(th)
            K, Ko = c.compute_K_train(X), PORT.K_train(X, th)
            assert np.array_equal(K, K.T)
            assert np.array_equal(np.diag(K), np.diag(Ko))                 # sf2*exp(0)+sn2, exact
            sf2 = np.exp(2 * th[1])
            with np.errstate(divide="ignore"):
                arg = np.where(Ko > 0, -np.log(np.maximum(Ko, 1e-320) / sf2), 745.0)
            np.fill_diagonal(arg, 0.0)
            ulps = 8 + (8 * np.abs(arg) if fast else 0)   # <= ~6 ulp on the exponent: 10 fused steps + the scaling
            # off-diagonal: exp() within a couple of ulp of glibc's (subnormals: absolute)
            assert np.all(np.abs(K - Ko) <= ulps * np.spacing(np.abs(Ko)) + 1e-320), (np.abs(K - Ko) / np.spacing(np.abs(Ko) + 1e-300)).max()
            xt = rng.uniform(-3, 3, d)
            k, ko = c.compute_k_test(X, xt), PORT.k_test(X, th, xt)
            with np.errstate(divide="ignore"):
                argk = np.where(ko > 0, -np.log(np.maximum(ko, 1e-320) / sf2), 745.0)
            assert np.all(np.abs(k - ko) <= (8 + (8 * argk if fast else 0)) * np.spacing(np.abs(ko)) + 1e-320)
            c.close()
    finally:
        lib().cugp_set_tuning(b"cov_fast", 1)


def test_goldens_hold_with_the_reference_arithmetic_too():
    """The strict covariance arithmetic (cov_fast = 0) is still there and still meets the goldens (C1 at both thetas)."""
    try:
        lib().cugp_set_tuning(b"cov_fast", 0)
        for name in ("C1_sine1024_thA_pred8", "C1_sine1024_thB_pred16"):
            c = GOLD[name]
            X, y, Xt, yt = case_inputs(c)
            g = cg.Covsum(c["n"], c["d"])
            g.set_loghyperparam(c["theta"])
            assert_ll(g.compute_loglikelihood(X, y), c["ll"])
            assert_grad(g.compute_gradient_loghyperparam(X, y), c["grad"])
            mu, var = g.compute_test_means_and_variances(X, y, Xt)
            assert_pred(mu, var, c["mean"], c["var"], yscale=np.abs(y).max())
            g.close()
    finally:
        lib().cugp_set_tuning(b"cov_fast", 1)


# ---------------------------------------------------------------------------------------------- golden cases
OVERLAP_DEFAULT = 2800   # library default of the tuning key overlap_inv_max_n

COVSUM = [n for n, c in GOLD.items() if c["kind"] == "covsum"]
BCMS = [n for n, c in GOLD.items() if c["kind"] == "bcm"]


@pytest.mark.parametrize("name", COVSUM)
def test_covsum_golden(name):
    c = GOLD[name]
    X, y, Xt, yt = case_inputs(c)
    g = cg.Covsum(c["n"], c["d"])
    g.set_loghyperparam(c["theta"])
    assert_ll(g.compute_loglikelihood(X, y), c["ll"])
    assert_grad(g.compute_gradient_loghyperparam(X, y), c["grad"])
    if "mean" in c:
        mu, var = g.compute_test_means_and_variances(X, y, Xt)
        assert_pred(mu, var, c["mean"], c["var"], yscale=np.abs(y).max())
        # NLPP: through the reference formula on our moments (gated like the moments)
        assert abs(g.get_negative_log_predprob(yt, mu, var) - c["nlpp"]) <= 1e-7 * max(1.0, abs(c["nlpp"]))
    if "alpha" in c:
        g.set_data(X, y)
        q, ld, ll = g.scalars_resident()
        assert abs(q - c["quad"]) <= 1e-9 * abs(c["quad"]) and abs(ld - c["logdet"]) <= 1e-10 * abs(c["logdet"])
        a = g.alpha_resident()
        assert np.linalg.norm(a - c["alpha"]) <= 1e-9 * np.linalg.norm(c["alpha"])
    g.close()


@pytest.mark.parametrize("name", BCMS)
def test_bcm_golden_single_gpu(name):
    c = GOLD[name]
    X, y, Xt, yt = case_inputs(c)
    b = cg.BCM(X, y, c["n"], c["d"], c["K"], rank=0, world=1)
    b.set_BCM_log_hyperparam(c["theta"])
    assert_ll(b.get_BCM_loglikelihood(), c["ll"])
    assert_grad(b.get_BCM_gradient_hyper(), c["grad"])
    if "mean" in c:
        mu, var = b.compute_BCM_test_means_and_var(Xt)
        assert_pred(mu, var, c["mean"], c["var"], yscale=np.abs(y).max())
        assert abs(b.get_BCM_negative_log_predprob(yt, mu, var) - c["nlpp"]) <= 1e-7 * max(1.0, abs(c["nlpp"]))
    b.close()


def test_bcm_rank_partials_sum_to_whole():
    """Emulate W=3 ranks in one process: the per-rank partial sums and moments must add up to the W=1 result
    (the allreduce is a plain sum)."""
    c = GOLD["si24000_first3000_bcm4_thB_pred16"]
    X, y, Xt, _ = case_inputs(c)
    whole = cg.BCM(X, y, K=4, rank=0, world=1)
    whole.set_BCM_log_hyperparam(c["theta"])
    ll, g = whole.loglik_and_gradient()
    mu, var = whole.compute_BCM_test_means_and_var(Xt)
    acc4, accPQ = np.zeros(4), np.zeros((2, Xt.shape[0]))
    for r in range(3):
        part = cg.BCM(X, y, K=4, rank=r, world=3)       # no process group: _allreduce is the identity
        part.set_BCM_log_hyperparam(c["theta"])
        acc4 += part._local.loglik_grad(True)
        accPQ += part._local.moments(np.ascontiguousarray(Xt))
        part.close()
    assert_ll(acc4[0], ll, 1e-13)
    assert_grad(acc4[1:], g, 1e-12)
    assert_pred(accPQ[1] / accPQ[0], 1.0 / accPQ[0], mu, var, rtol=1e-12)
    assert_ll(ll, c["ll"])
    whole.close()


@pytest.mark.parametrize("chunk", [0, 256])
def test_c4_prediction_at_600_points(chunk):
    """C4 at m = 600 test points (golden from the unmodified reference, tests/golden/make_golden.py --group c4pred):
    several row tiles of the variance GEMM's column-sum epilogue and -- with the chunk cap -- three test-set chunks."""
    c, c1 = GOLD["C4_si24000_bcm16_thC_pred600"], GOLD["C4_expert0_n1500_thB_pred600"]
    d = load_data("si24000")
    Xt = np.ascontiguousarray(load_data("c4_xtest600")["Xtest"])
    try:
        assert lib().cugp_set_tuning(b"pred_chunk", chunk) == 0
        b = cg.BCM(d["X"], d["y"], K=16, rank=0, world=1)
        b.set_BCM_log_hyperparam(c["theta"])
        mu, var = b.compute_BCM_test_means_and_var(Xt)
        b.close()
        assert_pred(mu, var, c["mean"], c["var"], yscale=np.abs(d["y"]).max())
        g = cg.Covsum(1500, 10)
        g.set_loghyperparam(c1["theta"])
        mu, var = g.compute_test_means_and_variances(d["X"][:1500], d["y"][:1500], Xt)
        g.close()
        assert_pred(mu, var, c1["mean"], c1["var"], yscale=np.abs(d["y"]).max())
    finally:
        lib().cugp_set_tuning(b"pred_chunk", 0)


@pytest.mark.parametrize("name", [n for n, c in GOLD.items() if c["kind"] == "kinv"])
def test_k_inverse_and_cholesky_golden(name):
    """compute_K_inverse / get_cholesky (matrixops.cpp:383-435, 68-108) at n = 1000 and 2048 against the unmodified
    reference: every 37th row and column, the whole diagonal and the Frobenius norm."""
    c = GOLD[name]
    n, st = c["n"], c["stride"]
    X = load_data("sine4096")["X"][:n]
    K = PORT.K_train(X, c["theta"])
    idx = np.arange(0, n, st)
    Ki = cg.compute_K_inverse(K)
    ref = np.array(c["Kinv_sample"]).reshape(len(idx), len(idx))
    assert np.linalg.norm(Ki[np.ix_(idx, idx)] - ref) <= 1e-10 * np.linalg.norm(ref)
    dref = np.array(c["Kinv_diag"])
    assert np.linalg.norm(np.diag(Ki) - dref) <= 1e-10 * np.linalg.norm(dref)
    assert abs(np.linalg.norm(Ki) - c["Kinv_fro"]) <= 1e-10 * c["Kinv_fro"]
    assert np.array_equal(Ki, Ki.T)
    L = cg.get_cholesky(K)
    lref = np.array(c["L_sample"]).reshape(len(idx), len(idx))
    assert np.linalg.norm(L[np.ix_(idx, idx)] - lref) <= 1e-12 * np.linalg.norm(lref)
    assert abs(np.linalg.norm(L) - c["L_fro"]) <= 1e-12 * c["L_fro"]


# ---------------------------------------------------------------------------------------------- API behaviour
def test_loglik_then_grad_share_one_factorisation():
    c = GOLD["sine300_thB_pred7"]
    X, y, _, _ = case_inputs(c)
    g = cg.Covsum(300, 10)
    g.set_loghyperparam(c["theta"])
    ll = g.compute_loglikelihood(X, y)
    lib().cugp_launch_count_reset()
    grad_cached = g.compute_gradient_loghyperparam(X, y)   # same (X, y, theta): the factor of the first call is reused
    n2 = lib().cugp_launch_count()
    g.set_loghyperparam([c["theta"][0] + 1e-12, c["theta"][1], c["theta"][2]])
    lib().cugp_launch_count_reset()
    grad_again = g.compute_gradient_loghyperparam(X, y)    # new theta: covariance + factorisation + inverse + trace
    n3 = lib().cugp_launch_count()
    assert_grad(grad_again, grad_cached, 1e-9)
    h = cg.Covsum(300, 10)
    h.set_loghyperparam(c["theta"])
    grad_fresh = h.compute_gradient_loghyperparam(X, y)
    # a handle that is asked for a gradient FIRST takes L^-T out of the factorisation itself (identity rows); the one that
    # factorised for the log-likelihood first inverts the factor afterwards: same numbers up to rounding
    assert_grad(grad_fresh, grad_cached, 1e-11)
    assert np.array_equal(h.compute_gradient_loghyperparam(X, y), grad_fresh)     # cached: bit-identical
    assert ll == h.compute_loglikelihood(X, y)
    # the cached call launched no factorisation: fewer launches than covariance + 3 block steps + the inverse chain
    assert 0 < n2 and n3 > 0
    g2 = cg.Covsum(300, 10)
    g2.set_loghyperparam(c["theta"])
    g2.compute_loglikelihood(X, y)
    lib().cugp_launch_count_reset()
    g2.compute_loglikelihood(X, y)                        # identical call: nothing but the upload
    assert lib().cugp_launch_count() == 0
    y2 = y.copy()
    y2[0] += 1.0
    assert g.compute_loglikelihood(X, y2) != g.compute_loglikelihood(X, y)   # changed data is noticed


@pytest.mark.parametrize("n,B", [(900, 1), (1500, 2), (2600, 1), (3400, 1)])
def test_identity_rows_match_the_inverse_chain(n, B):
    """n <= 3500: L^-T and K^-1 leave the factorisation itself (n appended identity rows, 64-row tiles once the step is
    wide; replayed as a graph up to 2048, direct launches above) against TRTRI + LAUUM after it (idrows_max_n = 0):
    LL identical (same factor), gradient, alpha and predictions to rounding -- matrixops.cpp:383-435, covkernel.cpp:277-302."""
    from cugp_b200.loaders import synthetic_sine
    X, y = synthetic_sine(n * B + 9, 10, seed=n + B)
    Xt = X[n * B:]
    out = {}
    try:
        for mode, bound in (("chain", 0), ("idrows", 3500)):
            lib().cugp_set_tuning(b"idrows_max_n", bound)
            res = []
            if B == 1:
                g = cg.Covsum(n, 10)
                g.set_data(X[:n], y[:n])
                for th in (TH_B, [2.0, 2.0, 2.0], TH_B):          # the second and third evaluation take the learned path
                    g.set_loghyperparam(th)
                    res.append((g.loglik_resident(), g.grad_resident().copy()))
                res.append(g.compute_test_means_and_variances(X[:n], y[:n], Xt))
                g.close()
            else:
                b = cg.BCM(X[:n * B], y[:n * B], K=B, rank=0, world=1)
                for th in (TH_B, [2.0, 2.0, 2.0], TH_B):
                    b.set_BCM_log_hyperparam(th)
                    ll, gr = b.loglik_and_gradient()
                    res.append((ll, gr.copy()))
                res.append(b.compute_BCM_test_means_and_var(Xt))
                b.close()
            out[mode] = res
    finally:
        lib().cugp_set_tuning(b"idrows_max_n", 3500)
    for a, b_ in zip(out["chain"][:3], out["idrows"][:3]):
        assert_ll(b_[0], a[0], 1e-12)
        assert_grad(b_[1], a[1], 1e-10)
    np.testing.assert_allclose(out["idrows"][3][0], out["chain"][3][0], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(out["idrows"][3][1], out["chain"][3][1], rtol=1e-10)
    assert_grad(out["idrows"][2][1], out["idrows"][0][1], 1e-10)    # same theta: first evaluation inverts afterwards, third carries the rows
    if n <= 1000 and B == 1:                                        # and the oracle agrees (seconds on the CPU)
        assert_ll(out["idrows"][0][0], PORT.loglik(X[:n], y[:n], TH_B), 1e-10)
        assert_grad(out["idrows"][0][1], PORT.grad(X[:n], y[:n], TH_B), 1e-9)


def test_cg_solve_replays_reference_log():
    """cuda_bettersinglenode_ver2/REF: LL trajectory (6 digits) and optimum of the 128x2 problem."""
    d = load_data("si128x2")
    log = GOLD["REF_log"]
    g = cg.Covsum(128, 2)
    g.set_loghyperparam(log["theta0"])
    trace = g.cg_solve(d["X"], d["y"])
    lls = [-f for f in trace]
    logged = log["ll_sequence"][1:1 + len(lls)]
    assert len(lls) >= 60
    bad = [(i, a, b) for i, (a, b) in enumerate(zip(lls, logged)) if abs(a - b) > 5e-6 * max(1.0, abs(b))]
    assert not bad, bad[:5]
    assert np.allclose(g.get_loghyperparam(), GOLD["si128_cg_from_th15"]["theta_final"], rtol=0, atol=1e-6)
    b = cg.BCM(d["X"], d["y"], K=4, rank=0, world=1)
    b.set_BCM_log_hyperparam(log["theta0"])
    b.cg_solve()
    assert np.allclose(b.get_loghyperparam(), GOLD["si128_bcm4_cg_from_th15"]["theta_final"], rtol=0, atol=1e-5)


# ---------------------------------------------------------------------------------------------- full sizes
def test_c3_size_properties():
    """n = 10000 (C3 shape, synthetic sine data): properties that do not need the O(n^3) CPU path.
    K alpha = y to 1e-10, logdet and quad against LAPACK (float64) to 1e-10, LL formula consistency."""
    from cugp_b200.loaders import synthetic_sine
    n = 10000
    X, y = synthetic_sine(n + 64, 10)
    Xt, X, y = X[n:], X[:n], y[:n]
    g = cg.Covsum(n, 10)
    g.set_loghyperparam(TH_B)
    K = g.compute_K_train(X)
    g.set_data(X, y)
    q, ld, ll = g.scalars_resident()
    a = g.alpha_resident()
    assert np.linalg.norm(K @ a - y) <= 1e-10 * np.linalg.norm(y)
    Lk = np.linalg.cholesky(K)
    ld_ref = 2.0 * np.log(np.diag(Lk)).sum()
    assert abs(ld - ld_ref) <= 1e-10 * abs(ld_ref)
    assert abs(q - y @ a) <= 1e-10 * abs(q)
    assert ll == -0.5 * (q + ld + n * 1.83787)
    # gradient against the closed form with LAPACK's inverse
    import scipy.linalg as sl
    Ki = sl.cho_solve((Lk, True), np.eye(n))
    W = Ki - np.outer(a, a)
    sf2, sn2, ell2 = np.exp(2 * TH_B[1]), np.exp(2 * TH_B[2]), np.exp(2 * TH_B[0])
    Ks = K - sn2 * np.eye(n)
    D = -2.0 * ell2 * np.log(np.maximum(Ks, 1e-300) / sf2)     # |xi-xj|^2 recovered from K
    np.fill_diagonal(D, 0.0)
    gref = np.array([0.5 * np.sum(W * Ks * D / ell2), np.sum(W * Ks), sn2 * np.trace(W)])
    assert_grad(g.grad_resident(), gref, 1e-8)
    # prediction against the same LAPACK factor
    mu, var = g.compute_test_means_and_variances(X, y, Xt)
    kst = np.stack([PORT.k_test(X, TH_B, x) for x in Xt])
    v = sl.solve_triangular(Lk, kst.T, lower=True)
    assert_pred(mu, var, kst @ a, sf2 + sn2 - (v * v).sum(0), yscale=1.0, rtol=1e-8)


@pytest.mark.parametrize("n", [130, 700, 3000, 5200])
def test_inplace_inverse_matches_three_buffer_path(n):
    """Large-n gradient path (default from n = 60 000: three n x n buffers no longer fit next to each other): T = L^-1 over
    L with a compact scratch, K^-1 = T^T T over T row block by row block (matrixops.cpp:383-435 in one buffer).  Forced
    here at small n and compared with the three-buffer path; the factor is rebuilt transparently afterwards."""
    from cugp_b200.loaders import synthetic_sine
    X, y = synthetic_sine(n, 10, seed=n)
    try:
        lib().cugp_set_tuning(b"idrows_max_n", 0)          # (handles created below keep L^-T out of the factorisation)
        g = cg.Covsum(n, 10)
        g.set_data(X, y)
        g.set_loghyperparam(TH_B)
        ll0, g0 = g.loglik_resident(), g.grad_resident()
        lib().cugp_set_tuning(b"inplace_inverse_min_n", 0)
        g.set_loghyperparam([TH_B[0], TH_B[1], TH_B[2] + 1e-13])
        ll1, g1 = g.loglik_resident(), g.grad_resident()
        a1 = g.alpha_resident()
        ll2 = g.loglik_resident()                          # K^-1 sits where L was: asking again must still be right
        mu, var = g.compute_test_means_and_variances(X, y, X[:5])
    finally:
        lib().cugp_set_tuning(b"inplace_inverse_min_n", 60000)
        lib().cugp_set_tuning(b"idrows_max_n", 3500)
    assert_ll(ll1, ll0, 1e-11)
    assert ll2 == ll1
    assert_grad(g1, g0, 1e-9)
    K = PORT.K_train(X, TH_B)
    assert np.linalg.norm(K @ a1 - y) <= 1e-9 * np.linalg.norm(y)
    assert np.all(np.isfinite(mu)) and np.all(var > 0)
    g.close()


def test_residual_beyond_int32_indexing():
    """n = 50 000 > 46 341: n * n no longer fits a 32-bit int -- where the reference's GPU flavour indexes with `int`
    (cuda_src/cuda_gp.cu:613-642).  The reference printed a Cholesky residual after every factorisation
    (cuda_src/cuda_gp.cu:1126-1139); here: ||K alpha - y|| / ||y|| with K rebuilt matrix-free, and the log-determinant and
    y' K^-1 y must not depend on the outer block width or on the look-ahead."""
    from cugp_b200.loaders import synthetic_sine
    n = 50000
    X, y = synthetic_sine(n, 10)
    g = cg.Covsum(n, 10)
    g.set_data(X, y)
    g.set_loghyperparam(TH_B)
    q, ld, ll = g.scalars_resident()
    r = g.residual_resident()
    assert np.linalg.norm(r) <= 1e-10 * np.linalg.norm(y), np.linalg.norm(r) / np.linalg.norm(y)
    a = g.alpha_resident()
    assert abs(q - y @ a) <= 1e-10 * abs(q)
    assert ll == -0.5 * (q + ld + n * 1.83787)
    try:
        for nb, la in ((512, 1), (1024, 0)):
            lib().cugp_set_tuning(b"potrf_nb", nb)
            lib().cugp_set_tuning(b"lookahead", la)
            g.set_loghyperparam([TH_B[0], TH_B[1], TH_B[2] + 0.0])   # same theta: force a fresh factorisation below
            g.factorize_resident()
            q2, ld2, _ = g.scalars_resident()
            assert abs(ld2 - ld) <= 1e-12 * abs(ld) and abs(q2 - q) <= 1e-11 * abs(q), (nb, la, ld, ld2, q, q2)
    finally:
        lib().cugp_set_tuning(b"potrf_nb", 0)
        lib().cugp_set_tuning(b"lookahead", 1)
    g.close()


def test_cholesky_against_cusolver_at_40000():
    """Test-only library cross-check (never on the product path): torch.linalg.cholesky (cuSOLVER) on the same K at
    n = 40 000 -- log-determinant, y' K^-1 y and a sample of L."""
    import torch
    from cugp_b200.loaders import synthetic_sine
    n = 40000
    X, y = synthetic_sine(n, 10)
    g = cg.Covsum(n, 10)
    g.set_data(X, y)
    g.set_loghyperparam(TH_B)
    q, ld, _ = g.scalars_resident()
    a = g.alpha_resident()
    g.close()
    del g
    Xd = torch.from_numpy(X).cuda()
    ell2, sf2, sn2 = np.exp(2 * TH_B[0]), np.exp(2 * TH_B[1]), np.exp(2 * TH_B[2])
    K = torch.cdist(Xd, Xd).pow_(2).mul_(-0.5 / ell2).exp_().mul_(sf2)
    K.diagonal().add_(sn2)
    Lc = torch.linalg.cholesky(K)
    del K
    ld_ref = 2.0 * torch.log(Lc.diagonal()).sum().item()
    yd = torch.from_numpy(y).cuda()
    a_ref = torch.cholesky_solve(yd[:, None], Lc)[:, 0]
    q_ref = float(yd @ a_ref)
    # cdist's |x-y|^2 differs from the reference's ordered sum in the last bits: cond(K) ~ 2e4 amplifies that to ~1e-11
    assert abs(ld - ld_ref) <= 1e-9 * abs(ld_ref), (ld, ld_ref)
    assert abs(q - q_ref) <= 1e-8 * abs(q_ref), (q, q_ref)
    assert np.linalg.norm(a - a_ref.cpu().numpy()) <= 1e-6 * np.linalg.norm(a)
