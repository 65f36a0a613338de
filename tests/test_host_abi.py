"""The parts of the C ABI that are pure host arithmetic, and its behaviour without a CUDA device (this suite runs on a
box with no GPU): there is NO CPU fallback -- compute entry points must fail loudly with CUGP_ERR_NODEVICE."""
import ctypes as C

import numpy as np
import pytest

from cugp_b200._lib import ERR_INVALID, ERR_NODEVICE, lib, ptr
from tests.conftest import _has_gpu, case_inputs, load_golden

GOLD = load_golden()


def test_nlpp_matches_reference_goldens():
    """Covsum::get_negative_log_predprob (covkernel.cpp:649-659, 2*pi truncated to 6.283185) on the golden predictions."""
    n = 0
    for name, c in GOLD.items():
        if "nlpp" not in c or "mean" not in c:
            continue
        X, y, Xt, yt = case_inputs(c)
        out = C.c_double()
        mu, var = np.array(c["mean"]), np.array(c["var"])
        assert lib().cugp_nlpp(ptr(np.ascontiguousarray(yt)), ptr(mu), ptr(var), len(mu), C.byref(out)) == 0
        assert abs(out.value - c["nlpp"]) <= 1e-12 * max(1.0, abs(c["nlpp"])), name
        n += 1
    assert n >= 3
    assert lib().cugp_nlpp(None, None, None, 0, None) == ERR_INVALID


def test_poe_finalize_is_the_reference_product_of_experts():
    """BCM.cpp:51-60: var = 1 / sum_e 1/var_e, mean = var * sum_e mean_e/var_e."""
    rng = np.random.default_rng(0)
    m, E = 17, 5
    mu, var = rng.standard_normal((E, m)), rng.uniform(0.1, 2.0, (E, m))
    PQ = np.concatenate([(1.0 / var).sum(0), (mu / var).sum(0)])
    mean, v = np.zeros(m), np.zeros(m)
    assert lib().cugp_poe_finalize(ptr(PQ), m, ptr(mean), ptr(v)) == 0
    tempvar = 1.0 / (1.0 / var).sum(0)
    assert np.array_equal(v, tempvar) and np.array_equal(mean, tempvar * (mu / var).sum(0))


@pytest.mark.skipif(_has_gpu(), reason="this box has a GPU")
def test_no_device_means_error_not_fallback():
    L = lib()
    cnt = C.c_int(-1)
    assert L.cugp_device_count(C.byref(cnt)) == ERR_NODEVICE and cnt.value == 0
    h = C.c_void_p()
    assert L.cugp_covsum_create(16, 2, C.byref(h)) == ERR_NODEVICE and not h.value
    assert b"no CPU fallback" in L.cugp_last_error()
    X, y = np.zeros((16, 2)), np.zeros(16)
    assert L.cugp_bcm_create(ptr(X), ptr(y), 16, 2, 2, 0, 1, C.byref(h)) == ERR_NODEVICE
    assert L.cugp_shardstream_open_memory(ptr(X), ptr(y), 2, 8, 2, 0, 1, 0, C.byref(h)) == ERR_NODEVICE
    A, Lm = np.eye(4), np.zeros((4, 4))
    assert L.cugp_cholesky(ptr(A), ptr(Lm), 4) == ERR_NODEVICE
    # argument checks come before the device check and never touch a device
    assert L.cugp_covsum_create(0, 2, C.byref(h)) == ERR_INVALID
    assert L.cugp_set_tuning(b"no_such_key", 1) == ERR_INVALID
    import cugp_b200 as cg
    with pytest.raises(cg.CugpError):
        cg.Covsum(8, 2)
