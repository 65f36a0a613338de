import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device here")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden():
    cases = {}
    for f in sorted(os.listdir(GOLD)):
        if f.startswith("golden_") and f.endswith(".json"):
            cases.update(json.load(open(os.path.join(GOLD, f)))["cases"])
    return cases


def load_data(name):
    return np.load(os.path.join(GOLD, f"data_{name}.npz"))


def case_inputs(c):
    """(X, y, Xtest, ytest) for a golden case, recovered from its name."""
    name, n = c["name"], c["n"]
    m = c.get("m", 0)
    if name.startswith("si24000_first") or name.startswith("C4_"):
        d = load_data("si24000")
        return d["X"][:n], d["y"][:n], d["Xtest"][:m], d["ytest"][:m]
    if name.startswith("C2_"):
        d = load_data("sine4096")
    elif name.startswith("sine") or name.startswith("C1_"):
        d = load_data("sine1024")
    else:
        d = load_data("si128x2")
    X, y = d["X"], d["y"]
    return X[:n], y[:n], X[n:n + m], y[n:n + m]


@pytest.fixture(scope="session")
def golden():
    return load_golden()


# Tolerance policy (BASELINE.json north_star):
#   log-likelihood, gradient: 1e-9 relative; every gradient component against max(|g_k|, 1e-3 * ||g||_inf)
#   predictive mean / variance: 1e-8 relative; means that are numerically zero in the reference
#   (K ~ diagonal at theta_A) are gated against max(|ref|, 1e-12 * ||y||_inf ... ) i.e. absolutely.
LL_RTOL = 1e-9
PRED_RTOL = 1e-8


def assert_ll(got, ref, rtol=LL_RTOL):
    assert np.isfinite(got), got
    assert abs(got - ref) <= rtol * abs(ref), (got, ref, abs(got - ref) / abs(ref))


GRAD_FLOOR = 1e-3   # a component is gated against max(|g_k|, GRAD_FLOOR * ||g||_inf)


def assert_grad(got, ref, rtol=LL_RTOL, floor=GRAD_FLOOR):
    """Every component on its own: |got_k - ref_k| <= rtol * max(|ref_k|, floor * ||ref||_inf).  The floor only
    protects components that are cancellation noise next to the others (g0 ~ 1e-10 against g1 ~ 1e3 at theta_A, where
    K is numerically diagonal: SURVEY.md section 8c); floor = 1 is the inf-norm gate SURVEY allows as the minimum."""
    got, ref = np.asarray(got, float), np.asarray(ref, float)
    scale = np.maximum(np.abs(ref), floor * np.max(np.abs(ref)))
    assert np.all(np.isfinite(got)), got
    assert np.all(np.abs(got - ref) <= rtol * scale), (got, ref, np.abs(got - ref) / scale)


def assert_pred(mean, var, ref_mean, ref_var, yscale=1.0, rtol=PRED_RTOL):
    mean, var, ref_mean, ref_var = (np.asarray(a, float) for a in (mean, var, ref_mean, ref_var))
    ms = np.maximum(np.abs(ref_mean), 1e-6 * yscale)   # means ~1e-28 where K ~ diag: absolute gate
    assert np.all(np.abs(mean - ref_mean) <= rtol * ms), (mean, ref_mean)
    assert np.all(np.abs(var - ref_var) <= rtol * np.abs(ref_var)), (var, ref_var)
