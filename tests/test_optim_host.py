"""The product's host-side optimisers (csrc/optim.cpp: cugp_cg_minimize, cugp_rprop_minimize -- the callers of the
hot loop) need no GPU: here they are driven with the CPU oracle as the evaluation
callback and must replay the reference's own run log (cuda_bettersinglenode_ver2/REF) and the oracle's port of
Covsum::cg_solve / rprop_solve (covkernel.cpp:337-647) step for step."""
import ctypes as C

import numpy as np

from cugp_b200._lib import EVAL_FN, lib, ptr
from oracle import oracle
from tests.conftest import load_data, load_golden

PORT = oracle.port()
GOLD = load_golden()


def _callback(X, y, calls):
    def _eval(_ctx, th_p, f_p, g_p):
        th = [th_p[0], th_p[1], th_p[2]]
        calls.append(th)
        f_p[0] = -1.0 * PORT.loglik(X, y, th)        # f = -LL (covkernel.cpp:441-443)
        g = PORT.grad(X, y, th)                      # already d(-LL)/dtheta
        g_p[0], g_p[1], g_p[2] = g
        return 0
    return EVAL_FN(_eval)


def test_cg_minimize_replays_reference_log():
    d = load_data("si128x2")
    log = GOLD["REF_log"]
    X, y = d["X"], d["y"]
    calls = []
    th = np.array(log["theta0"], dtype=float)
    tr = np.full(256, np.nan)
    ne = C.c_int()
    cb = _callback(X, y, calls)
    assert lib().cugp_cg_minimize(cb, None, ptr(th), ptr(tr), 256, C.byref(ne)) == 0
    lls = [-f for f in tr[:ne.value]]
    logged = log["ll_sequence"][1:1 + len(lls)]
    assert len(logged) >= 60 and len(calls) == ne.value + 1     # the start point + one evaluation per trial point
    bad = [(i, a, b) for i, (a, b) in enumerate(zip(lls, logged)) if abs(a - b) > 5e-6 * max(1.0, abs(b))]
    assert not bad, bad[:5]
    assert np.allclose(th, [0.882908, 0.098703, -2.971479], rtol=0, atol=2e-6)      # REF:3183
    # and the oracle's port of the same loop: identical trajectory (same arithmetic on the same evaluations)
    th_o, ne_o, tr_o = PORT.cg_solve(X, y, log["theta0"])
    assert ne_o == ne.value
    assert np.array_equal(tr_o, tr[:ne.value]) and np.array_equal(th_o, th)


def test_cg_minimize_bisects_on_nan_and_stops_on_error():
    """covkernel.cpp:509-524: a NaN / Inf evaluation shrinks the step instead of aborting; a non-zero return from the
    callback aborts with that status."""
    state = {"n": 0}

    def _eval(_ctx, th_p, f_p, g_p):
        state["n"] += 1
        x = np.array([th_p[0], th_p[1], th_p[2]])
        if np.linalg.norm(x) > 4.0:                  # a wall of NaN around the basin
            f_p[0] = float("nan")
            g_p[0] = g_p[1] = g_p[2] = float("nan")
            return 0
        f_p[0] = float(np.sum((x - 1.0) ** 2) + 0.1 * np.sum((x - 1.0) ** 4))
        g = 2.0 * (x - 1.0) + 0.4 * (x - 1.0) ** 3
        g_p[0], g_p[1], g_p[2] = g
        return 0
    th = np.array([-2.0, 2.5, 0.5])
    ne = C.c_int()
    assert lib().cugp_cg_minimize(EVAL_FN(_eval), None, ptr(th), None, 0, C.byref(ne)) == 0
    assert np.allclose(th, 1.0, atol=1e-5) and 3 < ne.value <= 101

    def _fail(_ctx, th_p, f_p, g_p):
        return 7
    th = np.zeros(3)
    assert lib().cugp_cg_minimize(EVAL_FN(_fail), None, ptr(th), None, 0, C.byref(ne)) != 0


def test_rprop_minimize_matches_oracle_port():
    d = load_data("si128x2")
    X, y = d["X"][:64], d["y"][:64]
    th0 = [0.5, 0.5, 0.5]
    calls = []
    th = np.array(th0)
    it = C.c_int()
    assert lib().cugp_rprop_minimize(_callback(X, y, calls), None, ptr(th), C.byref(it)) == 0
    th_o = PORT.rprop_solve(X, y, th0)
    assert np.allclose(th, th_o, rtol=1e-12, atol=1e-12)
    assert it.value == 100                            # covkernel.cpp:345: 100 iterations, no early exit
