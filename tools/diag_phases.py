"""Cycle breakdown of the blocked 128x128 diagonal-block kernel (clock64 stamps of thread 0)."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from cugp_b200._lib import lib, ptr

rng = np.random.default_rng(0)
X = rng.uniform(-3, 3, (128, 4))
D = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)
A = np.exp(-0.5 * D / 2.0) + 0.1 * np.eye(128)
st = (C.c_longlong * 32)()
rc = lib().cugp_debug_diag_phases(ptr(np.ascontiguousarray(A)), st, 32)
s = list(st)
names = ["load"] + [f"{w}{p}" for p in range(3) for w in ("chol", "trsm", "syrk")] + ["chol3", "inv_diag", "inv_l32", "inv_l64", "writeback"]
print("rc", rc, "total cycles", s[15] - s[0])
for i in range(1, 16):
    print(f"{names[i-1]:10s} {s[i] - s[i-1]:8d}")
