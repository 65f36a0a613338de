"""BCM evaluation (LL + gradient) time against the number of 1500-row experts on ONE GPU, optionally under a tuning key.
usage: r2_bcm_k.py [key v0,v1,...] [K list]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200._lib import lib

L = lib()
key = sys.argv[1].encode() if len(sys.argv) > 2 else None
vals = [int(v) for v in sys.argv[2].split(",")] if key else [None]
Ks = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 3, 4, 8]
d = np.load("tests/golden/data_si24000.npz")
for v in vals:
    if key:
        assert L.cugp_set_tuning(key, v) == 0, L.cugp_last_error()
    line = f"{key.decode() if key else 'defaults'}={v}:"
    for K in Ks:
        b = cg.BCM(d["X"][:1500 * K], d["y"][:1500 * K], K=K, rank=0, world=1)
        te = []
        for r in range(9):
            b.set_BCM_log_hyperparam([2.0 + 1e-7 * r, 2.0, 2.0])
            t = time.perf_counter()
            b.loglik_and_gradient()
            te.append(time.perf_counter() - t)
        b.close()
        line += f"  K={K}: {1e3 * min(te[2:]):.3f} ms"
    print(line, flush=True)
