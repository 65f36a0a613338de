"""Phase stamps of the fused Cholesky block steps (cugp_debug_step_stamps): where a step's time goes."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200._lib import check, lib
from cugp_b200.loaders import synthetic_sine

L = lib()
TH_B = [3.762111, -1.152105, -0.384461]
for n in [int(a) for a in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["1500", "4096"])]:
    for la in (1, 0):
        L.cugp_set_tuning(b"lookahead", la)
        X, y = synthetic_sine(n, 10)
        g = cg.Covsum(n, 10)
        g.set_data(X, y)
        g.set_loghyperparam(TH_B)
        g.loglik_resident()
        nblk = (n + 127) // 128
        st = (C.c_longlong * (nblk * 48))()
        ms = C.c_float()
        for rep in range(2):
            g.set_loghyperparam([TH_B[0] + 1e-7 * (rep + 1), TH_B[1], TH_B[2]])
            check(L.cugp_debug_step_stamps(g._h, st, nblk, C.byref(ms)))
        s = np.array(st[:], dtype=np.int64).reshape(nblk, 3, 16)
        t0 = s[0, 1, 0]
        print(f"== n={n} lookahead={la} chol {ms.value:.3f} ms; per step (us): diag-start-to-next-diag-start, then phases")
        for b in range(nblk):
            d, r, y_ = s[b, 1], s[b, 2], s[b, 0]
            nxt = (s[b + 1, 1, 0] - d[0]) / 1e3 if b + 1 < nblk else float("nan")
            dd = [(d[i] - d[0]) / 1e3 if d[i] else -1 for i in range(13)]
            rr = [(r[i] - d[0]) / 1e3 if r[i] else -1 for i in range(9)]
            yy = [(y_[i] - d[0]) / 1e3 if y_[i] else -1 for i in range(3)]
            cyc = d[14] - d[13]
            print(f"blk {b:2d} chol32 {cyc} cyc +{(d[0]-t0)/1e3:8.1f}us step {nxt:6.1f} | DIAG wait {dd[1]:5.1f} load {dd[2]:5.1f} p0 {dd[3]:5.1f}/{dd[4]:5.1f} p1 {dd[5]:5.1f}/{dd[6]:5.1f} "
                  f"p2 {dd[7]:5.1f}/{dd[8]:5.1f} p3 {dd[9]:5.1f} flag {dd[11]:5.1f} end {dd[12]:5.1f} | ROWS0 start {rr[0]:5.1f} ld {rr[1]:5.1f} pro {rr[2]:5.1f} "
                  f"flags {rr[3]:5.1f} {rr[4]:5.1f} {rr[5]:5.1f} {rr[6]:5.1f} trsm {rr[7]:5.1f} end {rr[8]:5.1f} | SYRKD {yy[0]:5.1f} {yy[1]:5.1f} {yy[2]:5.1f}")
            if b >= 13 and b < nblk - 3:
                continue
        g.close()
L.cugp_set_tuning(b"lookahead", 1)
