"""Times the reference's OWN GPU generation (cuSOLVER potrf + cuBLAS trsm / gemm, cuda_bettersinglenode_ver2/cuda_gp.cu,
compiled unchanged for sm_100 as oracle/_ref/libcugp_refgpu.so) on the synthetic set: one JSON line.
Run as a subprocess of bench.py (`library_baseline`): the reference keeps global device state, never frees its
workspaces and prints every step.   usage: refgpu_time.py <n> [reps]"""
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cugp_b200.loaders import synthetic_sine  # noqa: E402
from oracle import oracle  # noqa: E402

n = int(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
TH_B = [3.762111, -1.152105, -0.384461]
path = oracle.reference_gpu_path()
if path is None:
    print(json.dumps({"n": n, "unavailable": "oracle/_ref/libcugp_refgpu.so not built"}))
    sys.exit(0)
X, y = synthetic_sine(n, 10)
with tempfile.TemporaryDirectory() as t:
    fi, fl = os.path.join(t, "in.txt"), os.path.join(t, "lab.txt")
    with open(fi, "w") as f:
        f.write(f"{n} 10\n")
        np.savetxt(f, X, fmt="%.17g")
    np.savetxt(fl, y, fmt="%.17g")
    out_fd = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    if not os.environ.get("REFGPU_VERBOSE"):
        os.dup2(devnull, 1)                  # the reference prints from C: silence fd 1 while it runs
    try:
        lib = C.CDLL(path)
        lib.refgpu_loglik.restype = C.c_double
        lib.refgpu_setup(n, fi.encode(), fl.encode())
        th = (C.c_double * 3)(*TH_B)
        lib.refgpu_set_theta(th)
        lls, ts, tg = [], [], []
        for _ in range(reps + 1):
            t0 = time.perf_counter()
            lls.append(lib.refgpu_loglik())
            ts.append(time.perf_counter() - t0)
        g = (C.c_double * 3)()
        for _ in range(reps):
            t0 = time.perf_counter()
            lib.refgpu_grad(g)
            tg.append(time.perf_counter() - t0)
    finally:
        sys.stdout.flush()
        os.dup2(out_fd, 1)
print(json.dumps({"n": n, "loglik_ms": 1e3 * min(ts[1:]), "grad_ms": 1e3 * min(tg), "ll": lls[-1], "grad": list(g),
                  "what": "reference cuda_bettersinglenode_ver2/cuda_gp.cu unchanged (cusolverDnDpotrf + cublasDtrsm on I + "
                          "cublasDgemm for K^-1 even on the LL path, SURVEY 2.3), theta_B, best of %d" % reps}))
