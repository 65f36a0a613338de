"""DMMA GEMM probe: C[MxN] -= A B^T on resident operands, TFLOP/s per shape (args: M,N,K triples)."""
import ctypes as C
import sys

sys.path.insert(0, ".")
from cugp_b200._lib import lib

shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(8192, 8192, 128), (8192, 8192, 512), (8192, 8192, 1024), (16384, 16384, 1024)]
for M, N, K in shapes:
    t = C.c_double()
    rc = lib().cugp_probe_gemm(M, N, K, 5, C.byref(t))
    print(f"{M}x{N}x{K}: {t.value:.2f} TF (rc={rc})", flush=True)
