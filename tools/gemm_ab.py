"""DMMA GEMM on the probe shapes next to the DMMA-only probe (burst and sustained)."""
import ctypes as C
import sys

sys.path.insert(0, ".")
from cugp_b200._lib import lib

L = lib()
for v in (1,):
    for M, N, K in [(8192, 8192, 128), (8192, 8192, 512), (8192, 8192, 1024), (16384, 16384, 1024), (32768, 32768, 1024)]:
        t = C.c_double()
        rc = L.cugp_probe_gemm(M, N, K, 5, C.byref(t))
        print(f"gemm_kernel={v} {M}x{N}x{K}: {t.value:.2f} TF (rc={rc})", flush=True)
for ms in (10.0, 100.0, 1000.0, 3000.0):
    tf, mhz = C.c_double(), C.c_double()
    L.cugp_probe_dmma(ms, C.byref(tf), C.byref(mhz))
    print(f"dmma probe {ms:.0f} ms: {tf.value:.2f} TF at {mhz.value:.0f} MHz", flush=True)
