"""Build cugp_b200/libcugp.so (the C ABI of include/cugp.h) for sm_100a with nvcc, in tree.

    python -m cugp_b200.build [--force]

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libcugp.so")
SOURCES = ["gemm_dmma.cu", "kernels.cu", "cholstep.cu", "gp.cu", "capi.cu", "peerxchg.cu", "shardstream.cu", "probe.cu",
           "optim.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "cugp.h"))
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        if force or not _newer(o, [s] + headers):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = ["nvcc", *NVCC_FLAGS, "-x", "cu", "-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{r.stdout}\n{r.stderr}")
        return o

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(compile_one, jobs))
    objs = [os.path.join(OBJ, os.path.splitext(src)[0] + ".o") for src in SOURCES]
    if force or jobs or not _newer(LIB, objs):
        # extern "C" entry points are exported explicitly (visibility attribute below is applied through a
        # version script); libstdc++ is linked statically in this image, so keep it private (-Bsymbolic).
        vs = os.path.join(OBJ, "exports.map")
        with open(vs, "w") as f:
            f.write("{ global: cugp_*; local: *; };\n")
        cmd = ["nvcc", "-shared", "-o", LIB, *objs, "-cudart", "static",
               "-Xlinker", "-Bsymbolic", "-Xlinker", f"--version-script={vs}"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
