"""The drop-in boundary: the header shim (include/cugp_shim) re-declares the reference's C++ classes and free
functions with their exact signatures on top of the C ABI.

* CPU (here): tests/shim/shim_probe.cpp, which calls every member, compiles and links against libcugp.so; and, where
  /root/reference exists, the reference's OWN drivers (cpp_serial_gp/serial_gp.cpp, distributed_gp/
  distributed_ver1.cpp) compile UNCHANGED with the shim headers in place of covkernel.h / matrixops.h / BCM.h.
  The binaries land in oracle/_ref/drivers/ (git-ignored, travels to the GPU box).
* GPU: the binaries run; shim_probe's numbers equal the Python mirror's, and the reference driver's own optimiser
  loop (distributed_ver1.cpp:13-232, compiled from the reference source) reaches the oracle's optimum."""
import os
import re
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "cugp_b200")
OUT = os.path.join(ROOT, "oracle", "_ref", "drivers")
REF = "/root/reference"


def _gxx(args, cwd=None):
    r = subprocess.run(["g++", "-O2", "-w", *args], capture_output=True, text=True, cwd=cwd)
    assert r.returncode == 0, r.stderr[-4000:]


def _link_flags():
    return ["-L" + LIBDIR, "-lcugp", "-Wl,-rpath," + LIBDIR]


def _ensure_lib():
    from cugp_b200 import build
    build.build()


def _build_probe():
    _ensure_lib()
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "shim_probe")
    _gxx(["-I" + INC, os.path.join(ROOT, "tests", "shim", "shim_probe.cpp"), "-o", exe, *_link_flags()])
    return exe


def test_shim_probe_compiles_and_links():
    exe = _build_probe()
    assert os.path.exists(exe)
    # every undefined cugp_* symbol of the binary is exported by the library
    need = {l.split()[-1] for l in subprocess.run(["nm", "-u", exe], capture_output=True, text=True).stdout.splitlines()
            if "cugp_" in l}
    have = {l.split()[-1] for l in subprocess.run(["nm", "-D", "--defined-only", os.path.join(LIBDIR, "libcugp.so")],
                                                  capture_output=True, text=True).stdout.splitlines()}
    assert need and need <= have, need - have


def test_header_symbols_exported():
    """The C-ABI library exports every function include/cugp.h declares (no compute calls here)."""
    _ensure_lib()
    hdr = open(os.path.join(INC, "cugp.h")).read()
    declared = set(re.findall(r"\b(cugp_[A-Za-z0-9_]+)\s*\(", hdr)) - {"cugp_eval_fn"}
    have = {l.split()[-1] for l in subprocess.run(["nm", "-D", "--defined-only", os.path.join(LIBDIR, "libcugp.so")],
                                                  capture_output=True, text=True).stdout.splitlines()}
    assert declared <= have, declared - have
    from cugp_b200._lib import SIGNATURES
    assert declared == set(SIGNATURES), declared ^ set(SIGNATURES)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources only exist in the build container")
def test_reference_drivers_build_unchanged_against_shim():
    _ensure_lib()
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as t:
        # the reference's directory layout, with its three headers replaced by one-line forwards to the shim
        for d in ("common", "cpp_serial_gp", "distributed_gp"):
            os.makedirs(os.path.join(t, d))
        shutil.copy(os.path.join(REF, "common", "cycleTimer.h"), os.path.join(t, "common"))
        shutil.copy(os.path.join(REF, "cpp_serial_gp", "serial_gp.cpp"), os.path.join(t, "cpp_serial_gp"))
        shutil.copy(os.path.join(REF, "distributed_gp", "distributed_ver1.cpp"), os.path.join(t, "distributed_gp"))
        open(os.path.join(t, "common", "matrixops.h"), "w").write('#include "cugp_shim/matrixops.h"\n')
        for d in ("cpp_serial_gp", "distributed_gp"):
            open(os.path.join(t, d, "covkernel.h"), "w").write('#include "cugp_shim/covkernel.h"\n')
        open(os.path.join(t, "distributed_gp", "BCM.h"), "w").write('#include "cugp_shim/BCM.h"\n')
        _gxx(["-I" + INC, "serial_gp.cpp", "-o", os.path.join(OUT, "serial_gp"), *_link_flags()],
             cwd=os.path.join(t, "cpp_serial_gp"))
        # unqualified isnan/isinf (distributed_ver1.cpp:96,102,129,163) need the same compat pre-include as the oracle build
        _gxx(["-I" + INC, "-include", os.path.join(ROOT, "oracle", "ref_compat.h"), "distributed_ver1.cpp", "-o", os.path.join(OUT, "distributed_ver1"), *_link_flags()],
             cwd=os.path.join(t, "distributed_gp"))
    assert os.path.exists(os.path.join(OUT, "serial_gp")) and os.path.exists(os.path.join(OUT, "distributed_ver1"))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources only exist in the build container")
def test_reference_gpu_flavour_driver_builds_unchanged_against_facade():
    """Row f4: the reference's GPU-generation driver -- cuda_bettersinglenode_ver2/main.cpp (socket master/worker main,
    :132-224) and cg_solver.cpp (its Polack-Ribiere loop over compute_log_likelihood / compute_gradient_log_hyperparams,
    :191-425), plus its csapp.cpp -- compiled WHERE THEY LIE, unchanged, and linked against libcugp.so through
    include/cugp_shim/cuda_gp.h instead of the reference's cuda_gp.cu (cuSOLVER / cuBLAS)."""
    _ensure_lib()
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(REF, "cuda_bettersinglenode_ver2")
    exe = os.path.join(OUT, "cuda_bettersinglenode_main")
    _gxx(["-I" + INC, "-I" + os.path.join(REF, "cuda_src"), "-include", os.path.join(ROOT, "oracle", "ref_compat.h"),
          "-fpermissive", os.path.join(src, "main.cpp"), os.path.join(src, "cg_solver.cpp"), os.path.join(src, "csapp.cpp"),
          os.path.join(ROOT, "tests", "shim", "cuda_gp_link.cpp"), "-o", exe, "-lpthread", *_link_flags()])
    assert os.path.exists(exe)
    need = {l.split()[-1] for l in subprocess.run(["nm", "-u", exe], capture_output=True, text=True).stdout.splitlines() if "cugp_" in l}
    assert need, "the driver must reach the library through the facade"


def _run(exe, cwd=None, timeout=600, args=()):
    env = dict(os.environ, LD_LIBRARY_PATH=LIBDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe, *args], capture_output=True, text=True, cwd=cwd, env=env, timeout=timeout)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    return r.stdout


@pytest.mark.gpu
def test_shim_probe_runs_and_matches_python_mirror():
    import cugp_b200 as cg
    exe = os.path.join(OUT, "shim_probe")
    if not os.path.exists(exe):
        exe = _build_probe()
    out = _run(exe)
    val = {l.split()[0]: [float(x) for x in re.findall(r"-?\d+\.?\d*(?:e[-+]?\d+)?|nan|inf", l.split(None, 1)[1])]
           for l in out.splitlines() if l and l.split()[0].isupper()}
    # same data as shim_probe.cpp
    n, d, m = 96, 2, 8
    s, X = 12345, np.zeros((n + m, d))
    for i in range(n + m):
        for j in range(d):
            s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
            X[i, j] = -5.0 + 10.0 * (s >> 8) / 16777216.0
    y = np.sin(X[:, 0])
    th = [0.5, 0.25, -1.5]
    g = cg.Covsum(n, d)
    g.set_loghyperparam(th)
    assert val["LL"][0] == g.compute_loglikelihood(X[:n], y[:n])
    assert np.array_equal(val["GRAD"], g.compute_gradient_loghyperparam(X[:n], y[:n]))
    mu, var = g.compute_test_means_and_variances(X[:n], y[:n], X[n:])
    # (the probe predicts after compute_K_train dropped the factor, so it takes L^-T out of a fresh factorisation with the
    # identity rows; the mirror here inverts the factor of the log-likelihood call: same numbers up to rounding)
    assert np.allclose(val["PRED"], [mu[0], var[0]], rtol=1e-12, atol=0)
    assert val["NLPP"][1] == d == g.get_param_dim()   # Covsum::get_param_dim returns numdim (covkernel.cpp:661-663)
    assert val["CHOLDET"][2] < 1e-9          # forward + backward matrix substitution reproduce compute_K_inverse
    assert val["HOST"][0] < 1e-20            # K^-1 y (host helper on the GPU inverse) equals alpha
    b = cg.BCM(X[:n], y[:n], K=3, rank=0, world=1)
    b.set_BCM_log_hyperparam(th)
    bll, bg = b.loglik_and_gradient()
    assert val["BCM"][0] == bll and np.array_equal(val["BCM"][1:4], bg)
    assert val["BCM"][-1] == 3 * th[0]
    assert np.isfinite(val["CG"]).all() and np.isfinite(val["RPROP"]).all()


@pytest.mark.gpu
def test_reference_bcm_driver_runs_on_the_shim():
    """distributed_gp/distributed_ver1.cpp, compiled unchanged: 128 x 2 points, 4 experts, theta0 = 1.5, its own
    Polack-Ribiere loop.  Its final hyper-parameters ("PLEASE-SEE  3") must be the CPU oracle's to the 6 printed
    digits."""
    exe = os.path.join(OUT, "distributed_ver1")
    if not os.path.exists(exe):
        pytest.skip("driver binary not built (needs the reference sources; built in the CPU container)")
    from oracle import oracle
    from tests.conftest import load_data
    d = load_data("si128x2")
    with tempfile.TemporaryDirectory() as t:
        os.makedirs(os.path.join(t, "dataset"))
        os.makedirs(os.path.join(t, "run"))
        with open(os.path.join(t, "dataset", "input_128.txt"), "w") as f:
            f.write("128 2\n")
            for row in d["X"]:
                f.write(" ".join(repr(float(v)) for v in row) + "\n")
        with open(os.path.join(t, "dataset", "label_128.txt"), "w") as f:
            for v in d["y"]:
                f.write(repr(float(v)) + "\n")
        out = _run(exe, cwd=os.path.join(t, "run"))
    finals = re.findall(r"PLEASE-SEE\s+3:\s*(\S+), (\S+), (\S+)", out)
    assert finals, out[-2000:]
    got = [float(v) for v in finals[-1]]
    th, _, _ = oracle.port().cg_solve(d["X"], d["y"], [1.5, 1.5, 1.5], K=4)
    assert [f"{v:.6f}" for v in got] == [f"{v:.6f}" for v in th], (got, th)


@pytest.mark.gpu
def test_reference_gpu_flavour_driver_runs_on_the_facade():
    """The unchanged cuda_bettersinglenode_ver2 driver as master of a 1-worker run (main.cpp:189-210): setup() on the
    128 x 2 set (it asks for 8192 rows; setup clamps to the file), theta0 = 0.5, its own cg_solve.  The optimum
    it prints ("PLEASE-SEE  3", cg_solver.cpp:415) must be the CPU oracle's from the same start -- which is the optimum of
    the reference's own log REF:3183 (0.882908, 0.098703, -2.971479) to five digits."""
    exe = os.path.join(OUT, "cuda_bettersinglenode_main")
    if not os.path.exists(exe):
        pytest.skip("driver binary not built (needs the reference sources; built in the CPU container)")
    from oracle import oracle
    from tests.conftest import load_data
    d = load_data("si128x2")
    with tempfile.TemporaryDirectory() as t:
        os.makedirs(os.path.join(t, "chunked_dataset"))
        os.makedirs(os.path.join(t, "run"))
        with open(os.path.join(t, "chunked_dataset", "siproper_9192_10_chunk0.txt"), "w") as f:
            f.write("128 2\n")
            for row in d["X"]:
                f.write(" ".join(repr(float(v)) for v in row) + "\n")
        with open(os.path.join(t, "chunked_dataset", "siproper_9192_10_label0.txt"), "w") as f:
            for v in d["y"]:
                f.write(repr(float(v)) + "\n")
        out = _run(exe, cwd=os.path.join(t, "run"), args=["node0", "node0", "1"])
    finals = re.findall(r"PLEASE-SEE\s+3:\s*(\S+), (\S+), (\S+)", out)
    assert finals, out[-2000:]
    got = np.array([float(v) for v in finals[-1]])
    th, _, _ = oracle.port().cg_solve(d["X"], d["y"], [0.5, 0.5, 0.5], K=0)
    assert np.allclose(got, th, rtol=0, atol=2e-6), (got, th)
    assert np.allclose(got, [0.882908, 0.098703, -2.971479], rtol=0, atol=5e-6)      # REF:3183


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs (run under gpurun --gpus 2)")
def test_reference_bcm_driver_runs_on_two_gpus_without_python():
    """The same unchanged reference driver, started once per GPU with CUGP_RANK / CUGP_WORLD / CUGP_NCCL_ID_FILE: the shim's
    BCM shards the 4 experts over the ranks and the LIBRARY allreduces (ncclAllReduce, no torch, no Python) -- every rank
    must print the single-process optimum.  This is the socket master/worker layer of cuda_src/cg_solver.cpp:22-79
    replaced at the reference's own call surface."""
    exe = os.path.join(OUT, "distributed_ver1")
    if not os.path.exists(exe):
        pytest.skip("driver binary not built (needs the reference sources; built in the CPU container)")
    from oracle import oracle
    from tests.conftest import load_data
    d = load_data("si128x2")
    world = 2
    with tempfile.TemporaryDirectory() as t:
        os.makedirs(os.path.join(t, "dataset"))
        os.makedirs(os.path.join(t, "run"))
        with open(os.path.join(t, "dataset", "input_128.txt"), "w") as f:
            f.write("128 2\n")
            for row in d["X"]:
                f.write(" ".join(repr(float(v)) for v in row) + "\n")
        with open(os.path.join(t, "dataset", "label_128.txt"), "w") as f:
            for v in d["y"]:
                f.write(repr(float(v)) + "\n")
        procs = []
        for r in range(world):
            env = dict(os.environ, LD_LIBRARY_PATH=LIBDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""), CUGP_RANK=str(r),
                       CUGP_WORLD=str(world), CUGP_NCCL_ID_FILE=os.path.join(t, "nccl_id"))
            for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
                env.pop(k, None)
            procs.append(subprocess.Popen([exe], cwd=os.path.join(t, "run"), env=env, stdout=subprocess.PIPE,
                                          stderr=subprocess.PIPE, text=True))
        outs = [p.communicate(timeout=600) for p in procs]
    th, _, _ = oracle.port().cg_solve(d["X"], d["y"], [1.5, 1.5, 1.5], K=4)
    for r, (p, (out, err)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, (r, out[-1500:], err[-1500:])
        finals = re.findall(r"PLEASE-SEE\s+3:\s*(\S+), (\S+), (\S+)", out)
        assert finals, (r, out[-1500:], err[-1500:])
        got = [float(v) for v in finals[-1]]
        assert [f"{v:.6f}" for v in got] == [f"{v:.6f}" for v in th], (r, got, th)


@pytest.mark.gpu
def test_gpu_flavour_facade_reproduces_reference_logs():
    """cugp_shim/cuda_gp.h: setup / compute_log_likelihood / compute_gradient_log_hyperparams / set_loghyper_eigen /
    testing_phase (cuda_src/main.cpp:16-36).  setup() starts at theta = 0.5: the 128 x 2 set must give the LL of the
    reference's own log cuda_ref/seeee:19 (-203.386580), and at theta = 1.5 that of
    cuda_bettersinglenode_ver2/REF:33 (-319.512020), to the printed digits."""
    from tests.conftest import load_data, load_golden
    import cugp_b200 as cg
    exe = os.path.join(OUT, "shim_probe")
    if not os.path.exists(exe):
        exe = _build_probe()
    d = load_data("si128x2")
    gold = load_golden()
    with tempfile.TemporaryDirectory() as t:
        fi, fl = os.path.join(t, "in.txt"), os.path.join(t, "lab.txt")
        with open(fi, "w") as f:
            f.write("64 2\n")                               # the header lies about the row count
            for row in d["X"]:
                f.write(" ".join(repr(float(v)) for v in row) + " \n")
        with open(fl, "w") as f:
            for v in d["y"]:
                f.write(repr(float(v)) + "\n")
        out = _run(exe, args=[fi, fl, "128", "0"])
        out96 = _run(exe, args=[fi, fl, "96", "32"])
    val = {l.split()[0]: [float(x) for x in l.split()[1:] if x[0] in "-0123456789n"] for l in out.splitlines() if l.startswith("F")}
    assert f"{val['FLL'][0]:.6f}" == f"{gold['seeee_log']['ll']:.6f}"
    assert f"{val['FLL2'][0]:.6f}" == f"{gold['REF_log']['ll_sequence'][0]:.6f}"
    v96 = {l.split()[0]: [float(x) for x in l.split()[1:] if x[0] in "-0123456789n"] for l in out96.splitlines() if l.startswith("F")}
    g = cg.Covsum(96, 2)
    g.set_loghyperparam([1.5, 1.5, 1.5])
    mu, var = g.compute_test_means_and_variances(d["X"][:96], d["y"][:96], d["X"][96:])
    assert abs(v96["FNLPP"][0] - g.get_negative_log_predprob(d["y"][96:], mu, var)) <= 1e-12
