#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200, one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c3|c2|c4] [--n ROWS] [--impl reference]

Metric (BASELINE.json): "FP64 Cholesky TFLOP/s (% peak); loglik+grad evals/s; BCM pred pts/s at 1-8 GPU".
The default workload is C5, the configuration the Cholesky figure is quoted on and the largest that fits one
GPU: synthetic exact GP, n = 100 000, d = 10 (K is 80 GB of FP64).  A step is one training pass of the hot
path on device-resident inputs: covariance build (lower triangle) -> blocked Cholesky -> forward/backward
solves -> log-likelihood.  `value` = (n^3/3 flop) / step time, the flop count of the factorisation
(SURVEY.md section 8(d)); the `extra` block carries the other two parts of the metric (C2 evals/s, C4 pts/s).

  --workload c3  exact GP n = 10 000 (train + predict), same metric
  --workload c2  n = 4096 hyper-parameter loop: value = (LL + gradient) evaluations / s
  --workload c4  BCM 16 x 1500 on N ranks (experts sharded, NCCL allreduce of the moments): value = pts/s

Multi-GPU: the exact GP does not shard -- N ranks run N independent replicas
("weak"); c4 shards the 16 experts over the ranks ("strong").  Timing: W >= 3 warm-ups, then K steps between
barrier + cuda synchronize, device time with CUDA events, max over ranks.  Inputs of every workload are far
larger than the 126 MB L2 except c2/c4, where an L2 flush (256 MB write) runs between steps.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TH_B = [3.762111, -1.152105, -0.384461]   # trained values, cuda_src/main.cpp:191-193
# The reference's CPU path cannot run the GPU sizes (n = 100 000 needs 560 GB of host buffers and ~10^5 core-hours), so
# the CPU legs time it on the first REF_SAMPLE_N rows of the SAME synthetic set.  ONE size per workload, used by both the
# `cpu_baseline` leg and `--impl reference`; every CPU line says so (`same_config: false`) -- a flop-rate or
# evaluation-rate ratio across sizes is a stated baseline, not a like-for-like speed-up.
REF_SAMPLE_N = {"c5": 2048, "c3": 1024, "c2": 1024, "c4": 1500}
TH_C = [2.0, 2.0, 2.0]                    # cuda_scalingdist/main.cpp:298-301
NOMINAL_FP64_TFLOPS = 37.0                # HGX B200 data sheet, dense FP64 (tensor = vector)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def dist_setup(n_gpus: int):
    """One process per GPU (torchrun env).  Returns (rank, world, torch, dist or None)."""
    import torch
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local, torch, dist


def timed_steps(torch, dist, step, steps, warmup, flush=None):
    """W warm-ups, then K steps between barrier + synchronize; device time via CUDA events; max over ranks."""
    for _ in range(warmup):
        step()
        if flush is not None:
            flush()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total_ms = 0.0
    t_wall = time.perf_counter()
    for _ in range(steps):
        # the library runs on its own stream and synchronises before returning, so events recorded on the
        # current stream around the (blocking) call bracket the device work of the step
        e0.record()
        step()
        e1.record()
        e1.synchronize()
        total_ms += e0.elapsed_time(e1)
        if flush is not None:
            flush()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    ms = total_ms / steps
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, wall_ms / steps


def make_flush(torch):
    buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush():
        buf.add_(1)  # write 256 MB > 126 MB L2

    return flush


# --------------------------------------------------------------------------------------------------------------
def cpu_baseline(kind_pref: str, n_s: int, theta, reps: int = 1):
    """The reference's CPU path timed on this host: compute_loglikelihood (K build + Cholesky + solves) on a
    bounded sample of the same synthetic workload; TFLOP/s at n_s^3/3.  Single thread: the reference has no
    threading."""
    from cugp_b200.loaders import synthetic_sine
    from oracle import oracle
    impl = oracle.reference() if kind_pref == "reference" else None
    if impl is None:
        impl = oracle.port()
    X, y = synthetic_sine(n_s, 10)
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        ll = impl.loglik(X, y, theta)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return {"value": (n_s ** 3 / 3) / best / 1e12, "unit": "TFLOP/s", "cores": 1, "kind": impl.kind, "same_config": False,
            "sample_n": n_s,
            "sample": f"one compute_loglikelihood on the first {n_s} rows of the same synthetic set, theta_B, "
                      f"{best:.2f} s, LL={ll:.6f}; host has {os.cpu_count()} cores, reference is single-threaded"}


def cpu_c3_sample(impl, n_s: int):
    """C3 (train + predict) on a bounded sample: compute_loglikelihood, then compute_test_means_and_variances of n_s test
    points (the GPU arm's n : m = 1 : 1), with the GPU arm's flop count 2 n^3/3 + n^2 m."""
    from cugp_b200.loaders import synthetic_sine
    X, y = synthetic_sine(2 * n_s, 10)
    Xt, X, y = X[n_s:], X[:n_s], y[:n_s]
    t = time.perf_counter()
    impl.loglik(X, y, TH_B)
    impl.predict(X, y, TH_B, Xt)
    dt = time.perf_counter() - t
    fl = 2.0 * n_s ** 3 / 3 + float(n_s) ** 2 * n_s
    return {"value": fl / dt / 1e12, "unit": "TFLOP/s", "cores": 1, "kind": impl.kind, "same_config": False, "sample_n": n_s,
            "sample": f"compute_loglikelihood + compute_test_means_and_variances of {n_s} points at n={n_s} (of 10000 / 10000), "
                      f"theta_B, {dt:.2f} s; flops counted as on the GPU arm (2 n^3/3 + n^2 m)"}, dt


def cpu_c4_model(impl, m: int, m_lo: int = 4, m_hi: int = 260):
    """The reference's BCM prediction at the GPU arm's OWN m: every expert pays its training (three factorisations and
    the inverse, covkernel.cpp:277-296) once and then O(n^2) per test point (covkernel.cpp:297-302).  Both parts are
    timed on expert 0 (1500 rows): T(m_lo) and T(m_hi) give the per-point time and the training time; the ensemble
    is 16 such experts in sequence (BCM.cpp:64-83), so pts/s = m / (16 (T_train + m t_point))."""
    d4 = np.load(os.path.join(ROOT, "tests", "golden", "data_si24000.npz"))
    Xt = np.load(os.path.join(ROOT, "tests", "golden", "data_c4_xtest600.npz"))["Xtest"]
    X, y = d4["X"][:1500], d4["y"][:1500]
    t = time.perf_counter()
    impl.predict(X, y, TH_C, Xt[:m_lo])
    t_lo = time.perf_counter() - t
    t = time.perf_counter()
    impl.predict(X, y, TH_C, Xt[:m_hi])
    t_hi = time.perf_counter() - t
    t_point = max((t_hi - t_lo) / (m_hi - m_lo), 1e-9)
    t_train = max(t_lo - m_lo * t_point, 0.0)
    return {"value": m / (16.0 * (t_train + m * t_point)), "t_train_s": t_train, "t_point_s": t_point,
            "sample": f"expert 0 of 16 (1500 rows), theta_C: train {t_train:.2f} s + {1e3 * t_point:.2f} ms per test point "
                      f"(from {m_lo} and {m_hi} points); ensemble = 16 experts in sequence at m = {m}"}


def run_reference_arm(a):
    """--impl reference: the reference's own CPU implementation, bounded sample per step."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from cugp_b200.loaders import synthetic_sine
    from oracle import oracle
    impl = oracle.reference() or oracle.port()
    n_s = REF_SAMPLE_N[a.workload]
    lohi = (-20.0, 20.0, 0.05) if a.workload == "c2" else (-10.0, 10.0, 0.1)
    X, y = synthetic_sine(n_s + 64, 10, lo=lohi[0], hi=lohi[1], noise=lohi[2])
    Xt, X, y = X[n_s:], X[:n_s], y[:n_s]
    model = []

    def step():
        if a.workload == "c3":
            model.append(cpu_c3_sample(impl, n_s)[0])
        elif a.workload == "c5":
            impl.loglik(X, y, TH_B)
        elif a.workload == "c2":
            impl.loglik(X, y, TH_B)
            impl.grad(X, y, TH_B)
        else:
            model.append(cpu_c4_model(impl, a.m, 4, 68))

    for _ in range(a.warmup):
        step()
    model.clear()
    t = time.perf_counter()
    for _ in range(a.steps):
        step()
    ms = (time.perf_counter() - t) * 1e3 / a.steps
    if a.workload == "c3":
        metric, unit = "fp64_train_predict_tflops", "TFLOP/s"
        value = statistics.median(mm["value"] for mm in model)
        sample = model[-1]["sample"]
        same = False
    elif a.workload == "c5":
        metric, unit, value = "fp64_cholesky_tflops", "TFLOP/s", (n_s ** 3 / 3) / (ms * 1e-3) / 1e12
        sample = f"compute_loglikelihood, n={n_s} synthetic rows (of {a.n}), theta_B"
        same = False
    elif a.workload == "c2":
        metric, unit, value = "loglik_grad_evals_per_s", "evals/s", 1e3 / ms
        sample = f"compute_loglikelihood + compute_gradient_loghyperparam, n={n_s} (of 4096), theta_B"
        same = False
    else:
        metric, unit = "bcm_pred_pts_per_s", "pts/s"
        value = statistics.median(mm["value"] for mm in model)
        sample = model[-1]["sample"]
        same = True
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": a.workload, "sample": sample, "sample_n": n_s,
                                                              "same_config": same},
            "cpu_baseline": {"value": value, "unit": unit, "cores": 1, "kind": impl.kind, "sample": sample, "sample_n": n_s,
                             "same_config": same},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------
def fp64_probes(lib):
    """Measured FP64 denominators: DMMA burst (10 ms launch) and sustained (2 s launch) with the SM clock each ran
    at, DFMA for context, and a device copy."""
    out = {}
    for name, ms in (("dmma_burst", 10.0), ("dmma_sustained", 2000.0)):
        tf, mhz = C.c_double(), C.c_double()
        lib.cugp_probe_dmma(ms, C.byref(tf), C.byref(mhz))
        out[name] = tf.value
        out[name + "_sm_mhz"] = mhz.value
    dm, df, cp = C.c_double(), C.c_double(), C.c_double()
    lib.cugp_probe_fp64_peak(50.0, C.byref(dm), C.byref(df))
    lib.cugp_probe_copy(1 << 30, 10, C.byref(cp))
    out["dfma_probe"] = df.value
    return out, cp.value


def extra_c2(cg, torch, flush, n=4096, evals=6):
    """loglik+grad evals/s on the C2 shape (synthetic rows, theta perturbed per evaluation)."""
    from cugp_b200.loaders import synthetic_sine
    X, y = synthetic_sine(n, 10, lo=-20.0, hi=20.0, noise=0.05)
    g = cg.Covsum(n, 10)
    g.set_data(X, y)
    ts = []
    for i in range(evals + 2):
        g.set_loghyperparam([TH_B[0] + 1e-7 * i, TH_B[1], TH_B[2]])
        torch.cuda.synchronize()
        t = time.perf_counter()
        g.loglik_resident()
        g.grad_resident()
        ts.append(time.perf_counter() - t)
        flush()
    g.close()
    return 1.0 / statistics.median(ts[2:])


def bcm_parity(cg, torch, dist):
    """Same-run parity of the sharded ensemble on EVERY rank: C4 (si24000, 16 experts, theta_C) against the golden values
    generated from the unmodified reference (tests/golden/golden_c4.json) at the north-star tolerances."""
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_c4.json")))["cases"]["C4_si24000_bcm16_thC_pred16"]
    d = np.load(os.path.join(ROOT, "tests", "golden", "data_si24000.npz"))
    b = cg.BCM(d["X"], d["y"], K=16)
    b.set_BCM_log_hyperparam(gold["theta"])
    ll, g = b.loglik_and_gradient()
    mu, var = b.compute_BCM_test_means_and_var(d["Xtest"][:gold["m"]])
    in_lib = bool(getattr(b, "_in_library", False))
    kind = b.exchange_kind
    b.close()
    gref = np.array(gold["grad"])
    e = {"ll": abs(ll - gold["ll"]) / abs(gold["ll"]),
         "grad": float(np.max(np.abs(g - gref) / np.maximum(np.abs(gref), 1e-3 * np.abs(gref).max()))),
         "mean": float(np.max(np.abs(mu - gold["mean"]) / np.maximum(np.abs(gold["mean"]), 1e-6))),
         "var": float(np.max(np.abs(var - gold["var"]) / np.abs(gold["var"])))}
    ok = e["ll"] <= 1e-9 and e["grad"] <= 1e-9 and e["mean"] <= 1e-8 and e["var"] <= 1e-8
    worst = [e["ll"], e["grad"], e["mean"], e["var"], 0.0 if ok else 1.0]
    if dist is not None:
        t = torch.tensor(worst, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)          # the WORST rank decides
        worst = [float(v) for v in t.tolist()]
    return {"parity_ok_all_ranks": worst[4] == 0.0, "golden": "tests/golden/golden_c4.json (unmodified reference)",
            "max_rel_err_over_ranks": {"ll": worst[0], "grad": worst[1], "mean": worst[2], "var": worst[3]},
            "tolerance": {"ll": 1e-9, "grad": 1e-9, "mean": 1e-8, "var": 1e-8}, "exchange_inside_library": in_lib,
            "exchange": kind}


def extra_c4(cg, torch, dist, flush, m=10000, reps=6, experts=16, prefix="c4"):
    """BCM 16 x 1500 on ALL ranks of this run (experts e % world == rank, NCCL allreduce of the moments):
    prediction pts/s with factorised experts resident, and (LL, gradient) evaluations/s; theta_C.
    experts != 16: the same generator at `experts` x 1500 synthetic rows (an ensemble large enough that eight GPUs still
    hold several experts each)."""
    from cugp_b200.loaders import synthetic_sine
    if experts == 16:
        d = np.load(os.path.join(ROOT, "tests", "golden", "data_si24000.npz"))
        Xall, yall = d["X"], d["y"]
    else:
        Xall, yall = synthetic_sine(experts * 1500, 10, seed=11)
    Xt, _ = synthetic_sine(m, 10, seed=7)
    b = cg.BCM(Xall, yall, K=experts)
    b.set_BCM_log_hyperparam(TH_C)
    b.loglik_and_gradient()
    b.compute_BCM_test_means_and_var(Xt)

    def timed(fn):
        ts = []
        for _ in range(reps):
            flush()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            t = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t)
        v = statistics.median(ts[1:])
        if dist is not None:
            tt = torch.tensor([v], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            v = float(tt.item())
        return v

    k = [0]

    def ev():
        k[0] += 1
        b.set_BCM_log_hyperparam([TH_C[0] + 1e-7 * k[0], TH_C[1], TH_C[2]])
        b.loglik_and_gradient()

    t_pred = timed(lambda: b.compute_BCM_test_means_and_var(Xt))
    t_eval = timed(ev)
    world = b.world
    b.close()
    return {f"{prefix}_bcm_pred_pts_per_s": m / t_pred, f"{prefix}_bcm_loglik_grad_evals_per_s": 1.0 / t_eval,
            f"{prefix}_pred_ms": 1e3 * t_pred, f"{prefix}_eval_ms": 1e3 * t_eval,
            f"{prefix}_gpus": world, f"{prefix}_test_points": m, f"{prefix}_experts_x_rows": f"{experts} x 1500"}


def bcm_block(cg, torch, dist, flush, m=10000):
    """The part of the metric that shards (BASELINE.json: 'BCM pred pts/s at 1-8 GPU'), measured on the ranks of THIS run:
    C4 itself and the same shape at 64 experts, with the same-run parity flag."""
    out = {"workload": "C4 BCM si24000 16 experts x 1500, theta_C; experts e % world per rank; ONE ncclAllReduce per "
                       "operation (4 doubles per evaluation, 2 m doubles per prediction) issued by the library",
           "timing": "median of 5 calls, max over ranks, host wall clock around the blocking call (device work + exchange + "
                     "result copy), 256 MB L2 flush before every call"}
    out["parity"] = bcm_parity(cg, torch, dist)
    out.update(extra_c4(cg, torch, dist, flush, m=m))
    out.update(extra_c4(cg, torch, dist, flush, m=m, experts=64, prefix="c4x4"))
    return out


def library_baseline(cg, torch, sizes=(10000, 20000)):
    """Informative same-box LIBRARY baseline: the reference's own GPU generation -- cuSOLVER potrf,
    cuBLAS trsm / gemm, built unchanged for sm_100 (oracle/_ref/libcugp_refgpu.so) -- and a plain cuSOLVER Cholesky
    (torch.linalg.cholesky) next to this library at the same n and theta.  Each reference size runs in its own process
    (it keeps global device state and never frees its workspaces).  Never on the product path."""
    from cugp_b200.loaders import synthetic_sine
    out = []
    for n in sizes:
        row = {"n": n}
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "refgpu_time.py"), str(n)], capture_output=True,
                               text=True, timeout=600)
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            row["reference_gpu"] = json.loads(lines[-1]) if lines else {"error": (r.stderr or r.stdout)[-300:]}
        except Exception as e:
            row["reference_gpu"] = {"error": repr(e)}
        try:
            X, y = synthetic_sine(n, 10)
            g = cg.Covsum(n, 10)
            g.set_data(X, y)
            ts, tg, tc = [], [], []
            for i in range(4):
                g.set_loghyperparam([TH_B[0] + 1e-7 * i, TH_B[1], TH_B[2]])
                t = time.perf_counter()
                ll = g.loglik_resident()
                ts.append(time.perf_counter() - t)
                t = time.perf_counter()
                gr = g.grad_resident()
                tg.append(time.perf_counter() - t)
            for i in range(2):
                g.set_loghyperparam([TH_B[0] + 1e-7 * (i + 9), TH_B[1], TH_B[2]])
                tc.append(g.factorize_resident()[1])
            g.set_loghyperparam(TH_B)
            row["this_library"] = {"loglik_ms": 1e3 * min(ts[1:]), "grad_after_loglik_ms": 1e3 * min(tg[1:]),
                                   "cholesky_ms": min(tc), "ll": g.loglik_resident(), "grad": [float(v) for v in g.grad_resident()]}
            K = torch.from_numpy(g.compute_K_train(X)).cuda() if n <= 20000 else None
            g.close()
            if K is not None:
                torch.linalg.cholesky(K)
                torch.cuda.synchronize()
                t = time.perf_counter()
                torch.linalg.cholesky(K)
                torch.cuda.synchronize()
                row["cusolver_potrf_ms"] = 1e3 * (time.perf_counter() - t)
                del K
        except Exception as e:
            row["this_library"] = {"error": repr(e)}
        rg, tl = row.get("reference_gpu", {}), row.get("this_library", {})
        if "loglik_ms" in rg and "loglik_ms" in tl:
            row["speedup_loglik"] = rg["loglik_ms"] / tl["loglik_ms"]
            row["speedup_loglik_plus_grad"] = (rg["loglik_ms"] + rg["grad_ms"]) / (tl["loglik_ms"] + tl["grad_after_loglik_ms"])
            row["ll_rel_diff"] = abs(rg["ll"] - tl["ll"]) / abs(tl["ll"])
        out.append(row)
    return out


def correctness_block(cg, L, X, y, n):
    """Outside the timed region: is the factorisation of THIS size right?  (The reference printed a Cholesky residual
    after every factorisation, cuda_src/cuda_gp.cu:1126-1139.)  residual: ||K alpha - y|| / ||y|| with K rebuilt matrix-free
    (K itself was overwritten by L); logdet_crosscheck: log det K and y'K^-1 y from a second factorisation with another
    outer block width and no look-ahead -- a different launch schedule and update order over the same kernels."""
    g = cg.Covsum(n, 10)
    g.set_data(X, y)
    g.set_loghyperparam(TH_B)
    q, ld, ll = g.scalars_resident()
    r = g.residual_resident()
    res = float(np.linalg.norm(r) / np.linalg.norm(y))
    a = g.alpha_resident()
    nb_default = 1024 if n >= 24576 else 512 if n >= 9000 else 256 if n >= 5000 else 128
    nb_alt = 512 if nb_default != 512 else 256
    try:
        L.cugp_set_tuning(b"potrf_nb", nb_alt)
        L.cugp_set_tuning(b"lookahead", 0)
        g.factorize_resident()
        q2, ld2, _ = g.scalars_resident()
    finally:
        L.cugp_set_tuning(b"potrf_nb", 0)
        L.cugp_set_tuning(b"lookahead", 1)
    g.close()
    return {"n": n, "theta": TH_B, "c5_residual": res, "residual_bound": 1e-10,
            "quad_vs_y_dot_alpha_rel": float(abs(q - y @ a) / abs(q)),
            "c5_logdet_crosscheck": {"logdet": ld, "logdet_alt": ld2, "rel_diff": float(abs(ld - ld2) / abs(ld)),
                                     "quad": q, "quad_alt": q2, "quad_rel_diff": float(abs(q - q2) / abs(q)),
                                     "schedules": f"outer width {nb_default} + look-ahead vs outer width {nb_alt}, single stream"},
            "ok": bool(res <= 1e-10 and abs(ld - ld2) <= 1e-11 * abs(ld) and abs(q - q2) <= 1e-10 * abs(q))}


def extra_f3(cg, torch, dist, flush, chunks=32, n=2000, d=7, slots=8, passes=3):
    """Shard streaming on the shape of the reference's scaling harness (d = 7 'flight' shards,
    cuda_scalingdist/harness.sh:53-57; here 32 x 2000 synthetic rows): shard TEXT files -> reader thread -> pinned double
    buffers -> copy stream -> `slots` experts per launch.  Pass 1 parses the text, later passes hit the host cache; the
    reported rate is the median cached pass, reader_wait_ms says how much of the parse was NOT hidden behind the GPU."""
    import shutil
    import tempfile
    from cugp_b200.loaders import synthetic_sine
    rank = dist.get_rank() if dist is not None else 0
    tmp = os.path.join(tempfile.gettempdir(), "cugp_f3_shards")
    if rank == 0:
        shutil.rmtree(tmp, ignore_errors=True)
        os.makedirs(tmp)
        X, y = synthetic_sine(chunks * n, d, seed=3)
        for i in range(chunks):
            np.savetxt(os.path.join(tmp, f"in_{i}.txt"), X[i * n:(i + 1) * n], fmt="%.17g", header=f"{n} {d}", comments="")
            np.savetxt(os.path.join(tmp, f"lab_{i}.txt"), y[i * n:(i + 1) * n], fmt="%.17g")
    if dist is not None:
        dist.barrier()
    # dist is None: this process alone walks all shards (rank 0 / world 1 stated explicitly: a process group may still
    # be initialised when rank 0 runs its rank-0-only legs)
    rw = {} if dist is not None else {"rank": 0, "world": 1}
    s = cg.ShardStream.from_files(os.path.join(tmp, "in_"), os.path.join(tmp, "lab_"), chunks, n, d, slots=slots, **rw)
    ts = []
    for i in range(passes + 1):
        s.set_BCM_log_hyperparam([TH_C[0] + 1e-7 * i, TH_C[1], TH_C[2]])
        flush()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        s.loglik_and_gradient()
        ts.append(time.perf_counter() - t)
    st, lay = s.stats(), s.layout()
    s.close()
    v = statistics.median(ts[1:])
    if dist is not None:
        tt = torch.tensor([v, ts[0]], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        v, first = float(tt[0].item()), float(tt[1].item())
        dist.barrier()
    else:
        first = ts[0]
    if rank == 0:
        shutil.rmtree(tmp, ignore_errors=True)
    return {"f3_stream_shape": f"{chunks} shards x {n} rows x d={d}, {lay['slots']} slots, {lay['groups']} groups/pass on rank 0",
            "f3_stream_loglik_grad_evals_per_s": 1.0 / v, "f3_first_pass_s_incl_text_parse": first,
            "f3_parse_ms_rank0": st["parse_ms"], "f3_reader_wait_ms_rank0": st["reader_wait_ms"],
            "f3_h2d_bytes_per_pass_rank0": st["h2d_bytes"] / max(st["passes"], 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c5", choices=["c5", "c3", "c2", "c4"])
    ap.add_argument("--n", "--rows", dest="n", type=int, default=None,
                    help="override the row count of the workload (use --rows under torchrun: its parser claims --n)")
    ap.add_argument("--m", type=int, default=10000, help="test points (c3/c4)")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (scaling sweeps: it is rank-0 only work)")
    a = ap.parse_args()
    a.n = a.n or {"c5": 100000, "c3": 10000, "c2": 4096, "c4": 24000}[a.workload]
    a.warmup = max(a.warmup, 3) if a.impl == "native" else a.warmup
    if a.impl == "reference":
        return run_reference_arm(a)

    rank, world, local, torch, dist = dist_setup(a.gpus)
    import cugp_b200 as cg
    from cugp_b200._lib import check, lib, ptr
    from cugp_b200.loaders import synthetic_sine
    L = lib()
    check(L.cugp_set_device(local))
    pk, pk_src = peaks()
    flush = make_flush(torch)
    sampler = ClockSampler(local)
    n = a.n
    out = {}
    step_i = [0]

    if a.workload in ("c5", "c3", "c2"):
        lohi = (-20.0, 20.0, 0.05) if a.workload == "c2" else (-10.0, 10.0, 0.1)
        X, y = synthetic_sine(n + a.m, 10, lo=lohi[0], hi=lohi[1], noise=lohi[2])   # rng seed 15618
        Xt, X, y = X[n:], np.ascontiguousarray(X[:n]), np.ascontiguousarray(y[:n])
        g = cg.Covsum(n, 10)
        g.set_data(X, y)

        def theta_i():
            step_i[0] += 1
            return [TH_B[0] + 1e-7 * step_i[0], TH_B[1], TH_B[2]]     # new theta: nothing cached is reused

        if a.workload == "c2":
            def step():
                g.set_loghyperparam(theta_i())
                g.loglik_resident()
                g.grad_resident()

            def step_e2e():
                g.set_loghyperparam(theta_i())
                g.compute_loglikelihood(X, y)
                g.compute_gradient_loghyperparam(X, y)
            d2h = 8 + 24
        elif a.workload == "c3":
            def step():
                g.set_loghyperparam(theta_i())
                g.loglik_resident()
                g.compute_test_means_and_variances(X, y, Xt)

            step_e2e = step
            d2h = 8 + 16 * a.m
        else:
            def step():
                g.set_loghyperparam(theta_i())
                g.loglik_resident()
                g.solve_resident()      # alpha = K^-1 y (the reference's compute_loglikelihood solves for it too)

            def step_e2e():
                g.set_loghyperparam(theta_i())
                g.compute_loglikelihood(X, y)
                g.alpha_resident()
            d2h = 8 + 8 * n
        h2d = X.nbytes + y.nbytes + (Xt.nbytes if a.workload == "c3" else 0)
        use_flush = flush if a.workload == "c2" else None

        g.profile(True)
        sampler.start()
        L.cugp_launch_count_reset()
        ms, wall_ms = timed_steps(torch, dist, step, a.steps, a.warmup, use_flush)
        launches = L.cugp_launch_count()
        clocks = sampler.stop()
        # dominant kernel (SYRK trailing update): launches of the LAST (warmup + steps) passes were bracketed
        syrk_ms, syrk_flops, syrk_cnt = g.profile_read()
        g.profile(False)
        ms_e2e, _ = timed_steps(torch, dist, step_e2e, 1 if n >= 50000 else max(1, min(a.steps, 3)), 0, use_flush)

        if a.workload == "c2":
            metric, unit, per_rank, per_rank_e2e = "loglik_grad_evals_per_s", "evals/s", 1e3 / ms, 1e3 / ms_e2e
        elif a.workload == "c3":
            # the step is train + predict: Cholesky n^3/3, T = L^-1 n^3/3, V = T K*^T (triangular k-ranges) n^2 m
            metric, unit = "fp64_train_predict_tflops", "TFLOP/s"
            fl = 2.0 * n ** 3 / 3 + float(n) ** 2 * a.m
            per_rank, per_rank_e2e = fl / (ms * 1e-3) / 1e12, fl / (ms_e2e * 1e-3) / 1e12
        else:
            metric, unit = "fp64_cholesky_tflops", "TFLOP/s"
            per_rank, per_rank_e2e = (n ** 3 / 3) / (ms * 1e-3) / 1e12, (n ** 3 / 3) / (ms_e2e * 1e-3) / 1e12
        out.update(metric=metric, unit=unit, value=per_rank * world, ms_per_step=ms, scaling="weak",
                   e2e={"value": per_rank_e2e * world, "unit": unit, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h)},
                   gpu_launches=int(launches * a.steps / (a.steps + a.warmup)))
        # factorisation alone (events inside the library), for the record
        g.set_loghyperparam(theta_i())
        ms_cov, ms_chol = g.factorize_resident()
        ms_solve = g.solve_resident()
        tri_bytes = 4.0 * n * (n + 1)                  # lower triangle incl. diagonal, 8 B each
        out["phases_ms"] = {"covariance": ms_cov, "cholesky": ms_chol, "solves": ms_solve,
                            "cholesky_tflops": (n ** 3 / 3) / (ms_chol * 1e-3) / 1e12,
                            "covariance_gbs": tri_bytes / (ms_cov * 1e-3) / 1e9,       # K1 writes the lower triangle once
                            "solves_gbs": tri_bytes / (ms_solve * 1e-3) / 1e9}   # K3 backward sweep streams L once (the
                                                                                  # forward substitution is fused into K2)
        config = {"workload": {"c5": f"C5 synthetic exact GP n={n} d=10: covariance build + blocked Cholesky + solves + LL",
                               "c3": f"C3 exact GP n={n} d=10 train + predict {a.m} points per step (flops counted: "
                                     f"Cholesky n^3/3 + inverse factor n^3/3 + V = L^-1 K*^T n^2 m; the Cholesky alone is "
                                     f"phases_ms.cholesky_tflops)",
                               "c2": f"C2 exact GP n={n} d=10 hyper-parameter loop: LL + gradient per evaluation"}[a.workload],
                  "n": n, "d": 10, "theta": TH_B, "parallelism": f"replicas x{world} (exact GP does not shard)",
                  "l2": "256 MB flush between steps" if use_flush else "inputs (8 n^2 B) exceed L2"}
        g.close()
        del g
    else:  # c4: BCM, experts sharded over ranks, NCCL allreduce of the moments
        d = np.load(os.path.join(ROOT, "tests", "golden", "data_si24000.npz"))
        Xt, _ = synthetic_sine(a.m, 10, seed=7)
        b = cg.BCM(d["X"], d["y"], K=16)

        def step():
            step_i[0] += 1
            b.set_BCM_log_hyperparam([TH_C[0] + 1e-7 * step_i[0], TH_C[1], TH_C[2]])
            b.loglik_and_gradient()
            b.compute_BCM_test_means_and_var(Xt)

        sampler.start()
        L.cugp_launch_count_reset()
        ms, wall_ms = timed_steps(torch, dist, step, a.steps, a.warmup, flush)
        launches = L.cugp_launch_count()
        clocks = sampler.stop()
        syrk_ms = syrk_flops = syrk_cnt = 0
        out.update(metric="bcm_pred_pts_per_s", unit="pts/s", value=a.m / (ms * 1e-3), ms_per_step=ms, scaling="strong",
                   e2e={"value": a.m / (ms * 1e-3), "unit": "pts/s", "h2d_bytes_per_step": int(Xt.nbytes),
                        "d2h_bytes_per_step": int(16 * a.m + 32)},
                   gpu_launches=int(launches * a.steps / (a.steps + a.warmup)))
        config = {"workload": f"C4 BCM si24000 16 experts x 1500, theta_C: LL+gradient eval then predict {a.m} points per step",
                  "n": 24000, "d": 10, "experts": 16, "parallelism": f"experts e%{world} per rank, allreduce(4) + allreduce(2m)",
                  "l2": "256 MB flush between steps"}
        b.close()
        # the two halves of the step on their own: prediction with the factorised experts resident (the metric's name)
        # and the (LL, gradient) evaluation, each max-over-ranks
        # (also at 64 experts -- 96 000 rows -- where eight GPUs still hold eight experts each), with the same-run parity flag
        out["bcm"] = bcm_block(cg, torch, dist, flush, m=a.m)

    extra = None
    if not a.no_extra and a.workload == "c5":   # collective: every rank takes part in the sharded BCM
        try:
            out["bcm"] = bcm_block(cg, torch, dist, flush)
        except Exception as e:  # the headline must still print
            out["bcm"] = {"error": repr(e)}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    fp64, copy_gbs = fp64_probes(L)
    dmma = fp64["dmma_sustained"]
    # cuBLAS DGEMM as a library ceiling for context only (never on the product path)
    try:
        A = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        torch.matmul(A, A)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(3):
            torch.matmul(A, A)
        torch.cuda.synchronize()
        cublas = 3 * 2 * 8192 ** 3 / (time.perf_counter() - t) / 1e12
        del A
    except Exception:
        cublas = None
    if syrk_cnt:
        achieved = syrk_flops / (syrk_ms * 1e-3) / 1e12
        traffic = traffic_detail = None
        tp = os.path.join(ROOT, "profiles", "top_kernel_traffic.json")
        if os.path.exists(tp) and a.workload == "c5":   # the capture is of C5's own first full-width update
            # one ncu capture of ONE launch of this kernel (the first full-width update of C5 itself): its DRAM bytes next to
            # the algorithmic bytes of that same launch
            traffic_detail = json.load(open(tp))
            traffic = traffic_detail.get("dram_bytes_total")
        out["roofline"] = {"bound": "tensor", "kernel": "dgemm_ws_kernel<128,128> (Cholesky trailing update A22 -= P P^T, lower "
                                                        "tiles, K = outer block width)",
                           "achieved": achieved, "peak": dmma, "unit": "TFLOP/s", "frac": achieved / dmma, "traffic": traffic,
                           "traffic_detail": traffic_detail,
                           "peak_source": "measured here, sustained: one 2 s launch of register-resident mma.sync.m8n8k4.f64 "
                                          "(cugp_probe_dmma); MEASURED_PEAKS.json has no FP64 entry",
                           "frac_of_nominal_37": achieved / NOMINAL_FP64_TFLOPS,
                           "launches_timed": syrk_cnt, "share_of_step": syrk_ms / ((a.steps + a.warmup) * ms),
                           "algorithmic_flops_per_launch": "m(m+1)*K for an m x m trailing block (lower triangle, 2 flop/MAC)"}
    if a.workload in ("c5", "c3") and not a.no_extra:
        out["correctness"] = correctness_block(cg, L, X, y, n)
        try:
            out["library_baseline"] = library_baseline(cg, torch)
        except Exception as e:
            out["library_baseline"] = {"error": repr(e)}
    out["fp64_peaks_tflops"] = dict(fp64, cublas_dgemm_8192=cublas, nominal=NOMINAL_FP64_TFLOPS)
    out["hbm"] = {"copy_probe_gbs": copy_gbs, "peak_gbs": pk.get("hbm_gbs"), "peak_source": pk_src}
    if "phases_ms" in out and pk.get("hbm_gbs"):
        out["phases_ms"]["covariance_frac_of_hbm"] = out["phases_ms"]["covariance_gbs"] / pk["hbm_gbs"]
        out["phases_ms"]["solves_frac_of_hbm"] = out["phases_ms"]["solves_gbs"] / pk["hbm_gbs"]
    if not a.no_extra and a.workload == "c5":
        try:
            extra = dict(extra or {}, c2_loglik_grad_evals_per_s=extra_c2(cg, torch, flush))
        except Exception as e:
            extra = dict(extra or {}, c2_error=repr(e))
        try:
            # rank 0 only from here on (the other ranks have left): the stream runs un-sharded, no collective
            extra.update(extra_f3(cg, torch, None, flush))
        except Exception as e:
            extra["f3_error"] = repr(e)
        out["extra"] = extra
    n_s = REF_SAMPLE_N[a.workload]
    skip_cpu = a.no_cpu or world != 1          # the CPU baseline is a rank-0, N = 1 leg
    if skip_cpu:
        cpu = {"value": None, "unit": None, "cores": 0, "kind": "skipped",
               "sample": "--no-cpu" if a.no_cpu else "reported at N=1 only"}
    elif a.workload == "c5":
        cpu = cpu_baseline("reference", n_s, TH_B)
    elif a.workload == "c3":
        from oracle import oracle
        cpu = cpu_c3_sample(oracle.reference() or oracle.port(), n_s)[0]
    if skip_cpu:
        pass
    elif a.workload == "c2":
        from oracle import oracle
        impl = oracle.reference() or oracle.port()
        Xs, ys = synthetic_sine(n_s, 10, lo=-20.0, hi=20.0, noise=0.05)
        t = time.perf_counter()
        impl.loglik(Xs, ys, TH_B)
        impl.grad(Xs, ys, TH_B)
        dt = time.perf_counter() - t
        cpu = {"value": 1.0 / dt, "unit": "evals/s", "cores": 1, "kind": impl.kind, "same_config": False, "sample_n": n_s,
               "sample": f"one LL + gradient at n={n_s} (of {n}), theta_B, {dt:.1f} s"}
    elif a.workload == "c4":
        # like for like with the GPU arm's m: training once per expert + m per-point costs (cpu_c4_model)
        from oracle import oracle
        impl = oracle.reference() or oracle.port()
        mod = cpu_c4_model(impl, a.m)
        cpu = {"value": mod["value"], "unit": "pts/s", "cores": 1, "kind": impl.kind, "same_config": True, "sample_n": 1500,
               "sample": mod["sample"], "t_train_s": mod["t_train_s"], "t_point_s": mod["t_point_s"]}
    line = {"metric": out.pop("metric"), "value": out.pop("value"), "unit": out.pop("unit"), "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": out.pop("ms_per_step"), "higher_is_better": True, "scaling": out.pop("scaling"),
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": out.pop("e2e"), "gpu_launches": out.pop("gpu_launches"), "cpu_baseline": cpu}
    line.update(out)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
