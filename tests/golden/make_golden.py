"""Generate golden vectors from the UNMODIFIED reference objects (oracle/_ref/libcugp_ref.so).

Run in the build container (needs /root/reference to have built oracle/_ref):

    make -C oracle ref
    python tests/golden/make_golden.py --group small     # ~2 min   -> golden_small.json
    python tests/golden/make_golden.py --group c4        # ~7 min   BCM N=24000 K=16
    python tests/golden/make_golden.py --group c2        # ~30 min  n=4096 theta_B
    python tests/golden/make_golden.py --group logs      # instant  known answers from the reference's run logs
    python tests/golden/make_golden.py --group c2a       # ~45 min  n=4096 theta_A (the start point of serial_gp.cpp:49)
    python tests/golden/make_golden.py --group c4pred    # ~15 min  BCM N=24000 K=16, 600 test points
    python tests/golden/make_golden.py --group kinv      # ~3 min   compute_K_inverse at n=1000 and n=2048 (sampled)

Every number is produced by the reference's own compiled code (Covsum / BCM / matrixops) through
oracle/ref_shim.cpp; nothing here comes from the CUDA product or from the C restatement.
"""
import argparse
import json
import os
import re
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

TH_A = [0.5, 0.5, 0.5]                         # cpp_serial_gp/serial_gp.cpp:49
TH_B = [3.762111, -1.152105, -0.384461]        # cuda_src/main.cpp:191-193 (trained values)
TH_C = [2.0, 2.0, 2.0]                         # cuda_scalingdist/main.cpp:298-301
TH_15 = [1.5, 1.5, 1.5]                        # distributed_gp/distributed_ver1.cpp:274
TH_OPT = [0.882908, 0.098703, -2.971479]       # cuda_bettersinglenode_ver2/REF:3183


def L(a):
    return [float(v) for v in np.asarray(a).ravel()]


def covsum_case(ref, name, X, y, theta, Xt=None, yt=None, extras=False):
    t0 = time.time()
    c = {"name": name, "kind": "covsum", "n": int(X.shape[0]), "d": int(X.shape[1]), "theta": L(theta)}
    c["ll"] = ref.loglik(X, y, theta)
    c["grad"] = L(ref.grad(X, y, theta))
    if Xt is not None and len(Xt):
        mu, var = ref.predict(X, y, theta, Xt)
        c["m"] = int(Xt.shape[0])
        c["mean"], c["var"] = L(mu), L(var)
        if yt is not None:
            c["nlpp"] = ref.nlpp(yt, mu, var)
    if extras:
        K = ref.K_train(X, theta)
        q, ld = ref.chol_and_det(K, y)
        c["quad"], c["logdet"] = q, ld
        c["alpha"] = L(ref.kinv_y(K, y))
        Lm = ref.cholesky(K)
        Ki = ref.k_inverse(K)
        c["K_row0"], c["L_lastrow"], c["Kinv_diag"] = L(K[0]), L(Lm[-1]), L(np.diag(Ki))
        c["K_fro"], c["L_fro"], c["Kinv_fro"] = (float(np.linalg.norm(M)) for M in (K, Lm, Ki))
    c["seconds"] = round(time.time() - t0, 2)
    print(f"  {name}: ll={c['ll']!r} grad={c['grad']} ({c['seconds']} s)", flush=True)
    return c


def bcm_case(ref, name, X, y, K, theta, Xt=None, yt=None):
    t0 = time.time()
    c = {"name": name, "kind": "bcm", "n": int(X.shape[0]), "d": int(X.shape[1]), "K": int(K), "theta": L(theta)}
    c["ll"] = ref.bcm_loglik(X, y, K, theta)
    c["grad"] = L(ref.bcm_grad(X, y, K, theta))
    if Xt is not None and len(Xt):
        mu, var = ref.bcm_predict(X, y, K, theta, Xt)
        c["m"] = int(Xt.shape[0])
        c["mean"], c["var"] = L(mu), L(var)
        if yt is not None:
            c["nlpp"] = ref.nlpp(yt, mu, var)
    c["seconds"] = round(time.time() - t0, 2)
    print(f"  {name}: ll={c['ll']!r} grad={c['grad']} ({c['seconds']} s)", flush=True)
    return c


def group_small(ref):
    out = []
    d = np.load(f"{HERE}/data_si128x2.npz")
    X, y = d["X"], d["y"]
    out.append(covsum_case(ref, "si128_th15", X, y, TH_15, extras=True))
    out.append(covsum_case(ref, "si128_thA", X, y, TH_A, extras=True))
    out.append(covsum_case(ref, "si96_thOPT_pred32", X[:96], y[:96], TH_OPT, X[96:], y[96:], extras=True))
    out.append(bcm_case(ref, "si128_bcm4_th15", X, y, 4, TH_15))
    out.append(bcm_case(ref, "si96_bcm3_thOPT_pred32", X[:96], y[:96], 3, TH_OPT, X[96:], y[96:]))
    out.append(bcm_case(ref, "si100_bcm3_th15_ragged", X[:100], y[:100], 3, TH_15, X[100:], y[100:]))  # 33,33,34
    d = np.load(f"{HERE}/data_sine1024.npz")
    X, y = d["X"], d["y"]
    out.append(covsum_case(ref, "sine256_thA_pred4", X[:256], y[:256], TH_A, X[256:260], y[256:260]))
    out.append(covsum_case(ref, "sine256_thB_pred4", X[:256], y[:256], TH_B, X[256:260], y[256:260], extras=True))
    out.append(covsum_case(ref, "sine300_thB_pred7", X[:300], y[:300], TH_B, X[300:307], y[300:307]))  # ragged n
    out.append(covsum_case(ref, "C1_sine1024_thA_pred8", X[:1024], y[:1024], TH_A, X[1024:1032], y[1024:1032]))
    out.append(covsum_case(ref, "C1_sine1024_thB_pred16", X[:1024], y[:1024], TH_B, X[1024:1040], y[1024:1040]))
    d = np.load(f"{HERE}/data_si24000.npz")
    out.append(covsum_case(ref, "C4_expert0_n1500_thC_pred16", d["X"][:1500], d["y"][:1500], TH_C, d["Xtest"], d["ytest"]))
    out.append(bcm_case(ref, "si24000_first3000_bcm4_thB_pred16", d["X"][:3000], d["y"][:3000], 4, TH_B, d["Xtest"], d["ytest"]))
    # optimiser end points (reference Covsum::cg_solve / cg_solve(BCM) / rprop_solve), 128 x 2 data
    d = np.load(f"{HERE}/data_si128x2.npz")
    X, y = d["X"], d["y"]
    th, _, _ = ref.cg_solve(X, y, TH_15, K=0)
    out.append({"name": "si128_cg_from_th15", "kind": "cg", "K": 0, "theta0": TH_15, "theta_final": L(th)})
    th, _, _ = ref.cg_solve(X, y, TH_15, K=4)
    out.append({"name": "si128_bcm4_cg_from_th15", "kind": "cg", "K": 4, "theta0": TH_15, "theta_final": L(th)})
    th = ref.rprop_solve(X, y, TH_15)
    out.append({"name": "si128_rprop_from_th15", "kind": "rprop", "theta0": TH_15, "theta_final": L(th)})
    print("  optimiser end points:", [c["theta_final"] for c in out[-3:]])
    return out


def group_c4(ref):
    d = np.load(f"{HERE}/data_si24000.npz")
    return [bcm_case(ref, "C4_si24000_bcm16_thC_pred16", d["X"], d["y"], 16, TH_C, d["Xtest"], d["ytest"])]


def group_c2(ref):
    d = np.load(f"{HERE}/data_sine4096.npz")
    X, y = d["X"], d["y"]
    return [covsum_case(ref, "C2_sine4096_thB_pred16", X[:4096], y[:4096], TH_B, X[4096:4112], y[4096:4112])]


def group_c2a(ref):
    """C2's own start point: theta_A on the first 4096 rows (serial_gp.cpp:49; LL in SURVEY App. B)."""
    d = np.load(f"{HERE}/data_sine4096.npz")
    X, y = d["X"], d["y"]
    return [covsum_case(ref, "C2_sine4096_thA_pred16", X[:4096], y[:4096], TH_A, X[4096:4112], y[4096:4112])]


def group_c4pred(ref):
    """C4 ensemble at 600 test points (rows 0..599 of the C3 set): several row tiles of the variance
    GEMM's column-sum epilogue and, with a small chunk cap, several test-set chunks."""
    d = np.load(f"{HERE}/data_si24000.npz")
    Xt = np.load(f"{HERE}/data_c4_xtest600.npz")["Xtest"]
    t0 = time.time()
    mu, var = ref.bcm_predict(d["X"], d["y"], 16, TH_C, Xt)
    c = {"name": "C4_si24000_bcm16_thC_pred600", "kind": "bcm_pred", "n": 24000, "d": 10, "K": 16, "theta": L(TH_C),
         "m": 600, "mean": L(mu), "var": L(var), "seconds": round(time.time() - t0, 2)}
    # one expert alone at the same points (Covsum::compute_test_means_and_variances, m = 600)
    mu1, var1 = ref.predict(d["X"][:1500], d["y"][:1500], TH_B, Xt)
    c1 = {"name": "C4_expert0_n1500_thB_pred600", "kind": "covsum_pred", "n": 1500, "d": 10, "theta": L(TH_B), "m": 600,
          "mean": L(mu1), "var": L(var1)}
    return [c, c1]


def group_kinv(ref):
    """compute_K_inverse (matrixops.cpp:383-435) at n = 1000 and 2048 on the C2 data: every 37th row and column of the
    inverse, its diagonal and its Frobenius norm (the full matrices would be 8 and 33 MB)."""
    d = np.load(f"{HERE}/data_sine4096.npz")
    out = []
    for n, th in ((1000, TH_B), (2048, TH_B)):
        X = d["X"][:n]
        t0 = time.time()
        K = ref.K_train(X, th)
        Ki = ref.k_inverse(K)
        Lm = ref.cholesky(K)
        idx = list(range(0, n, 37))
        out.append({"name": f"kinv_sine{n}_thB", "kind": "kinv", "n": n, "d": 10, "theta": L(th), "stride": 37,
                    "Kinv_sample": L(Ki[np.ix_(idx, idx)]), "Kinv_diag": L(np.diag(Ki)), "Kinv_fro": float(np.linalg.norm(Ki)),
                    "L_sample": L(Lm[np.ix_(idx, idx)]), "L_fro": float(np.linalg.norm(Lm)),
                    "seconds": round(time.time() - t0, 2)})
        print(f"  kinv n={n}: fro={out[-1]['Kinv_fro']!r} ({out[-1]['seconds']} s)", flush=True)
    return out


def group_logs(_ref):
    """Known answers the reference itself ships: 6-digit run logs."""
    log = open("/root/reference/cuda_bettersinglenode_ver2/REF").read()
    lls = [float(v) for v in re.findall(r"The value of loglikelihood = (-?[0-9.]+)", log)]
    grads = [[float(a), float(b), float(c)] for a, b, c in
             re.findall(r"Final gradients of log hyperparams are (-?[0-9.]+), (-?[0-9.]+), (-?[0-9.]+)", log)]
    thetas = [[float(a), float(b), float(c)] for a, b, c in
              re.findall(r"PLEASE-SEE\s+\d\s*: (-?[0-9.]+), (-?[0-9.]+), (-?[0-9.]+)", log)]
    see = open("/root/reference/cuda_ref/seeee").read()
    return [{
        "name": "REF_log", "kind": "log", "source": "cuda_bettersinglenode_ver2/REF", "data": "si128x2", "theta0": TH_15,
        "ll_sequence": lls, "grad_sequence": grads, "theta_sequence": thetas,
    }, {
        "name": "seeee_log", "kind": "log", "source": "cuda_ref/seeee:3,19", "data": "si128x2", "theta": TH_A,
        "logdet": float(re.search(r"Determinant is (-?[0-9.]+)", see).group(1)),
        "ll": float(re.search(r"The value of loglikelihood = (-?[0-9.]+)", see).group(1)),
    }]


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default="small", choices=["small", "c4", "c2", "logs", "c2a", "c4pred", "kinv"])
    a = ap.parse_args()
    ref = oracle.reference()
    if ref is None:
        sys.exit("oracle/_ref/libcugp_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    path = f"{HERE}/golden_{a.group}.json"   # one file per group, so groups can be generated in parallel
    gold = {"generator": f"tests/golden/make_golden.py --group {a.group}",
            "source": "oracle/_ref (unmodified reference objects)", "cases": {}}
    for c in {"small": group_small, "c4": group_c4, "c2": group_c2, "logs": group_logs, "c2a": group_c2a,
              "c4pred": group_c4pred, "kinv": group_kinv}[a.group](ref):
        gold["cases"][c["name"]] = c
    json.dump(gold, open(path, "w"), indent=1)
    print("wrote", path, "with", len(gold["cases"]), "cases")
