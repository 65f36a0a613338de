"""Under torchrun: the BCM exchange step over NVLink peer memory (csrc/peerxchg.cu) against ncclAllReduce, same handle, same
data: (LL, gradient) evaluation and prediction of 10000 points with two experts per rank; results must agree to rounding."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200._lib import check, lib
from cugp_b200.loaders import synthetic_sine

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
check(lib().cugp_set_device(local))
L = lib()
d = np.load("tests/golden/data_si24000.npz")
K = 2 * world
Xt, _ = synthetic_sine(10000, 10, seed=7)
b = cg.BCM(d["X"][:1500 * K], d["y"][:1500 * K], K=K)
res = {}
for rnd in range(2):
    for peer in (1, 0):
        check(L.cugp_set_tuning(b"bcm_peer_exchange", peer))
        te, tp = [], []
        for r in range(12):
            b.set_BCM_log_hyperparam([2.0 + 1e-7 * r, 2.0, 2.0])
            dist.barrier()
            torch.cuda.synchronize()
            t = time.perf_counter()
            ll, g = b.loglik_and_gradient()
            te.append(time.perf_counter() - t)
            dist.barrier()
            t = time.perf_counter()
            mu, var = b.compute_BCM_test_means_and_var(Xt)
            tp.append(time.perf_counter() - t)
        res[peer] = (ll, g, mu, var)
        tt = torch.tensor([np.median(te[2:]), np.median(tp[2:])], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"world={world} experts={K} exchange={b.exchange_kind}: eval {1e3 * tt[0].item():.3f} ms  predict(10000) "
                  f"{1e3 * tt[1].item():.3f} ms (median, max over ranks)", flush=True)
a, c = res[1], res[0]
e = [abs(a[0] - c[0]) / abs(c[0]), float(np.max(np.abs(a[1] - c[1]) / np.abs(c[1]))),
     float(np.max(np.abs(a[2] - c[2]) / np.maximum(np.abs(c[2]), 1e-6))), float(np.max(np.abs(a[3] - c[3]) / np.abs(c[3])))]
print(f"[rank {rank}] peer vs nccl relerr: ll {e[0]:.1e} grad {e[1]:.1e} mean {e[2]:.1e} var {e[3]:.1e}", flush=True)
b.close()
dist.destroy_process_group()
sys.exit(0 if max(e) < 1e-12 else 1)
