"""Run under torchrun on N GPUs: the sharded BCM (experts e % world == rank, NCCL allreduce of 4 / 2m doubles) must
reproduce the golden C4 values (16 experts x 1500 points, generated from the unmodified reference) on every rank."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200._lib import check, lib

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
check(lib().cugp_set_device(local))
gold = json.load(open("tests/golden/golden_c4.json"))["cases"]
d = np.load("tests/golden/data_si24000.npz")
ok = True
for name, c in gold.items():
    if c["kind"] != "bcm":
        continue
    b = cg.BCM(d["X"][:c["n"]], d["y"][:c["n"]], K=c["K"])
    b.set_BCM_log_hyperparam(c["theta"])
    ll, g = b.loglik_and_gradient()
    m = c.get("m", 0)
    e_ll = abs(ll - c["ll"]) / abs(c["ll"])
    gref = np.array(c["grad"])
    e_g = float(np.max(np.abs(g - gref) / np.maximum(np.abs(gref), np.abs(gref).max())))
    line = f"[rank {rank}/{world}] {name}: K={c['K']} local experts={len(range(rank, c['K'], world))} LL relerr {e_ll:.2e} grad relerr {e_g:.2e}"
    good = e_ll <= 1e-9 and e_g <= 1e-9
    if m:
        mu, var = b.compute_BCM_test_means_and_var(d["Xtest"][:m])
        e_m = float(np.max(np.abs(mu - c["mean"]) / np.maximum(np.abs(c["mean"]), 1e-6)))
        e_v = float(np.max(np.abs(var - c["var"]) / np.abs(c["var"])))
        line += f" mean relerr {e_m:.2e} var relerr {e_v:.2e}"
        good = good and e_m <= 1e-8 and e_v <= 1e-8
    # every rank must hold the SAME bits after the exchange (rank-ordered sums over peer memory; NCCL gives that too)
    mine = torch.tensor([ll, *g], dtype=torch.float64, device="cuda")
    ref0 = mine.clone()
    dist.broadcast(ref0, src=0)
    same = bool(torch.equal(mine, ref0))
    good = good and same
    line += f" exchange={b.exchange_kind} same_bits_as_rank0={same} exchanges={b.exchanges}" + ("  OK" if good else "  FAIL")
    print(line, flush=True)
    ok = ok and good
    b.close()
t = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(t)
dist.destroy_process_group()
sys.exit(int(t.item() != 0))
