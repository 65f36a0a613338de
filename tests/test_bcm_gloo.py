"""Host logic of the sharded BCM on CPU: world_size 2 over gloo.  The per-rank expert maths is stood in by the
CPU oracle (test infrastructure only); what is under test is cugp_b200.bcm.BCM -- the expert -> rank
assignment, exactly one allreduce per operation with the payloads of SURVEY.md section 8(e), and that the
reduced results equal the reference's sequential sums (BCM.cpp:153-198) and its product of experts
(BCM.cpp:45-62)."""
import os
import socket

import numpy as np
import pytest

from tests.conftest import load_data

TH = [0.882908, 0.098703, -2.971479]


class OracleLocal:
    """This rank's experts evaluated with the CPU oracle: the seam `local_impl` of cugp_b200.bcm.BCM."""

    def __init__(self, X, y, K, rank, world):
        from cugp_b200.bcm import expert_partition, local_experts
        from oracle import oracle
        self.port = oracle.port()
        self.parts = [expert_partition(X.shape[0], K)[e] for e in local_experts(K, rank, world)]
        self.X, self.y, self.th = X, y, np.zeros(3)

    def set_theta(self, th):
        self.th = np.array(th, dtype=float)

    def loglik_grad(self, want_grad):
        out = np.zeros(4)
        for o, s in self.parts:
            out[0] += self.port.loglik(self.X[o:o + s], self.y[o:o + s], self.th)
            if want_grad:
                out[1:] += self.port.grad(self.X[o:o + s], self.y[o:o + s], self.th)
        return out

    def moments(self, Xt):
        PQ = np.zeros((2, Xt.shape[0]))
        for o, s in self.parts:
            mu, var = self.port.predict(self.X[o:o + s], self.y[o:o + s], self.th, Xt)
            PQ[0] += 1.0 / var
            PQ[1] += mu / var
        return PQ


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, K, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cugp_b200.bcm import BCM
        d = load_data("si128x2")
        X, y = d["X"][:96], d["y"][:96]
        Xt = d["X"][96:]
        b = BCM(X, y, K=K, local_impl=OracleLocal)          # rank / world from the process group
        assert (b.rank, b.world) == (rank, world)
        b.set_BCM_log_hyperparam(TH)
        ll, g = b.loglik_and_gradient()
        assert b.exchanges == 1                              # one allreduce of 4 doubles
        mu, var = b.compute_BCM_test_means_and_var(Xt)
        assert b.exchanges == 2                              # one allreduce of 2m doubles
        ll2 = b.get_BCM_loglikelihood()
        q.put((rank, ll, g.tolist(), mu.tolist(), var.tolist(), ll2, len(b._local.parts)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("K", [3, 4])
def test_bcm_two_ranks_gloo(K):
    import torch.multiprocessing as mp
    from oracle import oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, K, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    d = load_data("si128x2")
    X, y, Xt = d["X"][:96], d["y"][:96], d["X"][96:]
    ref = oracle.port()
    ll0, g0 = ref.bcm_loglik(X, y, K, TH), ref.bcm_grad(X, y, K, TH)
    mu0, var0 = ref.bcm_predict(X, y, K, TH, Xt)
    assert sorted(r[6] for r in res) == sorted([len(range(0, K, 2)), len(range(1, K, 2))])   # experts e % 2 == rank
    for _, ll, g, mu, var, ll2, _ in res:                    # every rank holds the reduced result
        assert abs(ll - ll0) <= 1e-12 * abs(ll0) and abs(ll2 - ll0) <= 1e-12 * abs(ll0)
        assert np.allclose(g, g0, rtol=1e-11, atol=0)
        assert np.allclose(mu, mu0, rtol=1e-11) and np.allclose(var, var0, rtol=1e-11)
    assert res[0][1:6] == res[1][1:6]                        # bitwise identical on both ranks


def test_expert_partition_matches_reference():
    """BCM.cpp:92-108: K chunks of floor(N/K) rows, the last takes the remainder."""
    from cugp_b200.bcm import expert_partition, local_experts
    assert expert_partition(128, 4) == [(0, 32), (32, 32), (64, 32), (96, 32)]
    assert expert_partition(100, 3) == [(0, 33), (33, 33), (66, 34)]
    assert expert_partition(24000, 16)[15] == (22500, 1500)
    assert local_experts(16, 3, 8) == [3, 11] and local_experts(3, 2, 4) == [2] and local_experts(3, 3, 4) == []
