"""Host-side mirror of the reference's ``class BCM`` (distributed_gp/BCM.h:2-27), one process per GPU.

The expert ensemble is the only part of the path that shards: expert e is a
contiguous chunk of floor(N/K) rows, the last takes the remainder (BCM.cpp:85-110), every expert shares
theta, the log-likelihood and gradient are plain sums over experts (BCM.cpp:153-198) and the prediction is a
product of experts (BCM.cpp:45-62).  Rank r of W owns experts e with e % W == r and keeps them device
resident; each operation has exactly ONE exchange step:

* training evaluation: allreduce(sum, f64, 4)  of (LL, g0, g1, g2)
* prediction:          allreduce(sum, f64, 2m) of (sum_e 1/var_e, sum_e mean_e/var_e)

On GPUs the exchange runs INSIDE libcugp: the library owns an NCCL communicator (``cugp_bcm_comm_init``; torch.distributed
only broadcasts the 128-byte unique id once) and enqueues ``ncclAllReduce`` on its own stream right behind the kernels
that produce the payload -- no host round trip, no torch tensor.  The gloo path (CPU tests of the host logic) and a
process group without NCCL go through ``torch.distributed``.  The reference moved the same payloads over blocking TCP
sockets (cuda_src/cg_solver.cpp:22-79).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import EVAL_FN, check, dp, f64, lib, ptr


def expert_partition(N: int, K: int):
    """(offset, size) of every expert -- BCM.cpp:92-108."""
    part = N // K
    out, start = [], 0
    for i in range(K):
        size = part if i < K - 1 else N - start
        out.append((start, size))
        start += part
    return out


def local_experts(K: int, rank: int, world: int):
    """Experts owned by `rank`: round-robin, as the reference assigns shards to nodes
    (cuda_scalingdist/main.cpp:101 `i = worker_id; i < chunks; i += W`)."""
    return list(range(rank, K, world))


NCCL_ID_BYTES = 128


class _CudaLocal:
    """This rank's experts on its GPU, through the C ABI.  With a communicator (``comm_init``) the exchange step of
    every operation runs inside the library: ``ncclAllReduce`` on the library's stream, no host round trip."""

    def __init__(self, X, y, K, rank, world, device=None):
        self._h = C.c_void_p()
        N, D = X.shape
        if device is not None:
            check(lib().cugp_set_device(int(device)))
        check(lib().cugp_bcm_create(ptr(X), ptr(y), N, D, K, rank, world, C.byref(self._h)))

    def close(self):
        if self._h.value:
            lib().cugp_bcm_destroy(self._h)
            self._h = C.c_void_p()

    # -- communicator ------------------------------------------------------------------------------------------
    @staticmethod
    def new_unique_id() -> bytes:
        buf = (C.c_ubyte * NCCL_ID_BYTES)()
        check(lib().cugp_nccl_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id: bytes):
        buf = (C.c_ubyte * NCCL_ID_BYTES).from_buffer_copy(unique_id)
        check(lib().cugp_bcm_comm_init(self._h, buf))

    def has_comm(self) -> bool:
        return bool(lib().cugp_bcm_has_comm(self._h))

    def collectives(self) -> int:
        return int(lib().cugp_bcm_collectives(self._h))

    def exchange_kind(self) -> str:
        """How the library's exchange step runs: 'none', 'nccl' or 'peer_memory' (csrc/peerxchg.cu)."""
        return ("none", "nccl", "peer_memory")[int(lib().cugp_bcm_exchange_kind(self._h))]

    def set_theta(self, th):
        check(lib().cugp_bcm_set_loghyper(self._h, ptr(th)))

    def loglik_grad(self, want_grad: bool):
        out = np.zeros(4)
        check(lib().cugp_bcm_loglik_grad_local(self._h, int(want_grad), ptr(out)))
        return out

    def loglik_grad_all(self, want_grad: bool):
        """(LL, g) over ALL experts: local sums + the library's own allreduce (collective)."""
        out = np.zeros(4)
        check(lib().cugp_bcm_loglik_grad(self._h, int(want_grad), ptr(out)))
        return out

    def predict_all(self, Xt):
        """PoE mean / variance over ALL experts: moments + allreduce + finalisation inside the library (collective)."""
        m = Xt.shape[0]
        mean, var = np.empty(m), np.empty(m)
        check(lib().cugp_bcm_predict(self._h, ptr(Xt), m, ptr(mean), ptr(var)))
        return mean, var

    def expert_logliks(self):
        cnt = C.c_int()
        check(lib().cugp_bcm_local_experts(self._h, C.byref(cnt), None, None))
        ids = (C.c_int * max(cnt.value, 1))()
        ll = np.zeros(max(cnt.value, 1))
        check(lib().cugp_bcm_local_experts(self._h, C.byref(cnt), ids, ptr(ll)))
        return list(ids[: cnt.value]), ll[: cnt.value]

    def moments(self, Xt):
        """(2, m) host array: sum_e 1/var_e and sum_e mean_e/var_e over the local experts."""
        PQ = np.zeros((2, Xt.shape[0]))
        check(lib().cugp_bcm_predict_moments(self._h, ptr(Xt), Xt.shape[0], ptr(PQ)))
        return PQ

    def moments_into(self, Xt, dev_ptr: int):
        check(lib().cugp_bcm_predict_moments_dev(self._h, ptr(Xt), Xt.shape[0], C.c_void_p(dev_ptr)))


class BCM:
    """``BCM(X, y, N, D, K)`` of the reference; ``group`` is a ``torch.distributed`` process group (or the
    default group when ``torch.distributed`` is initialised), ``None`` for a single process."""

    def __init__(self, X, y, N=None, D=None, K=1, rank=None, world=None, group=None, local_impl=None, device=None):
        X, y = f64(X), f64(y)
        N = X.shape[0] if N is None else int(N)
        D = X.shape[1] if D is None else int(D)
        X, y = np.ascontiguousarray(X[:N, :D]), np.ascontiguousarray(y[:N])
        self.N, self.D, self.num_experts = N, D, int(K)
        self._dist = None
        if rank is None or world is None:
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    self._dist = dist
                    rank, world = dist.get_rank(group), dist.get_world_size(group)
            except ImportError:
                pass
            if rank is None or world is None:
                rank, world = 0, 1
        elif world > 1:
            import torch.distributed as dist
            self._dist = dist
        self.rank, self.world, self.group = int(rank), int(world), group
        self.offset = [o for o, _ in expert_partition(N, self.num_experts)]
        self.log_hyper_bcm = np.zeros(3)
        self._in_library = False      # the exchange step runs inside libcugp (ncclAllReduce on its own stream)
        if local_impl is None:
            # one process per GPU: a rank of an initialised torchrun job drives GPU LOCAL_RANK unless told otherwise;
            # everything else (a single process, ranks emulated in one process) keeps the current device
            if device is None and self.world > 1 and self._dist is not None and self._dist.is_initialized():
                import os
                if "LOCAL_RANK" in os.environ:
                    device = int(os.environ["LOCAL_RANK"])
            self.device = device
            self._local = _CudaLocal(X, y, self.num_experts, self.rank, self.world, device)
            if self.world > 1 and self._dist is not None and self._dist.is_initialized() \
                    and self._dist.get_backend(self.group) == "nccl":
                # the library builds its OWN communicator; torch.distributed only carries the 128-byte id once
                box = [_CudaLocal.new_unique_id() if self.rank == 0 else None]
                src = self._dist.get_global_rank(self.group, 0) if self.group is not None else 0
                self._dist.broadcast_object_list(box, src=src, group=self.group)
                self._local.comm_init(box[0])
                self._in_library = True
        else:
            self.device = device
            self._local = local_impl(X, y, self.num_experts, self.rank, self.world)
        self._exchanges = 0  # collectives issued through torch.distributed (one per operation)

    @property
    def exchanges(self):
        """Collectives issued so far: one per operation, whichever layer carries them."""
        n = self._exchanges
        if self._in_library:
            n += self._local.collectives()
        return n

    @exchanges.setter
    def exchanges(self, v):
        self._exchanges = v

    @property
    def exchange_kind(self) -> str:
        """'peer_memory' / 'nccl' (inside the library), 'torch' (the caller's process group) or 'none' (one rank)."""
        if self._in_library:
            return self._local.exchange_kind()
        return "none" if self.world == 1 else "torch"

    def close(self):
        if getattr(self, "_local", None) is not None and hasattr(self._local, "close"):
            self._local.close()
            self._local = None

    __del__ = close

    # -- the single exchange step ---------------------------------------------------------------------------
    def _allreduce(self, arr: np.ndarray) -> np.ndarray:
        if self.world == 1 or self._dist is None:
            return arr
        import torch
        dist = self._dist
        self._exchanges += 1
        backend = dist.get_backend(self.group)
        t = torch.from_numpy(np.ascontiguousarray(arr))
        if backend == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    # -- hyper-parameters (BCM.cpp:123-130, 200-212) ------------------------------------------------------------
    def set_BCM_log_hyperparam(self, theta):
        th = f64(theta)
        assert th.shape == (3,)
        self.log_hyper_bcm = th.copy()
        self._local.set_theta(th)

    set_BCM_loghyper_eigen = set_BCM_log_hyperparam

    def get_loghyperparam(self):
        return self.log_hyper_bcm.copy()

    def get_BCM_log_hyperparam(self):
        """BCM.cpp:132-147 SUMS the experts' theta (unused by the driver); kept for parity."""
        return self.log_hyper_bcm * self.num_experts

    # -- training evaluation (BCM.cpp:153-198) ---------------------------------------------------------------------
    def loglik_and_gradient(self):
        """One factorisation per expert, one allreduce of 4 doubles: (LL, grad[3])."""
        if self._in_library:
            out = self._local.loglik_grad_all(True)
        else:
            out = self._allreduce(self._local.loglik_grad(True))
        return float(out[0]), out[1:4].copy()

    def get_BCM_loglikelihood(self):
        if self._in_library:
            return float(self._local.loglik_grad_all(False)[0])
        return float(self._allreduce(self._local.loglik_grad(False))[0])

    def get_BCM_gradient_hyper(self):
        return self.loglik_and_gradient()[1]

    # -- prediction (BCM.cpp:45-83) ----------------------------------------------------------------------------------
    def compute_BCM_test_means_and_var(self, Xtest):
        Xt = f64(Xtest).reshape(-1, self.D)
        m = Xt.shape[0]
        if m == 0:
            return np.empty(0), np.empty(0)
        if self._in_library or (self.world == 1 and isinstance(self._local, _CudaLocal)):
            return self._local.predict_all(Xt)
        nccl = self.world > 1 and self._dist is not None and self._dist.get_backend(self.group) == "nccl"
        if nccl and hasattr(self._local, "moments_into"):
            # device-resident exchange: the moments never leave HBM before the NCCL allreduce
            import torch
            PQ = torch.empty(2 * m, dtype=torch.float64, device="cuda")
            self._local.moments_into(Xt, PQ.data_ptr())
            self._exchanges += 1
            self._dist.all_reduce(PQ, op=self._dist.ReduceOp.SUM, group=self.group)
            torch.cuda.current_stream().synchronize()
            mean, var = np.empty(m), np.empty(m)
            check(lib().cugp_poe_finalize_dev(C.c_void_p(PQ.data_ptr()), m, ptr(mean), ptr(var)))
            return mean, var
        PQ = self._allreduce(self._local.moments(Xt).reshape(-1)).reshape(2, m)
        tempvar = 1.0 / PQ[0]          # BCM.cpp:56-57
        return tempvar * PQ[1], tempvar

    @staticmethod
    def get_BCM_negative_log_predprob(actual, predmean, predvar):
        a, mu, v = f64(actual), f64(predmean), f64(predvar)
        out = C.c_double()
        check(lib().cugp_nlpp(ptr(a), ptr(mu), ptr(v), a.shape[0], C.byref(out)))
        return out.value

    # -- optimiser (distributed_ver1.cpp:13-232) -----------------------------------------------------------------------
    def cg_solve(self, trace_cap: int = 256):
        """The free function cg_solve(BCM): every rank runs the same host loop on allreduced values."""
        err = []

        def _eval(_ctx, th_p, f_p, g_p):
            try:
                self.set_BCM_log_hyperparam(np.array([th_p[0], th_p[1], th_p[2]]))
                ll, g = self.loglik_and_gradient()
                f_p[0] = -1.0 * ll
                g_p[0], g_p[1], g_p[2] = g
                return 0
            except Exception as e:  # never unwind through the C frame
                err.append(e)
                return 2

        th = self.log_hyper_bcm.copy()
        tr = np.full(trace_cap, np.nan)
        ne = C.c_int()
        rc = lib().cugp_cg_minimize(EVAL_FN(_eval), None, ptr(th), ptr(tr), trace_cap, C.byref(ne))
        if err:
            raise err[0]
        check(rc)
        self.set_BCM_log_hyperparam(th)
        return tr[: min(ne.value, trace_cap)]
