"""Shard streaming: the expert ensemble of ``cuda_scalingdist`` whose shards do not
stay on the GPU.

The reference (``cuda_scalingdist/cg_solver.cpp:42-70``, ``main.cpp:94-125``) walks shard files
``<prefix><i>.txt`` for ``i = worker_id; i < numchunks; i += total_workers`` on every evaluation: a background
thread parses the next shard into one of two host buffers while the GPU factorises the current one, and the
per-shard log-likelihoods / gradients are summed (then sent to the master over a socket).

:class:`ShardStream` is that loop behind the :class:`~cugp_b200.bcm.BCM` interface: same partition of shards
over ranks, same sums, same single allreduce per operation -- but ``slots`` experts per launch, pinned double
buffers, a copy stream for the upload, and parsed text kept in a bounded host cache so only the first pass pays
for ``strtod``.  Every shard has ``numtrain`` rows (the reference's argv[5]).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import ShardStreamStats, check, f64, lib, ptr
from .bcm import BCM


class _StreamLocal:
    """This rank's shards, streamed through the C ABI (``cugp_shardstream_*``)."""

    def __init__(self, handle, keep=None):
        self._h = handle
        self._keep = keep  # arrays the library reads by pointer

    def close(self):
        if self._h is not None and self._h.value:
            lib().cugp_shardstream_close(self._h)
            self._h = None
            self._keep = None

    def layout(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        check(lib().cugp_shardstream_layout(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"local_shards": a.value, "slots": b.value, "groups": c.value}

    def set_theta(self, th):
        check(lib().cugp_shardstream_set_loghyper(self._h, ptr(th)))

    def loglik_grad(self, want_grad: bool):
        out = np.zeros(4)
        check(lib().cugp_shardstream_loglik_grad_local(self._h, int(want_grad), ptr(out), None))
        return out

    def shard_logliks(self):
        lay = self.layout()
        ll = np.zeros(max(lay["groups"] * lay["slots"], 1))
        out = np.zeros(4)
        check(lib().cugp_shardstream_loglik_grad_local(self._h, 0, ptr(out), ptr(ll)))
        return ll[: lay["local_shards"]]

    def moments(self, Xt):
        PQ = np.zeros((2, Xt.shape[0]))
        check(lib().cugp_shardstream_predict_moments(self._h, ptr(Xt), Xt.shape[0], ptr(PQ)))
        return PQ

    def moments_into(self, Xt, dev_ptr: int):
        check(lib().cugp_shardstream_predict_moments_dev(self._h, ptr(Xt), Xt.shape[0], C.c_void_p(dev_ptr)))

    def stats(self):
        s = ShardStreamStats()
        check(lib().cugp_shardstream_get_stats(self._h, C.byref(s)))
        return {name: getattr(s, name) for name, _ in ShardStreamStats._fields_}


class ShardStream(BCM):
    """An expert ensemble over ``numchunks`` shards of ``numtrain`` rows each, streamed through ``slots`` device
    slots.  Build it with :meth:`from_files` or :meth:`from_memory`; the evaluation / prediction / ``cg_solve``
    methods are :class:`BCM`'s (sum over shards, product of experts, one allreduce per operation)."""

    def __init__(self, local, numchunks, numtrain, dim, rank, world, group=None):
        self.N, self.D, self.num_experts = int(numchunks) * int(numtrain), int(dim), int(numchunks)
        self.rank, self.world, self.group = int(rank), int(world), group
        self._dist = None
        if self.world > 1:
            import torch.distributed as dist
            self._dist = dist
        self.offset = [i * int(numtrain) for i in range(int(numchunks))]
        self.log_hyper_bcm = np.zeros(3)
        self._local = local
        self._in_library = False     # the stream's exchange step goes through torch.distributed (BCM._allreduce)
        self._exchanges = 0

    @staticmethod
    def _rank_world(rank, world, group):
        if rank is None or world is None:
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    return dist.get_rank(group), dist.get_world_size(group)
            except ImportError:
                pass
            return 0, 1
        return int(rank), int(world)

    @classmethod
    def from_files(cls, input_prefix: str, label_prefix: str, numchunks: int, numtrain: int, dim: int, rank=None,
                   world=None, group=None, slots: int = 0, host_cache_bytes: int = 1 << 32):
        """Shard i is ``<input_prefix><i>.txt`` / ``<label_prefix><i>.txt`` (``main.cpp:247-252``)."""
        rank, world = cls._rank_world(rank, world, group)
        h = C.c_void_p()
        check(lib().cugp_shardstream_open_files(input_prefix.encode(), label_prefix.encode(), numchunks, numtrain, dim,
                                                rank, world, slots, host_cache_bytes, C.byref(h)))
        return cls(_StreamLocal(h), numchunks, numtrain, dim, rank, world, group)

    @classmethod
    def from_memory(cls, X, y, numchunks: int, rank=None, world=None, group=None, slots: int = 0):
        """Shards are consecutive blocks of ``len(y) // numchunks`` rows of (X, y), which stay on the host."""
        X, y = f64(X), f64(y)
        if numchunks <= 0 or X.shape[0] % numchunks:
            # the stream's shards all have numtrain rows (cuda_scalingdist/main.cpp:94-125); BCM's "last expert takes
            # the remainder" (BCM.cpp:92-108) is the resident cugp_b200.BCM, not the stream
            raise ValueError(f"ShardStream.from_memory: {X.shape[0]} rows do not split into {numchunks} equal shards; "
                             "trim the data or use cugp_b200.BCM (its last expert takes the remainder)")
        numtrain = X.shape[0] // numchunks
        rank, world = cls._rank_world(rank, world, group)
        h = C.c_void_p()
        check(lib().cugp_shardstream_open_memory(ptr(X), ptr(y), numchunks, numtrain, X.shape[1], rank, world, slots,
                                                 C.byref(h)))
        return cls(_StreamLocal(h, keep=(X, y)), numchunks, numtrain, X.shape[1], rank, world, group)

    def layout(self):
        return self._local.layout()

    def stats(self):
        return self._local.stats()

    def shard_logliks(self):
        return self._local.shard_logliks()
