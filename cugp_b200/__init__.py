"""cugp_b200 -- B200-native (sm_100a) implementation of the exact-GP regression hot path of
abhishekjoshi2/cuGP, behind the reference's own call surface.

The product is ``libcugp.so`` (hand-written CUDA kernels + the C ABI of ``include/cugp.h``); this package
is the thin host-side mirror of the reference's classes used by the tests and by ``bench.py``:

* :class:`Covsum`  -- ``cpp_serial_gp/covkernel.h:3-38``
* ``get_cholesky`` / ``compute_chol_and_det`` / ``vector_Kinvy_using_cholesky`` / ``compute_K_inverse``
  -- ``common/matrixops.h:5-25``
* :class:`BCM`     -- ``distributed_gp/BCM.h:2-27`` (one process per GPU, one allreduce per operation)
* :class:`ShardStream` -- the shard-streaming ensemble of ``cuda_scalingdist`` (``cg_solver.cpp:42-70``)
* ``loaders``      -- the reference's dataset text formats

There is no CPU fallback: without the built library or without a CUDA device, calls raise.
"""
from ._lib import CugpError, LIB_PATH, lib  # noqa: F401
from .bcm import BCM, expert_partition, local_experts  # noqa: F401
from .shardstream import ShardStream  # noqa: F401
from .covsum import (Covsum, compute_chol_and_det, compute_K_inverse, get_cholesky,  # noqa: F401
                     vector_Kinvy_using_cholesky)

__all__ = ["Covsum", "BCM", "ShardStream", "get_cholesky", "compute_chol_and_det", "vector_Kinvy_using_cholesky",
           "compute_K_inverse", "expert_partition", "local_experts", "CugpError", "lib", "LIB_PATH"]
