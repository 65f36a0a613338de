"""Sharded BCM on two real GPUs (skipped on a one-GPU box; `gpurun --gpus 2`): the library's exchange step -- one kernel over
NVLink peer memory (csrc/peerxchg.cu), ncclAllReduce as the fallback -- against the goldens of the unmodified reference and
against itself.  The CPU side of the same logic is tests/test_bcm_gloo.py."""
import ctypes as C
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _torchrun(script, port, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", script)]
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def test_exchange_kind_of_no_handle_is_none():
    from cugp_b200._lib import lib
    assert lib().cugp_bcm_exchange_kind(C.c_void_p()) == 0


@pytest.mark.gpu
@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs (run under gpurun --gpus 2)")
def test_sharded_bcm_reproduces_the_goldens_on_every_rank():
    """tools/bcm_nccl_check.py: C4 (16 experts x 1500 rows, BCM.cpp:64-83, 153-198) split over two ranks; LL / gradient to
    1e-9, mean / variance to 1e-8 of the reference's values on BOTH ranks, and the same bits on both."""
    r = _torchrun("bcm_nccl_check.py", 29571)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("[rank")]
    assert len(lines) >= 2 and all(l.rstrip().endswith("OK") for l in lines), lines
    assert all("same_bits_as_rank0=True" in l for l in lines), lines


@pytest.mark.gpu
@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs (run under gpurun --gpus 2)")
def test_peer_memory_exchange_equals_nccl():
    """tools/r2_peer_ab.py: the same handles with bcm_peer_exchange = 1 and 0: (LL, gradient, mean, variance) agree to 1e-12
    (exit code), and the first leg really took the peer-memory path where the box allows it."""
    r = _torchrun("r2_peer_ab.py", 29572)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "exchange=nccl" in r.stdout
