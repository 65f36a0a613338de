"""Static checks of bench.py that need no GPU.

Under torchrun every rank but 0 leaves main() once the collective legs are done; rank 0 alone then runs the probes,
the C2 / shard-streaming extras and the CPU baseline.  A collective issued after that point hangs the whole job
(it did, once: a streamed-ensemble extra that allreduced over ranks that had already left).  So: after the
`if rank != 0:` early return, main() must not hand `dist` to anything nor call a collective on it."""
import ast
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _main_source():
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "main")
    return src.splitlines()[fn.lineno - 1:fn.end_lineno]


def test_no_collective_after_nonzero_ranks_leave():
    lines = _main_source()
    cut = next(i for i, l in enumerate(lines) if re.match(r"\s*if rank != 0:\s*$", l))
    tail = "\n".join(lines[cut + 4:])          # skip the early-return block itself
    assert "dist.barrier" not in tail and "all_reduce" not in tail and "all_gather" not in tail
    # `dist` may only appear to tear the group down
    uses = [l.strip() for l in tail.splitlines() if re.search(r"\bdist\b", l) and not l.strip().startswith("#")]
    assert all("destroy_process_group" in u or u.startswith("if dist is not None") for u in uses), uses


def test_torchrun_safe_flags():
    """torch.distributed.run parses abbreviations of ITS options even after the script name (`--n` is ambiguous
    there): every flag the driver passes must not be a prefix of a torchrun option."""
    driver_flags = ["--gpus", "--steps", "--warmup", "--impl"]
    torchrun = ["--nnodes", "--nproc-per-node", "--nproc_per_node", "--rdzv-backend", "--rdzv-endpoint", "--rdzv-id",
                "--rdzv-conf", "--standalone", "--max-restarts", "--monitor-interval", "--start-method", "--role", "--module",
                "--no-python", "--run-path", "--log-dir", "--redirects", "--tee", "--local-ranks-filter", "--node-rank",
                "--master-addr", "--master-port", "--local-addr", "--logs-specs", "--signals-to-handle",
                "--virtual-local-rank", "--numa-binding", "--event-log-handler", "--duplicate-stdout-filters",
                "--duplicate-stderr-filters"]
    for f in driver_flags:
        assert not any(o.startswith(f) for o in torchrun), f
