"""Shard streaming (cuda_scalingdist/cg_solver.cpp:42-70).

GPU tests: the streamed ensemble must equal the oracle's BCM on the same rows for every slot count (whole
groups, a ragged last group, one shard at a time as the reference does), from text files and from host memory,
and a missing shard file must surface as an error, not a hang.  CPU test: the host logic over gloo."""
import os
import socket

import numpy as np
import pytest

from tests.conftest import assert_grad, assert_ll, assert_pred, load_data

TH_B = [3.762111, -1.152105, -0.384461]
TH = [0.882908, 0.098703, -2.971479]


def _write_shards(tmp_path, X, y, chunks):
    """The reference's shard layout: <prefix><i>.txt with an "n d" header, labels one per line."""
    n = X.shape[0] // chunks
    for i in range(chunks):
        with open(tmp_path / f"in_{i}.txt", "w") as f:
            f.write(f"{n} {X.shape[1]}\n")
            for row in X[i * n:(i + 1) * n]:
                f.write(" ".join(repr(float(v)) for v in row) + " \n")
        with open(tmp_path / f"lab_{i}.txt", "w") as f:
            for v in y[i * n:(i + 1) * n]:
                f.write(repr(float(v)) + "\n")
    return str(tmp_path / "in_"), str(tmp_path / "lab_"), n


def test_from_memory_rejects_a_remainder():
    """Shards of the stream are equally sized (cuda_scalingdist/main.cpp:94-125): rows that do not split evenly are an
    error, not a silent truncation."""
    import cugp_b200 as cg
    with pytest.raises(ValueError):
        cg.ShardStream.from_memory(np.zeros((10, 2)), np.zeros(10), 3)


@pytest.mark.gpu
@pytest.mark.parametrize("slots", [1, 2, 3, 0])
def test_stream_from_memory_matches_oracle_bcm(slots):
    import cugp_b200 as cg
    from oracle import oracle
    d = load_data("si24000")
    chunks, n = 5, 300
    X, y, Xt = d["X"][:chunks * n], d["y"][:chunks * n], d["Xtest"][:12]
    s = cg.ShardStream.from_memory(X, y, chunks, slots=slots)
    lay = s.layout()
    assert lay["local_shards"] == chunks and lay["groups"] == -(-chunks // lay["slots"])
    if slots:
        assert lay["slots"] == slots
    ref = oracle.port()
    for th in (TH_B, TH):
        s.set_BCM_log_hyperparam(th)
        ll, g = s.loglik_and_gradient()
        assert_ll(ll, ref.bcm_loglik(X, y, chunks, th))
        assert_grad(g, ref.bcm_grad(X, y, chunks, th))
        assert_ll(s.get_BCM_loglikelihood(), ref.bcm_loglik(X, y, chunks, th))
        mu, var = s.compute_BCM_test_means_and_var(Xt)
        mu0, var0 = ref.bcm_predict(X, y, chunks, th, Xt)
        assert_pred(mu, var, mu0, var0, yscale=np.abs(y).max())
    per = s.shard_logliks()
    for i in range(chunks):
        assert_ll(per[i], ref.loglik(X[i * n:(i + 1) * n], y[i * n:(i + 1) * n], TH))
    s.close()


@pytest.mark.gpu
def test_stream_equals_resident_bcm_bitwise():
    """Same kernels, same batch shapes per expert: streaming in groups must not change a single bit of the sums'
    terms (the per-shard log-likelihoods), whatever the slot count."""
    import cugp_b200 as cg
    d = load_data("si24000")
    chunks, n = 6, 256
    X, y = d["X"][:chunks * n], d["y"][:chunks * n]
    per = []
    for slots in (1, 4, 6):
        s = cg.ShardStream.from_memory(X, y, chunks, slots=slots)
        s.set_BCM_log_hyperparam(TH_B)
        per.append(s.shard_logliks().copy())
        s.close()
    assert np.array_equal(per[0], per[1]) and np.array_equal(per[0], per[2])


@pytest.mark.gpu
def test_stream_from_files_parses_once_when_cached(tmp_path):
    import cugp_b200 as cg
    from oracle import oracle
    d = load_data("si24000")
    chunks, n = 4, 200
    X, y = d["X"][:chunks * n], d["y"][:chunks * n]
    inp, lab, _ = _write_shards(tmp_path, X, y, chunks)
    ref = oracle.port()
    ll0, g0 = ref.bcm_loglik(X, y, chunks, TH_B), ref.bcm_grad(X, y, chunks, TH_B)
    for cache, parsed_after_3 in ((1 << 30, chunks), (0, 3 * chunks)):
        s = cg.ShardStream.from_files(inp, lab, chunks, n, X.shape[1], slots=2, host_cache_bytes=cache)
        s.set_BCM_log_hyperparam(TH_B)
        for _ in range(3):
            ll, g = s.loglik_and_gradient()
            assert_ll(ll, ll0)
            assert_grad(g, g0)
        st = s.stats()
        assert st["passes"] == 3 and st["groups"] == 6 and st["shards_parsed"] == parsed_after_3
        assert st["h2d_bytes"] == 3 * chunks * n * (X.shape[1] + 1) * 8
        s.close()


@pytest.mark.gpu
def test_stream_two_ranks_partials_sum_to_whole(tmp_path):
    """Shards i = rank, rank + world, ... (cg_solver.cpp:44): the two ranks' partial sums add up to the ensemble."""
    import cugp_b200 as cg
    from oracle import oracle
    d = load_data("si24000")
    chunks, n = 5, 200
    X, y, Xt = d["X"][:chunks * n], d["y"][:chunks * n], d["Xtest"][:8]
    inp, lab, _ = _write_shards(tmp_path, X, y, chunks)
    out, PQ = np.zeros(4), np.zeros((2, 8))
    for r in range(2):
        s = cg.ShardStream.from_files(inp, lab, chunks, n, X.shape[1], rank=r, world=1 + 1, slots=2)
        assert s.layout()["local_shards"] == len(range(r, chunks, 2))
        s._local.set_theta(np.array(TH_B))
        out += s._local.loglik_grad(True)
        PQ += s._local.moments(np.ascontiguousarray(Xt))
        s.close()
    ref = oracle.port()
    assert_ll(out[0], ref.bcm_loglik(X, y, chunks, TH_B))
    assert_grad(out[1:], ref.bcm_grad(X, y, chunks, TH_B))
    mu0, var0 = ref.bcm_predict(X, y, chunks, TH_B, Xt)
    assert_pred(PQ[1] / PQ[0], 1.0 / PQ[0], mu0, var0, yscale=np.abs(y).max())


@pytest.mark.gpu
def test_stream_missing_shard_is_an_error(tmp_path):
    import cugp_b200 as cg
    d = load_data("si24000")
    X, y = d["X"][:300], d["y"][:300]
    inp, lab, n = _write_shards(tmp_path, X, y, 3)
    os.remove(tmp_path / "in_2.txt")
    s = cg.ShardStream.from_files(inp, lab, 3, n, X.shape[1], slots=1)
    s.set_BCM_log_hyperparam(TH_B)
    with pytest.raises(cg.CugpError, match="cannot open"):
        s.loglik_and_gradient()
    # a short file too
    with open(tmp_path / "in_2.txt", "w") as f:
        f.write("100 10\n1.0 2.0\n")
    with pytest.raises(cg.CugpError, match="expected"):
        s.loglik_and_gradient()
    s.close()


# ---- the reader's parser (CPU) ------------------------------------------------------------------------------
def test_parser_matches_numpy_on_reference_and_generated_files(tmp_path):
    """from_chars must read exactly what fscanf("%lf") / numpy read: header skipped, exponents, signs, commas, CRLF."""
    import ctypes as C
    from cugp_b200._lib import lib, ptr
    rng = np.random.default_rng(5)
    X = rng.standard_normal((37, 5)) * 10.0 ** rng.integers(-30, 30, (37, 5))
    X[0, 0], X[1, 1], X[2, 2] = 0.0, -0.0, 5e-324
    inp, lab, n = _write_shards(tmp_path, X, X[:, 0], 1)
    out = np.full(X.size, np.nan)
    assert lib().cugp_shardstream_parse_file((inp + "0.txt").encode(), 2, X.size, ptr(out)) == 0
    assert np.array_equal(out.reshape(X.shape), X) and np.signbit(out[6])
    yl = np.full(n, np.nan)
    assert lib().cugp_shardstream_parse_file((lab + "0.txt").encode(), 0, n, ptr(yl)) == 0
    assert np.array_equal(yl, X[:, 0])
    # "+" signs, commas, CRLF, upper-case exponent
    with open(tmp_path / "odd.txt", "w", newline="") as f:
        f.write("3 2\r\n+1.5,2E3\r\n-3e-2\t4\r\n5 , 6\r\n")
    o = np.zeros(6)
    assert lib().cugp_shardstream_parse_file(str(tmp_path / "odd.txt").encode(), 2, 6, ptr(o)) == 0
    assert o.tolist() == [1.5, 2000.0, -0.03, 4.0, 5.0, 6.0]
    # short and missing files are errors that name the file
    assert lib().cugp_shardstream_parse_file(str(tmp_path / "odd.txt").encode(), 2, 7, ptr(np.zeros(7))) != 0
    assert b"odd.txt" in lib().cugp_last_error()
    assert lib().cugp_shardstream_parse_file(str(tmp_path / "nope.txt").encode(), 0, 1, ptr(np.zeros(1))) != 0
    assert b"cannot open" in lib().cugp_last_error()
    ref = "/root/reference/scaling_dataset/si24000_16sharded_chunk3.txt"
    if os.path.exists(ref):     # this container only: the reference's own shard file
        from cugp_b200.loaders import load_inputs
        Xr = load_inputs(ref)
        o = np.zeros(Xr.size)
        assert lib().cugp_shardstream_parse_file(ref.encode(), 2, Xr.size, ptr(o)) == 0
        assert np.array_equal(o.reshape(Xr.shape), Xr)


# ---- host logic over gloo (CPU) ---------------------------------------------------------------------------
class OracleStreamLocal:
    """Stand-in for the C-ABI stream of one rank, evaluated with the CPU oracle (test infrastructure)."""

    def __init__(self, X, y, chunks, rank, world):
        from oracle import oracle
        self.port = oracle.port()
        n = X.shape[0] // chunks
        self.parts = [(i * n, n) for i in range(rank, chunks, world)]
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


def _worker(rank, world, port, chunks, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cugp_b200.shardstream import ShardStream
        d = load_data("si128x2")
        X, y, Xt = d["X"][:96], d["y"][:96], d["X"][96:]
        r, w = ShardStream._rank_world(None, None, None)
        assert (r, w) == (rank, world)
        s = ShardStream(OracleStreamLocal(X, y, chunks, r, w), chunks, 96 // chunks, X.shape[1], r, w)
        s.set_BCM_log_hyperparam(TH)
        ll, g = s.loglik_and_gradient()
        mu, var = s.compute_BCM_test_means_and_var(Xt)
        assert s.exchanges == 2
        q.put((rank, ll, g.tolist(), mu.tolist(), var.tolist()))
    finally:
        dist.destroy_process_group()


def test_shardstream_two_ranks_gloo():
    import torch.multiprocessing as mp
    from oracle import oracle
    chunks = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, chunks, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    d = load_data("si128x2")
    X, y, Xt = d["X"][:96], d["y"][:96], d["X"][96:]
    ref = oracle.port()
    ll0, g0 = ref.bcm_loglik(X, y, chunks, TH), ref.bcm_grad(X, y, chunks, TH)
    mu0, var0 = ref.bcm_predict(X, y, chunks, TH, Xt)
    for _, ll, g, mu, var in res:
        assert abs(ll - ll0) <= 1e-12 * abs(ll0)
        assert np.allclose(g, g0, rtol=1e-11, atol=0)
        assert np.allclose(mu, mu0, rtol=1e-11) and np.allclose(var, var0, rtol=1e-11)
    assert res[0][1:] == res[1][1:]
