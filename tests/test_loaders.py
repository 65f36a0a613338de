"""Dataset formats of the reference: header that lies about the row count, rows with a
trailing space, header-less comma-separated source file, shard prefixes; and, where the reference tree is present,
its real files."""
import os

import numpy as np
import pytest

from cugp_b200 import loaders

REF = "/root/reference"


def _write(path, text):
    with open(path, "w") as f:
        f.write(text)


def test_header_can_lie_and_rows_are_read_to_eof(tmp_path):
    rng = np.random.default_rng(0)
    X = rng.uniform(-20, 20, (7, 3))
    y = np.sin(X[:, 0])
    fi, fl = tmp_path / "in.txt", tmp_path / "lab.txt"
    _write(fi, "4 3\n" + "".join(" ".join(f"{v:.5g}" for v in r) + " \n" for r in X))   # header says 4, file holds 7
    _write(fl, "".join(f"{v:.5g}\n" for v in y))
    Xa = loaders.load_inputs(str(fi))
    assert Xa.shape == (7, 3)
    assert np.allclose(Xa, X, rtol=1e-4)
    Xtr, ytr, Xte, yte = loaders.load_dataset(str(fi), str(fl), numtrain=4, numtest=2)
    assert Xtr.shape == (4, 3) and Xte.shape == (2, 3) and ytr.shape == (4,) and yte.shape == (2,)
    assert np.array_equal(Xte, Xa[4:6])                       # test rows follow the training rows (serial_gp.cpp:95)
    with pytest.raises(ValueError):
        loaders.load_dataset(str(fi), str(fl), numtrain=8)
    with pytest.raises(ValueError):
        loaders.load_inputs(str(fi), d=4)


def test_headerless_comma_separated_and_shards(tmp_path):
    X = np.arange(24, dtype=float).reshape(8, 3)
    y = np.arange(8, dtype=float)
    _write(tmp_path / "all_input.txt", "".join(",".join(repr(float(v)) for v in r) + "\n" for r in X))
    assert np.array_equal(loaders.load_inputs(str(tmp_path / "all_input.txt")), X)
    parts = np.array_split(np.arange(8), 3)                    # scaling_dataset/1.py: numpy.array_split
    for k, idx in enumerate(parts):
        _write(tmp_path / f"chunk{k}.txt", f"{len(idx)} 3\n" + "".join(" ".join(repr(float(v)) for v in X[i]) + " \n" for i in idx))
        _write(tmp_path / f"label{k}.txt", "".join(f"{float(y[i])!r}\n" for i in idx))
    Xs, ys, counts = loaders.load_shards(str(tmp_path / "chunk"), str(tmp_path / "label"), 3)
    assert counts == [3, 3, 2] and np.array_equal(Xs, X) and np.array_equal(ys, y)


def test_synthetic_generator_is_the_survey_recipe():
    X, y = loaders.synthetic_sine(100, 10)
    rng = np.random.default_rng(15618)
    X0 = rng.uniform(-10, 10, (100, 10))
    assert np.array_equal(X, X0) and np.allclose(y - np.sin(X[:, 0]), 0.1 * rng.standard_normal(100))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference data only exists in the build container")
def test_reference_files():
    d = os.path.join(REF, "chunked_dataset")
    X = loaders.load_inputs(os.path.join(d, "sine_dataset_1024_10_chunk0.txt"))
    assert X.shape == (2000, 10)                              # header says 1024
    Xtr, ytr, Xte, yte = loaders.load_dataset(os.path.join(d, "sine_dataset_1024_10_chunk0.txt"),
                                              os.path.join(d, "sine_dataset_1024_10_label0.txt"), 1024)
    assert Xtr.shape == (1024, 10) and Xte.shape == (976, 10) and yte.shape == (976,)
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "data_sine1024.npz"))
    assert np.array_equal(Xtr, fx["X"][:1024]) and np.array_equal(ytr, fx["y"][:1024])   # the committed fixture
    s = os.path.join(REF, "scaling_dataset")
    Xs, ys, counts = loaders.load_shards(os.path.join(s, "si24000_16sharded_chunk"), os.path.join(s, "si24000_16sharded_label"), 16)
    assert counts == [1500] * 16
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "data_si24000.npz"))
    assert np.array_equal(Xs, fx["X"]) and np.array_equal(ys, fx["y"])
