"""Dataset loaders for the reference's text formats.

Formats (reference ``chunked_dataset/`` and ``scaling_dataset/``):

* inputs: first line ``n d``; then rows of ``d`` whitespace-separated floats.  The header can LIE
  (``sine_dataset_1024_10_chunk0.txt`` says ``1024 10`` and holds 2000 rows): the drivers read
  ``numtrain`` from a literal and use the following rows as the test set
  (``cpp_serial_gp/serial_gp.cpp:33-54,95``; ``cuda_src/cuda_gp.cu:925-947``), so this loader reads
  rows until EOF and lets the caller split.
* labels: one float per line, no header.
* ``si24000_all_*``: no header, comma separated (the source the 16 shard files were split from with
  ``numpy.array_split``, ``scaling_dataset/1.py``).

Parsing happens ONCE; the product keeps the arrays device resident (the reference re-parses the shard
text on every log-likelihood and every gradient evaluation, ``cuda_scalingdist/cg_solver.cpp:45-52``).
"""
from __future__ import annotations

import numpy as np


def load_inputs(path: str, d: int | None = None) -> np.ndarray:
    """Read an input file -> (rows, d) float64.  Header ``n d`` supplies d; rows are read to EOF."""
    with open(path, "r") as f:
        text = f.read()
    first, _, rest = text.partition("\n")
    head = first.replace(",", " ").split()
    if len(head) == 2 and all(tok.lstrip("+-").isdigit() for tok in head):
        d_file = int(head[1])
        body = rest
    else:  # headerless (si24000_all_input.txt): d = number of fields on the first line
        d_file = len(head)
        body = text
    if d is not None and d != d_file:
        raise ValueError(f"{path}: file says d={d_file}, caller says d={d}")
    vals = np.array(body.replace(",", " ").split(), dtype=np.float64)
    if vals.size % d_file:
        raise ValueError(f"{path}: {vals.size} values is not a multiple of d={d_file}")
    return vals.reshape(-1, d_file)


def load_labels(path: str) -> np.ndarray:
    with open(path, "r") as f:
        return np.array(f.read().replace(",", " ").split(), dtype=np.float64)


def load_dataset(input_path: str, label_path: str, numtrain: int | None = None, numtest: int | None = None):
    """Returns (Xtrain, ytrain, Xtest, ytest): first ``numtrain`` rows train, the following rows test."""
    X, y = load_inputs(input_path), load_labels(label_path)
    rows = min(X.shape[0], y.shape[0])
    X, y = X[:rows], y[:rows]
    if numtrain is None:
        numtrain = rows
    if numtrain > rows:
        raise ValueError(f"numtrain={numtrain} exceeds the {rows} rows in {input_path}")
    stop = rows if numtest is None else min(rows, numtrain + numtest)
    return X[:numtrain], y[:numtrain], X[numtrain:stop], y[numtrain:stop]


def load_shards(input_prefix: str, label_prefix: str, chunks: int):
    """Read ``<prefix><k>.txt`` for k < chunks (``cuda_scalingdist/main.cpp:247-252`` argv convention)
    and return the concatenation plus the per-shard row counts."""
    Xs, ys = [], []
    for k in range(chunks):
        Xs.append(load_inputs(f"{input_prefix}{k}.txt"))
        ys.append(load_labels(f"{label_prefix}{k}.txt")[: Xs[-1].shape[0]])
    return np.concatenate(Xs), np.concatenate(ys), [x.shape[0] for x in Xs]


def synthetic_sine(n: int, d: int = 10, seed: int = 15618, lo: float = -10.0, hi: float = 10.0,
                   noise: float = 0.1):
    """The generator recovered from the shipped data: X ~ U(lo,hi)^d,
    y = sin(x_0) + noise * N(0,1)."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(lo, hi, (n, d))
    y = np.sin(X[:, 0]) + noise * rng.standard_normal(n)
    return X, y
