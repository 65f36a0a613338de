"""Re-pack the reference's text datasets the parity tests need as compressed float64 arrays.

Run in the build container (needs /root/reference):  python tests/golden/make_fixtures.py
The GPU box has no /root/reference, so tests read these .npz files instead of the text.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cugp_b200.loaders import load_inputs, load_labels  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
cd, sd = f"{REF}/chunked_dataset", f"{REF}/scaling_dataset"

# 128 x 2: the data behind the reference's run logs (cuda_bettersinglenode_ver2/REF, cuda_ref/seeee)
X = np.concatenate([load_inputs(f"{cd}/si_chunk{k}.txt") for k in (0, 1)])
y = np.concatenate([load_labels(f"{cd}/si_label{k}.txt") for k in (0, 1)])
assert X.shape == (128, 2) and y.shape == (128,)
np.savez_compressed(f"{OUT}/data_si128x2.npz", X=X, y=y)

# sine_dataset_{256,1024}_10 hold the same 2000 rows; _4096_10 holds 6000 rows with the same prefix
X1, y1 = load_inputs(f"{cd}/sine_dataset_1024_10_chunk0.txt"), load_labels(f"{cd}/sine_dataset_1024_10_label0.txt")
X4, y4 = load_inputs(f"{cd}/sine_dataset_4096_10_chunk0.txt"), load_labels(f"{cd}/sine_dataset_4096_10_label0.txt")
X2 = load_inputs(f"{cd}/sine_dataset_256_10_chunk0.txt")
assert X1.shape == (2000, 10) and X4.shape == (6000, 10) and np.array_equal(X1, X2)
same_prefix = np.array_equal(X1[:1024], X4[:1024]) and np.array_equal(y1[:1024], y4[:1024])
print("1024-file[:1024] == 4096-file[:1024]:", same_prefix, "| full 2000-row prefix:", np.array_equal(X1, X4[:2000]))
np.savez_compressed(f"{OUT}/data_sine1024.npz", X=X1[:1040], y=y1[:1040])       # 1024 train + 16 test
np.savez_compressed(f"{OUT}/data_sine4096.npz", X=X4[:4112], y=y4[:4112])       # 4096 train + 16 test

# C4: si24000 (16 shards x 1500) and the first 16 rows of siproper_10000_10 as test points
Xa, ya = load_inputs(f"{sd}/si24000_all_input.txt"), load_labels(f"{sd}/si24000_all_label.txt")
assert Xa.shape == (24000, 10) and ya.shape == (24000,)
for k in (0, 7, 15):
    Xk = load_inputs(f"{sd}/si24000_16sharded_chunk{k}.txt")
    yk = load_labels(f"{sd}/si24000_16sharded_label{k}.txt")
    assert np.array_equal(Xk, Xa[1500 * k:1500 * (k + 1)]) and np.array_equal(yk, ya[1500 * k:1500 * (k + 1)])
Xp, yp = load_inputs(f"{cd}/siproper_10000_10_chunk0.txt"), load_labels(f"{cd}/siproper_10000_10_label0.txt")
np.savez_compressed(f"{OUT}/data_si24000.npz", X=Xa, y=ya, Xtest=Xp[:16], ytest=yp[:16])
for f in sorted(os.listdir(OUT)):
    if f.endswith(".npz"):
        print(f, os.path.getsize(f"{OUT}/{f}"))

# C4 at 600 test points: rows 0..599 of siproper_10000_10 (the C4 test set is that file)
np.savez_compressed(f"{OUT}/data_c4_xtest600.npz", Xtest=Xp[:600], ytest=yp[:600])
print("data_c4_xtest600.npz", os.path.getsize(f"{OUT}/data_c4_xtest600.npz"))
