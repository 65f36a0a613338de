"""Gradient at C5 size (n = 100 000): the in-place inverse path (one 80 GB buffer + 21 GB scratch)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import cugp_b200 as cg
from cugp_b200.loaders import synthetic_sine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
TH_B = [3.762111, -1.152105, -0.384461]
X, y = synthetic_sine(n, 10)
g = cg.Covsum(n, 10)
g.set_data(X, y)
g.set_loghyperparam(TH_B)
t = time.perf_counter()
ll = g.loglik_resident()
t1 = time.perf_counter()
gr = g.grad_resident()
t2 = time.perf_counter()
import torch
free, total = torch.cuda.mem_get_info()
print(f"n={n}: LL={ll:.9f} ({t1-t:.2f} s)  grad={gr} ({t2-t1:.2f} s = {2*n**3/3/(t2-t1)/1e12:.1f} TF for the inverse)  device memory in use {(total-free)/2**30:.1f} GiB", flush=True)
# finite-difference check of the gradient component with respect to log sigma_n (cheap: two more LL evaluations)
eps = 1e-4
th = list(TH_B); th[2] += eps; g.set_loghyperparam(th); lp = g.loglik_resident()
th[2] -= 2 * eps; g.set_loghyperparam(th); lm = g.loglik_resident()
fd = -(lp - lm) / (2 * eps)
print(f"d(-LL)/d(log sigma_n): analytic {gr[2]:.6f}  central difference {fd:.6f}  rel diff {abs(fd-gr[2])/abs(gr[2]):.2e}", flush=True)
