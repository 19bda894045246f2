#!/usr/bin/env python
"""Covariance construction rates (HBM-store roofline for RBF; FP64-ALU for von Karman) at N = 40k."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treegp_b200 import backend, eval_kernel
from treegp_b200.kernels import lower_kernel
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts)
n = int(os.environ.get("PN", 40000)); m = 20000
rng = np.random.default_rng(0)
X = backend.as_points(rng.uniform(-80, 80, size=(n, 2))); Xs = backend.as_points(rng.uniform(-80, 80, size=(m, 2)))
e2 = backend.to_device(np.full(n, 1e-4))
ws = backend.alloc_matrix(n, n); wc = backend.alloc_matrix(m, n)
for name, s in (("AnisotropicRBF", "2.0 * AnisotropicRBF(invLam=array([[0.5, 0.1], [0.1, 0.4]]))"),
                ("AnisotropicVonKarman", "2.0 * AnisotropicVonKarman(invLam=array([[0.5, 0.1], [0.1, 0.4]]))"),
                ("Matern32", "2.0 * Matern(length_scale=1.5, nu=1.5)")):
    d = lower_kernel(eval_kernel(s), 2)
    tf = timed(lambda: backend.kmat_sym(X, d, e2, out=ws, lower_only=False))
    tl = timed(lambda: backend.kmat_sym(X, d, e2, out=ws, lower_only=True))
    tc = timed(lambda: backend.kmat_cross(Xs, X, d, out=wc))
    print("%-22s N=%d: full %.2f ms = %.0f GB/s (8 N^2) | lower %.2f ms = %.0f GB/s (4 N^2) | cross %dx%d %.2f ms = %.0f GB/s"
          % (name, n, tf * 1e3, 8.0 * n * n / tf / 1e9, tl * 1e3, 4.0 * n * n / tl / 1e9, m, n, tc * 1e3, 8.0 * m * n / tc / 1e9))
