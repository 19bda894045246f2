#!/usr/bin/env python
"""predict_mean at the configs[2] shape (N=40k train, M=1e6 test, AnisotropicVonKarman): full sum vs truncated."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treegp_b200 import backend, eval_kernel
from treegp_b200.kernels import lower_kernel
from treegp_b200.two_pcf import get_correlation_length_matrix
n, m = int(os.environ.get("PN", 40000)), int(os.environ.get("PM", 1000000))
L = 160.0 * np.sqrt(n / 40000.0)
inv = np.linalg.inv(get_correlation_length_matrix(1.5, 0.2, 0.2))
kstr = "4.0 * AnisotropicVonKarman(invLam=array([[%.17g, %.17g], [%.17g, %.17g]]))" % (inv[0, 0], inv[0, 1], inv[1, 0], inv[1, 1])
desc = lower_kernel(eval_kernel(kstr), 2)
rng = np.random.default_rng(42)
X = backend.as_points(rng.uniform(-L / 2, L / 2, size=(n, 2)))
Xs = backend.as_points(rng.uniform(-L / 2, L / 2, size=(m, 2)))
alpha = backend.to_device(rng.normal(size=n))
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts), r
tf, full = timed(lambda: backend.predict_mean(Xs, X, desc, alpha, truncate=False))
tt, tr = timed(lambda: backend.predict_mean(Xs, X, desc, alpha, truncate=True))
print("full  : %.1f ms  (%.3g kernel evals/s)" % (tf * 1e3, n * m / tf))
print("trunc : %.1f ms incl. Hilbert sort of both sets; max |diff| = %.3e (sum|amp alpha| = %.3e)" % (
    tt * 1e3, (full - tr).abs().max().item(), 4.0 * alpha.abs().sum().item()))
