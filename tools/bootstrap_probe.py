#!/usr/bin/env python
"""cProfile of the batched bootstrap (configs[4]: N=200k, 100 resamples, nbins=21, default max_sep)."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import treegp_b200 as treegp
n = int(os.environ.get("PN", 200000)); B = int(os.environ.get("PB", 100))
rng = np.random.default_rng(42); L = 1000.0 * np.sqrt(n / 1e6)
X = rng.uniform(0, L, size=(n, 2)); y = rng.normal(size=n); e = np.full(n, 0.1)
tp = treegp.two_pcf(X, y, e, 0.0, 0.5 * np.hypot(L, L), nbins=21, anisotropic=True)
def one():
    tp._rng = None
    r = tp._bootstrap_xi(B); torch.cuda.synchronize(); return r
one()
t0 = time.perf_counter(); one(); print("bootstrap wall: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
pr = cProfile.Profile(); pr.enable(); one(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
