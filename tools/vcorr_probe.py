#!/usr/bin/env python
"""E/B pair sums at the reference's cap (maxpts = 30000 points): device kernel vs the numpy oracle loop."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import eb_oracle
from treegp_b200 import backend
rng = np.random.default_rng(1); n = int(os.environ.get("PN", 30000))
x, y, dx, dy = rng.uniform(0, 1.5, n), rng.uniform(0, 1.5, n), rng.normal(size=n), rng.normal(size=n)
rmin, rmax, dlogr = 5.0 / 3600.0, 1.5, 0.05; bins = int(np.ceil(np.log(rmax / rmin) / dlogr))
backend.vcorr_sums(x[:1000], y[:1000], dx[:1000], dy[:1000], np.log(rmin), dlogr, bins); torch.cuda.synchronize()
t0 = time.perf_counter(); got = backend.vcorr_sums(x, y, dx, dy, np.log(rmin), dlogr, bins); t_gpu = time.perf_counter() - t0
m = 6000
t0 = time.perf_counter(); eb_oracle.pair_sums(x[:m], y[:m], dx[:m], dy[:m], np.log(rmin), dlogr, bins); t_cpu = (time.perf_counter() - t0) * (n / m) ** 2
print("n=%d (%.3g pairs, %d bins): device %.1f ms from host arrays | numpy oracle loop ~%.1f s (extrapolated from %d points)"
      % (n, n * (n - 1) / 2, bins, t_gpu * 1e3, t_cpu, m))
