#!/usr/bin/env python
"""Mean-function lookup at the configs[2] / configs[4] sizes: tgp_knn_mean vs sklearn's KD-tree on the host."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sklearn.neighbors import KNeighborsRegressor
from treegp_b200 import backend
rng = np.random.default_rng(0)
g = np.linspace(-80, 80, 50); X0 = np.array([[a, b] for a in g for b in g]); y0 = rng.normal(size=len(X0))
for m in (200000, 1000000):
    Xq = rng.uniform(-80, 80, size=(m, 2))
    t0 = time.perf_counter(); ref = KNeighborsRegressor(n_neighbors=4).fit(X0, y0).predict(Xq); t_cpu = time.perf_counter() - t0
    backend.knn_mean(X0, y0, Xq, 4); torch.cuda.synchronize()
    t0 = time.perf_counter(); got = backend.knn_mean(X0, y0, Xq, 4).cpu().numpy(); t_e2e = time.perf_counter() - t0
    Xd, X0d, y0d = backend.as_points(Xq), backend.as_points(X0), backend.to_device(y0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); backend.knn_mean(X0d, y0d, Xd, 4); b.record(); torch.cuda.synchronize()
    print("M=%d vs 2500 grid points: sklearn (host, %d cores) %.1f ms | device kernel %.2f ms, from host arrays %.1f ms | max diff %.1e"
          % (m, os.cpu_count(), t_cpu * 1e3, a.elapsed_time(b), t_e2e * 1e3, np.abs(got - ref).max()))
