#!/usr/bin/env python
"""Isotropic (Log) pair binning rate at N = 200k / 1M, block forms on and off."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treegp_b200 import _cabi, backend, binning
for n in (200000, 1000000):
    rng = np.random.default_rng(0); L = 1000.0 * np.sqrt(n / 1e6)
    x = backend.to_device(rng.uniform(0, L, n)); y = backend.to_device(rng.uniform(0, L, n)); k = backend.to_device(rng.normal(size=n))
    o = backend.hilbert_order(x, y); x, y, k = x[o].contiguous(), y[o].contiguous(), k[o].contiguous()
    off = backend.to_device(np.array([0, n]), torch.int64)
    mn, mx, nb = np.sqrt(1.0), 0.5 * np.hypot(L, L), 20      # the reference's isotropic defaults (two_pcf.py:400-424)
    edges = backend.to_device(binning.log_thresholds(mn, mx, nb))
    for mode in (1, 0):
        backend.set_option("pairbin_block_sums", mode)
        for i in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); res = backend.pairbin(x, y, k, None, off, n, _cabi.BIN_LOG, edges, nb, mn, mx); e1.record()
            torch.cuda.synchronize(); t = e0.elapsed_time(e1) * 1e-3
        print("Log bins n=%d block_sums=%d: %.4f s  %.1f Gpairs/s  in range: %d" % (n, mode, t, n * (n - 1) / 2 / t / 1e9, int(res[0].sum().item())), flush=True)
backend.set_option("pairbin_block_sums", 1)
