#!/usr/bin/env python
"""Minimal pair-binning driver (for ncu captures and quick timing): N points, default max_sep, 4 launches."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treegp_b200 import _cabi, backend, binning
n = int(os.environ.get("PB_N", "200000")); reps = int(os.environ.get("PB_REPS", "4"))
weighted = os.environ.get("PB_W", "0") == "1"
frac = float(os.environ.get("PB_MAXSEP_FRAC", str(np.sqrt(2) / 2)))
rng = np.random.default_rng(0); L = 1000.0
x = backend.to_device(rng.uniform(0, L, n)); y = backend.to_device(rng.uniform(0, L, n)); k = backend.to_device(rng.normal(size=n))
w = backend.to_device(rng.uniform(0.5, 2, n)) if weighted else None
if os.environ.get("PB_SORT", "1") == "1":
    o = backend.hilbert_order(x, y); x, y, k = x[o].contiguous(), y[o].contiguous(), k[o].contiguous()
    w = None if w is None else w[o].contiguous()
off = backend.to_device(np.array([0, n]), torch.int64); mx = L * frac
edges = backend.to_device(binning.twod_thresholds(mx, 21))
if "PB_FAST" in os.environ:   # bit mask of the short-cut paths (default 3), see tgp_set_option
    backend.set_option("pairbin_fast_paths", int(os.environ["PB_FAST"]))
if "PB_BLOCK_SUMS" in os.environ:
    backend.set_option("pairbin_block_sums", int(os.environ["PB_BLOCK_SUMS"]))
backend.pairbin_stats(reset=True)
for i in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); res = backend.pairbin(x, y, k, w, off, n, _cabi.BIN_TWOD, edges, 21, 0.0, mx); e1.record()
    torch.cuda.synchronize(); t = e0.elapsed_time(e1) * 1e-3
    print("n=%d t=%.4fs  %.1f Gpairs/s  in-range(x2)=%d" % (n, t, n * (n - 1) / 2 / t / 1e9, int(res[0].sum().item())), flush=True)
import ctypes
raw = (ctypes.c_ulonglong * 8)(); _cabi.check(_cabi.load().tgp_pairbin_stats(raw, 1), "stats"); raw = list(raw); tot = max(1, sum(raw[:5]))
print("paths:", dict(zip(("closed_form", "one_axis", "pairwise", "one_axis_sorted", "two_axis_sorted"), (round(v / tot, 4) for v in raw[:5]))),
      "checksum", int((res[0] * torch.arange(res[0].numel(), device=res[0].device).reshape(res[0].shape)).sum().item()), flush=True)
