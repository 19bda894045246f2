#!/usr/bin/env python
"""Where the fixed cost of one two_pcf.comp_2pcf call goes (N = 1e6, host arrays in, xi out): each stage timed with
a synchronize on both sides (so the stages do not overlap as they do in the real call)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import treegp_b200 as treegp
from treegp_b200 import _cabi, backend, binning
n = int(os.environ.get("PB_N", 1000000)); L = 1000.0
rng = np.random.default_rng(42)
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True); v = t.numpy(); v[...] = a; return v
X = pinned(rng.uniform(-L / 2, L / 2, size=(n, 2))); y = pinned(rng.normal(size=n)); ye = pinned(np.zeros(n))
mx = np.sqrt(2.0) * L / 2
tp = treegp.two_pcf(X, y, ye, 0.0, mx, nbins=21, anisotropic=True)
for _ in range(3): tp.comp_2pcf(X, y, ye)
torch.cuda.synchronize()
def T(name, fn, reps=5):
    out = None; ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print("%-46s %8.3f ms" % (name, 1e3 * min(ts)), flush=True); return out
T("whole comp_2pcf", lambda: tp.comp_2pcf(X, y, ye))
T("np.sum(y_err) + np.mean(y) (host)", lambda: (np.sum(ye), np.mean(y)))
Xd = T("H2D X (16 MB) + y (8 MB), pinned", lambda: (backend.to_device(X, non_blocking=True), backend.to_device(y, non_blocking=True)))
Xd, yd = Xd
px, py = T("split columns", lambda: (Xd[:, 0].contiguous(), Xd[:, 1].contiguous()))
pk = T("pk = y - mean", lambda: yd - 0.1)
order = T("hilbert_order (minmax + keys + argsort)", lambda: backend.hilbert_order(px, py))
T("  torch.aminmax x2 + tolist", lambda: torch.stack(torch.aminmax(px) + torch.aminmax(py)).tolist())
keys = torch.empty(n, dtype=torch.int64, device=px.device)
T("  torch.argsort(int64 keys)", lambda: torch.argsort(order))
sx, sy, sk = T("3 gathers", lambda: (px[order], py[order], pk[order]))
off = backend.to_device(np.array([0, n]), torch.int64); edges = backend.to_device(binning.twod_thresholds(mx, 21))
T("offsets + edges upload", lambda: (backend.to_device(np.array([0, n]), torch.int64), backend.to_device(binning.twod_thresholds(mx, 21))))
res = T("pairbin_packed (pre-pass + kernel)", lambda: backend.pairbin_packed(sx, sy, sk, None, off, n, _cabi.BIN_TWOD, edges, 21, 0.0, mx))
T("D2H packed bins", lambda: res.cpu().numpy())
