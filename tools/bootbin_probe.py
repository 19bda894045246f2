#!/usr/bin/env python
"""Timing of the bootstrap batch: shared-geometry kernel (tgp_bootbin_twod) against the per-catalogue batch
(tgp_pairbin on B weighted catalogues).  PN points, PB resamples, PMODES = comma list of bootbin_paths masks."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import treegp_b200 as treegp
from treegp_b200 import backend
n = int(os.environ.get("PN", 200000)); B = int(os.environ.get("PB", 100))
fracs = [float(f) for f in os.environ.get("PFRAC", "0.5").split(",")]
modes = [int(m) for m in os.environ.get("PMODES", "15").split(",")]
rng = np.random.default_rng(42); L = 1000.0 * np.sqrt(n / 1e6)
X = rng.uniform(0, L, size=(n, 2)); y = rng.normal(size=n); e = np.full(n, 0.1)
frac = fracs[0]
tp = treegp.two_pcf(X, y, e, 0.0, frac * np.hypot(L, L), nbins=21, anisotropic=True)
def run(shared):
    tp.SHARED_BOOTSTRAP = shared; tp._rng = None
    t0 = time.perf_counter(); r = tp._bootstrap_xi(B); torch.cuda.synchronize()
    return r, time.perf_counter() - t0
for frac in fracs[1:] if len(fracs) > 1 else []:
    tq = treegp.two_pcf(X, y, e, 0.0, frac * np.hypot(L, L), nbins=21, anisotropic=True)
    chunk = np.sqrt(32.0 * L * L / n); binw = 2 * tq.max_sep / 21
    res = []
    for shared in (False, True):
        tq.SHARED_BOOTSTRAP = shared; tq._rng = None; tq._bootstrap_xi(B); torch.cuda.synchronize()
        tq._rng = None; backend.bootbin_stats(reset=True); t0 = time.perf_counter(); r = tq._bootstrap_xi(B); torch.cuda.synchronize()
        res.append((time.perf_counter() - t0, r))
    print("frac %.3f  2*chunk/bin = %.2f: per-catalogue %.1f ms, shared %.1f ms, max diff %.2e  %s" % (
        frac, 2 * chunk / binw, res[0][0] * 1e3, res[1][0] * 1e3, np.max(np.abs(res[0][1] - res[1][1])), backend.bootbin_stats()), flush=True)
if len(fracs) > 1:
    sys.exit(0)
ref = None
if os.environ.get("POLD", "1") == "1":
    run(False); ref, t = run(False); print("per-catalogue batch: %.1f ms" % (t * 1e3), flush=True)
for m in modes:
    backend.set_option("bootbin_paths", m)
    run(True); backend.bootbin_stats(reset=True)
    got, t = run(True); st = backend.bootbin_stats()
    err = np.max(np.abs(got - ref)) if ref is not None else float("nan")
    print("shared geometry paths=%d: %.1f ms  max|xi - xi_ref| = %.2e  %s" % (m, t * 1e3, err, st), flush=True)
    # device time of the kernel chain alone
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x = backend.to_device(X[:, 0]); yy = backend.to_device(X[:, 1]); val = backend.to_device(y)
    order = backend.hilbert_order(x, yy); x, yy, val = x[order].contiguous(), yy[order].contiguous(), val[order].contiguous()
    w = torch.full_like(x, 100.0)
    mult = torch.from_numpy(np.random.default_rng(1).multinomial(n, np.full(n, 1.0 / n), size=B).astype(np.uint8)).cuda()
    ed = tp._device_edges(tp._bin_geometry()[1])
    backend.bootbin_sums(x, yy, val, w, mult, ed, 21, 0.0, tp.max_sep); torch.cuda.synchronize()
    ev0.record(); backend.bootbin_sums(x, yy, val, w, mult, ed, 21, 0.0, tp.max_sep); ev1.record(); torch.cuda.synchronize()
    print("   device time of tgp_bootbin_twod: %.1f ms" % ev0.elapsed_time(ev1), flush=True)
backend.set_option("bootbin_paths", 15)
