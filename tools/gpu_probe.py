#!/usr/bin/env python
"""Quick on-GPU probe: FP64 micro-peaks and first timings of every hot-path kernel (CUDA events).
Writes gpurun_out/probe.json.  Not a benchmark of record -- bench.py is."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treegp_b200 import _cabi, backend, binning, eval_kernel  # noqa: E402
from treegp_b200.kernels import lower_kernel  # noqa: E402


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts)


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    out["dfma_tflops"] = backend.microbench_fp64(0, 20000)
    out["dmma_tflops"] = backend.microbench_fp64(1, 20000)
    print(out, flush=True)
    rng = np.random.default_rng(0)
    sizes = [int(s) for s in os.environ.get("PROBE_N", "4096,10000,20000").split(",")]
    for name, s in (("rbf", "4.0 * AnisotropicRBF(invLam=array([[4.2, -0.8], [-0.8, 5.1]]))"),
                    ("vk", "4.0 * AnisotropicVonKarman(invLam=array([[0.46, -0.09], [-0.09, 0.55]]))")):
        desc = lower_kernel(eval_kernel(s), 2)
        for n in sizes:
            L = 0.2 * np.sqrt(n) * 2
            X = backend.to_device(rng.uniform(-L / 2, L / 2, size=(n, 2)))
            ws = backend.alloc_matrix(n, n)
            d = backend.to_device(np.full(n, 1e-2))
            t_full = timed(lambda: backend.kmat_sym(X, desc, d, out=ws))
            t_low = timed(lambda: backend.kmat_sym(X, desc, d, out=ws, lower_only=True))
            out["kmat_%s_%d" % (name, n)] = dict(full_s=t_full, full_GBs=8 * n * n / t_full / 1e9,
                                                 lower_s=t_low, lower_GBs=4 * n * n / t_low / 1e9)
            print(name, n, out["kmat_%s_%d" % (name, n)], flush=True)
            if name == "rbf":
                def chol():
                    backend.kmat_sym(X, desc, d, out=ws, lower_only=True)
                    backend.potrf(ws, n)
                t = timed(chol, reps=2) - t_low
                out["potrf_%d" % n] = dict(s=t, tflops=n ** 3 / 3 / t / 1e12)
                backend.kmat_sym(X, desc, d, out=ws)
                A = ws[:, :n].contiguous()
                tc = timed(lambda: torch.linalg.cholesky(A), reps=2)
                out["cusolver_potrf_%d" % n] = dict(s=tc, tflops=n ** 3 / 3 / tc / 1e12)
                del A
                # check factor vs cuSOLVER
                backend.kmat_sym(X, desc, d, out=ws, lower_only=True)
                info = backend.potrf(ws, n)
                b = backend.to_device(rng.normal(size=n))
                ts = timed(lambda: backend.potrs_vec(ws, n, b.clone()))
                out["potrs_vec_%d" % n] = dict(s=ts, GBs=8 * n * n / ts / 1e9, info=int(info.item()))
                y = backend.to_device(rng.normal(size=n))
                tl = timed(lambda: backend.loglike(X, y, d, desc, work=ws), reps=2)
                out["loglike_%d" % n] = dict(s=tl)
                print(n, out["potrf_%d" % n], out["cusolver_potrf_%d" % n], out["potrs_vec_%d" % n], out["loglike_%d" % n], flush=True)
            m = 200000
            Xs = backend.to_device(rng.uniform(-L / 2, L / 2, size=(m, 2)))
            alpha = backend.to_device(rng.normal(size=n))
            tp = timed(lambda: backend.predict_mean(Xs, X, desc, alpha), reps=2)
            out["predict_mean_%s_%d" % (name, n)] = dict(s=tp, gevals=m * n / tp / 1e9)
            print("predict", name, n, out["predict_mean_%s_%d" % (name, n)], flush=True)
            del ws
    # pair binning
    for n in [int(s) for s in os.environ.get("PROBE_PB", "50000,200000").split(",")]:
        Lf = 1000.0
        x = backend.to_device(rng.uniform(0, Lf, n))
        y = backend.to_device(rng.uniform(0, Lf, n))
        k = backend.to_device(rng.normal(size=n))
        order = backend.hilbert_order(x, y)
        x, y, k = x[order].contiguous(), y[order].contiguous(), k[order].contiguous()
        off = backend.to_device(np.array([0, n]), torch.int64)
        for label, mx in (("default", Lf * np.sqrt(2) / 2), ("small", Lf / 100)):
            edges = backend.to_device(binning.twod_thresholds(mx, 21))
            t = timed(lambda: backend.pairbin(x, y, k, None, off, n, _cabi.BIN_TWOD, edges, 21, 0.0, mx), reps=2)
            out["pairbin_twod_%s_%d" % (label, n)] = dict(s=t, gpairs=n * (n - 1) / 2 / t / 1e9)
            print("pairbin", label, n, out["pairbin_twod_%s_%d" % (label, n)], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe.json", "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
