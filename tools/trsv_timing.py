#!/usr/bin/env python
"""Where a step of the triangular sweep's serial chain goes: builds a private copy of the library with
-DTGP_TRSV_TIMING (globaltimer stamps per diagonal block) and prints the medians for the forward sweep."""
import ctypes, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = "/tmp/libtgp_trsv_timing.so"
srcs = ["kmat.cu", "dense.cu", "trsv.cu", "predict.cu", "pairbin.cu", "microbench.cu", "hostrng.cu", "vcorr.cu"]
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                       "-DTGP_TRSV_TIMING", "-Xcompiler", "-fPIC", "-shared", "-o", so] +
                      [os.path.join(ROOT, "treegp_b200/csrc", f) for f in srcs])
lib = ctypes.CDLL(so)
vp, i64 = ctypes.c_void_p, ctypes.c_int64
lib.tgp_trsv_only.argtypes = [vp, i64, i64, vp, ctypes.c_int, vp]
for n in (4096, 10000, 40000):
    ws = torch.randn((n, n), dtype=torch.float64, device="cuda") * (0.5 / np.sqrt(n))
    ws.diagonal().fill_(1.5)
    b = torch.randn(n, dtype=torch.float64, device="cuda")
    for _ in range(3):
        assert lib.tgp_trsv_only(ws.data_ptr(), n, n, b.clone().data_ptr(), 1, None) == 0
    torch.cuda.synchronize()
    nb = min((n + 63) // 64, 4096)
    buf = (ctypes.c_ulonglong * (3 * nb))()
    assert lib.tgp_debug_trsv_stamps(buf, 3 * nb) == 0
    t = np.array(buf[:], dtype=np.int64).reshape(nb, 3)[1:]
    step = np.diff(t[:, 2])
    hop = t[1:, 0] - t[:-1, 2]
    comp = t[:, 2] - t[:, 0]
    print("n=%d forward sweep: chain step median %.0f ns (p10 %.0f, p90 %.0f) | previous publish -> seen %.0f ns | seen -> published %.0f ns"
          " | total %.3f ms" % (n, np.median(step), np.percentile(step, 10), np.percentile(step, 90),
          np.median(hop), np.median(comp), (t[-1, 2] - t[0, 0]) * 1e-6))
    for which, name in ((1, "forward"), (5, "forward, dependencies off (streaming only)"), (2, "backward"), (6, "backward, dependencies off")):
        ts = []
        for _ in range(4):
            bb = b.clone(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); lib.tgp_trsv_only(ws.data_ptr(), n, n, bb.data_ptr(), which, None); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print("   %-48s %.3f ms  (%.0f GB/s of 4 N^2 bytes)" % (name, min(ts), 4.0 * n * n / min(ts) / 1e6))
