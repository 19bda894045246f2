#!/usr/bin/env python
"""Phase timestamps (globaltimer, ns) inside panel_left_kernel: builds a private copy of the library with
-DTGP_PANEL_TIMING and factorises a 512 x 512 block; prints the phases of the LAST launch (64 right-hand-side rows ride along, so the last launch (block column 7)
has CTA 0 = potf2 of block (7,7) and CTA 1 = the extra row block with a K = 448 left-looking update) and of a launch with K = 0."""
import ctypes, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = "/tmp/libtgp_timing.so"
srcs = ["kmat.cu", "dense.cu", "trsv.cu", "predict.cu", "pairbin.cu", "microbench.cu", "hostrng.cu", "vcorr.cu"]
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                       "-DTGP_PANEL_TIMING", "-Xcompiler", "-fPIC", "-shared", "-o", so] +
                      [os.path.join(ROOT, "treegp_b200/csrc", f) for f in srcs])
lib = ctypes.CDLL(so)
vp, i64 = ctypes.c_void_p, ctypes.c_int64
lib.tgp_potrf_rows.argtypes = [vp, i64, i64, i64, vp, vp]
def run(n, reps=20):
    A = torch.randn((n, 64), dtype=torch.float64, device="cuda")
    K = A @ A.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
    ws = torch.zeros((n + 64, n), dtype=torch.float64, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    for _ in range(reps):
        ws[:n].copy_(K); ws[n:].fill_(1.0)
        assert lib.tgp_potrf_rows(ws.data_ptr(), n, n, 64, info.data_ptr(), None) == 0
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 32)()
    assert lib.tgp_debug_panel_times(buf) == 0
    t = np.array(buf[:], dtype=np.int64)
    pf = (ctypes.c_ulonglong * 4)()
    assert lib.tgp_debug_potf2_cycles(pf) == 0
    print("potf2 phase cycles (last diagonal block): (i) 16x16 pivots %d, (ii) rows below %d, (iii) trailing update %d" % (pf[0], pf[1], pf[2]))
    err = (torch.linalg.cholesky(K) - torch.tril(ws[:n])).abs().max().item()
    return t, err
names = {0: "diag start", 1: "diag loaded", 2: "diag potf2 done", 3: "diag stored+flag",
         8: "row start", 9: "row T ready (gemm done)", 10: "row flag seen", 11: "row L loaded", 12: "row trsm done",
         13: "row own-diag update done", 14: "row stored"}
for n in (512, 128):
    t, err = run(n)
    t0 = min(t[0], t[8])
    print("n=%d (last launch with a row block), max|L-Lref| = %.2e" % (n, err))
    prev = {0: None, 8: None}
    for i in sorted(names):
        base = 0 if i < 8 else 8
        dc = (t[16 + i] - t[16 + prev[base]]) if prev[base] is not None else 0
        prev[base] = i
        print("  %-28s %8.2f us   (+%d cycles)" % (names[i], (t[i] - t0) * 1e-3, dc))
