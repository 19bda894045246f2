import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treegp_b200 import backend
def avg(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps * 1e3
n = 4096
A = torch.randn((n, 64), dtype=torch.float64, device="cuda"); K = A @ A.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
Lw = backend.alloc_matrix(n, n); Lw[:, :n] = torch.linalg.cholesky(K)
for M in (64, 1024):
    B = torch.randn((M, n), dtype=torch.float64, device="cuda")
    t = avg(lambda: backend.trsm_rows(Lw, n, B, M))
    print("trsm_rows N=4096 M=%d: %.1f us total; 64 trsm_panel + 63 gemm launches -> %.1f us per launch" % (M, t, t / 127))
n2 = 64
for nn in (64, 128, 256, 512):
    ws = backend.alloc_matrix(nn, nn); keep = ws.clone(); keep[:, :nn] = K[:nn, :nn]
    def f():
        ws.copy_(keep); backend.potrf(ws, nn)
    t = avg(f, reps=50); t0 = avg(lambda: ws.copy_(keep), reps=50)
    nl = {64: 1, 128: 4, 256: 10, 512: 22}[nn]
    print("potrf n=%d: %.1f us (copy %.1f us)" % (nn, t - t0, t0))
