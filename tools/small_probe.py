import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treegp_b200 import backend
def avg(fn, reps=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps * 1e3
n = 64
A = torch.randn((n, 64), dtype=torch.float64, device="cuda"); K = A @ A.T + 64 * torch.eye(n, dtype=torch.float64, device="cuda")
ws = backend.alloc_matrix(n, n); ws[:, :n] = K
print("potf2 (n=64): %.1f us" % avg(lambda: backend.potrf(ws, n)))   # includes memset of info
L = torch.linalg.cholesky(K).contiguous(); Lw = backend.alloc_matrix(n, n); Lw[:, :n] = L
for M in (64, 448, 4096, 40000):
    B = torch.randn((M, 64), dtype=torch.float64, device="cuda")
    print("trsm_panel M=%d: %.1f us" % (M, avg(lambda: backend.trsm_rows(Lw, n, B, M))))
for (m, nc, k) in ((448, 448, 64), (64, 64, 64), (4096, 448, 64), (4096, 256, 256), (4096, 512, 512), (3584, 3584, 512), (40000, 448, 64)):
    C = torch.zeros((m, backend.even(nc)), dtype=torch.float64, device="cuda"); Aa = torch.randn((m, k), dtype=torch.float64, device="cuda"); Bb = torch.randn((nc, k), dtype=torch.float64, device="cuda")
    t = avg(lambda: backend.gemm_nt_sub(C, m, nc, Aa, Bb, k), reps=50)
    print("gemm m=%d n=%d k=%d: %.1f us (%.1f TF)" % (m, nc, k, t, 2.0 * m * nc * k / t / 1e6))
n = 512
A = torch.randn((n, 64), dtype=torch.float64, device="cuda"); K = A @ A.T + 512 * torch.eye(n, dtype=torch.float64, device="cuda")
ws = backend.alloc_matrix(n, n); keep = ws.clone(); keep[:, :n] = K
def f():
    ws.copy_(keep); backend.potrf(ws, n)
print("potrf 512 block (24 launches + copy): %.1f us" % avg(f, reps=50))
print("empty launch pair (copy only): %.1f us" % avg(lambda: ws.copy_(keep), reps=50))
