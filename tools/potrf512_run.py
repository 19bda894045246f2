import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treegp_b200 import backend
n = int(os.environ.get("P_N", "1024"))
A = torch.randn((n, 64), dtype=torch.float64, device="cuda"); ws = backend.alloc_matrix(n, n); ws[:, :n] = A @ A.T; ws[:, :n].diagonal().add_(float(n)); keep = ws.clone()
for _ in range(2):
    ws.copy_(keep); backend.potrf(ws, n)
torch.cuda.synchronize(); print("ok")
