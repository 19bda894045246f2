import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treegp_b200 import backend
m = int(os.environ.get("G_M", "8192")); k = int(os.environ.get("G_K", "512"))
C = torch.zeros((m, m), dtype=torch.float64, device="cuda"); A = torch.randn((m, k), dtype=torch.float64, device="cuda")
for _ in range(3):
    backend.gemm_nt_sub(C, m, m, A, A, k, lower_only=True)
torch.cuda.synchronize(); print("ok")
