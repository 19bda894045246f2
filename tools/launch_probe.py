import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treegp_b200 import backend
for n in (2048, 4096, 10000):
    A = torch.randn((n, 64), dtype=torch.float64, device="cuda")
    ws = backend.alloc_matrix(n, n); ws[:, :n] = A @ A.T; ws[:, :n].diagonal().add_(float(n)); keep = ws.clone()
    for la in (1, 0):
        backend.set_option("potrf_lookahead", la)
        for rep in range(3):
            ws.copy_(keep); torch.cuda.synchronize()
            t0 = time.perf_counter(); backend.potrf(ws, n); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print("n=%d lookahead=%d enqueue %.2f ms, total %.2f ms" % (n, la, (t1 - t0) * 1e3, (t2 - t0) * 1e3), flush=True)
