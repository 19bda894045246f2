#!/usr/bin/env python
"""Where does GPInterpolation.solve(optimizer='anisotropic') spend its time at the configs[2] shape (N=40k)?
cProfile of one warm solve (the same call bench.py's gp_fit_predict times)."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import treegp_b200 as treegp
from treegp_b200 import backend
from treegp_b200.kernels import lower_kernel
from treegp_b200.two_pcf import get_correlation_length_matrix
n = int(os.environ.get("PN", 40000))
L = 160.0 * np.sqrt(n / 40000.0)
inv = np.linalg.inv(get_correlation_length_matrix(1.5, 0.2, 0.2))
kstr = "2.0**2 * AnisotropicVonKarman(invLam=array([[%.17g, %.17g], [%.17g, %.17g]]))" % (inv[0, 0], inv[0, 1], inv[1, 0], inv[1, 1])
rng = np.random.default_rng(42)
X = rng.uniform(-L / 2, L / 2, size=(n, 2))
desc = lower_kernel(treegp.eval_kernel(kstr), 2)
ws = backend.kmat_sym(X, desc, diag_add=backend.to_device(np.full(n, 1e-8)), lower_only=True)
assert int(backend.potrf(ws, n).item()) == 0
z = torch.as_tensor(rng.normal(size=n), device="cuda")
y = torch.zeros(n, dtype=torch.float64, device="cuda")
for r0 in range(0, n, 4096):
    r1 = min(n, r0 + 4096)
    y[r0:r1] = torch.tril(ws[r0:r1, :n], diagonal=r0) @ z
y = y.cpu().numpy() + rng.normal(scale=0.01, size=n)
y_err = np.full(n, 0.01)
del ws
def one():
    gp = treegp.GPInterpolation(kernel=kstr, optimizer="anisotropic", normalize=True, nbins=21, min_sep=0.0,
                                max_sep=1.0, p0=[1.0, 0.0, 0.0])
    gp.initialize(X, y, y_err=y_err)
    gp.solve()
    torch.cuda.synchronize()
    return gp
one()
t0 = time.perf_counter(); one(); print("solve wall: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
pr = cProfile.Profile(); pr.enable(); one(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
