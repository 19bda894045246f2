import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from oracle import pairbin_oracle as po
from test_gpu_pairbin import _gpu_pairbin
from treegp_b200 import backend
for m, mx, nb in ((300, 64.0, 16), (150, 64.0, 16), (300, 16.0, 4)):
    g = np.arange(0, m, dtype=np.float64) * 0.5
    X, Y = np.meshgrid(g, g); x, y = X.ravel(), Y.ravel(); n = len(x)
    k = np.random.default_rng(11).normal(size=n)
    ref = po.pairbin(x, y, k, None, 0.0, mx, nb, "TwoD")
    for mode in (1, 0):
        backend.set_option("pairbin_block_sums", mode)
        backend.pairbin_stats(reset=True)
        r = _gpu_pairbin(x, y, k, None, 0.0, mx, nb, "TwoD", hilbert=True)
        st = backend.pairbin_stats(reset=True)
        bad = (r[0][0] != ref["npairs"]).sum()
        print(m, mx, nb, "mode", mode, "bins wrong:", bad, "total diff", int(r[0][0].sum() - ref["npairs"].sum()), st)
backend.set_option("pairbin_block_sums", 1)
