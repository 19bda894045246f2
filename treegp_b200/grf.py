"""Synthetic Gaussian random fields on the device.

The reference's tests draw fields with ``np.random.multivariate_normal`` on the dense kernel matrix
(/root/reference/tests/treegp_test_helper.py:47-104: an O(N^3) SVD on the host, unusable beyond ~10^4 points).
Here the same distribution is sampled as y = L z with K = L L^T from the library's own K build and DMMA
Cholesky, so benchmark-sized fields (N = 40,000 and beyond) take a second.
"""
import numpy as np
import torch

from . import backend
from .kernels import lower_kernel


def sample_grf(kernel, X, noise=None, seed=42, jitter=1e-8, block=4096):
    """Draw y ~ N(0, kernel(X, X)) (+ white noise of standard deviation `noise`).

    kernel: kernel object (see eval_kernel); X: (n, 1|2) positions.  Returns (y, y_err) as numpy arrays, with
    y_err = None when noise is None -- the return convention of the reference's make_1d_grf / make_2d_grf.
    """
    rng = np.random.default_rng(seed)
    Xd = backend.as_points(X)
    n = Xd.shape[0]
    desc = lower_kernel(kernel, Xd.shape[1])
    ws = backend.kmat_sym(Xd, desc, diag_add=backend.to_device(np.full(n, jitter * desc.amp)), lower_only=True)
    info = backend.potrf(ws, n)
    if int(info.item()) != 0:
        raise np.linalg.LinAlgError("kernel matrix not positive definite (leading minor %d)" % int(info.item()))
    z = torch.as_tensor(rng.normal(size=n), device=ws.device)
    y = torch.empty(n, dtype=torch.float64, device=ws.device)
    for r0 in range(0, n, block):  # y = tril(L) z by row blocks (no N x N temporary)
        r1 = min(n, r0 + block)
        y[r0:r1] = torch.tril(ws[r0:r1, :n], diagonal=r0) @ z
    y = y.cpu().numpy()
    if noise is None:
        return y, None
    y = y + rng.normal(scale=noise, size=n)
    return y, np.full(n, float(noise))
