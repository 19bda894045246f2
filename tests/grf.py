"""Synthetic Gaussian random fields for the API tests (same recipe as the reference's
tests/treegp_test_helper.py:47-104: legacy np.random.seed, uniform coordinates in [-10, 10],
multivariate_normal on the dense kernel matrix, optional white noise)."""
import numpy as np


def corr_matrix(size, e1, e2):
    e = np.hypot(e1, e2)
    q = (1 - e) / (1 + e)
    phi = 0.5 * np.arctan2(e2, e1)
    R = np.array([[np.cos(phi), np.sin(phi)], [-np.sin(phi), np.cos(phi)]])
    return R.T @ np.diag([size ** 2, (size * q) ** 2]) @ R


def make_grf(kernel, ndim, npoints, noise=None, seed=42):
    np.random.seed(seed)
    if ndim == 1:
        x = np.random.uniform(-10, 10, npoints).reshape((npoints, 1))
    else:
        x = np.array([np.random.uniform(-10, 10, npoints), np.random.uniform(-10, 10, npoints)]).T
    K = kernel(x)
    y = np.random.multivariate_normal(np.zeros(npoints), K)
    if noise is None:
        return x, y, None
    y += np.random.normal(scale=noise, size=npoints)
    return x, y, np.ones_like(y) * noise
