"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy/scipy restatement of pieces (1), (2) and (4) of the hot path, following the reference line by
line.  Pinned against the reference itself: tests/golden/make_golden.py imports /root/reference
(kernels.py, log_likelihood.py, gp_interp.py) in the build container and stores its outputs in
tests/golden/*.npz; tests/test_oracle_golden.py checks this file against those vectors.

Reference citations are relative to /root/reference/treegp/.
"""
import numpy as np
from scipy import special
from scipy.linalg import cholesky, cho_solve, solve_triangular
from scipy.spatial.distance import cdist

# lim0 = Gamma(5/6) / (2 pi^(5/6))                                   kernels.py:260, :365
_VK_LIM0 = special.gamma(5.0 / 6.0) / (2.0 * np.pi ** (5.0 / 6.0))


def metric_from_theta(theta, ndim=2):
    """theta -> invLam = L L^T, L lower with exp(theta[:n]) on the diagonal.  kernels.py:173-179."""
    theta = np.asarray(theta, dtype=float)
    L = np.zeros((ndim, ndim))
    L[np.diag_indices(ndim)] = np.exp(theta[:ndim])
    L[np.tril_indices(ndim, -1)] = theta[ndim:]
    return L @ L.T


def theta_from_metric(invLam):
    """invLam -> theta (log of the Cholesky diagonal, then the strict lower part).  kernels.py:163-167."""
    L = np.linalg.cholesky(np.asarray(invLam, dtype=float))
    n = L.shape[0]
    return np.hstack([np.log(L[np.diag_indices(n)]), L[np.tril_indices(n, -1)]])


def _mahalanobis(X, Y, invLam):
    # scipy's 'mahalanobis' metric, as used at kernels.py:118,125,359,372
    return cdist(np.atleast_2d(X), np.atleast_2d(Y), metric="mahalanobis", VI=np.asarray(invLam, dtype=float))


def _vk_profile(d):
    """d^(5/6) K_{5/6}(2 pi d) / lim0 with the value 1 at d == 0.  kernels.py:254-262, :361-368."""
    out = np.zeros_like(d)
    nz = d != 0.0
    out[nz] = d[nz] ** (5.0 / 6.0) * special.kv(5.0 / 6.0, 2.0 * np.pi * d[nz])
    out[~nz] = _VK_LIM0
    return out / _VK_LIM0


def kmat(family, X, Y=None, amp=1.0, invLam=None, length_scale=None):
    """amp * f(dist) for the kernel families the reference's tests exercise.

    family: 'rbf' (sklearn RBF / AnisotropicRBF, kernels.py:114-126), 'vonkarman'
    (kernels.py:251-277, :358-381), 'matern12' | 'matern32' | 'matern52' (sklearn Matern).
    Exactly one of invLam (Mahalanobis metric) / length_scale (isotropic) is given.
    """
    X = np.atleast_2d(np.asarray(X, dtype=float))
    Y = X if Y is None else np.atleast_2d(np.asarray(Y, dtype=float))
    if invLam is not None:
        d = _mahalanobis(X, Y, invLam)
    else:
        d = cdist(X / length_scale, Y / length_scale, metric="euclidean")
    if family == "rbf":
        K = np.exp(-0.5 * d ** 2)
    elif family == "vonkarman":
        K = _vk_profile(d)
    elif family == "matern12":
        K = np.exp(-d)
    elif family == "matern32":
        s = np.sqrt(3.0) * d
        K = (1.0 + s) * np.exp(-s)
    elif family == "matern52":
        s = np.sqrt(5.0) * d
        K = (1.0 + s + s ** 2 / 3.0) * np.exp(-s)
    else:
        raise ValueError(family)
    return amp * K


def log_likelihood(K_noisy, y):
    """-1/2 y^T K^-1 y - N/2 ln 2pi - 1/2 ln|K|; -inf when K is not PD.  log_likelihood.py:29-39."""
    y = np.asarray(y, dtype=float)
    try:
        U = cholesky(K_noisy, lower=False)
    except np.linalg.LinAlgError:
        return -np.inf, None
    alpha = cho_solve((U, False), y)
    chi2 = float(np.dot(y, alpha))
    log_det = float(np.sum(2.0 * np.log(np.diag(U))))
    return -0.5 * chi2 - 0.5 * len(y) * np.log(2.0 * np.pi) - 0.5 * log_det, alpha


def gp_predict(K_noisy, K_star, K_starstar, y):
    """alpha, mean and covariance of gp_interp.py:177-191 from ONE factorisation.

    The reference re-factorises an already overwritten K under scipy >= 1.15; the
    intended algebra cov = K** - K* (K + s^2 I)^-1 K*^T is what is restated here.
    """
    Lc = cholesky(K_noisy, lower=True)
    alpha = cho_solve((Lc, True), np.asarray(y, dtype=float))
    mean = K_star @ alpha
    v = cho_solve((Lc, True), K_star.T)
    cov = K_starstar - K_star @ v
    return alpha, mean, cov


def predictive_variance(K_noisy, K_star, amp):
    """diag of the covariance above: amp - ||L^-1 k*||^2 per test point."""
    Lc = cholesky(K_noisy, lower=True)
    V = solve_triangular(Lc, K_star.T, lower=True)
    return amp - np.sum(V * V, axis=0)


def cholesky_envelope(K, row_end, block=512):
    """Restatement (numpy, small cases) of the ENVELOPE form of the factorisation the reference takes at
    gp_interp.py:181 / log_likelihood.py:30 -- not an algorithm of the reference, which always factorises the dense
    matrix, but the statement the device path (tgp_potrf_env) makes: a right-looking blocked Cholesky whose panel
    solve and trailing update of block column b only touch the rows [end of block, row_end[b]).  Everything outside
    is neither read nor written (it keeps the entries of K).  Used by the CPU tests to show that this equals the dense
    factor whenever K is (numerically) zero outside the envelope."""
    A = np.array(K, dtype=float, copy=True)
    n = A.shape[0]
    prev = 0
    for b, k in enumerate(range(0, n, block)):
        c1 = min(n, k + block)
        re = min(n, max(int(row_end[b]), prev, c1))
        prev = re
        A[k:c1, k:c1] = cholesky(np.tril(A[k:c1, k:c1]) + np.tril(A[k:c1, k:c1], -1).T, lower=True)
        if re > c1:
            A[c1:re, k:c1] = solve_triangular(A[k:c1, k:c1], A[c1:re, k:c1].T, lower=True).T
            upd = A[c1:re, k:c1] @ A[c1:re, k:c1].T
            A[c1:re, c1:re] -= np.tril(upd)
    return A
