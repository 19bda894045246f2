#!/usr/bin/env python
"""Generate tests/golden/reference_vectors.npz by running the UNMODIFIED reference
(/root/reference, PFLeget/treegp v1.4.1) on small seeded inputs.

Only runnable in the build container (the GPU box has no /root/reference); the produced .npz is
committed and is what the oracle (oracle/gp_oracle.py) and the CUDA path are pinned against.

The reference's missing third-party imports (treecorr, fitsio, iminuit -- none of which is touched by
pieces (1), (2), (4) of the hot path) are satisfied by empty stub modules, and
scipy.linalg.cholesky's `overwrite_a` is neutralised inside gp_interp (under
scipy >= 1.15 the reference otherwise re-factorises an overwritten buffer and returns garbage
covariances).  Piece (3) (TreeCorr) cannot be generated: parity unpinned.

Usage:  python tests/golden/make_golden.py
"""
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.npz")


def import_reference():
    stub = tempfile.mkdtemp(prefix="tgp_stubs_")
    for name, body in (("treecorr", ""), ("fitsio", ""), ("iminuit", "__version__ = '2.0.0'\n")):
        with open(os.path.join(stub, name + ".py"), "w") as fh:
            fh.write(body)
    sys.path.insert(0, stub)
    sys.path.insert(0, REF)
    import treegp  # noqa
    import scipy.linalg
    gpi = sys.modules["treegp.gp_interp"]

    def safe_cholesky(a, lower=False, overwrite_a=False, check_finite=True):
        return scipy.linalg.cholesky(a, lower=lower, overwrite_a=False, check_finite=check_finite)

    gpi.cholesky = safe_cholesky
    return treegp


def corr_matrix(size, e1, e2):
    # same parameterisation as tests/treegp_test_helper.py:24-44 (written out independently)
    e = np.hypot(e1, e2)
    q = (1 - e) / (1 + e)
    phi = 0.5 * np.arctan2(e2, e1)
    R = np.array([[np.cos(phi), np.sin(phi)], [-np.sin(phi), np.cos(phi)]])
    return R.T @ np.diag([size ** 2, (size * q) ** 2]) @ R


def main():
    treegp = import_reference()
    rng = np.random.default_rng(20261018)
    out = {}
    cases = []

    X2 = rng.uniform(-10, 10, size=(70, 2))
    X2[5] = X2[3]  # a duplicated point: off-diagonal zero distance
    Xs2 = rng.uniform(-12, 12, size=(33, 2))
    Xs2[7] = X2[11]
    X1 = rng.uniform(-10, 10, size=(45, 1))
    Xs1 = rng.uniform(-12, 12, size=(21, 1))
    out.update(X2=X2, Xs2=Xs2, X1=X1, Xs1=Xs1)

    inv_a = np.linalg.inv(corr_matrix(0.5, 0.2, 0.2))
    inv_b = np.linalg.inv(corr_matrix(30.0, -0.4, 0.4))
    inv_c = np.linalg.inv(corr_matrix(1.5, 0.2, -0.1))
    kernel_strings = {
        "arbf_a": "4.0 * AnisotropicRBF(invLam={0!r})".format(inv_a),
        "arbf_b": "1e-6 * AnisotropicRBF(invLam={0!r})".format(inv_b),
        "avk_c": "4.0 * AnisotropicVonKarman(invLam={0!r})".format(inv_c),
        "avk_b": "0.01 * AnisotropicVonKarman(invLam={0!r})".format(inv_b),
        "rbf2": "2.0**2 * RBF(0.45)",
        "vk2": "1.0**2 * VonKarman(length_scale=8.0)",
        "vk2s": "0.5 * VonKarman(length_scale=0.7)",
        "matern32": "1.5 * Matern(length_scale=0.8, nu=1.5)",
        "matern52": "1.5 * Matern(length_scale=2.0, nu=2.5)",
        "matern12": "Matern(length_scale=3.0, nu=0.5)",
    }
    for name, s in kernel_strings.items():
        k = treegp.eval_kernel(s)
        out["kstr_" + name] = np.array(s)
        out["K_" + name] = k(X2)
        out["Kx_" + name] = k(Xs2, Y=X2)
        out["theta_" + name] = np.array(k.theta)
        cases.append(name)
    for name in ("rbf2", "vk2", "vk2s", "matern32"):
        k = treegp.eval_kernel(kernel_strings[name])
        out["K1_" + name] = k(X1)
        out["K1x_" + name] = k(Xs1, Y=X1)
    out["cases"] = np.array(cases)

    # theta <-> invLam round trip (kernels.py:163-179)
    k = treegp.AnisotropicRBF(invLam=inv_a)
    out["rt_invLam"] = inv_a
    out["rt_theta"] = np.array(k.theta)
    th2 = np.array([0.3, -0.2, 0.45])
    k.theta = th2
    out["rt_theta2"] = th2
    out["rt_invLam2"] = np.array(k.invLam)

    # log-likelihood (log_likelihood.py:21-41) and predict (gp_interp.py:143-194)
    for name in ("arbf_a", "avk_c", "rbf2", "vk2s", "matern32"):
        kern = treegp.eval_kernel(kernel_strings[name])
        K = kern(X2)
        jitter = 1e-10 * np.eye(len(X2))
        K_pd = K + jitter
        K_pd[5, 3] = K_pd[3, 5] = K[3, 5] * (1 - 1e-9)  # duplicated point: keep the draw well defined
        y = rng.multivariate_normal(np.zeros(len(X2)), K_pd) + rng.normal(scale=0.05, size=len(X2)) + 0.7
        y_err = np.full(len(X2), 0.05) * rng.uniform(0.8, 1.2, size=len(X2))
        out["y_" + name] = y
        out["yerr_" + name] = y_err
        ll = treegp.log_likelihood(X2, y - np.mean(y), y_err)
        out["logL_" + name] = np.array(ll.log_likelihood(kern))
        gp = treegp.GPInterpolation(kernel=kernel_strings[name], optimizer="none", normalize=True,
                                    white_noise=0.0)
        gp.initialize(X2, y, y_err=y_err)
        gp.solve()
        mean, cov = gp.predict(Xs2, return_cov=True)
        out["pmean_" + name] = mean
        out["pcov_" + name] = cov
        out["alpha_" + name] = np.array(gp._alpha)
        gp2 = treegp.GPInterpolation(kernel=kernel_strings[name], optimizer="none", normalize=False,
                                     white_noise=0.03)
        gp2.initialize(X2, y, y_err=y_err)
        gp2.solve()
        out["pmean_wn_" + name] = gp2.predict(Xs2)
    # not-PD case -> -inf (log_likelihood.py:38-39)
    kern = treegp.eval_kernel("1.0 * RBF(50.0)")
    ll = treegp.log_likelihood(X2, np.ones(len(X2)), np.zeros(len(X2)))
    out["logL_notpd"] = np.array(ll.log_likelihood(kern))

    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
