"""CPU: the oracle restatement (oracle/gp_oracle.py) against vectors produced by the reference itself."""
import numpy as np
import pytest

from oracle import gp_oracle as go


from kspec import parse, vk_sym_quirk

NAMES = ["arbf_a", "arbf_b", "avk_c", "avk_b", "rbf2", "vk2", "vk2s", "matern32", "matern52", "matern12"]


def _args(golden, name):
    return parse(golden["kstr_" + name]).oracle_args()


def _vk_sym_quirk(K, X, args):
    return vk_sym_quirk(K, X, args["family"])


@pytest.mark.parametrize("name", NAMES)
def test_kernels_match_reference(golden, name):
    args = _args(golden, name)
    X, Xs = golden["X2"], golden["Xs2"]
    K = _vk_sym_quirk(go.kmat(X=X, **args), X, args)
    np.testing.assert_allclose(K, golden["K_" + name], rtol=0, atol=1e-13 * args["amp"])
    np.testing.assert_allclose(go.kmat(X=Xs, Y=X, **args), golden["Kx_" + name], rtol=0, atol=1e-13 * args["amp"])


@pytest.mark.parametrize("name", ["rbf2", "vk2", "vk2s", "matern32"])
def test_kernels_1d_match_reference(golden, name):
    args = _args(golden, name)
    X, Xs = golden["X1"], golden["Xs1"]
    np.testing.assert_allclose(go.kmat(X=X, **args), golden["K1_" + name], rtol=0, atol=1e-13 * args["amp"])
    np.testing.assert_allclose(go.kmat(X=Xs, Y=X, **args), golden["K1x_" + name], rtol=0, atol=1e-13 * args["amp"])


def test_theta_round_trip(golden):
    np.testing.assert_allclose(go.theta_from_metric(golden["rt_invLam"]), golden["rt_theta"], atol=1e-14)
    np.testing.assert_allclose(go.metric_from_theta(golden["rt_theta2"]), golden["rt_invLam2"], atol=1e-14)


@pytest.mark.parametrize("name", ["arbf_a", "avk_c", "rbf2", "vk2s", "matern32"])
def test_loglike_and_predict_match_reference(golden, name):
    args = _args(golden, name)
    X, Xs = golden["X2"], golden["Xs2"]
    y, yerr = golden["y_" + name], golden["yerr_" + name]
    K = _vk_sym_quirk(go.kmat(X=X, **args), X, args) + np.diag(yerr ** 2)
    resid = y - np.mean(y)
    logl, _ = go.log_likelihood(K, resid)
    np.testing.assert_allclose(logl, float(golden["logL_" + name]), rtol=1e-11)
    Ks = go.kmat(X=Xs, Y=X, **args)
    Kss = _vk_sym_quirk(go.kmat(X=Xs, **args), Xs, args)
    alpha, mean, cov = go.gp_predict(K, Ks, Kss, resid)
    np.testing.assert_allclose(alpha, golden["alpha_" + name], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(mean + np.mean(y), golden["pmean_" + name], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(cov, golden["pcov_" + name], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(go.predictive_variance(K, Ks, args["amp"]), np.diag(golden["pcov_" + name]),
                               rtol=1e-7, atol=1e-8)


def test_not_positive_definite_gives_minus_inf(golden):
    X = golden["X2"]
    K = go.kmat("rbf", X, amp=1.0, length_scale=50.0)
    logl, _ = go.log_likelihood(K, np.ones(len(X)))
    assert logl == -np.inf and float(golden["logL_notpd"]) == -np.inf


def test_eb_oracle_matches_reference_vcorr():
    """oracle/eb_oracle.py against the reference's own `vcorr` (utils.py:5-74) on a random field with a coincident
    pair and on a lattice (tests/golden/make_golden_r2.py): EQUAL per-bin counts, correlation functions 1e-12."""
    import os
    from oracle import eb_oracle

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors_r2.npz"))
    for pre, (rmin, rmax, dlogr) in (("eb", tuple(float(v) for v in g["eb_par"])), ("ebl", (0.01, 1.0, 0.05))):
        bins = int(np.ceil(np.log(rmax / rmin) / dlogr))
        c, slr, sp, sz2, sm = eb_oracle.pair_sums(g[pre + "_x"], g[pre + "_y"], g[pre + "_dx"], g[pre + "_dy"],
                                                  np.log(rmin), dlogr, bins)
        np.testing.assert_array_equal(c, g[pre + "_counts"])
        ok = c > 0
        for got, key in ((slr, "_logr"), (sp, "_xip"), (sm.real, "_xim"), (sm.imag, "_xix"), (sz2, "_xiz2")):
            ref = g[pre + key]
            np.testing.assert_allclose((got / np.where(ok, c, 1))[ok], ref[ok], rtol=0, atol=1e-12 * np.abs(ref[ok]).max())
