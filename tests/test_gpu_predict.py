"""GPU parity: predict mean / variance (csrc/predict.cu) vs the reference's golden vectors and the oracle."""
import numpy as np
import pytest

from kspec import parse, vk_sym_quirk
from oracle import gp_oracle as go

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["arbf_a", "avk_c", "rbf2", "vk2s", "matern32"])
def test_predict_mean_and_var_match_reference_golden(gpu_ready, golden, name):
    """gp_interp.py:168-194; tolerance 1e-8 (north_star: 1e-6 on predictions)."""
    from treegp_b200 import backend, eval_kernel
    from treegp_b200.kernels import lower_kernel

    X, Xs = golden["X2"], golden["Xs2"]
    n = len(X)
    y, yerr = golden["y_" + name], golden["yerr_" + name]
    desc = lower_kernel(eval_kernel(str(golden["kstr_" + name])), 2)
    ws = backend.kmat_sym(X, desc, diag_add=backend.to_device(yerr ** 2), lower_only=True)
    assert int(backend.potrf(ws, n).item()) == 0
    alpha = backend.potrs_vec(ws, n, backend.to_device(y - np.mean(y)).clone())
    mean = backend.predict_mean(Xs, X, desc, alpha).cpu().numpy() + np.mean(y)
    np.testing.assert_allclose(mean, golden["pmean_" + name], rtol=1e-8, atol=1e-8)
    var = backend.predict_var(Xs, X, desc, ws).cpu().numpy()
    np.testing.assert_allclose(var, np.diag(golden["pcov_" + name]), rtol=1e-7, atol=1e-8)


@pytest.mark.parametrize("m,n", [(1, 1), (5, 3000), (257, 1025), (5000, 700)])
def test_predict_mean_vs_oracle_sizes(gpu_ready, m, n):
    from treegp_b200 import backend, eval_kernel
    from treegp_b200.kernels import lower_kernel

    rng = np.random.default_rng(m + n)
    X = rng.uniform(-10, 10, size=(n, 2))
    Xs = rng.uniform(-10, 10, size=(m, 2))
    alpha = rng.normal(size=n)
    M = np.array([[0.8, -0.1], [-0.1, 0.5]])
    for fam, s in (("rbf", "3.0 * AnisotropicRBF(invLam=array([[0.8, -0.1], [-0.1, 0.5]]))"),
                   ("vonkarman", "3.0 * AnisotropicVonKarman(invLam=array([[0.8, -0.1], [-0.1, 0.5]]))")):
        desc = lower_kernel(eval_kernel(s), 2)
        got = backend.predict_mean(Xs, X, desc, backend.to_device(alpha)).cpu().numpy()
        ref = go.kmat(fam, Xs, X, amp=3.0, invLam=M) @ alpha
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-11 * np.abs(alpha).sum())


def test_predict_var_chunked(gpu_ready):
    from treegp_b200 import backend, eval_kernel
    from treegp_b200.kernels import lower_kernel

    rng = np.random.default_rng(9)
    n, m = 900, 700
    X = rng.uniform(-10, 10, size=(n, 2))
    Xs = rng.uniform(-10, 10, size=(m, 2))
    s = "2.0 * AnisotropicRBF(invLam=array([[0.8, -0.1], [-0.1, 0.5]]))"
    desc = lower_kernel(eval_kernel(s), 2)
    yerr2 = np.full(n, 0.01)
    ws = backend.kmat_sym(X, desc, diag_add=backend.to_device(yerr2), lower_only=True)
    assert int(backend.potrf(ws, n).item()) == 0
    var = backend.predict_var(Xs, X, desc, ws, chunk=256).cpu().numpy()
    Mi = np.array([[0.8, -0.1], [-0.1, 0.5]])
    K = go.kmat("rbf", X, amp=2.0, invLam=Mi) + np.diag(yerr2)
    ref = go.predictive_variance(K, go.kmat("rbf", Xs, X, amp=2.0, invLam=Mi), 2.0)
    np.testing.assert_allclose(var, ref, rtol=0, atol=1e-9)
