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


@pytest.mark.parametrize("ndim", [1, 2])
@pytest.mark.parametrize("fam", ["rbf", "vonkarman", "matern32"])
def test_predict_mean_truncated_support(gpu_ready, fam, ndim):
    """tgp_predict_mean_trunc (Hilbert-sorted, far blocks skipped) against the full sum and the oracle on a field
    much larger than the correlation length, where most blocks ARE skipped.  Bound: 1e-40 * sum|amp alpha|
    plus summation order (asserted at 1e-13 of sum|amp alpha|)."""
    from treegp_b200 import backend, eval_kernel
    from treegp_b200.kernels import lower_kernel

    rng = np.random.default_rng(17 + ndim)
    n, m, half = 4000, 3000, 400.0
    X = rng.uniform(-half, half, size=(n, ndim))
    Xs = rng.uniform(-half, half, size=(m, ndim))
    alpha = rng.normal(size=n) * 10.0 ** rng.uniform(-3, 3, size=n)
    if ndim == 2:
        Mi = np.array([[0.8, -0.1], [-0.1, 0.5]])
        s = {"rbf": "3.0 * AnisotropicRBF(invLam=array([[0.8, -0.1], [-0.1, 0.5]]))",
             "vonkarman": "3.0 * AnisotropicVonKarman(invLam=array([[0.8, -0.1], [-0.1, 0.5]]))",
             "matern32": "3.0 * Matern(length_scale=1.5, nu=1.5)"}[fam]
        okw = dict(length_scale=1.5) if fam == "matern32" else dict(invLam=Mi)
    else:
        s = {"rbf": "3.0 * RBF(length_scale=0.7)", "vonkarman": "3.0 * VonKarman(length_scale=0.7)",
             "matern32": "3.0 * Matern(length_scale=0.7, nu=1.5)"}[fam]
        okw = dict(length_scale=0.7)
    desc = lower_kernel(eval_kernel(s), ndim)
    a = backend.to_device(alpha)
    full = backend.predict_mean(Xs, X, desc, a, truncate=False).cpu().numpy()
    trunc = backend.predict_mean(Xs, X, desc, a, truncate=True).cpu().numpy()
    ref = go.kmat(fam, Xs, X, amp=3.0, **okw) @ alpha
    tol = 1e-13 * 3.0 * np.abs(alpha).sum()
    np.testing.assert_allclose(trunc, full, rtol=0, atol=tol)
    np.testing.assert_allclose(trunc, ref, rtol=0, atol=1e-11 * np.abs(alpha).sum())
    # any point order is valid input for the C entry point (the boxes just get large): unsorted call
    import ctypes, torch
    from treegp_b200 import _cabi
    lib = _cabi.load()
    Xd, Xsd = backend.as_points(X), backend.as_points(Xs)
    out = torch.empty(m, dtype=torch.float64, device=Xd.device)
    work = torch.empty(int(lib.tgp_predict_work_doubles(n)), dtype=torch.float64, device=Xd.device)
    _cabi.check(lib.tgp_predict_mean_trunc(ctypes.c_void_p(Xsd.data_ptr()), m, ctypes.c_void_p(Xd.data_ptr()), n,
                                           ctypes.byref(desc), ctypes.c_void_p(a.data_ptr()),
                                           ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(work.data_ptr()),
                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    np.testing.assert_allclose(out.cpu().numpy(), full, rtol=0, atol=tol)


@pytest.mark.parametrize("ndim", [1, 2])
@pytest.mark.parametrize("k", [1, 4, 7, 16, 20, 32, 45])
def test_knn_mean_matches_sklearn(gpu_ready, ndim, k):
    """tgp_knn_mean against what the reference calls (gp_interp.py:236-238): KNeighborsRegressor(k).fit(X0, y0)
    .predict(X).  Random queries have no distance ties, so the neighbour sets are identical and the means agree
    to rounding; also a two-column y0 (the reference then picks a column with indice_meanify) and a grid that
    spans several shared-memory tiles."""
    from sklearn.neighbors import KNeighborsRegressor
    from treegp_b200 import backend

    rng = np.random.default_rng(3 + k)
    for n0, m in ((2500, 20000), (20 if k <= 20 else 64, 300), (1025, 5)):
        X0 = rng.uniform(-1, 1, size=(n0, ndim))
        Xq = rng.uniform(-1.2, 1.2, size=(m, ndim))
        y0 = rng.normal(size=(n0, 2))
        ref = KNeighborsRegressor(n_neighbors=k).fit(X0, y0).predict(Xq)
        got = backend.knn_mean(X0, y0, Xq, k).cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-14 * k)
        got1 = backend.knn_mean(X0, y0[:, 1], Xq, k).cpu().numpy()
        np.testing.assert_allclose(got1, ref[:, 1], rtol=0, atol=1e-14 * k)
    with pytest.raises(ValueError):
        backend.knn_mean(X0[:3], y0[:3], Xq, 4)


def test_knn_mean_ties_take_the_lower_index(gpu_ready):
    """On a regular grid a query at a node has four neighbours at the same distance: the documented rule (lower
    grid index first) makes the result deterministic where sklearn's KD-tree leaves it open."""
    from treegp_b200 import backend

    g = np.arange(5.0)
    X0 = np.array([[a, b] for a in g for b in g])
    y0 = np.arange(25.0)
    got = backend.knn_mean(X0, y0, np.array([[2.0, 2.0]]), 4).cpu().numpy()
    # nearest: index 12 (d = 0), then 7, 11, 13, 17 at d = 1 -> 7, 11, 13 by index
    np.testing.assert_allclose(got, [(12 + 7 + 11 + 13) / 4.0])


@pytest.mark.parametrize("case", ["rbf2", "vk2", "rbf1"])
def test_windowed_variance_equals_plain_variance(gpu_ready, case):
    """backend.predict_var_windowed: chunks of neighbouring test points solve only the trailing sub-system of the
    ascending / descending factorisation they can be correlated with.  Same numbers as tgp_predict_var on the
    unsorted factor (only rounding differs: correlations below 1e-40 amp count as zero) and as the oracle."""
    from treegp_b200 import backend, eval_kernel
    from treegp_b200.kernels import lower_kernel

    rng = np.random.default_rng(17)
    n, m, field = 2500, 12000, 300.0
    if case == "rbf1":
        ndim, s, fam, kw = 1, "2.0 * RBF(0.8)", "rbf", dict(amp=2.0, invLam=np.array([[1.0 / 0.64]]))
    else:
        ndim = 2
        Mi = np.array([[0.9, -0.35], [-0.35, 0.6]])
        name, fam = ("AnisotropicRBF", "rbf") if case == "rbf2" else ("AnisotropicVonKarman", "vonkarman")
        s, kw = "2.0 * %s(invLam=array([[0.9, -0.35], [-0.35, 0.6]]))" % name, dict(amp=2.0, invLam=Mi)
    X = rng.uniform(0, field, size=(n, ndim))
    Xs = rng.uniform(0, field, size=(m, ndim))
    desc = lower_kernel(eval_kernel(s), ndim)
    yerr2 = rng.uniform(0.005, 0.02, size=n)
    e2 = backend.to_device(yerr2)
    ws = backend.kmat_sym(X, desc, diag_add=e2, lower_only=True)
    assert int(backend.potrf(ws, n).item()) == 0
    plain = backend.predict_var(Xs, X, desc, ws).cpu().numpy()
    stats = {}
    win = backend.predict_var_windowed(Xs, X, desc, e2, chunk=1024, stats=stats)
    assert win is not None and stats["used"] and stats["chunks"] == 12
    assert 0 < stats["chunks_descending"] < stats["chunks"]
    assert stats["flops_windowed"] + stats["flops_factors"] < 0.6 * stats["flops_full"]
    np.testing.assert_allclose(win.cpu().numpy(), plain, rtol=0, atol=1e-11)
    sub = rng.choice(m, 300, replace=False)
    K = go.kmat(fam, X, **kw) + np.diag(yerr2)
    ref = go.predictive_variance(K, go.kmat(fam, Xs[sub], X, **kw), 2.0)
    np.testing.assert_allclose(win.cpu().numpy()[sub], ref, rtol=0, atol=1e-9)


def test_windowed_variance_declines_when_it_does_not_pay(gpu_ready):
    """Support as wide as the field, or too few test points for two extra factorisations: None, and
    GPInterpolation.predict_var falls through to the plain solves with the same result."""
    import treegp_b200 as treegp
    from treegp_b200 import backend, eval_kernel
    from treegp_b200.kernels import lower_kernel

    rng = np.random.default_rng(3)
    n = 1500
    X = rng.uniform(0, 20.0, size=(n, 2))
    e2 = backend.to_device(np.full(n, 0.01))
    wide = lower_kernel(eval_kernel("1.0 * AnisotropicRBF(invLam=array([[0.5, 0.0], [0.0, 0.5]]))"), 2)
    st = {}
    assert backend.predict_var_windowed(rng.uniform(0, 20.0, size=(20000, 2)), X, wide, e2, chunk=1024, stats=st) is None
    assert st["used"] is False
    narrow = lower_kernel(eval_kernel("1.0 * AnisotropicRBF(invLam=array([[400.0, 0.0], [0.0, 400.0]]))"), 2)
    assert backend.predict_var_windowed(rng.uniform(0, 20.0, size=(100, 2)), X, narrow, e2, stats=st) is None
    # through the public class: both settings of the switch give the same variance
    kstr = "1.0 * AnisotropicRBF(invLam=array([[30.0, 0.0], [0.0, 30.0]]))"
    y = rng.normal(size=n)
    Xs = rng.uniform(0, 20.0, size=(9000, 2))
    out = []
    for flag in (True, False):
        gp = treegp.GPInterpolation(kernel=kstr, optimizer="none", normalize=True)
        gp.WINDOWED_VARIANCE, gp.VAR_CHUNK = flag, 1024
        gp.initialize(X, y, y_err=np.full(n, 0.1))
        out.append(gp.predict_var(Xs) + (gp._var_plan.get("used", False),))
    assert out[0][2] is True and out[1][2] is False
    np.testing.assert_allclose(out[0][0], out[1][0], rtol=0, atol=1e-12)
    np.testing.assert_allclose(out[0][1], out[1][1], rtol=0, atol=1e-11)
