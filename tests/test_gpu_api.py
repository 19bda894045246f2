"""GPU: the reference-facing API (GPInterpolation / two_pcf / log_likelihood) end to end.

These follow the reference's own test strategy (tests/test_gp_interp.py,
tests/test_hyp_search.py) with the same sizes, kernels and tolerances, written against treegp_b200."""
import numpy as np
import pytest

from grf import corr_matrix, make_grf

pytestmark = pytest.mark.gpu


def _interp_checks(treegp, kernel, x, y, y_err, new_x, amp, noise):
    gp = treegp.GPInterpolation(kernel=kernel, optimizer="none", normalize=False, white_noise=0.0)
    gp.initialize(x, y, y_err=y_err)
    gp.solve()
    y_predict, y_cov = gp.predict(x, return_cov=True)
    y_std = np.sqrt(np.abs(np.diag(y_cov)))
    if noise is None:
        # noiseless: the GP interpolates the data exactly (reference tolerance 3e-5)
        np.testing.assert_allclose(y, y_predict, atol=3e-5)
        np.testing.assert_allclose(np.zeros_like(y_std), y_std, atol=3e-5)
    else:
        pull = (y - y_predict) / np.sqrt(y_err ** 2 + np.diag(y_cov).clip(0))
        assert abs(np.mean(pull)) < 3.0 * np.std(pull) / np.sqrt(len(y))
        assert np.std(pull) <= 1.0
    # far from the data: prior mean 0 and prior sigma
    y_predict, y_cov = gp.predict(new_x, return_cov=True)
    np.testing.assert_allclose(np.zeros(len(new_x)), y_predict, atol=1e-5)
    np.testing.assert_allclose(amp * np.ones(len(new_x)), np.sqrt(np.diag(y_cov)), atol=1e-5)
    # extension: diagonal-only variance agrees with the full covariance
    yv, var = gp.predict_var(new_x)
    np.testing.assert_allclose(var, np.diag(y_cov), atol=1e-9)
    np.testing.assert_allclose(yv, y_predict, atol=1e-12)


@pytest.mark.parametrize("ker,length", [("RBF", 0.5), ("RBF", 0.8), ("VonKarman", 8.0), ("VonKarman", 10.0)])
@pytest.mark.parametrize("noise", [None, 0.1])
def test_gp_interp_1d(gpu_ready, ker, length, noise):
    import treegp_b200 as treegp

    sigma, npoints = 1.5, 40
    kernel = "%f**2 * %s(%f)" % (sigma, ker, length)
    x, y, y_err = make_grf(treegp.eval_kernel(kernel), 1, npoints, noise=noise)
    new_x = np.linspace(np.max(x) + 6.0 * length, np.max(x) + 7.0 * length, npoints).reshape((npoints, 1))
    _interp_checks(treegp, kernel, x, y, y_err, new_x, sigma, noise)


@pytest.mark.parametrize("ker,size", [("AnisotropicRBF", 0.5), ("AnisotropicVonKarman", 5.0)])
@pytest.mark.parametrize("noise", [None, 0.1])
def test_gp_interp_2d(gpu_ready, ker, size, noise):
    import treegp_b200 as treegp

    sigma, npoints = 2.0, 200
    inv = np.linalg.inv(corr_matrix(size, 0.2, 0.2))
    kernel = "%f**2 * %s(invLam=%s)" % (sigma, ker, np.array2string(inv, separator=",", floatmode="unique").replace("\n", ""))
    kernel = kernel.replace("invLam=", "invLam=array(").replace("]])", "]]))")
    x, y, y_err = make_grf(treegp.eval_kernel(kernel), 2, npoints, noise=noise)
    far = np.max(x) + 6.0 * size
    new_x = np.array([np.linspace(far, far + size, npoints), np.linspace(far, far + size, npoints)]).T
    _interp_checks(treegp, kernel, x, y, y_err, new_x, sigma, noise)


@pytest.mark.parametrize("i", range(4))
def test_hyperparameter_search_1d_loglikelihood(gpu_ready, i):
    """tests/test_hyp_search.py:12-89, optimizer='log-likelihood', N=100."""
    import treegp_b200 as treegp

    sigma = [1.0, 2.0, 1.0, 2.0][i]
    length = [0.5, 0.8, 8.0, 10.0][i]
    ker = ["RBF", "RBF", "VonKarman", "VonKarman"][i]
    kernel = "%f**2 * %s(%f)" % (sigma, ker, length)
    truth = treegp.eval_kernel(kernel)
    x, y, y_err = make_grf(truth, 1, 100, noise=0.01)
    gp = treegp.GPInterpolation(kernel=kernel, optimizer="log-likelihood", normalize=True)
    gp.initialize(x, y, y_err=y_err)
    gp.solve()
    np.testing.assert_allclose(truth.theta, gp.kernel.theta, atol=7e-1)
    np.testing.assert_allclose(gp.return_log_likelihood(), gp._optimizer._logL, atol=1e-10)
    # the optimum is a maximum of the reference's objective as well: compare with the oracle
    from oracle import gp_oracle as go
    fam = "rbf" if ker == "RBF" else "vonkarman"
    amp, ls = np.exp(gp.kernel.theta[0]), np.exp(gp.kernel.theta[1])
    resid = y - np.mean(y)
    K = go.kmat(fam, x, amp=amp, length_scale=ls) + np.diag(y_err ** 2)
    ref, _ = go.log_likelihood(K, resid)
    np.testing.assert_allclose(gp._optimizer._logL, ref, rtol=1e-8)
    y_predict, y_cov = gp.predict(x, return_cov=True)
    pull = y - y_predict
    assert abs(np.mean(pull)) < 3.0 * np.std(pull) / np.sqrt(100)
    assert np.std(pull) <= 1.0


@pytest.mark.parametrize("i", range(4))
def test_hyperparameter_search_1d_two_pcf(gpu_ready, i):
    """tests/test_hyp_search.py:12-89, optimizer='two-pcf', N=2000, nbins=15, min_sep=0.1."""
    import treegp_b200 as treegp

    sigma = [1.0, 2.0, 1.0, 2.0][i]
    length = [0.5, 0.8, 8.0, 10.0][i]
    ker = ["RBF", "RBF", "VonKarman", "VonKarman"][i]
    max_sep = [1.75, 1.75, 1.25, 1.25][i]
    kernel = "%f**2 * %s(%f)" % (sigma, ker, length)
    truth = treegp.eval_kernel(kernel)
    x, y, y_err = make_grf(truth, 1, 2000, noise=0.01)
    gp = treegp.GPInterpolation(kernel=kernel, optimizer="two-pcf", normalize=True, nbins=15, min_sep=0.1,
                                max_sep=max_sep)
    gp.initialize(x, y, y_err=y_err)
    gp.solve()
    np.testing.assert_allclose(truth.theta, gp.kernel.theta, atol=7e-1)
    xi, xi_weight, distance, coord, mask = gp.return_2pcf()
    np.testing.assert_allclose(xi, gp._optimizer._2pcf, atol=1e-10)
    new_x = np.linspace(np.max(x) + 6.0 * length, np.max(x) + 7.0 * length, 200).reshape((200, 1))
    y_predict, y_cov = gp.predict(new_x, return_cov=True)
    np.testing.assert_allclose(np.mean(y) * np.ones(200), y_predict, atol=1e-5)
    np.testing.assert_allclose(np.sqrt(np.exp(gp.kernel.theta[0])) * np.ones(200), np.sqrt(np.diag(y_cov)), atol=1e-5)


def _aniso_kernel_string(sigma, ker, size, g1, g2):
    inv = np.linalg.inv(corr_matrix(size, g1, g2))
    return "%f**2 * %s(invLam=array([[%.17g, %.17g], [%.17g, %.17g]]))" % (
        sigma, ker, inv[0, 0], inv[0, 1], inv[1, 0], inv[1, 1])


def test_hyperparameter_search_2d_loglikelihood(gpu_ready):
    """tests/test_hyp_search.py:93-184, optimizer='log-likelihood', N=600."""
    import treegp_b200 as treegp

    kernel = _aniso_kernel_string(2.0, "AnisotropicRBF", 0.5, 0.2, 0.2)
    truth = treegp.eval_kernel(kernel)
    x, y, y_err = make_grf(truth, 2, 600, noise=0.01)
    gp = treegp.GPInterpolation(kernel=kernel, optimizer="log-likelihood", normalize=True)
    gp.initialize(x, y, y_err=y_err)
    gp.solve()
    np.testing.assert_allclose(truth.theta, gp.kernel.theta, atol=5e-1)
    np.testing.assert_allclose(gp.return_log_likelihood(), gp._optimizer._logL, atol=1e-10)
    y_predict, y_cov = gp.predict(x, return_cov=True)
    pull = y - y_predict
    assert abs(np.mean(pull)) < 3.0 * np.std(pull) / np.sqrt(600)
    assert np.std(pull) <= 1.0


@pytest.mark.parametrize("ker,size", [("AnisotropicRBF", 0.5), ("AnisotropicVonKarman", 1.5)])
def test_hyperparameter_search_2d_anisotropic(gpu_ready, ker, size):
    """tests/test_hyp_search.py:93-184, optimizer='anisotropic', N=2000, nbins=21, min_sep=0, max_sep=1,
    p0=[0.3, 0, 0]: 1 + 444 full pair counts (bootstrap) then the robust fit."""
    import treegp_b200 as treegp

    kernel = _aniso_kernel_string(2.0, ker, size, 0.2, 0.2)
    truth = treegp.eval_kernel(kernel)
    x, y, y_err = make_grf(truth, 2, 2000, noise=0.01)
    gp = treegp.GPInterpolation(kernel=kernel, optimizer="anisotropic", normalize=True, nbins=21, min_sep=0.0,
                                max_sep=1.0, p0=[0.3, 0.0, 0.0])
    gp.initialize(x, y, y_err=y_err)
    gp.solve()
    np.testing.assert_allclose(truth.theta, gp.kernel.theta, atol=5e-1)
    assert gp._optimizer._2pcf.shape == (441,)
    assert gp._optimizer._2pcf_mask.sum() == 221
    assert gp._optimizer._2pcf_weight.shape == (221, 221)
    new_x = np.array([np.linspace(np.max(x) + 6.0 * size, np.max(x) + 7.0 * size, 100)] * 2).T
    y_predict, y_cov = gp.predict(new_x, return_cov=True)
    np.testing.assert_allclose(np.mean(y) * np.ones(100), y_predict, atol=1e-5)
    np.testing.assert_allclose(np.sqrt(np.exp(gp.kernel.theta[0])) * np.ones(100), np.sqrt(np.diag(y_cov)), atol=1e-5)


def test_two_pcf_matches_oracle_and_bootstrap_is_batched(gpu_ready):
    """comp_2pcf and the batched bootstrap against the oracle's sequential restatement
    (two_pcf.py:269-281, :283-340, :342-362)."""
    import treegp_b200 as treegp
    from oracle import pairbin_oracle as po

    rng = np.random.default_rng(1)
    n = 1500
    X = rng.uniform(-10, 10, size=(n, 2))
    y = rng.normal(size=n)
    y_err = np.full(n, 0.1)
    for aniso, mn, mx, nb in ((True, 0.0, 2.0, 9), (False, 0.2, 4.0, 12)):
        t = treegp.two_pcf(X, y, y_err, mn, mx, nbins=nb, anisotropic=aniso)
        xi, dist, coord, mask = t.comp_2pcf(X, y, y_err)
        rxi, rdist, rcoord, rmask = po.comp_2pcf(X, y, y_err, mn, mx, nb, aniso)
        np.testing.assert_allclose(xi, rxi, rtol=0, atol=1e-12)
        np.testing.assert_allclose(dist, rdist, rtol=1e-12)
        np.testing.assert_array_equal(mask, rmask)
        # bootstrap: sequential oracle with the same index stream
        B = 7
        cov = t.comp_xi_covariance(n_bootstrap=B, mask=mask, seed=610639139)
        r = np.random.default_rng(610639139)
        xis = []
        for _ in range(B):
            ind = r.integers(0, n - 1, size=n)
            bxi, _, _, _ = po.comp_2pcf(X[ind], y[ind], y_err[ind], mn, mx, nb, aniso)
            xis.append(bxi[mask])
        xis = np.array(xis)
        d = xis - xis.mean(axis=0)
        np.testing.assert_allclose(cov, d.T @ d / (B - 1.0), rtol=0, atol=1e-12)


def test_gpinterp_around_a_meanify_mean_function(gpu_ready, tmp_path):
    """tests/test_meanify.py:68-129 shape: a GP on top of a spatial average read from a meanify FITS table
    (KNN(4) lookup, gp_interp.py:229-243); the mean file is produced by treegp_b200.meanify itself."""
    import treegp_b200 as treegp

    rng = np.random.default_rng(11)

    def mean_fn(c):
        return 0.1 + 0.5 * (c[:, 0] ** 2 + c[:, 1] ** 2) / 100.0

    m = treegp.meanify(bin_spacing=1.0, statistics="mean")
    for _ in range(40):
        c = rng.uniform(-10, 10, (2000, 2))
        m.add_field(c, mean_fn(c) + rng.normal(0, 0.002, 2000))
    m.meanify()
    path = str(tmp_path / "mean_gp.fits")
    m.save_results(path)

    kernel = _aniso_kernel_string(0.1, "AnisotropicRBF", 0.8, 0.1, 0.1)
    truth = treegp.eval_kernel(kernel)
    x, y, y_err = make_grf(truth, 2, 1500, noise=0.005)
    y = y + mean_fn(x)
    gp = treegp.GPInterpolation(kernel=kernel, optimizer="anisotropic", normalize=True, average_fits=path, nbins=21,
                                min_sep=0.0, max_sep=1.6, p0=[0.5, 0.0, 0.0])
    gp.initialize(x, y, y_err=y_err)
    assert np.std(y - gp._spatial_average) < 0.7 * np.std(y)  # the mean function was removed
    gp.solve()
    np.testing.assert_allclose(truth.theta, gp.kernel.theta, atol=5e-1)
    y_predict, y_cov = gp.predict(x, return_cov=True)
    pull = y - y_predict
    assert abs(np.mean(pull)) < 3.0 * np.std(pull) / np.sqrt(len(y))
    assert np.std(pull) <= 1.0
    # far outside the field the GP term vanishes: prediction = normalising mean + looked-up mean function
    far = np.array([[40.0, 40.0], [-45.0, 38.0]])
    yp = gp.predict(far)
    np.testing.assert_allclose(yp, gp._mean + gp._build_average_meanify(far), atol=1e-6)
    with pytest.raises(NotImplementedError):
        treegp.GPInterpolation(kernel=kernel, optimizer="none").plot_fitted_kernel()


def test_sample_grf_has_the_kernel_covariance(gpu_ready):
    """Extension: y = L z from the library's own K build + Cholesky has covariance K."""
    import treegp_b200 as treegp

    kernel = treegp.eval_kernel(_aniso_kernel_string(1.5, "AnisotropicRBF", 2.0, 0.2, -0.1))
    rng = np.random.default_rng(0)
    X = rng.uniform(-5, 5, size=(60, 2))
    ys = np.array([treegp.sample_grf(kernel, X, seed=s)[0] for s in range(3000)])
    K = kernel(X)
    emp = ys.T @ ys / len(ys)
    assert np.abs(emp - K).max() < 0.2 * K.max()      # 3000 draws: ~4 sigma of the sampling noise
    y, y_err = treegp.sample_grf(kernel, X, noise=0.1, seed=1)
    assert y.shape == (60,) and np.all(y_err == 0.1)


def test_eb_pair_sums_and_api_on_the_device(gpu_ready):
    """csrc/vcorr.cu through treegp.comp_eb / comp_eb_treecorr / utils.vcorr (utils.py:5-155) against the oracle
    restatement of the reference's all-pairs loop (np.histogram on np.log(np.absolute(d)), utils.py:50-55; pinned
    to the reference's own output in tests/test_oracle_golden.py).  Pair counts EQUAL (the device settles pairs
    within 1e-14 of a bin threshold with the reference's expression); sums to 1e-10 of the bin's absolute sum."""
    import treegp_b200 as treegp
    from oracle import eb_oracle
    from treegp_b200 import backend
    from treegp_b200.utils import vcorr

    rng = np.random.default_rng(4)
    n = 3000
    x, y = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    dx, dy = rng.normal(size=n), rng.normal(size=n)
    x[10], y[10] = x[11], y[11]                      # a coincident pair (r = 0) is skipped
    rmin, rmax, dlogr = 0.002, 1.0, 0.05
    bins = int(np.ceil(np.log(rmax / rmin) / dlogr))
    ref = eb_oracle.pair_sums(x, y, dx, dy, np.log(rmin), dlogr, bins)
    got = backend.vcorr_sums(x, y, dx, dy, np.log(rmin), dlogr, bins)
    np.testing.assert_array_equal(got[0], ref[0])
    assert got[0].sum() > 0.9 * n * (n - 1) / 2 * 0.5
    for a, b in zip(got[1:], ref[1:]):
        assert np.all(np.abs(a - b) <= 1e-10 * np.abs(b) + 1e-9)
    lr, xp, xm, xc, xz = vcorr(x, y, dx, dy, rmin=rmin, rmax=rmax, dlogr=dlogr)
    ok = ref[0] > 0
    np.testing.assert_allclose(xp[ok], (ref[2] / ref[0])[ok], atol=1e-6)
    np.testing.assert_allclose(lr[ok], (ref[1] / ref[0])[ok], atol=1e-6)
    xie, xib, logr = treegp.comp_eb(x, y, dx, dy, rmin=rmin, rmax=rmax, dlogr=dlogr)
    xie2, xib2, logr2 = treegp.comp_eb_treecorr(x, y, dx, dy, rmin=rmin, rmax=rmax, dlogr=dlogr)
    assert xie.shape == xib.shape == logr.shape == xie2.shape == (bins,)
    np.testing.assert_allclose((xie + xib)[ok], xp[ok], atol=1e-12)


def test_cabi_collective_single_rank(gpu_ready):
    """tgp_comm_* / tgp_allreduce_bins (NCCL bound by the library at run time) with a one-rank communicator: the
    packed bin buffer comes back unchanged, plane 0 still holding int64 counts; two_pcf accepts the communicator as
    its group.  (The two- and eight-rank exchange runs in bench.py under torchrun.)"""
    import torch
    import treegp_b200 as treegp
    from treegp_b200 import dist

    comm = dist.CabiComm(0, 1)
    packed = torch.zeros((3, 2, 9), dtype=torch.float64, device="cuda")
    counts = torch.arange(18, dtype=torch.int64, device="cuda").reshape(2, 9) * 1234567891
    packed[0] = counts.view(torch.float64)
    packed[1:] = torch.randn((2, 2, 9), dtype=torch.float64, device="cuda")
    before = packed.clone()
    comm.allreduce_packed_bins(packed)
    torch.cuda.synchronize()
    assert torch.equal(packed[0].view(torch.int64), counts) and torch.equal(packed[1:], before[1:])
    rng = np.random.default_rng(0)
    X, y = rng.uniform(-5, 5, size=(500, 2)), rng.normal(size=500)
    t = treegp.two_pcf(X, y, np.full(500, 0.1), 0.0, 3.0, nbins=9, anisotropic=True)
    xi0 = t.comp_2pcf(X, y, np.full(500, 0.1))[0]
    n0 = t._last_npairs.copy()
    t.group = comm
    xi1 = t.comp_2pcf(X, y, np.full(500, 0.1))[0]
    np.testing.assert_array_equal(t._last_npairs, n0)
    np.testing.assert_allclose(xi1, xi0, rtol=0, atol=1e-13 * np.abs(xi0).max())   # FP64 sums: atomic order
    comm.close()
