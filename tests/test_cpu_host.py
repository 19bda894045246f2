"""CPU: host-side logic that needs no GPU -- bin thresholds, oracle self-consistency, the von Karman
profile (host compilation of the device header), the MIGRAD stand-in, FITS table I/O, kernel lowering."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- bin thresholds ---------------------------------------------------------------------------------
@pytest.mark.parametrize("max_sep,nbins", [(14.142135623730951, 21), (1.0, 21), (5.0, 20), (707.1067811865476, 21), (3.3, 1), (2.5, 2)])
def test_twod_thresholds_reproduce_the_formula(max_sep, nbins):
    from treegp_b200 import binning

    e = binning.twod_thresholds(max_sep, nbins)
    assert e[0] == -np.inf and e[-1] == np.inf and np.all(np.diff(e[1:-1]) > 0)
    rng = np.random.default_rng(0)
    v = rng.uniform(-max_sep, max_sep, 50000)
    inner = e[1:-1]
    v = np.concatenate([v, inner, np.nextafter(inner, -np.inf), np.nextafter(inner, np.inf)])
    v = v[np.abs(v) < max_sep]
    bin_size = 2.0 * max_sep / nbins
    f = ((v + max_sep) / bin_size).astype(np.int64)
    f[f == nbins] -= 1
    g = (v[:, None] >= inner[None, :]).sum(axis=1)
    np.testing.assert_array_equal(f, g)


def test_thresholds_are_memoised_and_returned_as_private_copies():
    """two_pcf asks for the thresholds on every comp_2pcf call (1.7 ms of Python bisections per geometry): the
    second request is served from the memo, and a caller writing into its array cannot poison it."""
    from treegp_b200 import binning

    a = binning.twod_thresholds(3.25, 7)
    a_copy = a.copy()
    a[3] = 123.0
    np.testing.assert_array_equal(binning.twod_thresholds(3.25, 7), a_copy)
    assert binning.twod_thresholds(np.float64(3.25), np.int64(7)) is not binning.twod_thresholds(3.25, 7)
    assert binning._twod_thresholds.cache_info().hits >= 2
    b = binning.log_thresholds(0.2, 9.0, 11)
    b_copy = b.copy()
    b[:] = 0.0
    np.testing.assert_array_equal(binning.log_thresholds(0.2, 9.0, 11), b_copy)


def test_log_thresholds_reproduce_the_formula():
    from treegp_b200 import binning

    mn, mx, nb = 0.1, 1.75, 15
    e = binning.log_thresholds(mn, mx, nb)
    rng = np.random.default_rng(1)
    r2 = np.concatenate([rng.uniform(mn * mn, mx * mx, 20000), e[1:-1], np.nextafter(e[1:-1], 0), np.nextafter(e[1:-1], 10)])
    bs = math.log(mx / mn) / nb
    f = np.array([min(max(int((0.5 * math.log(x) - math.log(mn)) / bs), 0), nb - 1) for x in r2])
    g = (r2[:, None] >= e[None, 1:-1]).sum(axis=1)
    np.testing.assert_array_equal(f, g)


@pytest.mark.parametrize("nbins", [4, 5, 15, 20, 21])
def test_mask_and_coords_match_the_oracle_restatement(nbins):
    from oracle import pairbin_oracle as po
    from treegp_b200 import binning

    m, c = po.twod_mask_and_coords(nbins, 1.37)
    np.testing.assert_array_equal(m, binning.twod_mask(nbins))
    np.testing.assert_array_equal(c, binning.twod_coords(nbins, 1.37))
    # number of kept pixels quoted in SURVEY.md section 3.3
    assert binning.twod_mask(21).sum() == 221 and binning.twod_mask(20).sum() == 200 and binning.twod_mask(15).sum() == 113


# ---- oracle: C vs independent numpy restatement -----------------------------------------------------
@pytest.mark.parametrize("cfg", [("TwoD", 0.0, 14.1, 21), ("TwoD", 0.3, 5.0, 20), ("Log", 0.1, 1.75, 15)])
@pytest.mark.parametrize("weighted", [False, True])
def test_c_oracle_matches_numpy_oracle(cfg, weighted):
    from oracle import pairbin_oracle as po

    bt, mn, mx, nb = cfg
    rng = np.random.default_rng(3)
    n = 600
    x, y, k = rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.normal(size=n)
    x[5], y[5] = x[9], y[9]
    w = rng.uniform(0.5, 2.0, n) if weighted else None
    a, b = po.pairbin(x, y, k, w, mn, mx, nb, bt), po.pairbin_numpy(x, y, k, w, mn, mx, nb, bt)
    np.testing.assert_array_equal(a["npairs"], b["npairs"])
    np.testing.assert_allclose(a["sumwkk"], b["sumwkk"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(a["weight"], b["weight"], rtol=1e-12)
    # row slabs add up (what the multi-GPU split relies on)
    parts = [po.pairbin(x, y, k, w, mn, mx, nb, bt, rows=r) for r in ((0, 100), (100, 350), (350, n))]
    np.testing.assert_array_equal(sum(p["npairs"] for p in parts), a["npairs"])
    if bt == "TwoD":
        c = a["npairs"].reshape(nb, nb)
        np.testing.assert_array_equal(c, c[::-1, ::-1])  # point symmetry two_pcf.py:306-321 relies on


# ---- von Karman profile: the device header compiled for the host ------------------------------------
@pytest.fixture(scope="module")
def vk_host(tmp_path_factory):
    d = tmp_path_factory.mktemp("vk")
    src = d / "h.cpp"
    src.write_text('#include "vk_profile.cuh"\nextern "C" void vk_eval(const double* q, long n, double* out)'
                   '{ for (long i = 0; i < n; i++) out[i] = tgp_vk_profile(q[i], tgp_vk_phi_host); }\n')
    so = d / "libvk.so"
    subprocess.check_call(["/usr/bin/g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC",
                           "-I", os.path.join(ROOT, "treegp_b200", "csrc"), str(src), "-o", str(so)])
    lib = ctypes.CDLL(str(so))

    def f(q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        out = np.empty_like(q)
        lib.vk_eval(q.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(q.size), out.ctypes.data_as(ctypes.c_void_p))
        return out

    return f


def test_vk_profile_against_mpmath(vk_host):
    import mpmath as mp

    mp.mp.dps = 40
    nu = mp.mpf(5) / 6
    rng = np.random.default_rng(1)
    d = np.concatenate([10 ** rng.uniform(-8, 1.2, 1500), rng.uniform(0, 3, 500), [1 / (2 * np.pi), 0.5 / np.pi]])
    got = vk_host(d * d)
    for di, gi in zip(d, got):
        q = mp.mpf(float(di)) ** 2
        z = 2 * mp.pi * mp.sqrt(q)
        truth = mp.mpf(2) ** (mp.mpf(1) / 6) / mp.gamma(nu) * z ** nu * mp.besselk(nu, z)
        assert abs(mp.mpf(float(gi)) - truth) < mp.mpf(4e-15) * max(truth, mp.mpf(1e-3)), (di, gi)
    assert vk_host(np.array([0.0]))[0] == 1.0
    assert vk_host(np.array([1e6]))[0] == 0.0


def test_vk_profile_against_scipy_reference_formula(vk_host):
    """The reference's own expression (kernels.py:255-262) evaluated with scipy: atol 1e-12 is the
    reference tests' tolerance (tests/test_kernels.py:122-123)."""
    from scipy import special

    d = np.concatenate([np.linspace(1e-6, 20, 20001), 10 ** np.linspace(-12, 2, 2000)])
    lim0 = special.gamma(5 / 6) / (2 * np.pi ** (5 / 6))
    ref = d ** (5 / 6) * special.kv(5 / 6, 2 * np.pi * d) / lim0
    np.testing.assert_allclose(vk_host(d * d), ref, rtol=0, atol=5e-14)


# ---- MIGRAD stand-in ----------------------------------------------------------------------------------
def test_migrad_finds_minima_and_flags_accuracy():
    from treegp_b200.migrad import Migrad

    def rosen(p):
        return (1 - p[0]) ** 2 + 100 * (p[1] - p[0] ** 2) ** 2 + (p[2] - 0.3) ** 2

    m = Migrad(rosen, [-1.2, 1.0, 0.0]).migrad()
    assert m.accurate and m.fval < 1e-3
    np.testing.assert_allclose(m.values, [1.0, 1.0, 0.3], atol=0.05)

    def walled(p):  # inf outside |g| <= 1, like two_pcf.py:105-106
        if abs(p[1]) > 1 or abs(p[2]) > 1:
            return np.inf
        return 50 * np.log(p[0] / 0.5) ** 2 + 30 * (p[1] - 0.2) ** 2 + 30 * (p[2] + 0.9) ** 2 + 3 * p[1] * p[2]

    m = Migrad(walled, [0.3, 0.0, 0.0]).migrad()
    assert m.accurate
    np.testing.assert_allclose(m.values, [0.5, 0.246, -0.912], atol=2e-2)
    m = Migrad(lambda p: np.inf, [1.0, 0.0, 0.0]).migrad()
    assert not m.accurate


# ---- FITS table I/O -------------------------------------------------------------------------------------
def test_fits_table_round_trip(tmp_path):
    from treegp_b200 import fitstable

    rng = np.random.default_rng(0)
    cols = {"COORDS0": rng.normal(size=(37, 2)), "PARAMS0": rng.normal(size=37), "WRMS0": np.zeros(37)}
    path = str(tmp_path / "mean.fits")
    fitstable.write_table(path, cols)
    assert os.path.getsize(path) % 2880 == 0
    back = fitstable.read_table(path)
    for k, v in cols.items():
        np.testing.assert_array_equal(back[k][0], v)


# ---- kernel objects / lowering ------------------------------------------------------------------------
def test_kernel_protocol_and_lowering(golden):
    import treegp_b200 as treegp
    from treegp_b200 import _cabi
    from treegp_b200.kernels import lower_kernel

    k = treegp.AnisotropicRBF(invLam=golden["rt_invLam"])
    np.testing.assert_allclose(k.theta, golden["rt_theta"], atol=1e-14)
    k.theta = golden["rt_theta2"]
    np.testing.assert_allclose(k.invLam, golden["rt_invLam2"], atol=1e-14)
    for name in [str(c) for c in golden["cases"]]:
        ker = treegp.eval_kernel(str(golden["kstr_" + name]))
        np.testing.assert_allclose(ker.theta, golden["theta_" + name], atol=1e-12)  # same theta layout as the reference
        clone = ker.clone_with_theta(ker.theta + 0.1)
        np.testing.assert_allclose(clone.theta, ker.theta + 0.1, atol=1e-12)
        assert ker.bounds.shape == (len(ker.theta), 2)
    d = lower_kernel(treegp.eval_kernel("2.0**2 * RBF(0.5)"), 2)
    assert (d.family, d.ndim, d.amp, d.m00, d.m01, d.m11) == (_cabi.FAM_RBF, 2, 4.0, 4.0, 0.0, 4.0)
    d = lower_kernel(treegp.eval_kernel("3.0 * VonKarman(length_scale=2.0)"), 1)
    assert (d.family, d.ndim, d.amp, d.m00) == (_cabi.FAM_VONKARMAN, 1, 3.0, 0.25)
    d = lower_kernel(treegp.eval_kernel("Matern(length_scale=2.0, nu=2.5)"), 2)
    assert d.family == _cabi.FAM_MATERN52 and d.amp == 1.0
    with pytest.raises(_cabi.TgpError):
        lower_kernel(treegp.eval_kernel("Matern(length_scale=2.0, nu=0.7)"), 2)
    with pytest.raises(TypeError):
        treegp.AnisotropicRBF(invLam=np.eye(2), scale_length=[1.0, 1.0])
    with pytest.raises(RuntimeError):
        treegp.eval_kernel("NoSuchKernel(1)")
    with pytest.raises(TypeError):
        treegp.eval_kernel("RBF")


def test_gpinterpolation_argument_errors():
    import treegp_b200 as treegp

    with pytest.raises(TypeError):
        treegp.GPInterpolation(kernel=3)
    with pytest.raises(ValueError):
        treegp.GPInterpolation(optimizer="nope")
    with pytest.raises(ValueError):
        treegp.two_pcf(np.zeros((4, 3)), np.zeros(4), np.zeros(4), 0.0, 1.0)
    gp = treegp.GPInterpolation(kernel="RBF(1)", optimizer="none")
    assert gp.robust_fit is False and gp.nbins == 20 and gp.n_neighbors == 4


def test_product_path_never_falls_back_to_cpu():
    """Without a CUDA device every compute entry must raise (no silent CPU path)."""
    import torch
    import treegp_b200 as treegp
    from treegp_b200 import _cabi

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    k = treegp.eval_kernel("1.0 * RBF(1.0)")
    with pytest.raises(_cabi.TgpError):
        k.k2(np.zeros((3, 2))) if False else treegp.AnisotropicRBF(scale_length=[1.0, 1.0])(np.zeros((3, 2)))
    gp = treegp.GPInterpolation(kernel="1.0 * AnisotropicRBF(scale_length=[1.0, 1.0])", optimizer="none")
    gp.initialize(np.zeros((3, 2)), np.zeros(3))
    with pytest.raises(_cabi.TgpError):
        gp.predict(np.zeros((2, 2)))


# ---- mean function producer and E/B diagnostics (SURVEY 8f "next" rows, host-side) ---------------------
@pytest.mark.parametrize("stat", ["mean", "median", "weighted"])
def test_meanify_matches_binned_statistic(stat, tmp_path):
    """tests/test_meanify.py:43-64 shape: many fields, paraboloid mean, bin_spacing 40."""
    from scipy.stats import binned_statistic_2d
    import treegp_b200 as treegp
    from treegp_b200.meanify import read_average

    rng = np.random.default_rng(0)
    m = treegp.meanify(bin_spacing=40.0, statistics=stat)
    C, P, E = [], [], []
    for _ in range(30):
        c = rng.uniform(-1000, 1000, (500, 2))
        p = 1e-6 * (c[:, 0] ** 2 + c[:, 1] ** 2) + rng.normal(0, 0.01, 500)
        e = 0.01 * rng.uniform(0.5, 2, 500)
        m.add_field(c, p, params_err=e if stat == "weighted" else None)
        C.append(c); P.append(p); E.append(e)
    m.meanify()
    c, p, e = np.concatenate(C), np.concatenate(P), np.concatenate(E)
    b = [np.linspace(c[:, 0].min(), c[:, 0].max(), int((c[:, 0].max() - c[:, 0].min()) / 40.0)),
         np.linspace(c[:, 1].min(), c[:, 1].max(), int((c[:, 1].max() - c[:, 1].min()) / 40.0))]
    if stat == "weighted":
        w = 1 / e ** 2
        with np.errstate(invalid="ignore"):
            ref = (binned_statistic_2d(c[:, 0], c[:, 1], w * p, bins=b, statistic="sum")[0]
                   / binned_statistic_2d(c[:, 0], c[:, 1], w, bins=b, statistic="sum")[0])
    else:
        ref = binned_statistic_2d(c[:, 0], c[:, 1], p, bins=b, statistic=stat)[0]
    np.testing.assert_allclose(m._average, ref.T, rtol=0, atol=1e-14, equal_nan=True)
    truth = 1e-6 * (m.coords0[:, 0] ** 2 + m.coords0[:, 1] ** 2)
    np.testing.assert_allclose(m.params0, truth, atol=0.2)   # reference tolerance, tests/test_meanify.py:59-60
    path = str(tmp_path / "mean.fits")
    m.save_results(path)
    X0, y0 = read_average(path)
    np.testing.assert_array_equal(X0, m.coords0)
    np.testing.assert_array_equal(y0, m.params0)
    with pytest.raises(ValueError):
        treegp.meanify(statistics="mode")
    with pytest.raises(ValueError):
        treegp.meanify().add_field(np.zeros((3, 1)), np.zeros(3))


def test_eb_oracle_matches_all_pairs_formula():
    """oracle/eb_oracle.py (the checker of csrc/vcorr.cu) against the reference's own all-pairs formulas
    (utils.py:38-74), and the threshold form of the log-radius binning against the floor formula."""
    from oracle import eb_oracle
    from treegp_b200 import binning

    rng = np.random.default_rng(2)
    n = 500
    x, y, dx, dy = rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.normal(size=n), rng.normal(size=n)
    rmin, dlogr = 0.01, 0.2
    bins = int(np.ceil(np.log(1 / rmin) / dlogr))
    cnt_o, slr, sp, sz2, sm = eb_oracle.pair_sums(x, y, dx, dy, np.log(rmin), dlogr, bins)
    i1, i2 = np.triu_indices(n, 1)
    dr = (x[i2] - x[i1]) + 1j * (y[i2] - y[i1])
    ld = np.log(np.abs(dr))
    hr = (np.log(rmin), np.log(rmin) + bins * dlogr)
    cnt = np.histogram(ld, bins=bins, range=hr)[0]
    np.testing.assert_array_equal(cnt_o, cnt)
    v = dx + 1j * dy
    np.testing.assert_allclose(sp, np.histogram(ld, bins=bins, range=hr, weights=dx[i1] * dx[i2] + dy[i1] * dy[i2])[0], atol=1e-11)
    vv = v[i1] * v[i2] * np.conj(dr) ** 2 / np.abs(dr) ** 2
    np.testing.assert_allclose(sm.real, np.histogram(ld, bins=bins, range=hr, weights=vv.real)[0], atol=1e-11)
    np.testing.assert_allclose(sm.imag, np.histogram(ld, bins=bins, range=hr, weights=vv.imag)[0], atol=1e-11)
    # thresholds on r^2: outside the undecided band (1e-14 relative, settled on the host) they give np.histogram's bin
    ed = binning.hist_thresholds_r2(np.log(rmin), dlogr, bins)
    r2 = (x[i2] - x[i1]) ** 2 + (y[i2] - y[i1]) ** 2
    k_hist = binning.hist_bin(ld, binning.hist_edges(np.log(rmin), dlogr, bins))
    k_thr = np.searchsorted(ed, r2, side="right") - 1
    k_thr[(k_thr < 0) | (k_thr >= bins)] = -1
    near = np.min(np.abs(r2[:, None] / ed[None, :] - 1.0), axis=1) <= 1e-14
    np.testing.assert_array_equal(k_thr[~near], k_hist[~near])
    assert near.sum() < 5
    edges = binning.hist_edges(np.log(rmin), dlogr, bins)
    for k in range(bins + 1):
        h = np.sqrt(ed[k])
        assert abs(np.log(h) - edges[k]) < 1e-14 * max(1.0, abs(edges[k]))


def test_truncation_thresholds_are_below_1e_minus_40():
    """tgp_profile_qcut (host function of the library; csrc/predict.cu): f(q_cut) <= 1e-40 for every family, f decreasing
    beyond -- the bound tgp_predict_mean_trunc's skip rule rests on.  Checked with mpmath (60 digits)."""
    import mpmath as mp
    from treegp_b200 import _cabi

    lib = _cabi.load()
    mp.mp.dps = 60

    def vk(q):
        z = 2 * mp.pi * mp.sqrt(q)
        lim0 = mp.gamma(mp.mpf(5) / 6) / (2 * mp.pi ** (mp.mpf(5) / 6))
        return mp.sqrt(q) ** (mp.mpf(5) / 6) * mp.besselk(mp.mpf(5) / 6, z) / lim0

    fams = {_cabi.FAM_RBF: lambda q: mp.exp(-q / 2), _cabi.FAM_VONKARMAN: vk,
            _cabi.FAM_MATERN12: lambda q: mp.exp(-mp.sqrt(q)),
            _cabi.FAM_MATERN32: lambda q: (1 + mp.sqrt(3 * q)) * mp.exp(-mp.sqrt(3 * q)),
            _cabi.FAM_MATERN52: lambda q: (1 + mp.sqrt(5 * q) + 5 * q / 3) * mp.exp(-mp.sqrt(5 * q))}
    for fam, f in fams.items():
        qc = mp.mpf(lib.tgp_profile_qcut(fam))
        assert f(qc) <= mp.mpf("1e-40")
        assert f(qc * 0.98) > mp.mpf("1e-41")     # not wastefully large
        assert f(qc * 1.5) < f(qc) and f(qc * 4) < f(qc * 1.5)


def _pcg_state(rng):
    st = rng.bit_generator.state
    v, inc = st["state"]["state"], st["state"]["inc"]
    m64 = (1 << 64) - 1
    return np.array([v >> 64, v & m64, inc >> 64, inc & m64, st["has_uint32"], st["uinteger"]], dtype=np.uint64)


@pytest.mark.parametrize("seed", [610639139, 1, 2])
@pytest.mark.parametrize("n,b", [(2, 3), (3, 5), (7, 11), (1000, 13), (40001, 7), (40000, 60), (1 << 22, 2)])
def test_bootstrap_generator_is_numpy_bit_for_bit(seed, n, b):
    """csrc/hostrng.cu against numpy itself: the multiplicities of b resamples equal the bincount of b successive
    default_rng(seed).integers(0, n - 1, size=n) calls (two_pcf.py:266,275), also from an odd position in the
    32-bit stream and through a storage permutation, and the generator state afterwards is numpy's."""
    from treegp_b200 import _cabi

    lib = _cabi.load()
    r1, r2 = np.random.default_rng(seed), np.random.default_rng(seed)
    r1.integers(0, 5, size=3)
    r2.integers(0, 5, size=3)          # leaves half a 64-bit word buffered
    idx = np.stack([r1.integers(0, n - 1, size=n) for _ in range(b)])
    perm = np.random.default_rng(5).permutation(n)
    ref = np.stack([np.bincount(perm[row], minlength=n) for row in idx])
    st = _pcg_state(r2)
    mult = np.empty((b, n), dtype=np.uint8)
    assert lib.tgp_bootstrap_multiplicities(st.ctypes.data, n, b, perm.ctypes.data, mult.ctypes.data) == 0
    assert np.array_equal(ref, mult)
    assert np.array_equal(st, _pcg_state(r1))
    assert not (ref[:, perm[n - 1]] != 0).any()      # the reference's quirk: index n-1 is never drawn


def test_draw_multiplicities_continues_the_python_generator():
    """two_pcf._draw_multiplicities hands the advanced PCG64 state back, so resample_bootstrap() afterwards
    returns what it would have after the same number of numpy draws."""
    import treegp_b200 as treegp

    rng = np.random.default_rng(3)
    n = 500
    X, y, e = rng.uniform(size=(n, 2)), rng.normal(size=n), np.full(n, 0.1)
    a = treegp.two_pcf(X, y, e, 0.0, 0.3, nbins=5, anisotropic=True)
    b = treegp.two_pcf(X, y, e, 0.0, 0.3, nbins=5, anisotropic=True)
    mult = a._draw_multiplicities(4, n, None).numpy()
    for r in range(4):
        u, v, yy, ee = b.resample_bootstrap()
        assert np.array_equal(np.bincount(np.searchsorted(np.sort(y), yy), minlength=n)[np.argsort(np.argsort(y))],
                              mult[r])
    ua, _, ya, _ = a.resample_bootstrap()
    ub, _, yb, _ = b.resample_bootstrap()
    assert np.array_equal(ua, ub) and np.array_equal(ya, yb)


# ---- windowed predictive variance: host-side plan ------------------------------------------------------
@pytest.mark.parametrize("ndim", [1, 2])
def test_variance_window_plan_only_skips_uncorrelated_points(ndim):
    """backend.support_cutoffs / plan_var_windows (pure host logic): every training point a chunk skips lies at
    q = d^T M d >= q_cut from EVERY test point of the chunk, in the ascending and in the descending order; the
    skips are multiples of 64 and leave a non-empty system; the flop counts are what the skips imply."""
    from treegp_b200 import _cabi, backend, eval_kernel
    from treegp_b200.kernels import lower_kernel

    rng = np.random.default_rng(5 + ndim)
    if ndim == 2:
        Mi = np.array([[0.9, -0.35], [-0.35, 0.6]])
        desc = lower_kernel(eval_kernel("2.0 * AnisotropicVonKarman(invLam=array([[0.9, -0.35], [-0.35, 0.6]]))"), 2)
    else:
        Mi = np.array([[1.0 / 0.7 ** 2]])
        desc = lower_kernel(eval_kernel("2.0 * RBF(0.7)"), 1)
    qcut = float(_cabi.load().tgp_profile_qcut(int(desc.family)))
    dcut = backend.support_cutoffs(desc)
    assert len(dcut) == ndim
    # the per-axis cut-off is exactly the smallest |d_a| that guarantees q >= q_cut: minimise q over the other axis
    if ndim == 2:
        for a in (0, 1):
            o = 1 - a
            d = np.zeros(2)
            d[a] = dcut[a]
            d[o] = -Mi[a, o] * d[a] / Mi[o, o]
            assert abs(d @ Mi @ d - qcut) < 1e-9 * qcut
    n, m, chunk = 5000, 20000, 1500
    field = 400.0
    X = rng.uniform(0, field, size=(n, ndim))
    Xs = rng.uniform(0, field, size=(m, ndim))
    axis = int(np.argmin(np.asarray(dcut)))          # same extent on both axes
    ot, os_ = np.argsort(X[:, axis], kind="stable"), np.argsort(Xs[:, axis], kind="stable")
    Xt, Xss = X[ot], Xs[os_]
    starts = np.arange(0, m, chunk)
    ends = np.minimum(starts + chunk, m) - 1
    sizes = ends - starts + 1
    sa, sd, use_desc, f_win, f_full = backend.plan_var_windows(Xt[:, axis], Xss[starts, axis], Xss[ends, axis], dcut[axis], sizes)
    assert np.all(sa % 64 == 0) and np.all(sd % 64 == 0) and np.all(sa < n) and np.all(sd < n)
    assert use_desc.any() and (~use_desc).any() and f_win < 0.5 * f_full
    Xdesc = Xt[::-1]
    left = 0.0
    for c in range(len(starts)):
        P = Xss[starts[c]:ends[c] + 1]
        for skipped in (Xt[:sa[c]], Xdesc[:sd[c]]):
            if len(skipped):
                d = P[:, None, :] - skipped[None, :, :]
                q = np.einsum("pta,ab,ptb->pt", d, Mi, d)
                assert q.min() >= qcut
        left += sizes[c] * float(n - (sd[c] if use_desc[c] else sa[c])) ** 2
    assert f_win == left and f_full == float(m) * n * n


# ---- envelope factorisation: host-side plan ---------------------------------------------------------------
@pytest.mark.parametrize("n,dcut", [(5000, 7.0), (1537, 0.5), (4096, 1000.0), (700, 3.0)])
def test_envelope_rows_bound_every_correlated_pair(n, dcut):
    """backend.envelope_rows / envelope_flops / envelope_trsm_flops (pure host logic behind tgp_potrf_env): a row at or
    beyond row_end[b] is at least d_cut away from EVERY column of block b and of all earlier blocks (so K is below
    1e-40 amp there and the factor stays zero), the bound is tight to one row, ends are non-decreasing and reach N,
    sub-systems (start > 0) get the same envelope shifted, and the flop counts are the sums they claim to be."""
    from treegp_b200 import backend

    ob = backend.envelope_block()
    assert ob == 512
    rng = np.random.default_rng(n)
    x = np.sort(np.concatenate([rng.uniform(0, 100, n - n // 5), rng.normal(50, 2, n // 5)]))   # with a dense clump
    re = backend.envelope_rows(x, dcut)
    nb = (n + ob - 1) // ob
    assert re.dtype == np.int64 and len(re) == nb and re[-1] == n and np.all(np.diff(re) >= 0)
    for b in range(nb):
        c1 = min(n, (b + 1) * ob)
        assert c1 <= re[b] <= n
        if re[b] < n:
            assert x[re[b]] - x[c1 - 1] >= dcut           # first row outside: far from the block's last column ...
            assert np.all(x[re[b]:] - x[c1 - 1] >= dcut)  # ... and so is every later row, from every earlier column
        if re[b] > c1:
            assert x[re[b] - 1] - x[c1 - 1] < dcut * (1 + 2e-9)   # tight: the last row inside is within reach
    # a trailing sub-system has the same envelope, shifted
    s = 192
    if n > s + ob:
        re_s = backend.envelope_rows(x, dcut, start=s)
        np.testing.assert_array_equal(re_s, backend.envelope_rows(x[s:], dcut))
    # flop counts
    k = np.arange(0, n, ob)
    w = np.minimum(ob, n - k)
    below = re - (k + w)
    assert backend.envelope_flops(re, n) == pytest.approx(float(np.sum(w ** 3 / 3 + w * w * below + w * below * below)))
    assert backend.envelope_trsm_flops(re, n) == pytest.approx(float(np.sum(w * w + 2 * w * below)))
    if dcut >= 1000.0:   # support wider than the data: the envelope is the whole triangle
        assert np.all(re == n)
        assert backend.envelope_flops(re, n) == pytest.approx(n ** 3 / 3, rel=0.3)


def test_plan_envelope_declines_small_and_wide_problems():
    """backend.plan_envelope returns None before touching the device for N < ENVELOPE_MIN_N (the reference's tests and
    the golden vectors never take the envelope path)."""
    import torch
    from treegp_b200 import backend, eval_kernel
    from treegp_b200.kernels import lower_kernel

    desc = lower_kernel(eval_kernel("1.0 * AnisotropicRBF(invLam=array([[400.0, 0.0], [0.0, 400.0]]))"), 2)
    X = torch.rand((backend.ENVELOPE_MIN_N - 1, 2), dtype=torch.float64)
    assert backend.plan_envelope(X, desc) is None
    # wide support: declined from the cut-offs alone (CPU tensors are enough for the extent)
    wide = lower_kernel(eval_kernel("1.0 * AnisotropicRBF(invLam=array([[0.5, 0.0], [0.0, 0.5]]))"), 2)
    assert backend.plan_envelope(torch.rand((5000, 2), dtype=torch.float64) * 10.0, wide) is None
    # a metric that is not positive definite has no cut-off: dense path
    bad = lower_kernel(eval_kernel("1.0 * AnisotropicRBF(invLam=array([[1.0, 0.0], [0.0, 1.0]]))"), 2)
    bad.m01 = 5.0
    assert not np.all(np.isfinite(backend.support_cutoffs(bad)))
    assert backend.plan_envelope(torch.rand((5000, 2), dtype=torch.float64) * 100.0, bad) is None
    # and a short support on CPU tensors gives a plan (host logic only: sort, searchsorted, flop model)
    narrow = lower_kernel(eval_kernel("1.0 * AnisotropicRBF(invLam=array([[4.0, 0.0], [0.0, 1.0]]))"), 2)
    plan = backend.plan_envelope(torch.rand((6000, 2), dtype=torch.float64) * 300.0, narrow)
    assert plan is not None and plan["axis"] == 0 and plan["flops"] * 2 <= plan["flops_dense"]
    assert np.all(np.diff(plan["x"]) >= 0) and len(plan["row_end"]) == 12


@pytest.mark.parametrize("family,ndim", [("rbf", 2), ("vonkarman", 2), ("rbf", 1)])
def test_envelope_factorisation_equals_the_dense_one_on_the_oracle(family, ndim):
    """The algorithm of tgp_potrf_env restated in numpy (oracle.gp_oracle.cholesky_envelope) with the envelope the
    host plans (backend.support_cutoffs / envelope_rows, 64-column blocks here so that a small case has many block
    columns): same factor as scipy's dense Cholesky of the same sorted matrix, K below 1e-40 amp outside the envelope,
    and the likelihood of log_likelihood.py:29-37 from either factor agrees to rounding."""
    from oracle import gp_oracle as go
    from treegp_b200 import backend, eval_kernel
    from treegp_b200.kernels import lower_kernel

    rng = np.random.default_rng(3 + ndim)
    n, field, blk = 900, 400.0, 64
    Mi = np.array([[0.9, -0.35], [-0.35, 0.6]]) if ndim == 2 else np.array([[1.0 / 0.64]])
    name = {"rbf": "AnisotropicRBF", "vonkarman": "AnisotropicVonKarman"}[family]
    kstr = ("2.0 * %s(invLam=array([[0.9, -0.35], [-0.35, 0.6]]))" % name) if ndim == 2 else "2.0 * RBF(0.8)"
    desc = lower_kernel(eval_kernel(kstr), ndim)
    X = rng.uniform(0, field, size=(n, ndim))
    dcut = backend.support_cutoffs(desc)
    axis = int(np.argmin(dcut))
    X = X[np.argsort(X[:, axis], kind="stable")]
    x = X[:, axis]
    # envelope_rows with the product's block width, restated for 64-column blocks
    c1 = np.minimum(np.arange(blk, n + blk, blk), n)
    row_end = np.maximum(np.searchsorted(x, x[c1 - 1] + dcut[axis] * (1 + 1e-9), side="left"), c1)
    assert row_end[0] < n // 2 and row_end[-1] == n          # a real envelope, many block columns
    e2 = rng.uniform(0.005, 0.02, size=n)
    K = go.kmat(family, X, amp=2.0, invLam=Mi) + np.diag(e2)
    outside = np.zeros((n, n), dtype=bool)
    for b, r in enumerate(row_end):
        outside[r:, blk * b:blk * (b + 1)] = True
    assert outside.sum() > 0.5 * n * n / 2 and np.max(np.abs(K[outside])) <= 1e-40 * 2.0
    from scipy.linalg import cholesky
    Ld = cholesky(K, lower=True)
    Le = np.tril(go.cholesky_envelope(K, row_end, block=blk))
    assert np.max(np.abs(Ld[outside])) < 1e-35
    np.testing.assert_allclose(Le[~outside], Ld[~outside], rtol=0, atol=1e-13 * np.max(np.abs(Ld)))
    np.testing.assert_array_equal(Le[outside], np.tril(K)[outside])   # never touched
    y = rng.normal(size=n)
    logl_d, alpha_d = go.log_likelihood(K, y)
    from scipy.linalg import cho_solve
    alpha_e = cho_solve((Le, True), y)
    logl_e = -0.5 * y @ alpha_e - np.sum(np.log(np.diag(Le))) - 0.5 * n * np.log(2 * np.pi)
    assert abs(logl_e - logl_d) <= 1e-12 * abs(logl_d)
    np.testing.assert_allclose(alpha_e, alpha_d, rtol=0, atol=1e-10 * np.max(np.abs(alpha_d)))
