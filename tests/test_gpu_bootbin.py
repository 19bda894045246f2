"""GPU: the shared-geometry bootstrap batch (tgp_bootbin_twod; two_pcf.py:342-362 of the reference) against
* the per-catalogue batch (tgp_pairbin on the compacted weighted resamples, itself bit-exact against the oracle), and
* the sequential CPU oracle (one brute-force weighted count per resample),
for every combination of the kernel's paths, weighted and unweighted, odd and even bin counts, partial chunks,
lattices (displacements exactly on bin edges) and tiny catalogues."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _field(n, seed, lattice=False, L=100.0):
    rng = np.random.default_rng(seed)
    if lattice:
        side = int(np.ceil(np.sqrt(n)))
        gx, gy = np.meshgrid(np.arange(side, dtype=float), np.arange(side, dtype=float))
        X = np.column_stack([gx.ravel(), gy.ravel()])[:n] * (L / side)
    else:
        X = rng.uniform(0, L, size=(n, 2))
    y = np.sin(X[:, 0] / 7.0) + 0.5 * np.cos(X[:, 1] / 11.0) + 0.3 * rng.normal(size=n) + 0.7
    return X, y


def _xi(tp, B, shared, paths=None):
    from treegp_b200 import backend

    tp.SHARED_BOOTSTRAP = shared
    tp._rng = None
    if paths is not None:
        backend.set_option("bootbin_paths", paths)
    try:
        return tp._bootstrap_xi(B)
    finally:
        if paths is not None:
            backend.set_option("bootbin_paths", 15)


def _close(a, b, scale, tol=2e-11):
    assert a.shape == b.shape
    assert np.all(np.isfinite(a)) and np.all(np.isfinite(b))
    err = np.max(np.abs(a - b))
    assert err <= tol * scale, (err, scale)


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("nbins,max_frac", [(21, 0.5), (20, 0.5), (21, 0.08), (7, 0.25)])
def test_shared_batch_equals_catalogue_batch(weighted, nbins, max_frac):
    import treegp_b200 as treegp
    from treegp_b200 import backend

    n, B = 2999, 40
    X, y = _field(n, 3)
    rng = np.random.default_rng(5)
    err = rng.uniform(0.05, 0.2, size=n) if weighted else np.zeros(n)
    L = 100.0
    tp = treegp.two_pcf(X, y, err, 0.0, max_frac * np.hypot(L, L), nbins=nbins, anisotropic=True)
    ref = _xi(tp, B, shared=False)
    scale = np.var(y)
    for paths in (0, 2, 3, 6, 7, 15):
        backend.bootbin_stats(reset=True)
        got = _xi(tp, B, shared=True, paths=paths)
        st = backend.bootbin_stats()
        _close(got, ref, scale)
        if paths == 0:
            assert st["closed_form"] == 0 and st["sweeps"] == 0 and st["pairwise"] == 0 and st["exact_per_pair"] > 0
        if paths == 15:
            assert st["closed_form"] + st["sweeps"] + st["pairwise"] + st["bin_by_bin"] > 0


def test_unweighted_pair_weights_are_exact_integers():
    """sum a_i a_j of an unweighted resample is an integer below 2^53: equal to the catalogue batch bit for bit."""
    import torch
    import treegp_b200 as treegp
    from treegp_b200 import backend

    n, B = 4100, 33
    X, y = _field(n, 11)
    L = 100.0
    tp = treegp.two_pcf(X, y, np.zeros(n), 0.0, 0.5 * np.hypot(L, L), nbins=21, anisotropic=True)
    _xi(tp, B, shared=False)
    sw_ref = np.array(tp._last_sumw)
    # the same draws through the shared kernel, sums read back
    tp._rng = None
    x = backend.to_device(X[:, 0]); yy = backend.to_device(X[:, 1]); val = backend.to_device(y)
    order = backend.hilbert_order(x, yy)
    pos = np.empty(n, dtype=np.int64)
    pos[order.cpu().numpy()] = np.arange(n)
    mult = tp._draw_multiplicities(B, n, pos).to(x.device)
    z = (val - val.mean())[order].contiguous()
    sums, delta = backend.bootbin_sums(x[order].contiguous(), yy[order].contiguous(), z, None, mult,
                                       tp._device_edges(tp._bin_geometry()[1]), 21, 0.0, tp.max_sep)
    xi, sw = backend.bootbin_xi(sums, delta, 21, B, want_sumw=True)
    assert np.array_equal(sw.cpu().numpy(), sw_ref)
    # the resample means
    m = mult.cpu().numpy().astype(np.float64)
    zz = z.cpu().numpy()
    assert np.allclose(delta.cpu().numpy()[:B], m @ zz / n, rtol=0, atol=1e-14)
    torch.cuda.synchronize()


@pytest.mark.parametrize("nbins", [21, 20])
def test_lattice_displacements_on_bin_edges(nbins):
    """Points on a lattice whose spacing divides the bin size: many displacements sit exactly on thresholds, where
    the mirrored bin is not the mirror image of the forward bin (the correction histogram)."""
    import treegp_b200 as treegp
    from treegp_b200 import backend

    n, B = 64 * 64, 34
    X, y = _field(n, 7, lattice=True, L=64.0)
    tp = treegp.two_pcf(X, y, np.zeros(n), 0.0, 21.0 if nbins == 21 else 20.0, nbins=nbins, anisotropic=True)
    ref = _xi(tp, B, shared=False)
    for paths in (0, 7, 15):
        backend.bootbin_stats(reset=True)
        got = _xi(tp, B, shared=True, paths=paths)
        _close(got, ref, np.var(y))


@pytest.mark.parametrize("n", [2, 3, 31, 32, 33, 65])
def test_tiny_catalogues(n):
    import treegp_b200 as treegp

    X, y = _field(n, 20 + n, L=10.0)
    tp = treegp.two_pcf(X, y, np.full(n, 0.1), 0.0, 6.0, nbins=5, anisotropic=True)
    ref = _xi(tp, 5, shared=False)
    got = _xi(tp, 5, shared=True)
    _close(got, ref, max(np.var(y), 1e-3))


def test_against_the_sequential_oracle():
    """Each resample as the reference draws it (indices with repetition, every copy its own point), counted by the
    CPU oracle."""
    import treegp_b200 as treegp
    from oracle import pairbin_oracle as po

    n, B, nbins = 700, 6, 9
    X, y = _field(n, 31, L=50.0)
    err = np.full(n, 0.25)
    max_sep = 20.0
    tp = treegp.two_pcf(X, y, err, 0.0, max_sep, nbins=nbins, anisotropic=True)
    got = _xi(tp, B, shared=True)
    rng = np.random.default_rng(tp.seed)
    for b in range(B):
        idx = rng.integers(0, n - 1, size=n)
        xi, _, _, _ = po.comp_2pcf(X[idx], y[idx], err[idx], 0.0, max_sep, nbins, True)
        assert np.max(np.abs(got[b] - np.asarray(xi).ravel())) <= 1e-11 * np.var(y)


def test_large_batch_matches_catalogue_batch():
    """Benchmark density (N = 60k, default max_sep, 64 resamples): all paths at production sizes."""
    import treegp_b200 as treegp
    from treegp_b200 import backend

    n, B = 60000, 64
    L = 1000.0 * np.sqrt(n / 1e6)
    X, y = _field(n, 41, L=L)
    tp = treegp.two_pcf(X, y, np.full(n, 0.1), 0.0, 0.5 * np.hypot(L, L), nbins=21, anisotropic=True)
    ref = _xi(tp, B, shared=False)
    backend.bootbin_stats(reset=True)
    got = _xi(tp, B, shared=True)
    st = backend.bootbin_stats()
    _close(got, ref, np.var(y))
    assert st["closed_form"] > 0 and st["sweeps"] > 0 and st["exact_per_pair"] < st["closed_form"]


def test_small_remainder_goes_through_the_catalogue_path_with_the_same_stream():
    """n_bootstrap = 32 k + (a few): the remainder is drawn from the same random stream and counted as independent
    catalogues; the result equals the all-catalogue batch resample by resample."""
    import treegp_b200 as treegp

    n, B = 1500, 35
    X, y = _field(n, 51)
    tp = treegp.two_pcf(X, y, np.full(n, 0.3), 0.0, 30.0, nbins=9, anisotropic=True)
    ref = _xi(tp, B, shared=False)
    got = _xi(tp, B, shared=True)
    _close(got, ref, np.var(y))
    # and the generator ends in the same state: the next draw agrees
    a = tp.resample_bootstrap()
    tp.SHARED_BOOTSTRAP = False
    tp._rng = None
    tp._bootstrap_xi(B)
    b = tp.resample_bootstrap()
    assert np.array_equal(a[0], b[0])
