"""GPU parity: pair binning (csrc/pairbin.cu) through the C ABI vs the brute-force oracle.

Counts must be bit-exact; FP64 sums agree up to summation order (stated tolerance: 1e-12 relative to
the sum of |contributions| in the bin)."""
import numpy as np
import pytest

from oracle import pairbin_oracle as po

pytestmark = pytest.mark.gpu


def _gpu_pairbin(x, y, k, w, min_sep, max_sep, nbins, bin_type, offsets=None, nranks=1, hilbert=False):
    import torch
    from treegp_b200 import _cabi, backend, binning

    if hilbert:  # spatially sorted input exercises the register-accumulation path of the kernel
        assert offsets is None
        order = backend.hilbert_order(backend.to_device(x), backend.to_device(y)).cpu().numpy()
        x, y, k = x[order], y[order], k[order]
        w = None if w is None else w[order]

    bt = _cabi.BIN_TWOD if bin_type == "TwoD" else _cabi.BIN_LOG
    edges = binning.twod_thresholds(max_sep, nbins) if bin_type == "TwoD" else binning.log_thresholds(min_sep, max_sep, nbins)
    if offsets is None:
        offsets = np.array([0, len(x)], dtype=np.int64)
    maxlen = int(np.diff(offsets).max())
    args = (backend.to_device(x), backend.to_device(y), backend.to_device(k),
            None if w is None else backend.to_device(w), backend.to_device(offsets, torch.int64), maxlen, bt,
            backend.to_device(edges), nbins, min_sep, max_sep)
    tot = None
    for r in range(nranks):
        res = backend.pairbin(*args, rank=r, nranks=nranks)
        res = [None if t is None else t.cpu().numpy() for t in res]
        tot = res if tot is None else [None if a is None else a + b for a, b in zip(tot, res)]
    return tot


def _kk_atol(sumwkk, weight):
    """Tolerance on sum w k k' per bin: 1e-11 of the largest bin, plus the rounding floor of a sum whose terms cancel --
    the FP64 additions happen in a different order on every launch (atomics) and in the oracle, and a bin of n pairs
    with |k| ~ 1 carries partial sums of size sqrt(n): a few 1e-15 of sum |w| (seen: 2.5e-8 on -2446 over 4e8 pairs)."""
    return 1e-11 * max(1.0, np.abs(sumwkk).max()) + 4e-15 * np.abs(weight).max()


def _check(res, ref, c=0):
    npairs, sumw, sumwkk, sumwr = res
    np.testing.assert_array_equal(npairs[c], ref["npairs"])
    scale = max(1.0, np.abs(ref["weight"]).max())
    np.testing.assert_allclose(sumw[c], ref["weight"], rtol=1e-12, atol=1e-12 * scale)
    np.testing.assert_allclose(sumwkk[c], ref["sumwkk"], rtol=0, atol=_kk_atol(ref["sumwkk"], ref["weight"]))
    if sumwr is not None:
        np.testing.assert_allclose(sumwr[c], ref["sumwr"], rtol=1e-12, atol=1e-12 * scale)


@pytest.mark.parametrize("n", [2, 255, 256, 257, 1000, 3001])
@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("cfg", [("TwoD", 0.0, 14.142135623730951, 21), ("TwoD", 0.5, 3.0, 20), ("Log", 0.1, 1.75, 15),
                                 ("Log", 0.3, 12.0, 20)])
def test_pairbin_matches_oracle(gpu_ready, n, weighted, cfg):
    bin_type, mn, mx, nb = cfg
    rng = np.random.default_rng(n + 17 * nb)
    x, y = rng.uniform(-10, 10, n), rng.uniform(-10, 10, n)
    k = rng.normal(size=n)
    w = rng.uniform(0.5, 2.0, n) if weighted else None
    if n > 10:  # coincident points must be skipped (r2 == 0), as bootstrap duplicates are
        x[7], y[7] = x[3], y[3]
    ref = po.pairbin(x, y, k, w, mn, mx, nb, bin_type)
    _check(_gpu_pairbin(x, y, k, w, mn, mx, nb, bin_type), ref)
    _check(_gpu_pairbin(x, y, k, w, mn, mx, nb, bin_type, hilbert=True), ref)


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("n,mx,nb,mn", [(30000, 70.0, 21, 0.0), (30000, 12.0, 6, 0.0), (50000, 40.0, 9, 3.0),
                                        (20011, 200.0, 1, 0.0), (40000, 25.0, 2, 0.0)])
def test_pairbin_register_path_matches_oracle(gpu_ready, weighted, n, mx, nb, mn):
    """Hilbert-sorted input at sizes where 32 x 32 pair blocks fit a 2 x 2 bin window: the
    register-accumulation path (full-in, partial and out-of-range blocks) must stay bit-exact."""
    rng = np.random.default_rng(n + nb)
    x, y = rng.uniform(-50, 50, n), rng.uniform(-50, 50, n)
    k = rng.normal(size=n)
    w = rng.uniform(0.5, 2.0, n) if weighted else None
    x[100], y[100] = x[5000], y[5000]  # coincident points far apart in memory
    x[7], y[7] = x[8], y[8]            # and adjacent
    ref = po.pairbin(x, y, k, w, mn, mx, nb, "TwoD")
    _check(_gpu_pairbin(x, y, k, w, mn, mx, nb, "TwoD", hilbert=True), ref)


def test_pairbin_values_on_bin_edges(gpu_ready):
    """Lattice points put many displacements exactly on bin edges: the threshold rule must agree with
    the floor((d+max_sep)/bin_size) rule bit for bit."""
    g = np.arange(-10, 11, dtype=np.float64)
    X, Y = np.meshgrid(g, g)
    x, y = X.ravel() * 0.5, Y.ravel() * 0.5
    k = np.cos(x) * np.sin(y)
    for mx, nb in ((5.0, 20), (5.25, 21), (4.0, 16)):
        ref = po.pairbin(x, y, k, None, 0.0, mx, nb, "TwoD")
        _check(_gpu_pairbin(x, y, k, None, 0.0, mx, nb, "TwoD"), ref)
        _check(_gpu_pairbin(x, y, k, None, 0.0, mx, nb, "TwoD", hilbert=True), ref)
    # a dense lattice: many points, displacements exactly on edges, register path
    g = np.arange(-60, 61, dtype=np.float64)
    X, Y = np.meshgrid(g, g)
    xl, yl = X.ravel() * 0.25, Y.ravel() * 0.25
    kl = np.cos(xl) * np.sin(yl)
    for mx, nb in ((5.0, 20), (5.25, 21), (16.0, 4)):
        ref = po.pairbin(xl, yl, kl, None, 0.0, mx, nb, "TwoD")
        _check(_gpu_pairbin(xl, yl, kl, None, 0.0, mx, nb, "TwoD", hilbert=True), ref)
    ref = po.pairbin(x, y, k, None, 0.5, 8.0, 16, "Log")
    _check(_gpu_pairbin(x, y, k, None, 0.5, 8.0, 16, "Log"), ref)


def test_pairbin_twod_is_point_symmetric(gpu_ready):
    rng = np.random.default_rng(5)
    n, nb = 4000, 21
    x, y, k = rng.uniform(0, 30, n), rng.uniform(0, 30, n), rng.normal(size=n)
    npairs, sumw, sumwkk, _ = _gpu_pairbin(x, y, k, None, 0.0, 10.0, nb, "TwoD")
    c = npairs[0].reshape(nb, nb)
    np.testing.assert_array_equal(c, c[::-1, ::-1])
    assert c.sum() % 2 == 0


def test_pairbin_batched_catalogues_and_rank_sharding(gpu_ready):
    """A bootstrap batch (ragged catalogues) in one launch, and the multi-GPU split: the per-rank
    partial results must add up to the single-rank answer (counts exactly)."""
    rng = np.random.default_rng(8)
    sizes = [700, 1, 513, 0, 1290]
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = offsets[-1]
    x, y, k, w = rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.normal(size=n), rng.uniform(0.5, 2, n)
    for bin_type, mn, mx, nb in (("TwoD", 0.0, 9.0, 21), ("Log", 0.2, 9.0, 12)):
        one = _gpu_pairbin(x, y, k, w, mn, mx, nb, bin_type, offsets)
        three = _gpu_pairbin(x, y, k, w, mn, mx, nb, bin_type, offsets, nranks=3)
        for c, sz in enumerate(sizes):
            s = slice(offsets[c], offsets[c + 1])
            ref = po.pairbin(x[s], y[s], k[s], w[s], mn, mx, nb, bin_type)
            _check(one, ref, c)
            _check(three, ref, c)


def test_pairbin_n20000_vs_oracle(gpu_ready):
    rng = np.random.default_rng(2)
    n, nb = 20000, 21
    x, y, k = rng.uniform(-50, 50, n), rng.uniform(-50, 50, n), rng.normal(size=n)
    mx = np.sqrt(2) * 100 / 2
    ref = po.pairbin(x, y, k, None, 0.0, mx, nb, "TwoD")
    res = _gpu_pairbin(x, y, k, None, 0.0, mx, nb, "TwoD")
    _check(res, ref)
    assert res[0].sum() == ref["npairs"].sum()
    _check(_gpu_pairbin(x, y, k, None, 0.0, mx, nb, "TwoD", hilbert=True), ref)


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("kind", ["uniform", "lattice"])
def test_pairbin_block_forms_equal_pair_by_pair(gpu_ready, weighted, kind):
    """The block forms (whole block booked from chunk sums; one-axis blocks answered by a rank query on the sorted
    chunk) against the same kernel with them switched off (every pair evaluated individually), at a size where
    most blocks take them.  Counts identical; sums to summation order.  The lattice puts thousands of
    displacements exactly on bin edges (spacing 0.5, bin width 8), which is where the exact mirrored-bin check of
    the rank query has to hand blocks back to the pair-by-pair path."""
    from treegp_b200 import backend

    rng = np.random.default_rng(11)
    if kind == "uniform":
        n = 150000
        x, y = rng.uniform(0, 400, n), rng.uniform(0, 400, n)
        mx, nb = 280.0, 21
    else:
        g = np.arange(0, 300, dtype=np.float64) * 0.5
        X, Y = np.meshgrid(g, g)
        x, y = X.ravel(), Y.ravel()
        n = len(x)
        mx, nb = 64.0, 16
    k = rng.normal(size=n)
    w = rng.uniform(0.5, 2.0, n) if weighted else None
    backend.pairbin_stats(reset=True)
    fast = _gpu_pairbin(x, y, k, w, 0.0, mx, nb, "TwoD", hilbert=True)
    stats = backend.pairbin_stats(reset=True)
    assert stats["closed_form"] > 0
    if kind == "uniform":
        assert stats["one_axis_sorted"] > 0 and stats["two_axis_sorted"] > 0
    else:   # every chunk of a lattice has columns on a bin edge: the rank query hands its blocks back
        assert stats["one_axis"] > 0
    # the same block forms through the general dispatch only (no short cuts): identical counts
    backend.set_option("pairbin_fast_paths", 0)
    try:
        general = _gpu_pairbin(x, y, k, w, 0.0, mx, nb, "TwoD", hilbert=True)
        stg = backend.pairbin_stats(reset=True)
    finally:
        backend.set_option("pairbin_fast_paths", 7)
    assert stg["two_axis_sorted"] == 0 and stg["closed_form"] == stats["closed_form"]
    np.testing.assert_array_equal(fast[0], general[0])
    np.testing.assert_allclose(fast[2], general[2], rtol=0, atol=_kk_atol(general[2], general[1]))
    backend.set_option("pairbin_block_sums", 0)
    try:
        slow = _gpu_pairbin(x, y, k, w, 0.0, mx, nb, "TwoD", hilbert=True)
        st0 = backend.pairbin_stats(reset=True)
        # ... and the pair-by-pair kernel with the mirrored bits evaluated for every pair (no per-block check)
        backend.set_option("pairbin_fast_paths", 0)
        slow_pp = _gpu_pairbin(x, y, k, w, 0.0, mx, nb, "TwoD", hilbert=True)
        backend.pairbin_stats(reset=True)
    finally:
        backend.set_option("pairbin_block_sums", 1)
        backend.set_option("pairbin_fast_paths", 7)
    np.testing.assert_array_equal(slow[0], slow_pp[0])
    np.testing.assert_allclose(slow[2], slow_pp[2], rtol=0, atol=_kk_atol(slow_pp[2], slow_pp[1]))
    assert st0["closed_form"] == 0 and st0["one_axis_sorted"] == 0 and st0["one_axis"] == 0
    assert st0["two_axis_sorted"] == 0
    np.testing.assert_array_equal(fast[0], slow[0])
    np.testing.assert_allclose(fast[1], slow[1], rtol=1e-12, atol=1e-12 * np.abs(slow[1]).max())
    np.testing.assert_allclose(fast[2], slow[2], rtol=0, atol=_kk_atol(slow[2], slow[1]))
    if kind == "uniform":   # point symmetric (on the lattice displacements sit ON bin edges, where the formula is not)
        c = fast[0][0].reshape(nb, nb)
        np.testing.assert_array_equal(c, c[::-1, ::-1])


@pytest.mark.parametrize("weighted", [False, True])
def test_pairbin_log_one_bin_blocks_match_oracle(gpu_ready, weighted):
    """Log (isotropic) bins on Hilbert-sorted input: blocks whose smallest and largest r^2 share a radial bin are
    booked from the row / chunk sums, with only sum w r evaluated pair by pair (two_pcf.py:330-338: npairs, xi and
    meanr of the TreeCorr Log binning).  Counts bit-exact, sums incl. sum w r to summation order."""
    from treegp_b200 import backend

    rng = np.random.default_rng(21)
    n = 30000
    x, y = rng.uniform(0, 100, n), rng.uniform(0, 100, n)
    k = rng.normal(size=n)
    w = rng.uniform(0.5, 2.0, n) if weighted else None
    for mn, mx, nb in ((0.5, 70.0, 12), (2.0, 40.0, 5)):
        ref = po.pairbin(x, y, k, w, mn, mx, nb, "Log")
        backend.pairbin_stats(reset=True)
        _check(_gpu_pairbin(x, y, k, w, mn, mx, nb, "Log", hilbert=True), ref)
        assert backend.pairbin_stats(reset=True)["closed_form"] > 0
        backend.set_option("pairbin_block_sums", 0)
        try:
            _check(_gpu_pairbin(x, y, k, w, mn, mx, nb, "Log", hilbert=True), ref)
        finally:
            backend.set_option("pairbin_block_sums", 1)


def test_hilbert_keys_auto_equal_keys_from_host_bounds(gpu_ready):
    """tgp_hilbert_keys_auto (bounding square found on the device, no read-back) gives exactly the keys of
    tgp_hilbert_keys with the extrema taken on the host."""
    import ctypes
    import torch
    from treegp_b200 import _cabi, backend

    rng = np.random.default_rng(11)
    for n in (1, 2, 1000, 300_001):
        x = backend.to_device(rng.uniform(-7.0, 13.0, n))
        y = backend.to_device(rng.uniform(100.0, 103.0, n))
        xmin, xmax, ymin, ymax = float(x.min()), float(x.max()), float(y.min()), float(y.max())
        extent = max(xmax - xmin, ymax - ymin)
        extent = extent * (1.0 + 1e-9) if extent > 0 else 1.0
        order = int(min(16, max(1, np.ceil(np.log2(max(np.sqrt(n), 2.0))) + 1)))
        k1 = torch.empty(n, dtype=torch.int64, device=x.device)
        k2 = torch.empty(n, dtype=torch.int64, device=x.device)
        scratch = torch.empty(4, dtype=torch.int64, device=x.device)
        lib = _cabi.load()
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _cabi.check(lib.tgp_hilbert_keys(p(x), p(y), n, xmin, ymin, extent, order, p(k1), st), "keys")
        _cabi.check(lib.tgp_hilbert_keys_auto(p(x), p(y), n, order, p(scratch), p(k2), st), "keys_auto")
        assert torch.equal(k1, k2)


def test_pairbin_block_forms_match_oracle_at_benchmark_density(gpu_ready):
    """N = 300 000 points at the benchmark's density and binning (field 548, nbins = 21, default max_sep = half the
    diagonal): the closed-form / rank-query / two-query block forms and their short cuts against the INDEPENDENT
    checker (the OpenMP oracle, 4.5e10 pairs, ~25 s on 16 cores) -- at this size most pairs no longer go through the
    pair-by-pair loop, so this is the comparison that exercises the block forms the way bench.py does."""
    import os

    os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count()))
    n = 300_000
    L = 1000.0 * np.sqrt(n / 1e6)
    rng = np.random.default_rng(2026)
    x, y = rng.uniform(-L / 2, L / 2, n), rng.uniform(-L / 2, L / 2, n)
    k = rng.normal(size=n)
    mx = np.sqrt(2.0) * L / 2.0
    from treegp_b200 import backend

    ref = po.pairbin(x, y, k, None, 0.0, mx, 21, "TwoD")
    backend.pairbin_stats(reset=True)
    res = _gpu_pairbin(x, y, k, None, 0.0, mx, 21, "TwoD", hilbert=True)
    _check(res, ref)
    st = backend.pairbin_stats(reset=True)
    tot = max(1, sum(st.values()))
    assert st["closed_form"] / tot > 0.3 and (st["one_axis_sorted"] + st["two_axis_sorted"]) / tot > 0.2
