"""GPU: race stress tests in lieu of compute-sanitizer racecheck (not available on the pool).

The pair-binning kernel privatises histograms per warp (__syncwarp + shared atomics + flushes to global atomics),
takes work items from a global counter, and the Cholesky panel / triangular sweeps synchronise CTAs through words in
global memory.  A race in any of those shows up as run-to-run differences, so every case below repeats one launch many
times on the same input and demands BIT-IDENTICAL integer counts (pair binning) / factors and solutions (dense)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _catalogue(n, weighted, seed=3):
    from treegp_b200 import backend

    rng = np.random.default_rng(seed)
    L = 300.0
    x, y = backend.to_device(rng.uniform(0, L, n)), backend.to_device(rng.uniform(0, L, n))
    k = backend.to_device(rng.normal(size=n))
    w = backend.to_device(rng.uniform(0.5, 2.0, n)) if weighted else None
    o = backend.hilbert_order(x, y)
    x, y, k = x[o].contiguous(), y[o].contiguous(), k[o].contiguous()
    w = None if w is None else w[o].contiguous()
    return x, y, k, w, L


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("nranks", [1, 3])
def test_pairbin_repeated_counts_are_identical(gpu_ready, weighted, nranks):
    """200 repetitions of the N = 1e5 count (default max_sep: every kernel path is taken), alone and as the sum of 3
    "ranks" (tile_rank / tile_nranks): int64 counts identical every time, and identical between 1 and 3 ranks."""
    from treegp_b200 import _cabi, backend, binning

    n = 100_000
    x, y, k, w, L = _catalogue(n, weighted)
    mx = np.sqrt(2.0) * L / 2.0
    edges = backend.to_device(binning.twod_thresholds(mx, 21))
    off = backend.to_device(np.array([0, n]), torch.int64)

    def count():
        tot = None
        for r in range(nranks):
            c = backend.pairbin(x, y, k, w, off, n, _cabi.BIN_TWOD, edges, 21, 0.0, mx, rank=r, nranks=nranks)[0]
            tot = c.clone() if tot is None else tot + c
        return tot

    first = count()
    assert int(first.sum().item()) > 0
    single = backend.pairbin(x, y, k, w, off, n, _cabi.BIN_TWOD, edges, 21, 0.0, mx)[0]
    assert torch.equal(first, single)
    reps = 200 if nranks == 1 else 70
    for _ in range(reps):
        assert torch.equal(count(), first)


def test_pairbin_concurrent_streams(gpu_ready):
    """Two host threads launch tgp_pairbin on two streams at the same time (different catalogues): the work-counter
    ring and the per-stream scratch keep the launches apart -- both results equal their serial values, 30 times."""
    import threading
    from treegp_b200 import _cabi, backend, binning

    cats = [_catalogue(60_000, False, seed=s) for s in (5, 6)]
    mx = 80.0
    edges = backend.to_device(binning.twod_thresholds(mx, 15))
    offs = [backend.to_device(np.array([0, 60_000]), torch.int64) for _ in cats]

    def count(i):
        x, y, k, w, _ = cats[i]
        return backend.pairbin(x, y, k, w, offs[i], 60_000, _cabi.BIN_TWOD, edges, 15, 0.0, mx)[0].clone()

    serial = [count(0), count(1)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for _ in range(30):
        out = [None, None]

        def work(i):
            with torch.cuda.stream(streams[i]):
                out[i] = count(i)
            streams[i].synchronize()

        ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert torch.equal(out[0], serial[0]) and torch.equal(out[1], serial[1])


@pytest.mark.parametrize("n,reps", [(2500, 100), (14336, 30)])
def test_potrf_and_sweeps_are_bit_reproducible(gpu_ready, n, reps):
    """Repeated tgp_potrf (fused panel kernel with inter-CTA flags; at N = 14336 the first 1024-wide look-ahead block
    on a second stream) and tgp_potrs_vec (persistent sweeps, payload polling between CTAs): bit-identical factors
    and solutions every time, no device error word."""
    from treegp_b200 import _cabi, backend

    g = torch.Generator(device="cuda").manual_seed(n)
    A = torch.randn((n, 96), dtype=torch.float64, device="cuda", generator=g)
    K = backend.alloc_matrix(n, n)
    K[:, :n] = A @ A.T
    K[:, :n].diagonal().add_(float(n) * 0.01)
    b = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    ws = K.clone()
    assert int(backend.potrf(ws, n).item()) == 0
    x = backend.potrs_vec(ws, n, b.clone())
    ref_l = torch.tril(ws[:, :n]).clone()
    work = torch.empty_like(K)
    for _ in range(reps):
        work.copy_(K)
        assert int(backend.potrf(work, n).item()) == 0
        assert torch.equal(torch.tril(work[:, :n]), ref_l)
        assert torch.equal(backend.potrs_vec(work, n, b.clone()), x)
    assert _cabi.load().tgp_device_error(0) == 0
    resid = K[:, :n] @ x - b      # the strict upper triangle of K is intact in the pristine copy
    assert float(resid.abs().max()) < 1e-9 * float(b.abs().max()) * n
