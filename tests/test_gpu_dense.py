"""GPU parity: FP64 Cholesky / solves / log-likelihood (csrc/dense.cu) through the C ABI.

Tolerances: north_star asks 1e-6 relative on the log-likelihood and predictions; we assert much
tighter on well-conditioned inputs (stated per test)."""
import numpy as np
import pytest
import scipy.linalg as sla

from kspec import parse, vk_sym_quirk
from oracle import gp_oracle as go

pytestmark = pytest.mark.gpu


def _spd(n, seed, cond=1e3):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(n, n))
    Q, _ = np.linalg.qr(A)
    ev = np.geomspace(1.0, cond, n)
    return (Q * ev) @ Q.T


def _to_ws(A):
    from treegp_b200 import backend

    n = A.shape[0]
    ws = backend.alloc_matrix(n, n)
    ws.zero_()
    ws[:, :n] = backend.to_device(A)
    return ws


@pytest.mark.parametrize("n", [1, 2, 31, 63, 64, 65, 127, 200, 511, 512, 513, 777, 1100, 2500])
def test_potrf_matches_lapack(gpu_ready, n):
    from treegp_b200 import backend

    A = _spd(n, n)
    ws = _to_ws(A)
    info = backend.potrf(ws, n)
    assert int(info.item()) == 0
    L = np.tril(ws[:, :n].cpu().numpy())
    Lref = sla.cholesky(A, lower=True)
    np.testing.assert_allclose(L, Lref, rtol=0, atol=1e-11 * np.abs(Lref).max())
    np.testing.assert_allclose(L @ L.T, A, rtol=0, atol=1e-12 * np.abs(A).max() * n)


def test_potrf_leaves_upper_triangle_untouched(gpu_ready):
    import torch
    from treegp_b200 import backend

    n = 700
    A = _spd(n, 5)
    ws = _to_ws(A)
    iu = torch.triu_indices(n, n, 1, device="cuda")
    ws[iu[0], iu[1]] = 12345.0
    backend.potrf(ws, n)
    assert bool((ws[iu[0], iu[1]] == 12345.0).all())


@pytest.mark.parametrize("n,bad", [(100, 37), (900, 600), (900, 10)])
def test_potrf_reports_first_non_pd_minor(gpu_ready, n, bad):
    from treegp_b200 import backend

    A = _spd(n, 11)
    A[bad, bad] = -1.0  # leading minor of order bad+1 is not PD
    ws = _to_ws(A)
    info = backend.potrf(ws, n)
    assert int(info.item()) == bad + 1  # LAPACK convention


@pytest.mark.parametrize("n", [1, 5, 64, 255, 256, 257, 1000, 2311])
def test_potrs_vec(gpu_ready, n):
    from treegp_b200 import backend

    A = _spd(n, 100 + n)
    b = np.random.default_rng(n).normal(size=n)
    ws = _to_ws(A)
    backend.potrf(ws, n)
    x = backend.potrs_vec(ws, n, backend.to_device(b).clone()).cpu().numpy()
    xref = sla.cho_solve(sla.cho_factor(A, lower=True), b)
    np.testing.assert_allclose(x, xref, rtol=1e-9, atol=1e-10 * np.abs(xref).max())


@pytest.mark.parametrize("n", [9601, 20011])
def test_potrs_vec_persistent_sweep_many_slabs_per_cta(gpu_ready, n):
    """csrc/trsv.cu with more 64-row slabs than CTAs (every CTA owns several, ring stages wrap many times), a short
    last block, against the residual of the factor itself: || L L^T x - b || small, and bit-identical repeats
    (fixed accumulation order)."""
    import torch
    from treegp_b200 import _cabi, backend

    g = torch.Generator(device="cuda").manual_seed(n)
    ws = backend.alloc_matrix(n, n)
    ws.normal_(generator=g)
    ws.mul_(0.5 / np.sqrt(n))
    ws[:, :n].diagonal().copy_(1.0 + torch.rand(n, dtype=torch.float64, device="cuda", generator=g))
    b = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    x = backend.potrs_vec(ws, n, b.clone())
    Lm = torch.tril(ws[:, :n])
    resid = Lm @ (Lm.T @ x) - b
    assert float(resid.abs().max()) <= 1e-11 * float(b.abs().max()) * np.sqrt(n)
    for _ in range(3):
        assert torch.equal(backend.potrs_vec(ws, n, b.clone()), x)
    assert _cabi.load().tgp_device_error(0) == 0


def test_potrs_vec_unaligned_matrix(gpu_ready):
    """Odd leading dimension / base address: the sweep falls back to plain loads and still solves."""
    import ctypes
    import torch
    from treegp_b200 import _cabi, backend

    n, ld = 777, 779
    A = _spd(n, 5)
    Lh = np.linalg.cholesky(A)
    buf = torch.zeros(n * ld + 1, dtype=torch.float64, device="cuda")
    view = buf[1:].reshape(n, ld)                 # 8-byte aligned base, odd pitch
    view[:, :n] = backend.to_device(Lh)
    b = np.random.default_rng(1).normal(size=n)
    bd = backend.to_device(b).clone()
    _cabi.check(_cabi.load().tgp_potrs_vec(ctypes.c_void_p(view.data_ptr()), n, ld, ctypes.c_void_p(bd.data_ptr()),
                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "tgp_potrs_vec")
    xref = sla.cho_solve((Lh, True), b)
    np.testing.assert_allclose(bd.cpu().numpy(), xref, rtol=1e-9, atol=1e-10 * np.abs(xref).max())


@pytest.mark.parametrize("m,nc,kd", [(1, 1, 1), (128, 128, 16), (130, 70, 33), (300, 300, 64), (257, 129, 512), (64, 500, 7)])
def test_gemm_nt_sub(gpu_ready, m, nc, kd):
    from treegp_b200 import backend

    rng = np.random.default_rng(m * 31 + nc)
    C, A, B = rng.normal(size=(m, nc)), rng.normal(size=(m, kd)), rng.normal(size=(nc, kd))

    def ws(a):
        t = backend.alloc_matrix(a.shape[0], a.shape[1])
        t.zero_()
        t[:, : a.shape[1]] = backend.to_device(a)
        return t

    Cd, Ad, Bd = ws(C), ws(A), ws(B)
    backend.gemm_nt_sub(Cd, m, nc, Ad, Bd, kd)
    np.testing.assert_allclose(Cd[:, :nc].cpu().numpy(), C - A @ B.T, rtol=0, atol=1e-12 * kd)
    if m == nc:
        Cd = ws(C)
        backend.gemm_nt_sub(Cd, m, nc, Ad, Bd, kd, lower_only=True)
        got = Cd[:, :nc].cpu().numpy()
        ref = C - A @ B.T
        il, iu = np.tril_indices(m), np.triu_indices(m, 1)
        np.testing.assert_allclose(got[il], ref[il], rtol=0, atol=1e-12 * kd)
        np.testing.assert_array_equal(got[iu], C[iu])


@pytest.mark.parametrize("n,m", [(64, 3), (200, 129), (700, 64), (1300, 300)])
def test_trsm_rows(gpu_ready, n, m):
    from treegp_b200 import backend

    A = _spd(n, 7 * n)
    Bm = np.random.default_rng(m).normal(size=(m, n))
    ws = _to_ws(A)
    backend.potrf(ws, n)
    Bd = backend.alloc_matrix(m, n)
    Bd.zero_()
    Bd[:, :n] = backend.to_device(Bm)
    backend.trsm_rows(ws, n, Bd, m)
    L = sla.cholesky(A, lower=True)
    ref = sla.solve_triangular(L, Bm.T, lower=True).T
    np.testing.assert_allclose(Bd[:, :n].cpu().numpy(), ref, rtol=0, atol=1e-10 * np.abs(ref).max())


@pytest.mark.parametrize("name", ["arbf_a", "avk_c", "rbf2", "vk2s", "matern32"])
@pytest.mark.parametrize("want_alpha", [False, True])
def test_loglike_matches_reference_golden(gpu_ready, golden, name, want_alpha):
    """log_likelihood.py:21-41 through tgp_loglike; tolerance 1e-10 relative (north_star: 1e-6)."""
    from treegp_b200 import backend, eval_kernel
    from treegp_b200.kernels import lower_kernel

    X, y, yerr = golden["X2"], golden["y_" + name], golden["yerr_" + name]
    desc = lower_kernel(eval_kernel(str(golden["kstr_" + name])), 2)
    resid = backend.to_device(y - np.mean(y))
    out, info, alpha, _ = backend.loglike(X, resid, backend.to_device(yerr ** 2), desc, want_alpha=want_alpha)
    assert int(info.item()) == 0
    np.testing.assert_allclose(float(out[0].item()), float(golden["logL_" + name]), rtol=1e-10)
    if want_alpha:
        np.testing.assert_allclose(alpha.cpu().numpy(), golden["alpha_" + name], rtol=1e-6, atol=1e-8)


def test_loglike_not_pd_is_minus_inf(gpu_ready, golden):
    """log_likelihood.py:38-39: any failure -> -inf."""
    from treegp_b200 import backend, eval_kernel
    from treegp_b200.kernels import lower_kernel

    X = golden["X2"]
    desc = lower_kernel(eval_kernel("1.0 * RBF(50.0)"), 2)
    n = len(X)
    out, info, _, _ = backend.loglike(X, backend.to_device(np.ones(n)), backend.to_device(np.zeros(n)), desc)
    assert int(info.item()) > 0
    assert float(out[0].item()) == -np.inf


def test_loglike_n2000_vs_oracle(gpu_ready):
    from treegp_b200 import backend, eval_kernel
    from treegp_b200.kernels import lower_kernel

    rng = np.random.default_rng(42)
    n = 2000
    X = rng.uniform(-20, 20, size=(n, 2))
    y = rng.normal(size=n)
    yerr = np.full(n, 0.1)
    s = "2.0 * AnisotropicRBF(invLam=array([[1.5, 0.3], [0.3, 0.9]]))"
    desc = lower_kernel(eval_kernel(s), 2)
    out, info, _, _ = backend.loglike(X, backend.to_device(y), backend.to_device(yerr ** 2), desc)
    K = go.kmat("rbf", X, amp=2.0, invLam=np.array([[1.5, 0.3], [0.3, 0.9]])) + np.diag(yerr ** 2)
    ref, _ = go.log_likelihood(K, y)
    assert int(info.item()) == 0
    np.testing.assert_allclose(float(out[0].item()), ref, rtol=1e-10)
