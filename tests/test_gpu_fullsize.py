"""GPU, BASELINE.json's full sizes: properties that do not need the (too slow) CPU oracle.

* 2PCF at N = 1e6 (configs[3]): the register path (Hilbert-sorted input) and the generic shared-atomic path
  (unsorted input) are different code and must give identical counts; counts are point symmetric; per-rank
  partial results add up exactly.
* GP at N = 40,000 (configs[2]): the solve is checked through the defining equation (K + diag(s^2)) alpha = y,
  with K applied by the fused predict kernel (a code path independent of the Cholesky), and the two
  log-likelihood variants (forward sweep only / both sweeps) agree.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_2pcf_full_size_paths_agree_and_shard_exactly(gpu_ready):
    import torch
    from treegp_b200 import _cabi, backend, binning

    n, nb, L = 1_000_000, 21, 1000.0
    rng = np.random.default_rng(42)
    x = backend.to_device(rng.uniform(-L / 2, L / 2, n))
    y = backend.to_device(rng.uniform(-L / 2, L / 2, n))
    k = backend.to_device(rng.normal(size=n))
    off = backend.to_device(np.array([0, n]), torch.int64)
    mx = np.sqrt(2.0) * L / 2.0
    edges = backend.to_device(binning.twod_thresholds(mx, nb))

    def run(px, py, pk, rank=0, nranks=1):
        return backend.pairbin(px, py, pk, None, off, n, _cabi.BIN_TWOD, edges, nb, 0.0, mx, rank=rank, nranks=nranks)

    order = backend.hilbert_order(x, y)
    xs, ys, ks = x[order].contiguous(), y[order].contiguous(), k[order].contiguous()
    c_reg, w_reg, s_reg, _ = run(xs, ys, ks)          # register path
    c_gen, w_gen, s_gen, _ = run(x, y, k)             # generic path (unsorted input)
    assert torch.equal(c_reg, c_gen)
    grid = c_reg[0].reshape(nb, nb)
    assert torch.equal(grid, torch.flip(grid, dims=(0, 1)))            # point symmetry
    assert int(c_reg.sum().item()) % 2 == 0 and int(c_reg.sum().item()) > 0.8 * n * (n - 1)
    assert torch.equal(w_reg, c_reg.to(torch.float64))                 # unit weights
    scale = float(s_reg.abs().max().item())
    assert float((s_reg - s_gen).abs().max().item()) <= 1e-9 * scale   # same sums up to summation order
    # rank sharding: partial results add up to the single-launch answer (counts exactly)
    parts = [run(xs, ys, ks, rank=r, nranks=3) for r in range(3)]
    assert torch.equal(sum(p[0] for p in parts), c_reg)
    assert float((sum(p[2] for p in parts) - s_reg).abs().max().item()) <= 1e-10 * scale
    assert all(int(p[0].sum().item()) > 0.25 * int(c_reg.sum().item()) for p in parts)  # balanced split


def test_gp_solve_full_size_satisfies_its_defining_equation(gpu_ready):
    import torch
    from treegp_b200 import backend, eval_kernel
    from treegp_b200.kernels import lower_kernel

    n = 40_000
    rng = np.random.default_rng(7)
    Lf = 160.0
    X = backend.as_points(rng.uniform(-Lf / 2, Lf / 2, size=(n, 2)))
    y = backend.to_device(rng.normal(size=n))
    noise2 = backend.to_device(np.full(n, 0.05 ** 2))
    desc = lower_kernel(eval_kernel("4.0 * AnisotropicVonKarman(invLam=array([[0.46, -0.09], [-0.09, 0.55]]))"), 2)
    work = backend.alloc_matrix(n + 1, n)
    out_a, info_a, alpha, _ = backend.loglike(X, y, noise2, desc, work=work, want_alpha=True)
    assert int(info_a.item()) == 0
    # (K + diag(s^2)) alpha = y, with K alpha from the fused predict kernel
    resid = backend.predict_mean(X, X, desc, alpha) + noise2 * alpha - y
    assert float(resid.abs().max().item()) <= 1e-8 * float(y.abs().max().item())
    out_b, info_b, _, _ = backend.loglike(X, y, noise2, desc, work=work, want_alpha=False)
    assert int(info_b.item()) == 0
    np.testing.assert_allclose(float(out_b[0].item()), float(out_a[0].item()), rtol=1e-12)
    np.testing.assert_allclose(float(out_b[1].item()), float(out_a[1].item()), rtol=1e-9)   # chi2 two ways
    np.testing.assert_allclose(float(out_b[2].item()), float(out_a[2].item()), rtol=0, atol=0)  # same factor


def test_envelope_solve_full_size_satisfies_its_defining_equation(gpu_ready):
    """configs[2] through the envelope factorisation (tgp_loglike_env, what GPInterpolation runs at this size): the
    solve satisfies (K + diag(s^2)) alpha = y with K applied by the fused predict kernel, and logL / chi2 / log-det
    equal the dense evaluation of the same sorted system."""
    from treegp_b200 import backend, eval_kernel
    from treegp_b200.kernels import lower_kernel

    n = 40_000
    rng = np.random.default_rng(7)
    Lf = 160.0
    X = backend.as_points(rng.uniform(-Lf / 2, Lf / 2, size=(n, 2)))
    y = backend.to_device(rng.normal(size=n))
    noise2 = backend.to_device(rng.uniform(0.03, 0.07, size=n) ** 2)
    desc = lower_kernel(eval_kernel("4.0 * AnisotropicVonKarman(invLam=array([[0.46, -0.09], [-0.09, 0.55]]))"), 2)
    plan = backend.plan_envelope(X, desc)
    assert plan is not None and plan["flops"] * 10 < plan["flops_dense"]
    o = plan["order"]
    Xs, ys, es = X[o].contiguous(), y[o].contiguous(), noise2[o].contiguous()
    work = backend.alloc_matrix(n + 1, n)
    out_e, info_e, alpha, _ = backend.loglike(Xs, ys, es, desc, work=work, want_alpha=True, row_end=plan["row_end"])
    assert int(info_e.item()) == 0
    resid = backend.predict_mean(Xs, Xs, desc, alpha, truncate=False) + es * alpha - ys
    assert float(resid.abs().max().item()) <= 1e-8 * float(ys.abs().max().item())
    out_f, info_f, _, _ = backend.loglike(Xs, ys, es, desc, work=work, want_alpha=False, row_end=plan["row_end"])
    np.testing.assert_allclose(float(out_f[0].item()), float(out_e[0].item()), rtol=1e-12)
    out_d, info_d, alpha_d, _ = backend.loglike(Xs, ys, es, desc, work=work, want_alpha=True)
    assert int(info_d.item()) == 0
    np.testing.assert_allclose(out_e.cpu().numpy(), out_d.cpu().numpy(), rtol=1e-11)
    assert float((alpha - alpha_d).abs().max().item()) <= 1e-9 * float(alpha_d.abs().max().item())
