#!/usr/bin/env python
"""Generate tests/golden/reference_vectors_r2.npz: outputs of the UNMODIFIED reference
(/root/reference, PFLeget/treegp v1.4.1) for every part of the hot path's HOST logic that runs in the build
container, on small seeded inputs.  Complements make_golden.py (K / logL / predict vectors).

What is pinned here ("What's missing" item 1):
  (i)   log_likelihood.optimizer (log_likelihood.py:43-62): theta-hat, _logL and the number of likelihood
        evaluations of the reference's own L-BFGS-B fits (1-D RBF and VonKarman at N = 100, 2-D AnisotropicRBF
        at N = 600);
  (ii)  get_correlation_length_matrix (two_pcf.py:12-31) and robust_2dfit.chi2 / alpha / residuals
        (two_pcf.py:115-148) on a fixed (xi, W, mask);
  (iii) utils.vcorr / xiB / comp_eb (utils.py:5-107) incl. the per-bin pair counts of its np.histogram binning,
        and meanify.meanify (meanify.py:49-137) for the 'mean' and 'median' statistics;
  (iv)  two_pcf.comp_2pcf mask / coordinates (two_pcf.py:283-340), comp_xi_covariance (:342-362), return_2pcf
        (:364-391) and the non-robust optimizer (:393-464), run UNMODIFIED on top of a generator-only fake
        `treecorr` module whose KKCorrelation.process is the brute-force oracle (oracle/pairbin_oracle.c).

What (iv) does and does not pin: everything the reference does AROUND the pair sums (weights, mean subtraction,
bootstrap index stream, resample bookkeeping, mask, coordinates, covariance, de-biasing, chi-square, the two
scipy minimisers) is the reference's own code; the pair-sum primitive inside `process` is the oracle's
restatement of TreeCorr, which stays PARITY UNPINNED (TreeCorr is not installable here).

Only runnable in the build container (the GPU box has no /root/reference).
Usage:  python tests/golden/make_golden_r2.py
"""
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
OUT = os.path.join(HERE, "reference_vectors_r2.npz")
sys.path.insert(0, ROOT)


def make_fake_treecorr():
    """Generator-only stand-in for the `treecorr` module: just the surface two_pcf.py touches
    (Catalog(x, y, k, w), KKCorrelation(min_sep, max_sep, nbins[, bin_type, bin_slop]).process(cat) and the
    attributes xi, meanr, bottom_edges, top_edges, npairs, weight), backed by the brute-force oracle."""
    from oracle import pairbin_oracle as po

    mod = types.ModuleType("treecorr")
    mod.__version__ = "0.0-oracle-backed-fake"
    mod.calls = []

    class Catalog(object):
        def __init__(self, x=None, y=None, k=None, w=None, **kw):
            self.x, self.y, self.k, self.w = x, y, k, w

    class KKCorrelation(object):
        def __init__(self, min_sep=None, max_sep=None, nbins=None, bin_type="Log", bin_slop=None, **kw):
            self.min_sep, self.max_sep, self.nbins, self.bin_type = min_sep, max_sep, nbins, bin_type

        def process(self, cat):
            res = po.pairbin(cat.x, cat.y, cat.k, cat.w, self.min_sep, self.max_sep, self.nbins, self.bin_type)
            nb = self.nbins
            if self.bin_type == "TwoD":
                shape = (nb, nb)
                edges = np.linspace(-self.max_sep, self.max_sep, nb + 1)
                # TreeCorr: left/right edges vary along axis 1 (dx), bottom/top edges along axis 0 (dy)
                self.left_edges, self.bottom_edges = np.meshgrid(edges[:-1], edges[:-1])
                self.right_edges, self.top_edges = np.meshgrid(edges[1:], edges[1:])
            else:
                shape = (nb,)
                bs = np.log(self.max_sep / self.min_sep) / nb
                self.rnom = np.exp(np.log(self.min_sep) + (np.arange(nb) + 0.5) * bs)
                meanr = res["meanr"].copy()
                meanr[res["weight"] == 0] = self.rnom[res["weight"] == 0]   # TreeCorr's finalize()
                self.meanr = meanr
            self.xi = res["xi"].reshape(shape)
            self.npairs = res["npairs"].astype(float).reshape(shape)
            self.weight = res["weight"].reshape(shape)
            mod.calls.append(self)

    mod.Catalog, mod.KKCorrelation = Catalog, KKCorrelation
    return mod


def import_reference():
    stub = tempfile.mkdtemp(prefix="tgp_stubs_")
    for name, body in (("fitsio", ""), ("iminuit", "__version__ = '2.0.0'\n")):
        with open(os.path.join(stub, name + ".py"), "w") as fh:
            fh.write(body)
    sys.modules["treecorr"] = make_fake_treecorr()
    sys.path.insert(0, stub)
    sys.path.insert(0, REF)
    import treegp  # noqa
    import scipy.linalg

    gpi = sys.modules["treegp.gp_interp"]

    def safe_cholesky(a, lower=False, overwrite_a=False, check_finite=True):
        return scipy.linalg.cholesky(a, lower=lower, overwrite_a=False, check_finite=check_finite)

    gpi.cholesky = safe_cholesky
    return treegp


def draw_field(rng, kernel, X, noise):
    K = kernel(X) + 1e-10 * np.eye(len(X))
    y = rng.multivariate_normal(np.zeros(len(X)), K)
    y_err = np.full(len(X), noise)
    return y + rng.normal(scale=noise, size=len(X)), y_err


def main():
    treegp = import_reference()
    treecorr = sys.modules["treecorr"]
    two_pcf_mod = sys.modules["treegp.two_pcf"]
    ll_mod = sys.modules["treegp.log_likelihood"]
    utils_mod = sys.modules["treegp.utils"]
    rng = np.random.default_rng(20261019)
    out = {}

    # ---------------------------------------------------------------- (i) L-BFGS-B likelihood fits
    fits = {
        "rbf1d": ("1.0**2 * RBF(0.5)", "0.7**2 * RBF(0.9)", 1, 100, 10.0),
        "vk1d": ("1.0**2 * VonKarman(length_scale=2.0)", "0.8**2 * VonKarman(length_scale=3.0)", 1, 100, 10.0),
        "arbf2d": (None, None, 2, 600, 12.0),
    }
    for name, (truth_s, start_s, ndim, n, half) in fits.items():
        if name == "arbf2d":
            Lt = two_pcf_mod.get_correlation_length_matrix(0.9, 0.2, -0.15)
            Ls = two_pcf_mod.get_correlation_length_matrix(1.3, 0.0, 0.0)
            truth_s = "2.0**2 * AnisotropicRBF(invLam={0!r})".format(np.linalg.inv(Lt))
            start_s = "1.5**2 * AnisotropicRBF(invLam={0!r})".format(np.linalg.inv(Ls))
        truth = treegp.eval_kernel(truth_s)
        X = rng.uniform(-half, half, size=(n, ndim))
        y, y_err = draw_field(rng, truth, X, 0.05)
        y = y - np.mean(y)
        ll = treegp.log_likelihood(X, y, y_err)
        count = [0]
        inner = ll.log_likelihood

        def counted(kernel, inner=inner, count=count):
            count[0] += 1
            return inner(kernel)

        ll.log_likelihood = counted
        start = treegp.eval_kernel(start_s)
        fitted = ll.optimizer(start)
        out["fit_%s_truth" % name] = np.array(truth_s)
        out["fit_%s_start" % name] = np.array(start_s)
        out["fit_%s_X" % name] = X
        out["fit_%s_y" % name] = y
        out["fit_%s_yerr" % name] = y_err
        out["fit_%s_theta0" % name] = np.array(start.theta)
        out["fit_%s_theta" % name] = np.array(fitted.theta)
        out["fit_%s_logL" % name] = np.array(ll._logL)
        out["fit_%s_logL0" % name] = np.array(inner(start))
        out["fit_%s_nevals" % name] = np.array(count[0])
        print("fit", name, "theta-hat", fitted.theta, "logL", ll._logL, "evals", count[0])

    # ---------------------------------------------------------------- (ii) correlation-length matrix
    clm_params = np.array([[0.5, 0.2, 0.2], [1.5, -0.3, 0.1], [3000.0, 0.0, 0.0], [2.0, 0.0, -0.6], [0.8, 0.95, 0.0]])
    out["clm_params"] = clm_params
    out["clm_values"] = np.array([two_pcf_mod.get_correlation_length_matrix(*p) for p in clm_params])

    # ---------------------------------------------------------------- (iv) two_pcf host logic on the fake treecorr
    Lt = two_pcf_mod.get_correlation_length_matrix(2.5, 0.25, 0.1)
    truth = treegp.eval_kernel("1.5**2 * AnisotropicRBF(invLam={0!r})".format(np.linalg.inv(Lt)))
    n = 500
    X = rng.uniform(-10, 10, size=(n, 2))
    y, y_err = draw_field(rng, truth, X, 0.1)
    y_err = y_err * rng.uniform(0.7, 1.3, size=n)
    out.update(pcf_X=X, pcf_y=y, pcf_yerr=y_err)

    # TwoD, explicit separations
    t = treegp.two_pcf(X, y, y_err, 0.0, 5.0, nbins=11, anisotropic=True)
    xi, dist, coord, mask = t.comp_2pcf(X, y, y_err)
    kk = treecorr.calls[-1]
    out.update(twod_xi=xi, twod_dist=dist, twod_coord=coord, twod_mask=mask,
               twod_npairs=kk.npairs.reshape(-1).astype(np.int64), twod_weight=kk.weight.reshape(-1))
    # unweighted branch (sum(y_err) == 0 -> w = None, two_pcf.py:291-294)
    xi0, _, _, _ = t.comp_2pcf(X, y, np.zeros(n))
    out["twod_xi_unweighted"] = xi0
    # first resamples of the seeded stream + covariance over 24 resamples
    t.seed, t._rng = 610639139, None
    u, v, yr, er = t.resample_bootstrap()
    out.update(boot_u0=u, boot_v0=v, boot_y0=yr, boot_e0=er)
    out["twod_cov24"] = t.comp_xi_covariance(n_bootstrap=24, mask=mask, seed=12345)
    out["twod_cov24_nomask"] = t.comp_xi_covariance(n_bootstrap=6, mask=None, seed=7)
    # return_2pcf: resample count from fsolve, de-biased inverse covariance
    xi_r, w_r, dist_r, coord_r, mask_r = t.return_2pcf()
    out.update(twod_ret_xi=xi_r, twod_ret_weight=w_r)
    ncalls = len(treecorr.calls)
    # non-robust optimizer (fmin + L-BFGS-B on the chi-square), default separations
    start = treegp.eval_kernel("1.4**2 * AnisotropicRBF(invLam={0!r})".format(
        np.linalg.inv(two_pcf_mod.get_correlation_length_matrix(2.2, 0.2, 0.05))))
    # explicit separations (bin width < correlation length): with the default max_sep = half the field diagonal the
    # reference's own non-robust TwoD fit runs away along a degenerate direction, which pins nothing
    t2 = treegp.two_pcf(X, y, y_err, 0.0, 5.0, nbins=9, anisotropic=True)
    fitted = t2.optimizer(start)
    out["twod_opt_start"] = np.array("1.4**2 * AnisotropicRBF(invLam={0!r})".format(
        np.linalg.inv(two_pcf_mod.get_correlation_length_matrix(2.2, 0.2, 0.05))))
    out.update(twod_opt_theta0=np.array(start.theta), twod_opt_theta=np.array(fitted.theta),
               twod_opt_min_sep=np.array(t2.min_sep), twod_opt_max_sep=np.array(t2.max_sep),
               twod_opt_xi=t2._2pcf, twod_opt_weight=t2._2pcf_weight, twod_opt_fit=t2._2pcf_fit,
               twod_opt_mask=t2._2pcf_mask, twod_opt_nresamples=np.array(len(treecorr.calls) - ncalls - 1))
    print("two_pcf TwoD fit theta-hat", fitted.theta, "resamples", len(treecorr.calls) - ncalls - 1)
    # default separations of the anisotropic path (two_pcf.py:404-424): min_sep = 0, max_sep = half the diagonal
    t2d = treegp.two_pcf(X, y, y_err, None, None, nbins=5, anisotropic=True)
    t2d.optimizer(start)
    out.update(twod_def_min_sep=np.array(t2d.min_sep), twod_def_max_sep=np.array(t2d.max_sep), twod_def_xi=t2d._2pcf,
               twod_def_weight=t2d._2pcf_weight)

    # robust_2dfit.chi2 on the measured map (two_pcf.py:115-148); the Minuit search itself needs iminuit
    rob = two_pcf_mod.robust_2dfit(start, t2._2pcf, t2._2pcf_dist[:, 0], t2._2pcf_dist[:, 1], t2._2pcf_weight,
                                   mask=t2._2pcf_mask)
    chi_params = np.array([[2.5, 0.25, 0.1], [1.8, 0.0, 0.0], [2.0, -0.3, 0.3], [0.6, 0.1, -0.2], [1.5, 1.2, 0.0],
                           [1.0, np.nan, 0.0]])
    chi_vals, chi_alpha, chi_res = [], [], []
    for p in chi_params:
        c = rob.chi2(p)
        chi_vals.append(c)
        if np.isfinite(c):
            chi_alpha.append(np.array(rob.alpha).reshape(-1))
            chi_res.append(np.array(rob.residuals))
    out.update(chi_params=chi_params, chi_values=np.array(chi_vals), chi_alpha=np.array(chi_alpha),
               chi_residuals=np.array(chi_res))
    for vk in (False, True):
        cls = treegp.AnisotropicVonKarman if vk else treegp.AnisotropicRBF
        rv = two_pcf_mod.robust_2dfit(cls(invLam=np.eye(2)), t2._2pcf, t2._2pcf_dist[:, 0], t2._2pcf_dist[:, 1],
                                      t2._2pcf_weight, mask=t2._2pcf_mask)
        out["chi_model_%s" % ("avk" if vk else "arbf")] = rv._model_skl(1.3, 1.7, 0.2, -0.1)

    # Log (isotropic) binning: comp_2pcf, return_2pcf and the optimizer with default separations
    t3 = treegp.two_pcf(X, y, y_err, 0.3, 8.0, nbins=12, anisotropic=False)
    xi, dist, coord, mask = t3.comp_2pcf(X, y, y_err)
    kk = treecorr.calls[-1]
    out.update(log_xi=xi, log_dist=dist, log_coord=coord, log_mask=mask,
               log_npairs=kk.npairs.astype(np.int64), log_weight=kk.weight)
    xi_r, w_r, _, _, _ = t3.return_2pcf()
    out.update(log_ret_weight=w_r)
    t4 = treegp.two_pcf(X, y, y_err, None, None, nbins=20, anisotropic=False)
    s_iso = "1.0**2 * RBF(0.8)"
    fitted = t4.optimizer(treegp.eval_kernel(s_iso))
    out.update(log_opt_start=np.array(s_iso), log_opt_theta=np.array(fitted.theta),
               log_opt_min_sep=np.array(t4.min_sep), log_opt_max_sep=np.array(t4.max_sep),
               log_opt_xi=t4._2pcf, log_opt_dist=t4._2pcf_dist, log_opt_fit=t4._2pcf_fit)
    print("two_pcf Log fit theta-hat", fitted.theta)
    # 1-D input is embedded as (x, 0) (two_pcf.py:250-251)
    X1 = rng.uniform(-40, 40, size=(300, 1))
    y1, e1 = draw_field(rng, treegp.eval_kernel("1.0**2 * RBF(1.5)"), X1, 0.05)
    t5 = treegp.two_pcf(X1, y1, e1, None, None, nbins=15, anisotropic=False)
    fitted = t5.optimizer(treegp.eval_kernel("0.8**2 * RBF(1.0)"))
    out.update(one_X=X1, one_y=y1, one_yerr=e1, one_opt_theta=np.array(fitted.theta), one_opt_xi=t5._2pcf,
               one_opt_dist=t5._2pcf_dist, one_opt_min_sep=np.array(t5.min_sep), one_opt_max_sep=np.array(t5.max_sep))

    # ---------------------------------------------------------------- (iii) E/B diagnostics and meanify
    m = 700
    ex, ey = rng.uniform(0, 1, m), rng.uniform(0, 1, m)
    edx, edy = rng.normal(size=m), rng.normal(size=m)
    ex[10], ey[10] = ex[11], ey[11]              # a coincident pair: log(0) = -inf falls outside the range
    rmin, rmax, dlogr = 0.002, 1.0, 0.05
    with np.errstate(divide="ignore", invalid="ignore"):
        logr, xip, xim, xix, xiz2 = utils_mod.vcorr(ex, ey, edx, edy, rmin=rmin, rmax=rmax, dlogr=dlogr)
        xie, xib, logr2 = utils_mod.comp_eb(ex, ey, edx, edy, rmin=rmin, rmax=rmax, dlogr=dlogr)
        # the per-bin counts, with the reference's own expressions (utils.py:37-55)
        i1, i2 = np.triu_indices(m)
        use = i1 != i2
        i1, i2 = i1[use], i2[use]
        dr = 1j * (ey[i2] - ey[i1])
        dr += ex[i2] - ex[i1]
        logdr = np.log(np.absolute(dr))
    bins = int(np.ceil(np.log(rmax / rmin) / dlogr))
    hrange = (np.log(rmin), np.log(rmin) + bins * dlogr)
    counts = np.histogram(logdr, bins=bins, range=hrange)[0]
    out.update(eb_x=ex, eb_y=ey, eb_dx=edx, eb_dy=edy, eb_par=np.array([rmin, rmax, dlogr]), eb_counts=counts,
               eb_logr=logr, eb_xip=xip, eb_xim=xim, eb_xix=xix, eb_xiz2=xiz2, eb_xie=xie, eb_xib=xib)
    # a lattice: many displacements share a radius, some of them sit on a bin edge to the last bit
    gx, gy = np.meshgrid(np.arange(24) * 0.04, np.arange(24) * 0.04)
    gx, gy = gx.reshape(-1), gy.reshape(-1)
    gdx, gdy = rng.normal(size=gx.size), rng.normal(size=gx.size)
    with np.errstate(divide="ignore", invalid="ignore"):
        lat = utils_mod.vcorr(gx, gy, gdx, gdy, rmin=0.01, rmax=1.0, dlogr=0.05)
        i1, i2 = np.triu_indices(gx.size, 1)
        ldr = np.log(np.absolute((gx[i2] - gx[i1]) + 1j * (gy[i2] - gy[i1])))
    lbins = int(np.ceil(np.log(1.0 / 0.01) / 0.05))
    lcounts = np.histogram(ldr, bins=lbins, range=(np.log(0.01), np.log(0.01) + lbins * 0.05))[0]
    out.update(ebl_x=gx, ebl_y=gy, ebl_dx=gdx, ebl_dy=gdy, ebl_counts=lcounts, ebl_logr=lat[0], ebl_xip=lat[1],
               ebl_xim=lat[2], ebl_xix=lat[3], ebl_xiz2=lat[4])
    out["xiB_direct"] = utils_mod.xiB(logr[np.isfinite(logr)], xip[np.isfinite(logr)], xim[np.isfinite(logr)])

    mfy_coords = [rng.uniform(-1000, 1000, size=(800, 2)) for _ in range(3)]
    mfy_params = [np.sin(c[:, 0] / 300.0) + 0.1 * rng.normal(size=len(c)) for c in mfy_coords]
    out["mfy_coords"] = np.array(mfy_coords)
    out["mfy_params"] = np.array(mfy_params)
    for stat in ("mean", "median"):
        mf = treegp.meanify(bin_spacing=200.0, statistics=stat)
        for c, p in zip(mfy_coords, mfy_params):
            mf.add_field(c, p)
        mf.meanify()
        out["mfy_%s_coords0" % stat] = mf.coords0
        out["mfy_%s_params0" % stat] = mf.params0
        out["mfy_%s_wrms0" % stat] = mf.wrms0
        out["mfy_%s_average" % stat] = mf._average
    mf = treegp.meanify(bin_spacing=250.0, statistics="mean")
    mf.add_field(mfy_coords[0], mfy_params[0])
    mf.meanify(lu_min=-900.0, lu_max=950.0, lv_min=-800.0, lv_max=1000.0)
    out.update(mfy_lim_coords0=mf.coords0, mfy_lim_params0=mf.params0)

    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
