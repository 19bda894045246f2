"""E/B-mode diagnostics of a vector field's 2-point correlation function.

Mirror of /root/reference/treegp/utils.py (``vcorr``, ``xiB``, ``comp_eb``, ``comp_eb_treecorr``).  These are
diagnostics that GPInterpolation never calls (ranked NEXT, same pair-tile shape as the
pair-binning kernel); the O(N^2) pair sums run on the device (`tgp_vcorr`), so the N(N-1)/2 index pairs the
reference materialises (utils.py:38-47) never exist; everything else (E/B combination, xiB integral) is O(bins)
host arithmetic.
"""
import numpy as np


def _pair_sums(x, y, dx, dy, logrmin, dlogr, bins):
    """Per log-r bin: counts, sum log r, sum v1.v2*, sum v1 v2, sum v1 v2 exp(-2 i phi) over all pairs -- on the
    device (csrc/vcorr.cu); the bins are np.histogram's of utils.py:52-55, decided by thresholds on r^2
    (binning.hist_thresholds_r2) with the undecided pairs settled on the host (backend.vcorr_sums)."""
    from . import backend

    return backend.vcorr_sums(x, y, dx, dy, logrmin, dlogr, bins)


def vcorr(x, y, dx, dy, rmin=5.0 / 3600.0, rmax=1.5, dlogr=0.05, maxpts=30000):
    """
    Angle-averaged 2-point correlation functions of a vector field (brute-force pair counting).

    Returns logr (mean log radius per bin), xi_+ = <vr1 vr2 + vt1 vt2>, xi_- = <vr1 vr2 - vt1 vt2>,
    xi_x = <vr1 vt2 + vt1 vr2>, xi_z2 = <vx1 vx2 - vy1 vy2 + 2 i vx1 vy2>.

    :param x, y:    positions of objects.
    :param dx, dy:  vector field (e.g. astrometric shift).
    :param rmin, rmax: separation range.  :param dlogr: bin size in log(r).
    :param maxpts:  maximum number of points used (random subsample beyond).
    """
    x, y, dx, dy = (np.asarray(a, dtype=float) for a in (x, y, dx, dy))
    if len(x) > maxpts:
        use = np.random.random(len(x)) <= float(maxpts) / len(x)
        x, y, dx, dy = x[use], y[use], dx[use], dy[use]
    logrmin = np.log(rmin)
    bins = int(np.ceil(np.log(rmax / rmin) / dlogr))
    counts, s_logr, s_plus, s_z2, s_minus = _pair_sums(x, y, dx, dy, logrmin, dlogr, bins)
    with np.errstate(invalid="ignore", divide="ignore"):
        logr = s_logr / counts
        xiplus = s_plus / counts
        xiz2 = s_z2 / counts
        xim = s_minus / counts
    return logr, xiplus, np.real(xim), np.imag(xim), xiz2


def xiB(logr, xiplus, ximinus):
    """
    Estimate of the pure B-mode correlation function: (xi+ - xi-)/2 + int_r^inf dlog r' xi-(r').
    """
    dlogr = np.zeros_like(logr)
    dlogr[1:-1] = 0.5 * (logr[2:] - logr[:-2])
    integral = np.cumsum((np.array(ximinus) * dlogr)[::-1])[::-1]
    return 0.5 * (xiplus - ximinus) + integral


def comp_eb(u, v, du, dv, **kwargs):
    """
    E/B decomposition of a vector field's correlation function.

    :returns: xie, xib, logr
    """
    logr, xiplus, ximinus, xicross, xiz2 = vcorr(u, v, du, dv, **kwargs)
    xib = xiB(logr, xiplus, ximinus)
    return xiplus - xib, xib, logr


def comp_eb_treecorr(u, v, du, dv, rmin=5.0 / 3600.0, rmax=1.5, dlogr=0.05):
    """
    Same decomposition with TreeCorr's conventions for the binning (the reference calls
    treecorr.VVCorrelation(min_sep, max_sep, bin_size), utils.py:110-155): nbins = ceil(ln(max/min)/bin_size),
    logr = nominal bin centres.  Brute force (the bin_slop -> 0 limit of the tree code).

    :returns: xie, xib, logr
    """
    u, v, du, dv = (np.asarray(a, dtype=float) for a in (u, v, du, dv))
    logrmin = np.log(rmin)
    bins = int(np.ceil(np.log(rmax / rmin) / dlogr))
    counts, _, s_plus, _, s_minus = _pair_sums(u, v, du, dv, logrmin, dlogr, bins)
    with np.errstate(invalid="ignore", divide="ignore"):
        xip = np.where(counts > 0, s_plus / counts, 0.0)
        xim = np.where(counts > 0, s_minus.real / counts, 0.0)
    logr = logrmin + (np.arange(bins) + 0.5) * dlogr
    xib = xiB(logr, xip, xim)
    return xip - xib, xib, logr
