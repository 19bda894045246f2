"""Host-side bin geometry for the 2-point correlation function.

The device never divides or takes logarithms to place a pair in a bin: it compares against
*decision thresholds* computed here once per call.  Threshold k is the smallest binary64 value v for
which the binning formula, evaluated in IEEE double arithmetic exactly as TreeCorr states it,

    TwoD:  int((v + max_sep) / bin_size)                  bin_size = 2 max_sep / nbins
    Log :  int((0.5 ln(v) - ln(min_sep)) / bin_size)      bin_size = ln(max_sep / min_sep) / nbins,  v = r^2

returns >= k.  Because both formulas are monotone in v the set {v >= threshold_k} is exactly
{bin(v) >= k}, so counting thresholds reproduces the formula bit for bit
(replaces the binning inside treecorr.KKCorrelation used at
/root/reference/treegp/two_pcf.py:297-305 and :330-334).

Also builds the half-plane mask and bin-centre coordinates of two_pcf.py:306-328.
"""
import functools
import math
import struct
import numpy as np


def _key(x):
    """Monotone map double -> int (total order of the non-NaN doubles)."""
    (u,) = struct.unpack("<q", struct.pack("<d", x))
    return u if u >= 0 else -(u & 0x7FFFFFFFFFFFFFFF)


def _unkey(k):
    u = k if k >= 0 else ((-k) | (1 << 63)) - (1 << 64)
    (x,) = struct.unpack("<d", struct.pack("<q", u))
    return x


def _smallest_true(pred, lo, hi):
    """Smallest double v in (lo, hi] with pred(v), given pred(lo) False, pred(hi) True, pred monotone."""
    klo, khi = _key(lo), _key(hi)
    while khi - klo > 1:
        mid = (klo + khi) // 2
        if pred(_unkey(mid)):
            khi = mid
        else:
            klo = mid
    return _unkey(khi)


def twod_thresholds(max_sep, nbins):
    """edges[0] = -inf, edges[nbins] = +inf, edges[k] = smallest dx with int((dx+max_sep)/bin_size) >= k.
    Memoised per (max_sep, nbins): the bisections cost ~2 ms of Python per geometry."""
    return _twod_thresholds(float(max_sep), int(nbins)).copy()


@functools.lru_cache(maxsize=64)
def _twod_thresholds(max_sep, nbins):
    bin_size = 2.0 * max_sep / nbins
    edges = np.empty(nbins + 1)
    edges[0], edges[nbins] = -np.inf, np.inf
    for k in range(1, nbins):
        edges[k] = _smallest_true(lambda v: int((v + max_sep) / bin_size) >= k, -max_sep, max_sep)
    edges.setflags(write=False)
    return edges


def log_thresholds(min_sep, max_sep, nbins):
    """edges[k] (1 <= k < nbins) = smallest r^2 whose Log bin index is >= k; edges[0], edges[nbins] are
    min_sep^2 and max_sep^2 (informational: the range test uses those two numbers directly).  Memoised like
    twod_thresholds."""
    return _log_thresholds(float(min_sep), float(max_sep), int(nbins)).copy()


@functools.lru_cache(maxsize=64)
def _log_thresholds(min_sep, max_sep, nbins):
    bin_size = math.log(max_sep / min_sep) / nbins
    logminsep = math.log(min_sep)
    lo, hi = min_sep * min_sep, max_sep * max_sep
    edges = np.empty(nbins + 1)
    edges[0], edges[nbins] = lo, hi
    for k in range(1, nbins):
        edges[k] = _smallest_true(lambda v: int((0.5 * math.log(v) - logminsep) / bin_size) >= k,
                                  lo * 0.5, hi * 2.0)
    edges.setflags(write=False)
    return edges


def hist_edges(logrmin, dlogr, bins):
    """bin_edges of np.histogram(logdr, bins=bins, range=(logrmin, logrmin + bins * dlogr)) (utils.py:50-55)."""
    return np.linspace(float(logrmin), float(logrmin) + bins * float(dlogr), bins + 1)


def hist_bin(logdr, edges):
    """Bin index np.histogram gives each value for equal-width `edges` (-1: outside): edges[k] <= v < edges[k+1],
    the last bin closed on the right."""
    logdr = np.asarray(logdr, dtype=np.float64)
    k = np.searchsorted(edges, logdr, side="right") - 1
    k[logdr == edges[-1]] = len(edges) - 2
    k[(logdr < edges[0]) | (logdr > edges[-1]) | ~np.isfinite(logdr)] = -1
    return k


def hist_thresholds_r2(logrmin, dlogr, bins):
    """r^2 thresholds for the log-radius binning of utils.py:50-55 (`vcorr`): out[k] = h_k^2 with h_k the smallest
    double whose np.log (numpy's own, the function the reference bins with) is >= bin_edges[k]; for k = bins the
    smallest with np.log(h) > bin_edges[bins] (np.histogram closes the last bin).  h^2 is rounded; the device
    treats r^2 within 1e-14 of a threshold as undecided (csrc/vcorr.cu), so that rounding is immaterial."""
    edges = hist_edges(logrmin, dlogr, bins)
    out = np.empty(bins + 1)
    for k in range(bins + 1):
        e = edges[k]
        lo, hi = math.exp(e - 1e-6 - 1e-9 * abs(e)), math.exp(e + 1e-6 + 1e-9 * abs(e))
        if k < bins:
            h = _smallest_true(lambda v: bool(np.log(np.float64(v)) >= e), lo, hi)
        else:
            h = _smallest_true(lambda v: bool(np.log(np.float64(v)) > e), lo, hi)
        out[k] = h * h
    return out


def twod_mask(nbins):
    """Boolean mask keeping one half-plane of the point-symmetric nbins x nbins grid (flattened,
    rows = dy).  Same selection as two_pcf.py:309-321: all rows below the centre, plus -- for odd
    nbins -- the left part of the centre row including the centre pixel."""
    mask = np.zeros((nbins, nbins), dtype=bool)
    half = nbins // 2 + nbins % 2
    mask[:half, :] = True
    if nbins % 2:
        mask[half - 1, half:] = False
    return mask.reshape(-1)


def twod_coords(nbins, max_sep):
    """(nbins^2, 2) array of bin-centre lags (dx, dy), flat index = row(dy) * nbins + column(dx)
    (two_pcf.py:323-327)."""
    edges = np.linspace(-float(max_sep), float(max_sep), nbins + 1)
    centres = (edges[:-1] + edges[1:]) / 2.0
    dy, dx = np.meshgrid(centres, centres, indexing="ij")
    return np.stack([dx.reshape(-1), dy.reshape(-1)], axis=1)
