"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Python front-end of oracle/pairbin_oracle.c plus an independent all-pairs numpy restatement used to
cross-check the C code on small inputs.  PARITY UNPINNED against TreeCorr (not installable; no golden
vectors in the reference) -- see the header of pairbin_oracle.c.

Also restates the host-side assembly of two_pcf.comp_2pcf (/root/reference/treegp/two_pcf.py:283-340):
weights, mean subtraction, the half-plane mask and the bin-centre coordinates.
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "libpairbin_oracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        lib = ctypes.CDLL(so)
        dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)
        i64, f64, i32 = ctypes.c_int64, ctypes.c_double, ctypes.c_int
        lib.oracle_pairbin_twod.argtypes = [dp, dp, dp, dp, i64, i64, i64, f64, f64, i32, ip, dp, dp]
        lib.oracle_pairbin_log.argtypes = [dp, dp, dp, dp, i64, i64, i64, f64, f64, i32, ip, dp, dp, dp]
        lib.oracle_num_threads.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def num_threads():
    return int(_lib().oracle_num_threads())


def set_threads(n):
    """Override OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1 to every rank)."""
    _lib().oracle_set_threads(int(n))


def _ptr(a, t=ctypes.c_double):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def pairbin(x, y, k, w, min_sep, max_sep, nbins, bin_type="TwoD", rows=None):
    """Returns dict(npairs, weight, sumwkk, xi[, sumwr, meanr]) for one catalogue (C, OpenMP)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    k = np.ascontiguousarray(k, dtype=np.float64)
    w = None if w is None else np.ascontiguousarray(w, dtype=np.float64)
    n = len(x)
    i0, i1 = (0, n) if rows is None else rows
    nb = nbins * nbins if bin_type == "TwoD" else nbins
    npairs = np.zeros(nb, dtype=np.int64)
    sumw = np.zeros(nb)
    sumwkk = np.zeros(nb)
    lib = _lib()
    out = {}
    if bin_type == "TwoD":
        lib.oracle_pairbin_twod(_ptr(x), _ptr(y), _ptr(k), _ptr(w), n, i0, i1, float(min_sep), float(max_sep),
                                int(nbins), _ptr(npairs, ctypes.c_int64), _ptr(sumw), _ptr(sumwkk))
    else:
        sumwr = np.zeros(nb)
        lib.oracle_pairbin_log(_ptr(x), _ptr(y), _ptr(k), _ptr(w), n, i0, i1, float(min_sep), float(max_sep),
                               int(nbins), _ptr(npairs, ctypes.c_int64), _ptr(sumw), _ptr(sumwkk), _ptr(sumwr))
        out["sumwr"] = sumwr
        with np.errstate(invalid="ignore", divide="ignore"):
            out["meanr"] = np.where(sumw > 0, sumwr / sumw, 0.0)
    with np.errstate(invalid="ignore", divide="ignore"):
        xi = np.where(sumw != 0, sumwkk / sumw, 0.0)
    out.update(npairs=npairs, weight=sumw, sumwkk=sumwkk, xi=xi)
    return out


def pairbin_numpy(x, y, k, w, min_sep, max_sep, nbins, bin_type="TwoD"):
    """Independent all-pairs numpy restatement (O(N^2) memory; small N only)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    k = np.asarray(k, dtype=np.float64)
    w = np.ones_like(x) if w is None else np.asarray(w, dtype=np.float64)
    i, j = np.triu_indices(len(x), 1)
    dx, dy = x[j] - x[i], y[j] - y[i]
    rsq = dx * dx + dy * dy
    ww = w[i] * w[j]
    kk = (w[i] * k[i]) * (w[j] * k[j])
    if bin_type == "TwoD":
        nb = nbins * nbins
        bin_size = 2.0 * max_sep / nbins
        keep = (rsq != 0.0) & (rsq >= min_sep * min_sep) & (np.maximum(np.abs(dx), np.abs(dy)) < max_sep)
        idx = []
        for sx, sy in ((dx[keep], dy[keep]), (-dx[keep], -dy[keep])):
            ii = ((sx + max_sep) / bin_size).astype(np.int64)
            jj = ((sy + max_sep) / bin_size).astype(np.int64)
            ii[ii == nbins] -= 1
            jj[jj == nbins] -= 1
            idx.append(jj * nbins + ii)
        idx = np.concatenate(idx)
        ww = np.concatenate([ww[keep], ww[keep]])
        kk = np.concatenate([kk[keep], kk[keep]])
        r = None
    else:
        nb = nbins
        bin_size = np.log(max_sep / min_sep) / nbins
        keep = (rsq >= min_sep * min_sep) & (rsq < max_sep * max_sep)
        import math
        lr = np.array([0.5 * math.log(v) for v in rsq[keep]])  # libm log, as the C oracle
        idx = ((lr - math.log(min_sep)) / bin_size).astype(np.int64)
        idx = np.clip(idx, 0, nbins - 1)
        ww, kk, r = ww[keep], kk[keep], np.sqrt(rsq[keep])
    npairs = np.bincount(idx, minlength=nb).astype(np.int64)
    sumw = np.bincount(idx, weights=ww, minlength=nb)
    sumwkk = np.bincount(idx, weights=kk, minlength=nb)
    out = dict(npairs=npairs, weight=sumw, sumwkk=sumwkk)
    with np.errstate(invalid="ignore", divide="ignore"):
        out["xi"] = np.where(sumw != 0, sumwkk / sumw, 0.0)
        if r is not None:
            out["sumwr"] = np.bincount(idx, weights=ww * r, minlength=nb)
            out["meanr"] = np.where(sumw > 0, out["sumwr"] / sumw, 0.0)
    return out


def twod_mask_and_coords(nbins, max_sep):
    """Half-plane mask and bin-centre coordinates of two_pcf.py:306-328."""
    npix = nbins * nbins
    mask = np.ones((nbins, nbins), dtype=bool)
    even = nbins % 2 == 0
    nmask = int(nbins / 2 + nbins % 2)
    mask[nmask:, :] = False
    mask[nmask - 1][nmask:] = even
    edges = np.linspace(-max_sep, max_sep, nbins + 1)
    centres = 0.5 * (edges[:-1] + edges[1:])
    # kk.bottom_edges / top_edges vary along axis 0 (rows = dy); dx = dy.T     two_pcf.py:323-324
    dy = np.repeat(centres[:, None], nbins, axis=1)
    dx = dy.T
    coord = np.array([dx.reshape(npix), dy.reshape(npix)]).T
    return mask.reshape(npix), coord


def comp_2pcf(X, y, y_err, min_sep, max_sep, nbins, anisotropic):
    """two_pcf.comp_2pcf (two_pcf.py:283-340): returns xi, distance, coord, mask."""
    X = np.asarray(X, dtype=float)
    y = np.asarray(y, dtype=float)
    w = None if np.sum(y_err) == 0 else 1.0 / np.asarray(y_err, dtype=float) ** 2
    k = y - np.mean(y)
    if anisotropic:
        res = pairbin(X[:, 0], X[:, 1], k, w, min_sep, max_sep, nbins, "TwoD")
        mask, coord = twod_mask_and_coords(nbins, max_sep)
        return res["xi"], coord, coord, mask
    res = pairbin(X[:, 0], X[:, 1], k, w, min_sep, max_sep, nbins, "Log")
    distance = res["meanr"]
    coord = np.array([distance, np.zeros_like(distance)]).T
    return res["xi"], distance, coord, np.ones(nbins, dtype=bool)
