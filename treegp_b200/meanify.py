"""Mean-function ("meanify") support.

SURVEY.md section 8f-1 ranks this path NEXT after the GP hot path.  Provided here on the host (it is O(N)
binning that runs once, offline, in the reference as well):
* the consumer side GPInterpolation needs (gp_interp.py:97-107, :229-243): reading the spatial-average
  table and the k-nearest-neighbour lookup (`tgp_knn_mean` on the device; sklearn's KD-tree in the reference);
* the producer, class ``meanify`` (mirror of /root/reference/treegp/meanify.py:12-165): 2-D binned mean /
  median / weighted mean over many fields and a FITS table writer -- written with numpy bin counts
  instead of scipy.stats.binned_statistic_2d, and with the in-repo FITS writer instead of fitsio.
"""
import copy

import numpy as np

from . import fitstable


def read_average(path):
    """COORDS0 (n, 2) and PARAMS0 (n,) of the 'average_solution' table (gp_interp.py:100-102)."""
    tab = fitstable.read_table(path, ext=1)
    return tab["COORDS0"][0], tab["PARAMS0"][0]


def knn_average(X0, y0, X, n_neighbors):
    """Uniform mean of the `n_neighbors` nearest mean-grid values (gp_interp.py:236-238: sklearn's
    KNeighborsRegressor in the reference; here every query scans the small grid on the device)."""
    from . import backend

    return backend.knn_mean(np.asarray(X0, dtype=np.float64), np.asarray(y0, dtype=np.float64),
                            np.asarray(X, dtype=np.float64), int(n_neighbors)).cpu().numpy()


def _bin_index(v, edges):
    """Bin of each value for edges e_0 < ... < e_m: [e_k, e_k+1), last bin closed on the right; -1 if
    outside (the convention of scipy.stats.binned_statistic_2d used at meanify.py:79-111)."""
    v = np.ascontiguousarray(v, dtype=np.float64)
    edges = np.asarray(edges, dtype=np.float64)

    def by_search(x):
        idx = np.searchsorted(edges, x, side="right") - 1
        idx[x == edges[-1]] = len(edges) - 2
        idx[(x < edges[0]) | (x > edges[-1])] = -1
        return idx

    # The edges come from np.linspace: guess the bin arithmetically, keep the guess where the defining inequality
    # e_k <= v < e_k+1 holds for the ACTUAL edge values, and settle the rest (points outside, on the last edge, or
    # one bin off through rounding) by the binary search -- the same bins in every case, without a binary search per
    # point (2 x 1.2 million of them are 2/3 of a 6-exposure meanify at N = 200 000).
    m = len(edges) - 1
    span = edges[-1] - edges[0] if m >= 1 else 0.0
    if not (m >= 1 and span > 0 and np.isfinite(span)):
        return by_search(v)
    with np.errstate(invalid="ignore", over="ignore"):
        t = v - edges[0]
        t *= m / span
        np.clip(t, 0, m - 1, out=t)
        t[np.isnan(t)] = 0
        k = t.astype(np.intp)
    good = edges.take(k) <= v
    good &= v < edges.take(k + 1)
    bad = np.flatnonzero(~good)
    if len(bad):
        k[bad] = by_search(v[bad])
    return k


class meanify(object):
    """Take data, build a spatial average, and write output average.

    :param bin_spacing: Bin_size, resolution on the mean function. (default=120.)
    :param statistics:  Statisitics used to compute the mean: "mean", "median" or "weighted". (default=mean)
    """

    def __init__(self, bin_spacing=120.0, statistics="mean"):
        self.bin_spacing = bin_spacing
        if statistics not in ["mean", "median", "weighted"]:
            raise ValueError(
                "%s is not a suported statistic (only mean, weighted, and median are currently suported)"
                % (statistics)
            )
        self.stat_used = statistics
        self.coords = []
        self.params = []
        self.params_err = []

    def add_field(self, coord, param, params_err=None):
        """
        Add new data to compute the mean function.

        :param coord: Array of coordinate of the parameter.
        :param param: Array of parameter.
        """
        if np.shape(coord)[1] != 2:
            raise ValueError("meanify is supported only in 2d for the moment.")
        self.coords.append(coord)
        self.params.append(param)
        if self.stat_used == "weighted":
            if params_err is None:
                raise ValueError("Need an associated error to params")
            self.params_err.append(params_err)

    def meanify(self, lu_min=None, lu_max=None, lv_min=None, lv_max=None):
        """
        Compute the mean function on a regular (u, v) grid of pitch ~bin_spacing.
        """
        params = np.concatenate(self.params)
        coords = np.concatenate(self.coords, axis=0)
        u, v = coords[:, 0], coords[:, 1]
        lu_min = np.min(u) if lu_min is None else lu_min
        lu_max = np.max(u) if lu_max is None else lu_max
        lv_min = np.min(v) if lv_min is None else lv_min
        lv_max = np.max(v) if lv_max is None else lv_max
        # int((max-min)/spacing) EDGES per axis, i.e. one bin fewer (meanify.py:68-74)
        xedge = np.linspace(lu_min, lu_max, int((lu_max - lu_min) / self.bin_spacing))
        yedge = np.linspace(lv_min, lv_max, int((lv_max - lv_min) / self.bin_spacing))
        nu, nv = len(xedge) - 1, len(yedge) - 1
        iu, iv = _bin_index(u, xedge), _bin_index(v, yedge)
        inside = (iu >= 0) & (iv >= 0)
        flat = iu[inside] * nv + iv[inside]
        vals = params[inside]

        def binned_sum(weights):
            return np.bincount(flat, weights=weights, minlength=nu * nv).reshape(nu, nv)

        with np.errstate(invalid="ignore", divide="ignore"):
            if self.stat_used == "weighted":
                w = 1.0 / np.concatenate(self.params_err)[inside] ** 2
                sum_w, sum_wp, sum_wpp = binned_sum(w), binned_sum(w * vals), binned_sum(w * vals * vals)
                average = sum_wp / sum_w
                wrms = np.sqrt((sum_wpp - 2.0 * average * sum_wp + average * average * sum_w) / sum_w)
            elif self.stat_used == "mean":
                average = binned_sum(vals) / binned_sum(np.ones_like(vals))
                wrms = np.zeros_like(average)
            else:  # median
                average = np.full(nu * nv, np.nan)
                order = np.argsort(flat, kind="stable")
                fs, vs = flat[order], vals[order]
                starts = np.flatnonzero(np.r_[True, fs[1:] != fs[:-1]])
                ends = np.r_[starts[1:], len(fs)]
                for a, b in zip(starts, ends):
                    average[fs[a]] = np.median(vs[a:b])
                average = average.reshape(nu, nv)
                wrms = np.zeros_like(average)
        average, wrms = average.T, wrms.T       # rows = v, columns = u, as meanify.py:112-113
        self._average = copy.deepcopy(average)
        self._wrms = wrms
        keep = np.isfinite(average).reshape(-1) & np.isfinite(wrms).reshape(-1)

        # centre of each bin
        u0 = xedge[:-1] + (xedge[1] - xedge[0]) / 2.0
        v0 = yedge[:-1] + (yedge[1] - yedge[0]) / 2.0
        u0, v0 = np.meshgrid(u0, v0)
        self._u0, self._v0 = u0, v0
        self._xedge, self._yedge = xedge, yedge
        coords0 = np.array([u0.reshape(-1), v0.reshape(-1)]).T
        # bins without data (non-finite statistic) are dropped
        self.coords0 = coords0[keep]
        self.params0 = average.reshape(-1)[keep]
        self.wrms0 = wrms.reshape(-1)[keep]

    def save_results(self, name_output="mean_gp.fits"):
        """
        Write output mean function (one-row binary table 'average_solution', meanify.py:139-165).

        :param name_output: Name of the output fits file. (default: 'mean_gp.fits')
        """
        fitstable.write_table(name_output, {
            "COORDS0": self.coords0, "PARAMS0": self.params0, "WRMS0": self.wrms0,
            "_AVERAGE": self._average, "_WRMS": self._wrms, "_U0": self._u0, "_V0": self._v0,
        }, extname="average_solution")
