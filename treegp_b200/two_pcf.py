"""Hyper-parameter estimation from the 2-point correlation function, pair-binned on the B200.

Host-side mirror of /root/reference/treegp/two_pcf.py: ``get_correlation_length_matrix`` (:12-31),
``get_kernel_class`` (:34-65), ``robust_2dfit`` (:68-206) and ``two_pcf`` (:209-464) keep their names,
signatures and the attributes the reference's tests read (``_2pcf``, ``_2pcf_weight``, ``_2pcf_dist``,
``_2pcf_fit``, ``_2pcf_mask``, ``_kernel``, ``_results_robust``).

What changed:
* ``comp_2pcf`` -- the TreeCorr ``KKCorrelation.process`` call (:297-305, :330-334) is the
  brute-force pair-binning kernel of csrc/pairbin.cu behind ``tgp_pairbin``;
* ``comp_xi_covariance`` -- the loop of ``n_bootstrap`` sequential TreeCorr runs (:342-362) is ONE
  batched launch over all resampled catalogues.  A resample with replacement is the same catalogue
  with integer multiplicities m_i: coincident copies are at r = 0 and never pair (TreeCorr skips
  rsq == 0), copies of i and j contribute m_i m_j w_i w_j, so each distinct drawn point enters once
  with weight m_i w_i.  Index generation stays on the host with numpy so that the reference's stream
  (``default_rng(seed).integers(0, N-1, N)``: index N-1 is never drawn, :275) is reproduced;
* the Minuit minimiser of the robust fit is the in-repo variable-metric ``migrad`` (iminuit is not
  part of this image); the chi-square machinery around it is unchanged.
"""
from __future__ import print_function

import copy
import warnings

import numpy as np
import sklearn
import torch
from scipy import optimize

from . import _cabi, backend, binning, kernels
from .migrad import Migrad


def get_correlation_length_matrix(size, e1, e2):
    """2 x 2 correlation-length matrix of an elliptical kernel in the weak-lensing shear parameterisation:
    `size` along the major axis, axis ratio (1-e)/(1+e) with e = |(e1, e2)|, position angle atan2(e2, e1)/2."""
    if abs(e1) > 1 or abs(e2) > 1:
        raise ValueError("abs value of e1 and e2 must be lower than one")
    e = np.sqrt(e1 ** 2 + e2 ** 2)
    q = (1 - e) / (1 + e)
    phi = 0.5 * np.arctan2(e2, e1)
    c, s = np.cos(phi), np.sin(phi)
    rot = np.array([[c, s], [-s, c]])
    ell = np.array([[size ** 2, 0], [0, (size * q) ** 2]])
    return np.dot(rot.T, ell.dot(rot))


_ANISOTROPIC = (kernels.AnisotropicVonKarman, kernels.AnisotropicRBF)


def get_kernel_class(A):
    """Class of the anisotropic factor (AnisotropicRBF / AnisotropicVonKarman) of kernel A, which may be
    that factor itself or a Product containing it; ValueError otherwise."""
    msg = "Work only with treegp.kernels.AnisotropicVonKarman and treegp.kernels.AnisotropicRBF"
    if isinstance(A, sklearn.gaussian_process.kernels.Product):
        found = [v.__class__ for v in vars(A).values() if v.__class__ in _ANISOTROPIC]
        if not found:
            raise ValueError(msg)
        return found[-1]
    if A.__class__ in _ANISOTROPIC:
        return A.__class__
    raise ValueError(msg)


class robust_2dfit(object):
    """Chi-square fit of (size, g1, g2) of an anisotropic kernel to a measured 2-D correlation function.

    kernel: template kernel (selects the family); flat_data: flattened xi; x, y: lag of every pixel;
    W: inverse covariance of the masked pixels; mask: pixels used (one half-plane of the symmetric map).
    """

    def __init__(self, kernel, flat_data, x, y, W, mask=None):
        self.mask = np.ones(len(x), dtype=bool) if mask is None else mask
        self.kernel_class = get_kernel_class(kernel)
        self.flat_data = flat_data
        self.x = x
        self.y = y
        self.coord = np.array([x, y]).T
        self.W = W
        self.N = int(np.sqrt(len(self.x)))
        self._origin = np.zeros((1, 2))

    def _model_skl(self, sigma, corr_length, g1, g2):
        """
        Analytical two point correlation function of the kernel at the bin lags
        (two_pcf.py:96-113); None outside |g| <= 1.
        """
        if abs(g1) > 1 or abs(g2) > 1:
            return None
        invLam = np.linalg.inv(get_correlation_length_matrix(corr_length, g1, g2))
        kernel_used = sigma ** 2 * self.kernel_class(invLam=invLam)
        # the reference evaluates the P x P matrix against P zero points and keeps column 0
        # (two_pcf.py:111); one zero point gives the same column
        pcf = kernel_used(self.coord, Y=self._origin)[:, 0]
        self.kernel_fit = kernel_used
        return pcf

    def chi2(self, param):
        """
        Chi2 over the non-linear parameters (correlation length, e1, e2); the amplitude and the
        constant offset are linear and solved analytically at every call (two_pcf.py:115-148).
        """
        if not np.isfinite(np.sum(param)):
            self.chi2_value = [np.inf]
            return np.inf
        model = self._model_skl(1.0, param[0], param[1], param[2])
        if model is None:
            self.chi2_value = [np.inf]
            return np.inf
        model = model[self.mask]
        F = np.array([model, np.ones_like(model)]).T
        FtW = np.dot(F.T, self.W)
        Y = self.flat_data[self.mask].reshape((len(model), 1))
        self.alpha = np.linalg.inv(FtW.dot(F)).dot(FtW.dot(Y))
        self.alpha[0] = abs(self.alpha[0])
        self.residuals = self.flat_data[self.mask] - ((self.alpha[0] * model) + self.alpha[1])
        self.chi2_value = self.residuals.dot(self.W).dot(self.residuals.reshape((len(model), 1)))
        return self.chi2_value[0]

    def chi2_batch(self, params):
        """chi2 of several parameter sets [size, g1, g2] in one device launch (the probes of a numerical gradient or
        Hessian): same objective as `chi2`, without its side effects (alpha, residuals, kernel_fit)."""
        if getattr(self, "_dev", None) is None:
            fam = _cabi.FAM_VONKARMAN if self.kernel_class is kernels.AnisotropicVonKarman else _cabi.FAM_RBF
            self._dev = (backend.to_device(np.ascontiguousarray(self.coord[self.mask])),
                         backend.to_device(np.ascontiguousarray(self.flat_data[self.mask])),
                         backend.to_device(np.ascontiguousarray(self.W)), fam)
        coord_d, y_d, W_d, fam = self._dev
        return backend.robust_chi2_batch(coord_d, y_d, W_d, fam, params)[:, 0]

    def _minimize_minuit(self, p0=[3000.0, 0.2, 0.2]):
        """One variable-metric minimisation started at p0 = [size, g1, g2]."""
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            batch = self.chi2_batch if 2 <= int(np.sum(self.mask)) <= 1024 else None
            self.m = Migrad(self.chi2, p0, fcn_batch=batch)
            self.m.migrad()
            results = list(self.m.values)
            self._fit_ok = self.m.accurate
        # leave alpha / kernel_fit consistent with the returned point
        self.chi2(results)
        self._minuit_result = results
        self.result = [np.sqrt(self.alpha[0][0]), results[0], results[1], results[2], self.alpha[1][0]]

    def minimize_minuit(self, p0=[3000.0, 0.2, 0.2]):
        """Minimise from p0; if that did not converge retry from a 3 x 3 x 3 grid of starts
        (two_pcf.py:178-206)."""
        self._minimize_minuit(p0=p0)

        if not self._fit_ok:
            n = 3
            g = np.linspace(-0.3, 0.3, n)
            size = np.linspace(p0[0] - p0[0] / 10.0, 2 * p0[0], n)
            g1, g2, size = np.meshgrid(g, g, size)
            for s_, a_, b_ in zip(size.ravel(), g1.ravel(), g2.ravel()):
                print("restart fit because failure")
                new_p0 = [s_, a_, b_]
                print(new_p0)
                self._minimize_minuit(p0=new_p0)
                if self._fit_ok:
                    break
        _ = self._model_skl(self.result[0], self.result[1], self.result[2], self.result[3])


class two_pcf(object):
    """2-point correlation function of a scalar field, its bootstrap covariance, and the kernel fit to it.

    X (n, 1|2), y (n,), y_err (n,): the field; min_sep, max_sep, nbins: binning (nbins per axis when
    anisotropic); anisotropic: two-dimensional (dx, dy) map instead of log-r bins; robust_fit: variable-metric
    fit of (size, g1, g2) with analytic amplitude/offset (anisotropic only), started at p0; seed: bootstrap seed.
    """

    def __init__(
        self,
        X,
        y,
        y_err,
        min_sep,
        max_sep,
        nbins=20,
        anisotropic=False,
        robust_fit=False,
        p0=[3000.0, 0.0, 0.0],
        seed=610639139,
    ):
        self.ndim = np.shape(X)[1]
        if self.ndim not in [1, 2]:
            raise ValueError(
                "two-pcf support only 1d and 2d modeling for the moment. curent ndim: %i" % (self.ndim)
            )
        if self.ndim == 2:
            self.X = X
        else:  # embed a 1-D field as (x, 0)                                  two_pcf.py:250-251
            self.X = np.column_stack([np.asarray(X)[:, 0], np.zeros(len(X))])
        self.y = y
        self.y_err = y_err
        self.min_sep = min_sep
        self.max_sep = max_sep
        self.nbins = nbins
        self.anisotropic = anisotropic
        self.robust_fit = robust_fit
        self.p0_robust_fit = p0
        self.seed = seed
        self._rng = None
        # Multi-GPU is OPT-IN: False = this process counts all pairs itself (the default: in a data-parallel job
        # every rank fits its own field).  Set to treegp_b200.dist.WORLD (or a ProcessGroup) to deal the pair
        # tiles to the ranks of that group and all-reduce the bin sums; all ranks must then hold IDENTICAL
        # data and call comp_2pcf / return_2pcf / optimizer together.
        self.group = False

    @property
    def rng(self):
        if self._rng is None:
            self._rng = np.random.default_rng(self.seed)
        return self._rng

    def resample_bootstrap(self):
        """One resample with replacement: (u, v, y, y_err) of the drawn points (two_pcf.py:269-281)."""
        npsfs = len(self.y)
        ind_object = self.rng.integers(0, npsfs - 1, size=npsfs)
        return (self.X[:, 0][ind_object], self.X[:, 1][ind_object], self.y[ind_object],
                self.y_err[ind_object])

    # ---- device pair binning ------------------------------------------------------------------
    def _bin_geometry(self):
        if self.anisotropic:
            edges = binning.twod_thresholds(self.max_sep, self.nbins)
            return _cabi.BIN_TWOD, edges
        return _cabi.BIN_LOG, binning.log_thresholds(self.min_sep, self.max_sep, self.nbins)

    def _pairbin(self, px, py, pk, pw, offsets, max_len):
        """Launch tgp_pairbin on (possibly many) catalogues and return xi (ncat, nb) and meanr|None
        as numpy arrays.  With a process group the pair tiles are sharded and the sums all-reduced."""
        from . import dist

        bt, edges = self._bin_geometry()
        rank, world = dist.rank_world(self.group)
        packed = backend.pairbin_packed(
            px, py, pk, pw, offsets, max_len, bt, self._device_edges(edges), self.nbins,
            self.min_sep, self.max_sep, rank=rank, nranks=world)
        values = False
        if world > 1:                                                 # ONE collective for all bin arrays
            if isinstance(self.group, dist.CabiComm):
                packed = self.group.allreduce_packed_bins(packed)     # behind the C ABI; counts stay int64 words
            else:
                packed = dist.allreduce_packed_bins(self.group, packed)
                values = True
        host = packed.cpu().numpy()                                   # ONE device->host transfer
        # plane 0: pair counts -- raw int64 words from the kernel, FP64 values after a torch all-reduce
        counts = host[0].astype(np.int64) if values else np.ascontiguousarray(host[0]).view(np.int64)
        sw, swkk = host[1], host[2]
        has_wr = host.shape[0] == 4
        with np.errstate(invalid="ignore", divide="ignore"):
            xi = np.where(sw != 0, swkk / sw, 0.0)
            meanr = None
            if has_wr:
                # TreeCorr reports the nominal bin centre exp(ln min_sep + (k + 1/2) bin_size) where a bin is empty
                bs = np.log(self.max_sep / self.min_sep) / self.nbins
                rnom = np.exp(np.log(self.min_sep) + (np.arange(self.nbins) + 0.5) * bs)
                meanr = np.where(sw != 0, host[3] / sw, rnom[None, :])
        self._last_npairs = counts
        self._last_sumw = sw
        return xi, meanr

    def _device_edges(self, edges):
        """Bin thresholds on the device, uploaded once per geometry (a fit makes hundreds of calls with one)."""
        key = (backend.require_cuda(), self.anisotropic, self.min_sep, self.max_sep, self.nbins)
        cached = getattr(self, "_edges_dev", None)
        if cached is None or cached[0] != key:
            self._edges_dev = (key, backend.to_device(edges))
        return self._edges_dev[1]

    def _assemble(self, xi, meanr):
        if self.anisotropic:
            mask = binning.twod_mask(self.nbins)
            coord = binning.twod_coords(self.nbins, self.max_sep)
            return xi, coord, coord, mask
        distance = meanr
        coord = np.array([distance, np.zeros_like(distance)]).T
        return xi, distance, coord, np.ones_like(xi, dtype=bool)

    def comp_2pcf(self, X, y, y_err):
        """xi, separations, bin coordinates and mask for the catalogue (X (n, 2), y, y_err)
        (two_pcf.py:283-340)."""
        X = np.asarray(X, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        y_err = np.asarray(y_err, dtype=np.float64)
        n = len(y)
        # The host only takes the two reductions the reference takes (sum(y_err), mean(y): two_pcf.py:291-297);
        # the arrays go up as they are and the elementwise arithmetic (1 / y_err^2, y - mean) runs on the device
        # -- the same IEEE operations, without three passes over host memory per call.
        # The uploads are enqueued first: from page-locked arrays they run while the host takes its reductions.
        Xd = backend.to_device(X, non_blocking=True)   # one upload of the (n, 2) array; the columns are split on the device
        yd = backend.to_device(y, non_blocking=True)
        pw = None
        if np.sum(y_err) != 0:
            ed = backend.to_device(y_err, non_blocking=True)
            pw = 1.0 / (ed * ed)
        pk = yd - float(np.mean(y))
        # spatially sorted input lets the kernel keep 32 x 32 pair blocks inside a 2 x 2 bin window (TwoD) or
        # inside one radial bin (Log)
        px, py = Xd[:, 0].contiguous(), Xd[:, 1].contiguous()
        order = backend.hilbert_order(px, py)
        px, py, pk = px[order], py[order], pk[order]
        pw = None if pw is None else pw[order]
        xi, meanr = self._pairbin(px, py, pk, pw, backend.to_device(np.array([0, n]), torch.int64), n)
        return self._assemble(xi[0], None if meanr is None else meanr[0])

    def comp_xi_covariance(self, n_bootstrap=1000, mask=None, seed=610639139):
        """Sample covariance of xi[mask] over n_bootstrap resamples (two_pcf.py:342-362); all resamples go
        through one batched launch."""
        self.seed = seed
        self._rng = None
        xi_bootstrap = self._bootstrap_xi(int(n_bootstrap))
        if mask is None:
            mask = np.ones(xi_bootstrap.shape[1], dtype=bool)
        xi_bootstrap = xi_bootstrap[:, mask]
        dxi = xi_bootstrap - np.mean(xi_bootstrap, axis=0)
        return 1.0 / (len(dxi) - 1.0) * np.dot(dxi.T, dxi)

    def _draw_multiplicities(self, b, n, pos):
        """(b, n) uint8 host tensor: how often each point (column = storage position pos[i]) occurs in each of
        the next `b` resamples -- the same draws as `b` calls of resample_bootstrap(), i.e.
        rng.integers(0, n - 1, size=n) (two_pcf.py:275).  The C generator (csrc/hostrng.cu) continues numpy's
        PCG64 stream bit for bit and hands the advanced state back to self.rng."""
        bg = self.rng.bit_generator
        st = bg.state
        if st.get("bit_generator") == "PCG64" and 2 <= n <= (1 << 22):
            v, inc = st["state"]["state"], st["state"]["inc"]
            m64 = (1 << 64) - 1
            cs = np.array([v >> 64, v & m64, inc >> 64, inc & m64, st["has_uint32"], st["uinteger"]], dtype=np.uint64)
            mult = torch.empty((b, n), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
            _cabi.check(_cabi.load().tgp_bootstrap_multiplicities(
                cs.ctypes.data, n, b, None if pos is None else pos.ctypes.data, mult.data_ptr()),
                "tgp_bootstrap_multiplicities")
            st["state"]["state"] = (int(cs[0]) << 64) | int(cs[1])
            st["has_uint32"], st["uinteger"] = int(cs[4]), int(cs[5])
            bg.state = st
            return mult
        # other generators / very large catalogues: numpy draws, counted on the host
        idx = self.rng.integers(0, n - 1, size=(b, n))       # same stream as b successive calls
        if pos is not None:
            idx = pos[idx]
        rows = np.repeat(np.arange(b, dtype=np.int64), n)
        cnt = np.bincount(rows * n + idx.reshape(-1), minlength=b * n).reshape(b, n)
        if cnt.max() > 255:
            raise _cabi.TgpError("a point was drawn more than 255 times in one resample")
        return torch.as_tensor(cnt.astype(np.uint8))

    # Bootstrap batches share the pair geometry across resamples (tgp_bootbin_twod); False: every resample is an
    # independent weighted catalogue of one tgp_pairbin launch (the round-1 batch, kept as the cross-check).
    SHARED_BOOTSTRAP = True

    def _bootstrap_xi(self, n_bootstrap, batch_points=1 << 26):
        """xi of `n_bootstrap` resamples, shape (n_bootstrap, nb)."""
        from . import dist

        n = len(self.y)
        dev = backend.require_cuda()
        x = backend.to_device(np.asarray(self.X[:, 0], dtype=np.float64))
        yy = backend.to_device(np.asarray(self.X[:, 1], dtype=np.float64))
        val = backend.to_device(np.asarray(self.y, dtype=np.float64))
        err = np.asarray(self.y_err, dtype=np.float64)
        err_d = backend.to_device(err)
        # Hilbert-sort the base catalogue once: every resample is a sub-multiset in the same order
        order = backend.hilbert_order(x, yy) if self.anisotropic else None
        if order is not None:
            x, yy, val, err_d = x[order], yy[order], val[order], err_d[order]
        # storage position of every original point (identity without the Hilbert sort)
        pos = None
        if order is not None:
            pos = np.empty(n, dtype=np.int64)
            pos[order.cpu().numpy()] = np.arange(n, dtype=np.int64)
        # The reference drops the weights of a resample iff its errors sum to 0 (two_pcf.py:291-294): with all
        # errors zero every resample is unweighted, with all errors positive none is -- the two cases the shared
        # kernel covers; mixed inputs keep the per-catalogue batch.
        all_zero, all_pos = bool(np.all(err == 0)), bool(np.all(err > 0))
        if (self.anisotropic and self.SHARED_BOOTSTRAP and n >= 2 and (all_zero or all_pos)
                and not isinstance(self.group, dist.CabiComm)):
            return self._bootstrap_xi_shared(n_bootstrap, x, yy, val, None if all_zero else 1.0 / (err_d * err_d), pos, err_d)
        return self._bootstrap_xi_catalogues(n_bootstrap, x, yy, val, err_d, pos, batch_points)

    def _bootstrap_xi_shared(self, n_bootstrap, x, yy, val, w, pos, err_d, batch_bytes=1 << 31):
        """All resamples of a batch in ONE pass over the pairs of the base catalogue (tgp_bootbin_twod): the bin of a
        pair does not depend on the resample, only its weight m_b[i] m_b[j] does."""
        from . import dist

        n = int(x.numel())
        dev = x.device
        _, edges = self._bin_geometry()
        edges_d = self._device_edges(edges)
        rank, world = dist.rank_world(self.group)
        z = (val - val.mean()).contiguous()       # any fixed centring; the resample means are applied after the sums
        per_batch = max(32, min(int(n_bootstrap), int(batch_bytes // max(n, 1)) // 32 * 32))
        out, done = [], 0
        # a warp carries 32 resamples: a few left over beyond a multiple of 32 would cost a whole group of lanes, more
        # than the same resamples cost as independent catalogues (measured break-even: ~7 at N = 200k)
        tail = n_bootstrap % 32 if (n_bootstrap > 32 and world == 1) else 0
        tail = tail if tail <= 6 else 0
        n_shared = n_bootstrap - tail
        while done < n_shared:
            b = min(per_batch, n_shared - done)
            mult = self._draw_multiplicities(b, n, pos).to(dev, non_blocking=True)
            sums, delta = backend.bootbin_sums(x, yy, z, w, mult, edges_d, self.nbins, self.min_sep, self.max_sep,
                                               rank=rank, nranks=world)
            if world > 1:
                dist.allreduce_bins(self.group, sums)
            out.append(backend.bootbin_xi(sums, delta, self.nbins, b).cpu().numpy())
            done += b
        if tail:   # the random stream simply continues: resamples n_shared .. n_bootstrap-1
            out.append(self._bootstrap_xi_catalogues(tail, x, yy, val, err_d, pos))
        return np.concatenate(out, axis=0)

    def _bootstrap_xi_catalogues(self, n_bootstrap, x, yy, val, err_d, pos, batch_points=1 << 26):
        """Every resample as its own weighted catalogue (points with multiplicity > 0, weight m w) of one batched
        tgp_pairbin launch."""
        n = int(x.numel())
        dev = x.device
        out = []
        per_batch = max(1, int(batch_points // max(n, 1)))
        done = 0
        while done < n_bootstrap:
            b = min(per_batch, n_bootstrap - done)
            # multiplicities of `b` successive resample_bootstrap() draws, in storage order
            mult = self._draw_multiplicities(b, n, pos).to(dev, non_blocking=True).to(torch.float64)
            ybar = (mult * val).sum(dim=1) / n                      # mean of the resampled values
            # weights: None in the reference iff the resampled errors sum to 0 (two_pcf.py:291-294)
            esum = (mult * err_d).sum(dim=1)
            w_all = torch.where(err_d > 0, 1.0 / (err_d * err_d), torch.zeros_like(err_d))
            rows, cols = torch.nonzero(mult > 0, as_tuple=True)      # row-major: catalogues are contiguous
            lens = (mult > 0).sum(dim=1)
            offsets = torch.zeros(b + 1, dtype=torch.int64, device=dev)
            offsets[1:] = torch.cumsum(lens, 0)
            m = mult[rows, cols]
            unit = (esum == 0)[rows]
            pw = torch.where(unit, m, m * w_all[cols])
            pk = val[cols] - ybar[rows]
            xi, _ = self._pairbin(x[cols].contiguous(), yy[cols].contiguous(), pk.contiguous(), pw.contiguous(),
                                  offsets, int(lens.max().item()))
            out.append(xi)
            done += b
        return np.concatenate(out, axis=0)

    def return_2pcf(self, seed=610639139):
        """xi, its weight matrix (de-biased inverse bootstrap covariance, or 1/var(y) when isotropic),
        separations, bin coordinates and mask (two_pcf.py:364-391)."""
        xi, distance, coord, mask = self.comp_2pcf(self.X, self.y, self.y_err)
        if self.anisotropic:
            # number of resamples from Taylor et al. 2012 (https://doi.org/10.1093/mnras/stt270) eq. 35:
            # the resample count at which the de-biasing factor of the inverse covariance equals 2
            npixel = len(xi[mask])

            def f_bias(x):
                return (x - 1.0) / (x - npixel - 2.0) - 2.0

            n_bootstrap = int(optimize.fsolve(f_bias, npixel + 10)[0])
            xi_cov = self.comp_xi_covariance(n_bootstrap=n_bootstrap, mask=mask, seed=seed)
            bias_factor = (n_bootstrap - 1.0) / (n_bootstrap - npixel - 2.0)
            xi_weight = np.linalg.inv(xi_cov) * bias_factor
        else:
            xi_weight = np.eye(len(xi)) * 1.0 / np.var(self.y)
        return xi, xi_weight, distance, coord, mask

    def optimizer(self, kernel):
        """Fit `kernel` to the measured correlation function and return it (two_pcf.py:393-464)."""
        size_x = np.max(self.X[:, 0]) - np.min(self.X[:, 0])
        if self.ndim == 2:
            size_y = np.max(self.X[:, 1]) - np.min(self.X[:, 1])
            rho = float(len(self.X[:, 0])) / (size_x * size_y)
        else:
            size_y = 0.0
            rho = float(len(self.X[:, 0])) / size_x
        # defaults: min_sep = mean inter-point distance (isotropic) or 0 (anisotropic);
        # max_sep = half of the field diagonal
        if self.min_sep is None:
            self.min_sep = 0.0 if self.anisotropic else np.sqrt(1.0 / rho)
        if self.max_sep is None:
            self.max_sep = np.sqrt(size_x ** 2 + size_y ** 2) / 2.0

        xi, xi_weight, distance, coord, mask = self.return_2pcf()
        origin = np.zeros((1, 2))

        def PCF(param, k=kernel):
            return k.clone_with_theta(param)(coord, Y=origin)[:, 0]

        xi_mask = xi[mask]

        def chi2(param):
            residual = xi_mask - PCF(param)[mask]
            return residual.dot(xi_weight.dot(residual))

        if self.robust_fit:
            robust = robust_2dfit(kernel, xi, coord[:, 0], coord[:, 1], xi_weight, mask=mask)
            robust.minimize_minuit(p0=self.p0_robust_fit)
            kernel = copy.deepcopy(robust.kernel_fit)
            cst = robust.result[-1]
            self._results_robust = robust.result
        else:
            p0 = kernel.theta
            candidates = [optimize.fmin(chi2, p0, disp=False),
                          optimize.minimize(chi2, p0, method="L-BFGS-B")["x"]]
            values = [chi2(c) for c in candidates]
            kernel = kernel.clone_with_theta(candidates[values.index(min(values))])
            cst = 0

        self._2pcf = xi
        self._2pcf_weight = xi_weight
        self._2pcf_dist = distance
        self._2pcf_fit = PCF(kernel.theta) + cst
        self._2pcf_mask = mask
        self._kernel = copy.deepcopy(kernel)
        return kernel
