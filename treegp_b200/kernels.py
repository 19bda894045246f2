"""Covariance functions with the scikit-learn ``Kernel`` protocol, evaluated on the B200.

Host-side mirror of /root/reference/treegp/kernels.py: the same public names (``eval_kernel``,
``AnisotropicRBF``, ``VonKarman``, ``AnisotropicVonKarman``), constructor arguments, ``theta``
parameterisation (log-Cholesky of the inverse metric, kernels.py:163-179 / :397-413) and ``bounds``,
so kernel strings such as ``"4.0 * AnisotropicRBF(invLam=array([[...]]))"`` keep working.  What
changes is where the numbers come from: ``__call__`` lowers the kernel to the POD descriptor of
include/treegp_b200.h and runs the tiled CUDA build (csrc/kmat.cu) instead of
pdist/cdist + scipy.special.kv.

``lower_kernel`` is the bridge used by GPInterpolation / log_likelihood / two_pcf to keep whole
solves on the device.
"""
import numpy as np
from sklearn.gaussian_process import kernels as _skk
from sklearn.gaussian_process.kernels import (
    ConstantKernel,
    Hyperparameter,
    Kernel,
    Matern,
    NormalizedKernelMixin,
    Product,
    RBF,
    StationaryKernelMixin,
)

from . import _cabi
from ._cabi import TgpKernel

__all__ = ["eval_kernel", "AnisotropicRBF", "VonKarman", "AnisotropicVonKarman", "lower_kernel"]


# ------------------------------------------------------------------------------------------------
# device evaluation shared by the three kernels
# ------------------------------------------------------------------------------------------------
def _device_call(desc, X, Y):
    """Evaluate amp * f on the GPU and return a numpy array, (N, N) or (N, M)."""
    from . import backend

    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    if Y is None:
        n = X.shape[0]
        out = backend.kmat_sym(X, desc)
        return out[:, :n].cpu().numpy()
    Y = np.atleast_2d(np.asarray(Y, dtype=np.float64))
    out = backend.kmat_cross(X, Y, desc)
    return out[:, : Y.shape[0]].cpu().numpy()


def _descriptor(family, ndim, amp, metric):
    m = np.asarray(metric, dtype=np.float64)
    if ndim == 1:
        return TgpKernel(family, 1, float(amp), float(m[0, 0]), 0.0, 0.0)
    if ndim != 2:
        raise _cabi.TgpError("the B200 backend handles 1-D and 2-D coordinates; got ndim=%d" % ndim)
    if abs(m[0, 1] - m[1, 0]) > 1e-12 * max(1.0, abs(m[0, 1])):
        raise ValueError("inverse metric must be symmetric")
    return TgpKernel(family, 2, float(amp), float(m[0, 0]), float(0.5 * (m[0, 1] + m[1, 0])), float(m[1, 1]))


# ------------------------------------------------------------------------------------------------
# kernels parameterised by a full inverse metric
# ------------------------------------------------------------------------------------------------
class _MetricKernel(StationaryKernelMixin, NormalizedKernelMixin, Kernel):
    """Stationary kernel k(x, y) = f((x-y)^T invLam (x-y)) with theta = log-Cholesky(invLam).

    theta[:n] are the logs of the diagonal of L, theta[n:] its strict lower triangle (row-major),
    invLam = L L^T -- the unconstrained parameterisation of kernels.py:71-83.
    """

    _family = None

    def __init__(self, invLam=None, scale_length=None, bounds=(-5, 5)):
        if scale_length is not None:
            if invLam is not None:
                raise TypeError("Cannot set both invLam and scale_length in %s." % type(self).__name__)
            invLam = np.diag(1.0 / np.array(scale_length) ** 2)
        if invLam is None:
            raise TypeError("Exactly one of invLam and scale_length must be provided.")
        self.ndim = invLam.shape[0]
        self.ntheta = self.ndim * (self.ndim + 1) // 2
        self._d = np.diag_indices(self.ndim)
        self._t = np.tril_indices(self.ndim, -1)
        self.set_params(invLam)
        b = np.array(bounds)
        if b.ndim == 1:
            b = np.tile(b, (self.ntheta, 1))
        assert b.shape == (self.ntheta, 2)
        self._bounds = b

    # -- sklearn plumbing ---------------------------------------------------------------------
    @property
    def hyperparameter_cholesky_factor(self):
        return Hyperparameter("CholeskyFactor", "numeric", (1e-5, 1e5), int(self.ntheta))

    def get_params(self, deep=True):
        return {"invLam": self.invLam}

    def set_params(self, invLam=None):
        if invLam is None:
            return
        self.invLam = invLam
        self._L = np.linalg.cholesky(self.invLam)
        self._theta = np.concatenate([np.log(self._L[self._d]), self._L[self._t]])

    @property
    def theta(self):
        return self._theta

    @theta.setter
    def theta(self, theta):
        theta = np.asarray(theta, dtype=float)
        L = np.zeros((self.ndim, self.ndim))
        L[self._d] = np.exp(theta[: self.ndim])
        L[self._t] = theta[self.ndim:]
        self._theta = theta
        self._L = L
        self.invLam = L @ L.T

    @property
    def bounds(self):
        return self._bounds

    def __repr__(self):
        return "{0}(invLam={1!r})".format(type(self).__name__, self.invLam)

    # -- evaluation ---------------------------------------------------------------------------
    def descriptor(self, amp=1.0):
        return _descriptor(self._family, self.ndim, amp, self.invLam)

    def __call__(self, X, Y=None, eval_gradient=False):
        if eval_gradient:
            return self._with_gradient(X, Y)
        return _device_call(self.descriptor(), X, Y)

    def _with_gradient(self, X, Y):
        raise ValueError("Gradient can not be evaluated.")


class AnisotropicRBF(_MetricKernel):
    """Squared-exponential kernel with an arbitrary inverse covariance:
    k = exp(-1/2 (x-y)^T invLam (x-y)).  Mirror of kernels.py:62-186.

    :param invLam:        inverse covariance matrix (exactly one of invLam / scale_length).
    :param scale_length:  per-axis scale lengths; invLam = diag(1 / scale_length^2).
    :param bounds:        bounds on theta, a pair or an (ntheta, 2) array.
    """

    _family = _cabi.FAM_RBF

    def _with_gradient(self, X, Y):
        # dK/dtheta_k = -1/2 K * (dx^T dInvLam/dtheta_k dx)   (kernels.py:128-150).  Not on the hot
        # path (log_likelihood.py:57 passes no jac): K comes from the device, the contraction with
        # the ntheta small matrices is done on the host.
        if Y is not None:
            raise ValueError("Gradient can only be evaluated when Y is None.")
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        K = _device_call(self.descriptor(), X, None)
        n, nth = self.ndim, self.ntheta
        dL = np.zeros((nth, n, n))
        for a in range(n):
            dL[a, a, a] = self._L[a, a]
        for a, (i, j) in enumerate(zip(*self._t)):
            dL[n + a, i, j] = 1.0
        half = dL @ self._L.T
        dM = half + np.transpose(half, (0, 2, 1))
        dX = X[:, None, :] - X[None, :, :]
        quad = np.einsum("pqi,kij,pqj->pqk", dX, dM, dX)
        return K, -0.5 * K[:, :, None] * quad


class AnisotropicVonKarman(_MetricKernel):
    """von Karman (Kolmogorov-turbulence) correlation with an arbitrary inverse covariance:
    k = d^(5/6) K_{5/6}(2 pi d) / lim0 with d the Mahalanobis distance.  Mirror of kernels.py:304-420.
    """

    _family = _cabi.FAM_VONKARMAN


# ------------------------------------------------------------------------------------------------
# isotropic von Karman with a length scale
# ------------------------------------------------------------------------------------------------
class VonKarman(StationaryKernelMixin, NormalizedKernelMixin, Kernel):
    """k = (r/l)^(5/6) K_{5/6}(2 pi r / l) / lim0.  Mirror of kernels.py:189-301.

    :param length_scale:         float (isotropic) -- an array of per-axis scales is accepted for
                                 scalar-equivalent shapes as in sklearn.
    :param length_scale_bounds:  bounds on length_scale.
    """

    def __init__(self, length_scale=1.0, length_scale_bounds=(1e-5, 1e5)):
        self.length_scale = length_scale
        self.length_scale_bounds = length_scale_bounds

    @property
    def anisotropic(self):
        return np.iterable(self.length_scale) and len(self.length_scale) > 1

    @property
    def hyperparameter_length_scale(self):
        if self.anisotropic:
            return Hyperparameter("length_scale", "numeric", self.length_scale_bounds, len(self.length_scale))
        return Hyperparameter("length_scale", "numeric", self.length_scale_bounds)

    def descriptor(self, ndim, amp=1.0):
        ls = np.ravel(np.asarray(self.length_scale, dtype=float))
        if ls.size == 1:
            ls = np.repeat(ls, ndim)
        if ls.size != ndim:
            raise ValueError("length_scale has %d entries for %d-D coordinates" % (ls.size, ndim))
        return _descriptor(_cabi.FAM_VONKARMAN, ndim, amp, np.diag(1.0 / ls ** 2))

    def __call__(self, X, Y=None, eval_gradient=False):
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        K = _device_call(self.descriptor(X.shape[1]), X, Y)
        if not eval_gradient:
            return K
        if Y is not None:
            raise ValueError("Gradient can only be evaluated when Y is None.")
        if self.hyperparameter_length_scale.fixed:
            return K, np.empty((X.shape[0], X.shape[0], 0))
        if self.anisotropic:
            raise ValueError("Gradient can only be evaluated with isotropic VonKarman kernel for the moment.")
        # the reference's (approximate) expression K * r  (kernels.py:283-285)
        r = np.sqrt(((X[:, None, :] - X[None, :, :]) ** 2).sum(-1))
        return K, (K * r)[:, :, None]

    def __repr__(self):
        if self.anisotropic:
            return "{0}(length_scale=[{1}])".format(
                type(self).__name__, ", ".join("{0:.3g}".format(v) for v in self.length_scale))
        return "{0}(length_scale={1:.3g})".format(type(self).__name__, np.ravel(self.length_scale)[0])


# ------------------------------------------------------------------------------------------------
# kernel strings
# ------------------------------------------------------------------------------------------------
def _kernel_namespace():
    ns = {}
    stack = list(Kernel.__subclasses__())
    while stack:
        cls = stack.pop()
        ns.setdefault(cls.__name__, cls)
        stack.extend(cls.__subclasses__())
    for name in dir(_skk):
        obj = getattr(_skk, name)
        if isinstance(obj, type) and issubclass(obj, Kernel):
            ns.setdefault(name, obj)
    ns["array"] = np.array
    return ns


def eval_kernel(kernel):
    """Turn a kernel string (a sklearn / treegp kernel ``repr`` or any expression over them, e.g.
    ``"2.0**2 * RBF(0.5)"``) into a kernel object.  Mirror of kernels.py:17-59: every subclass of
    sklearn's ``Kernel`` plus numpy's ``array`` is in scope."""
    try:
        k = eval(kernel, _kernel_namespace())
    except Exception as e:
        raise RuntimeError("Failed to evaluate kernel string {0!r}.  Original exception: {1}".format(kernel, e))
    if isinstance(getattr(k, "theta", None), property) or not isinstance(k, Kernel):
        raise TypeError("String provided was not initialized properly")
    return k


# ------------------------------------------------------------------------------------------------
# lowering sklearn kernel trees to the device descriptor
# ------------------------------------------------------------------------------------------------
def _iso_metric(length_scale, ndim, what):
    ls = np.ravel(np.asarray(length_scale, dtype=float))
    if ls.size == 1:
        ls = np.repeat(ls, ndim)
    if ls.size != ndim:
        raise ValueError("%s length_scale has %d entries for %d-D coordinates" % (what, ls.size, ndim))
    return np.diag(1.0 / ls ** 2)


def lower_kernel(kernel, ndim):
    """sklearn kernel tree -> ``TgpKernel``.

    Supported: an optional product with any number of ``ConstantKernel`` factors around one of
    RBF, Matern(nu in {0.5, 1.5, 2.5, inf}), VonKarman, AnisotropicRBF, AnisotropicVonKarman --
    the surface the reference's tests, docs and notebooks exercise.  Anything
    else raises: there is no CPU fallback to silently take over.
    """
    amp = 1.0
    base = None
    stack = [kernel]
    while stack:
        k = stack.pop()
        if isinstance(k, Product):
            stack.extend([k.k1, k.k2])
        elif isinstance(k, ConstantKernel):
            amp *= float(k.constant_value)
        elif base is None:
            base = k
        else:
            raise _cabi.TgpError("unsupported kernel for the B200 backend (more than one non-constant factor): %r"
                                 % (kernel,))
    if base is None:
        raise _cabi.TgpError("unsupported kernel for the B200 backend (no stationary factor): %r" % (kernel,))
    if isinstance(base, _MetricKernel):
        if base.ndim != ndim:
            raise ValueError("kernel is %d-D but coordinates are %d-D" % (base.ndim, ndim))
        return base.descriptor(amp)
    if isinstance(base, VonKarman):
        return base.descriptor(ndim, amp)
    if isinstance(base, Matern):  # Matern subclasses RBF: test it first
        fam = {0.5: _cabi.FAM_MATERN12, 1.5: _cabi.FAM_MATERN32, 2.5: _cabi.FAM_MATERN52,
               np.inf: _cabi.FAM_RBF}.get(float(base.nu))
        if fam is None:
            raise _cabi.TgpError("Matern(nu=%r) is not supported by the B200 backend (nu in 0.5, 1.5, 2.5, inf)"
                                 % (base.nu,))
        return _descriptor(fam, ndim, amp, _iso_metric(base.length_scale, ndim, "Matern"))
    if isinstance(base, RBF):
        return _descriptor(_cabi.FAM_RBF, ndim, amp, _iso_metric(base.length_scale, ndim, "RBF"))
    raise _cabi.TgpError("unsupported kernel for the B200 backend: %r" % (kernel,))
