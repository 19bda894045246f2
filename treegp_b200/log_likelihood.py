"""Marginal log-likelihood of a Gaussian process and its maximisation, evaluated on the B200.

Host-side mirror of /root/reference/treegp/log_likelihood.py (class ``log_likelihood``: same
constructor, ``log_likelihood(kernel)``, ``optimizer(kernel)``, attributes ``_kernel`` and ``_logL``).
The body of one evaluation (log_likelihood.py:29-37: K build, Cholesky, solve, chi2, log-det) is a
single C-ABI call, ``tgp_loglike``; X, y, y_err^2 and the N x N workspace stay resident on the
device for the whole L-BFGS-B search (40-120 evaluations per fit).
"""
import copy

import numpy as np
from scipy import optimize

from . import backend
from .kernels import lower_kernel


class log_likelihood(object):
    """Return and optimize (if requested) the log likelihood of gaussian process.

    :param X:      Coordinates of the field.  (n_samples, 1 or 2)
    :param y:      Values of the field.  (n_samples)
    :param y_err:  Error of y. (n_samples)
    """

    def __init__(self, X, y, y_err):
        self.X = X
        self.ndata = len(self.X[:, 0])
        self.y = y
        self.y_err = y_err
        self._dev = None
        self.n_evaluations = 0

    def _device_state(self):
        if self._dev is None:
            Xd = backend.as_points(self.X)
            yd = backend.to_device(np.asarray(self.y, dtype=np.float64).reshape(-1))
            e2 = backend.to_device(np.asarray(self.y_err, dtype=np.float64).reshape(-1) ** 2)
            work = backend.alloc_matrix(self.ndata + 1, self.ndata, Xd.device)  # +1 row: y rides through potrf
            self._dev = (Xd, yd, e2, work)
        return self._dev

    def log_likelihood(self, kernel):
        """
        Return of log likehood of gaussian process
        for given hyperparameters.

        :param kernel: Sklearn kernel object.
        """
        Xd, yd, e2, work = self._device_state()
        desc = lower_kernel(kernel, Xd.shape[1])
        # a non-finite hyper-parameter cannot be factorised: the reference's try/except returns -inf
        # (log_likelihood.py:38-39); same outcome here without launching anything
        if not np.all(np.isfinite([desc.amp, desc.m00, desc.m01, desc.m11])):
            return -np.inf
        out, info, _, _ = backend.loglike(Xd, yd, e2, desc, work=work, want_alpha=False)
        self.n_evaluations += 1
        return float(out[0].item())  # -inf when the factorisation failed (info != 0)

    def optimizer(self, kernel):
        """
        Fit hyperparameter using maximum likelihood fit.
        Used minimization with L-BFGS-B method from scipy.

        :param kernel: sklearn.gaussian_process kernel.
        """

        def minus_logl(theta):
            return -self.log_likelihood(kernel.clone_with_theta(theta))

        # unbounded L-BFGS-B with scipy's forward-difference gradient, as log_likelihood.py:56-57
        best = optimize.minimize(minus_logl, kernel.theta, method="L-BFGS-B")["x"]
        kernel = kernel.clone_with_theta(best)
        self._kernel = copy.deepcopy(kernel)
        self._logL = self.log_likelihood(self._kernel)
        return kernel
