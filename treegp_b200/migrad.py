"""A small variable-metric minimiser standing in for Minuit's MIGRAD.

The reference drives its robust 2-D fit with ``iminuit.Minuit(chi2, p0).migrad()`` and reads
``params[...].value`` and ``accurate`` (/root/reference/treegp/two_pcf.py:150-176).  iminuit (Minuit2,
C++) is a third-party dependency that is not part of this image, and the fit is host control flow over
three parameters and a <= 441-point model ("stays on host"), so this module
provides the same contract in ~150 lines of numpy:

* numerical first derivatives (central differences, per-parameter adaptive steps);
* inverse-Hessian estimate V updated with the Davidon-Fletcher-Powell / BFGS rank-two formulas;
* line search by successive parabolic interpolation along -V g, tolerant of ``inf`` plateaus (the
  chi-square returns inf for |g| > 1, two_pcf.py:105-106,128-133);
* convergence on the estimated distance to minimum  EDM = g^T V g / 2 < 0.002 * tol * errordef
  (Minuit's criterion, tol = 0.1 -> 2e-4), verified against a full finite-difference Hessian;
* ``accurate`` is True when the run converged and that Hessian is positive definite -- the flag the
  caller uses to decide on a restart from another starting point.
"""
import numpy as np


class Migrad(object):
    def __init__(self, fcn, start, errordef=1.0, tol=0.1, max_calls=4000, fcn_batch=None):
        self.fcn = fcn
        self.fcn_batch = fcn_batch          # optional: values of several points at once (one device launch)
        self.values = np.array(start, dtype=float)
        self.errordef = float(errordef)
        self.tol = float(tol)
        self.max_calls = int(max_calls)
        self.nfcn = 0
        self.fval = None
        self.edm = np.inf
        self.valid = False
        self.accurate = False
        self.covariance = None

    # -- helpers -------------------------------------------------------------------------------
    def _f(self, x):
        self.nfcn += 1
        v = self.fcn(np.asarray(x, dtype=float))
        v = float(v)
        return v if np.isfinite(v) else np.inf

    def _fmany(self, xs):
        """Values at several points: one call of fcn_batch if there is one, else point by point."""
        if self.fcn_batch is None or len(xs) < 2:
            return [self._f(x) for x in xs]
        self.nfcn += len(xs)
        vals = np.asarray(self.fcn_batch(np.asarray(xs, dtype=float)), dtype=float)
        return [float(v) if np.isfinite(v) else np.inf for v in vals]

    def _steps(self, x):
        # like Minuit's default initial errors: 1 % of the value, 0.01 for parameters at zero
        return np.where(x != 0.0, 1e-2 * np.abs(x), 1e-2)

    def _gradient(self, x, f0, h):
        n = len(x)
        g, g2 = np.zeros(n), np.zeros(n)
        first = None
        if self.fcn_batch is not None:      # the 2 n probes of the first attempt in one launch
            pts = []
            for i in range(n):
                xp, xm = x.copy(), x.copy()
                xp[i] += h[i]
                xm[i] -= h[i]
                pts += [xp, xm]
            first = self._fmany(pts)
        for i in range(n):
            hi = h[i]
            if first is not None and np.isfinite(first[2 * i]) and np.isfinite(first[2 * i + 1]):
                fp, fm = first[2 * i], first[2 * i + 1]
                g[i] = (fp - fm) / (2.0 * hi)
                g2[i] = (fp + fm - 2.0 * f0) / (hi * hi)
                continue
            for _ in range(8):  # shrink until both probes are finite
                xp, xm = x.copy(), x.copy()
                xp[i] += hi
                xm[i] -= hi
                fp, fm = self._f(xp), self._f(xm)
                if np.isfinite(fp) and np.isfinite(fm):
                    break
                hi *= 0.25
            if not (np.isfinite(fp) and np.isfinite(fm)):
                return None, None
            g[i] = (fp - fm) / (2.0 * hi)
            g2[i] = (fp + fm - 2.0 * f0) / (hi * hi)
        return g, g2

    def _hessian(self, x, f0, h):
        n = len(x)
        H = np.zeros((n, n))
        # all 2 n + n (n - 1) / 2 probes first (one launch with fcn_batch), in the order they are used below
        pts = []
        for i in range(n):
            e = np.zeros(n)
            e[i] = h[i]
            pts += [x + e, x - e]
        for i in range(n):
            for j in range(i):
                e = np.zeros(n)
                e[i], e[j] = h[i], h[j]
                pts.append(x + e)
        vals = self._fmany(pts)
        fp, fm = np.array(vals[0:2 * n:2]), np.array(vals[1:2 * n:2])
        for i in range(n):
            H[i, i] = (fp[i] + fm[i] - 2.0 * f0) / (h[i] * h[i])
        k = 2 * n
        for i in range(n):
            for j in range(i):
                H[i, j] = H[j, i] = (vals[k] - fp[i] - fp[j] + f0) / (h[i] * h[j])
                k += 1
        return H

    def _line_search(self, x, f0, p, slope):
        """Minimise f(x + t p) for t > 0 by parabolic interpolation; returns (t, f)."""
        best_t, best_f = 0.0, f0
        t = 1.0
        ft = self._f(x + t * p)
        tries = 0
        while not np.isfinite(ft) and tries < 12:  # walked out of the allowed region
            t *= 0.25
            ft = self._f(x + t * p)
            tries += 1
        if not np.isfinite(ft):
            return 0.0, f0
        if ft < best_f:
            best_t, best_f = t, ft
        for _ in range(6):
            # parabola through (0, f0) with slope `slope` and (t, ft)
            denom = ft - f0 - slope * t
            if denom <= 0:  # no curvature information: expand
                t_new = 2.5 * t
            else:
                t_new = -slope * t * t / (2.0 * denom)
                t_new = min(max(t_new, 0.05 * t), 5.0 * t)
            f_new = self._f(x + t_new * p)
            if np.isfinite(f_new) and f_new < best_f:
                improved = best_f - f_new
                best_t, best_f = t_new, f_new
                if improved < 1e-3 * max(self.edm_goal, abs(best_f) * 1e-12):
                    break
            elif np.isfinite(f_new) and abs(t_new - t) < 0.1 * t:
                break
            if np.isfinite(f_new):
                t, ft = t_new, f_new
            else:
                t = 0.5 * (best_t + t_new) if best_t > 0 else 0.25 * t_new
                ft = self._f(x + t * p)
                if not np.isfinite(ft):
                    break
                if ft < best_f:
                    best_t, best_f = t, ft
        return best_t, best_f

    # -- driver --------------------------------------------------------------------------------
    def migrad(self):
        x = self.values.copy()
        n = len(x)
        self.edm_goal = 0.002 * self.tol * self.errordef
        f = self._f(x)
        if not np.isfinite(f):
            self.fval = f
            return self
        h = self._steps(x)
        g, g2 = self._gradient(x, f, h)
        if g is None:
            self.fval = f
            return self
        V = np.diag([1.0 / v if v > 0 else (hh * hh) / (2.0 * self.errordef) for v, hh in zip(g2, h)])
        converged = False
        verified = False
        for _ in range(200):
            if self.nfcn > self.max_calls:
                break
            p = -V.dot(g)
            slope = float(g.dot(p))
            if slope >= 0:  # V lost positive definiteness: restart from the diagonal
                V = np.diag([1.0 / v if v > 0 else 1.0 for v in g2])
                p = -V.dot(g)
                slope = float(g.dot(p))
                if slope >= 0:
                    break
            t, f_new = self._line_search(x, f, p, slope)
            if t == 0.0:
                # no progress along p: accept only if the gradient says we are there
                self.edm = 0.5 * float(g.dot(V.dot(g)))
                converged = self.edm < self.edm_goal
                if converged or verified:
                    break
                verified = True
                H = self._hessian(x, f, np.maximum(1e-3 * h, 1e-7))
                try:
                    V = np.linalg.inv(H)
                    np.linalg.cholesky(H)
                except np.linalg.LinAlgError:
                    V = np.diag([1.0 / v if v > 0 else 1.0 for v in g2])
                continue
            dx = t * p
            x_new = x + dx
            h = np.maximum(np.minimum(h, np.abs(dx) + 1e-3 * h), 1e-8 * np.maximum(np.abs(x_new), 1e-3))
            g_new, g2_new = self._gradient(x_new, f_new, h)
            if g_new is None:
                break
            dg = g_new - g
            dxdg = float(dx.dot(dg))
            Vdg = V.dot(dg)
            gVg = float(dg.dot(Vdg))
            if dxdg > 0 and gVg > 0:
                V = V + np.outer(dx, dx) / dxdg - np.outer(Vdg, Vdg) / gVg
                if dxdg > gVg:  # BFGS correction term (Minuit's "delgam > gvg" branch)
                    u = dx / dxdg - Vdg / gVg
                    V = V + gVg * np.outer(u, u)
            x, f, g, g2 = x_new, f_new, g_new, g2_new
            self.edm = 0.5 * float(g.dot(V.dot(g)))
            if self.edm < self.edm_goal:
                # verify with a full Hessian (Minuit strategy 1 runs HESSE when the DFP estimate is unsure)
                H = self._hessian(x, f, np.maximum(1e-2 * self._steps(x), 1e-7))
                try:
                    np.linalg.cholesky(H)
                    Vh = np.linalg.inv(H)
                    edm_h = 0.5 * float(g.dot(Vh.dot(g)))
                    V = Vh
                    self.edm = edm_h
                    if edm_h < 10.0 * self.edm_goal:
                        converged = True
                        verified = True
                        break
                except np.linalg.LinAlgError:
                    pass
        self.values = x
        self.fval = f
        self.valid = bool(converged)
        ok = False
        if converged:
            try:
                H = np.linalg.inv(V)
                np.linalg.cholesky(0.5 * (H + H.T))
                ok = bool(np.all(np.isfinite(V)))
            except np.linalg.LinAlgError:
                ok = False
        self.accurate = ok
        self.covariance = 2.0 * self.errordef * V
        return self
