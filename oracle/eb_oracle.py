"""TEST INFRASTRUCTURE (oracle): pair sums of a vector field's 2-point functions in log-radius bins.

CPU restatement of the pair loop of /root/reference/treegp/utils.py:5-74 (`vcorr`), written as a row-blocked
accumulation so that the N(N-1)/2 index pairs the reference materialises (utils.py:38-47) never exist at once.
Every pair is placed exactly as the reference places it: logdr = np.log(np.absolute(dr)) (utils.py:50) binned by
np.histogram(logdr, bins=bins, range=(logrmin, logrmin + bins * dlogr)) (utils.py:52-55) -- numpy's own functions,
so the per-bin counts are the reference's bit for bit.  PINNED: tests/test_oracle_golden.py compares it with
outputs of the reference's `vcorr` stored by tests/golden/make_golden_r2.py.  Only tests/ (and the CPU stand-in
of the reference-suite test) may import this; the product path is csrc/vcorr.cu.
"""
import numpy as np


def pair_sums(x, y, dx, dy, logrmin, dlogr, bins, block=512):
    """Per log-r bin: counts, sum log r, sum v1.v2*, sum v1 v2, sum v1 v2 exp(-2 i phi)."""
    x, y, dx, dy = (np.asarray(a, dtype=np.float64) for a in (x, y, dx, dy))
    v = dx + 1j * dy
    hrange = (logrmin, logrmin + bins * dlogr)                      # utils.py:53
    counts = np.zeros(bins)
    s_logr = np.zeros(bins)
    s_plus = np.zeros(bins)
    s_z2 = np.zeros(bins, dtype=complex)
    s_minus = np.zeros(bins, dtype=complex)
    n = len(x)

    def hist(logdr, w=None):
        return np.histogram(logdr, bins=bins, range=hrange, weights=w)[0]

    for a in range(0, n, block):
        b = min(n, a + block)
        ii, jj = np.nonzero(np.arange(n)[None, :] > np.arange(a, b)[:, None])   # all pairs i < j of this row block
        i1, i2 = ii + a, jj
        dr = 1j * (y[i2] - y[i1])                                   # utils.py:46-47
        dr += x[i2] - x[i1]
        with np.errstate(divide="ignore"):
            logdr = np.log(np.absolute(dr))                         # utils.py:50 (-inf for coincident points)
        fin = np.isfinite(logdr)                                    # np.histogram rejects non-finite weights * 0
        dr, logdr, i1, i2 = dr[fin], logdr[fin], i1[fin], i2[fin]
        counts += hist(logdr)
        s_logr += hist(logdr, logdr)
        s_plus += hist(logdr, dx[i1] * dx[i2] + dy[i1] * dy[i2])    # utils.py:59-60
        vv = v[i1] * v[i2]
        s_z2 += hist(logdr, vv)                                     # utils.py:61-62
        vv = vv * np.conj(dr) * np.conj(dr) / (dr.real * dr.real + dr.imag * dr.imag)   # utils.py:65-68
        s_minus += hist(logdr, vv)
    return counts, s_logr, s_plus, s_z2, s_minus
