"""TEST INFRASTRUCTURE (oracle): pair sums of a vector field's 2-point functions in log-radius bins.

CPU restatement of the pair loop of /root/reference/treegp/utils.py:5-74 (`vcorr`), written as a row-blocked
accumulation so that the N(N-1)/2 index pairs the reference materialises (utils.py:38-47) never exist at once.
Checked against the reference's own all-pairs formulas in tests/test_cpu_host.py.  Only tests/ (and the CPU
stand-in of the reference-suite test) may import this; the product path is csrc/vcorr.cu.
"""
import numpy as np


def pair_sums(x, y, dx, dy, logrmin, dlogr, bins, block=512):
    """Per log-r bin: counts, sum log r, sum v1.v2*, sum v1 v2, sum v1 v2 exp(-2 i phi)."""
    z = x + 1j * y
    v = dx + 1j * dy
    counts = np.zeros(bins)
    s_logr = np.zeros(bins)
    s_plus = np.zeros(bins)
    s_z2 = np.zeros(bins, dtype=complex)
    s_minus = np.zeros(bins, dtype=complex)
    n = len(z)
    for a in range(0, n, block):
        b = min(n, a + block)
        dr = z[None, :] - z[a:b, None]                    # z_j - z_i
        jj = np.arange(n)[None, :] > np.arange(a, b)[:, None]
        dr = dr[jj]
        r2 = dr.real ** 2 + dr.imag ** 2
        ok = r2 > 0
        logdr = 0.5 * np.log(r2[ok])
        k = np.floor((logdr - logrmin) / dlogr).astype(np.int64)
        inb = (k >= 0) & (k < bins)
        k = k[inb]
        vi = np.broadcast_to(v[a:b, None], (b - a, n))[jj][ok][inb]
        vj = np.broadcast_to(v[None, :], (b - a, n))[jj][ok][inb]
        d = dr[ok][inb]
        counts += np.bincount(k, minlength=bins)
        s_logr += np.bincount(k, weights=logdr[inb], minlength=bins)
        s_plus += np.bincount(k, weights=(vi * np.conj(vj)).real, minlength=bins)
        vv = vi * vj
        s_z2 += np.bincount(k, weights=vv.real, minlength=bins) + 1j * np.bincount(k, weights=vv.imag, minlength=bins)
        rot = vv * np.conj(d) ** 2 / r2[ok][inb]
        s_minus += np.bincount(k, weights=rot.real, minlength=bins) + 1j * np.bincount(k, weights=rot.imag, minlength=bins)
    return counts, s_logr, s_plus, s_z2, s_minus
