"""TEST INFRASTRUCTURE: CPU stand-ins for backend.bootbin_sums / backend.bootbin_xi (the shared-geometry bootstrap
batch, tgp_bootbin_twod) built on the oracle, so that the host logic of two_pcf._bootstrap_xi_shared -- batching,
multiplicity stream, centring, the optional all-reduce -- runs in a container without a GPU.  The packed `sums` here
are additive per-rank partial sums {sum w, sum w k k} per resample and bin (k already centred on the resample mean);
the product never imports this module."""
import numpy as np
import torch


def install(backend, po, slab=None):
    def bootbin_sums(px, py, pz, pw, mult, edges, nbins, min_sep, max_sep, rank=0, nranks=1):
        x, y, z = (np.asarray(t.numpy(), dtype=float) for t in (px, py, pz))
        w = np.ones_like(x) if pw is None else np.asarray(pw.numpy(), dtype=float)
        m = np.asarray(mult.numpy(), dtype=float)
        nboot, n = m.shape
        nb = nbins * nbins
        out = np.zeros((2, nboot, nb))
        delta = m @ z / n
        for b in range(nboot):
            sel = m[b] > 0
            cnt = int(sel.sum())
            if cnt < 2:
                continue
            rows = None if slab is None else slab(cnt, rank, nranks)
            r = po.pairbin(x[sel], y[sel], z[sel] - delta[b], m[b][sel] * w[sel], min_sep, max_sep, nbins, "TwoD",
                           rows=rows)
            out[0, b], out[1, b] = r["weight"], r["sumwkk"]
        return torch.as_tensor(out), torch.as_tensor(delta)

    def bootbin_xi(sums, delta, nbins, nboot, want_sumw=False):
        sw, swkk = sums[0].numpy(), sums[1].numpy()
        with np.errstate(invalid="ignore", divide="ignore"):
            xi = np.where(sw != 0, swkk / sw, 0.0)
        return (torch.as_tensor(xi), torch.as_tensor(sw)) if want_sumw else torch.as_tensor(xi)

    backend.bootbin_sums = bootbin_sums
    backend.bootbin_xi = bootbin_xi
