"""CPU, world_size 2 over gloo: the multi-GPU host logic (rank discovery, slabs, all-reduce of the bin
sums before xi is formed, gathering sharded predictions).  The device kernel is replaced by the oracle
restricted to this rank's share of the pair-matrix rows -- the same additive split tgp_pairbin's
(tile_rank, tile_nranks) arguments implement on the GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import treegp_b200 as treegp
        from oracle import pairbin_oracle as po
        from treegp_b200 import _cabi, backend, dist

        # ---- stand-ins for the device: CPU tensors, oracle restricted to this rank's rows ----
        backend.require_cuda = lambda: torch.device("cpu")
        backend.hilbert_order = lambda px, py: torch.arange(px.numel())

        def fake_pairbin(px, py, pk, pw, cat_off, max_len, bin_type, edges, nbins, min_sep, max_sep, rank=0, nranks=1):
            n = px.numel()
            lo, hi = dist.slab(n, rank, nranks)
            bt = "TwoD" if bin_type == _cabi.BIN_TWOD else "Log"
            r = po.pairbin(px.numpy(), py.numpy(), pk.numpy(), None if pw is None else pw.numpy(), min_sep, max_sep,
                           nbins, bt, rows=(lo, hi))
            t = lambda a: torch.as_tensor(a).reshape(1, -1).clone()
            return t(r["npairs"]), t(r["weight"]), t(r["sumwkk"]), (t(r["sumwr"]) if bt == "Log" else None)

        def fake_packed(*a, **kw):
            c, sw, swkk, swr = fake_pairbin(*a, **kw)
            return torch.stack([c.to(torch.int64).view(torch.float64), sw, swkk] + ([] if swr is None else [swr]))

        backend.pairbin_packed = fake_packed
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from refsuite import bootbin_standin

        bootbin_standin.install(backend, po, slab=dist.slab)
        assert dist.rank_world(None) == (rank, world)
        assert dist.rank_world(False) == (0, 1) and dist.rank_world(dist.WORLD) == (rank, world)
        # sharding is opt-in: by default every process counts all pairs of its own catalogue
        assert treegp.two_pcf(np.zeros((3, 2)), np.zeros(3), np.zeros(3), 0.0, 1.0).group is False

        rng = np.random.default_rng(5)
        n = 900
        X = rng.uniform(-10, 10, size=(n, 2))
        y = rng.normal(size=n)
        y_err = np.full(n, 0.2)
        for aniso, mn, mx, nb in ((True, 0.0, 3.0, 11), (False, 0.2, 4.0, 10)):
            t = treegp.two_pcf(X, y, y_err, mn, mx, nbins=nb, anisotropic=aniso)
            t.group = dist.WORLD
            xi, dist_, coord, mask = t.comp_2pcf(X, y, y_err)
            rxi, rdist, _, rmask = po.comp_2pcf(X, y, y_err, mn, mx, nb, aniso)
            np.testing.assert_allclose(xi, rxi, rtol=0, atol=1e-12)
            np.testing.assert_allclose(dist_, rdist, rtol=1e-12)
            full = po.pairbin(X[:, 0], X[:, 1], y - y.mean(), 1 / y_err ** 2, mn, mx, nb, "TwoD" if aniso else "Log")
            np.testing.assert_array_equal(t._last_npairs[0], full["npairs"])  # counts exact for any world size
            if aniso:
                # the bootstrap batch shards the same way: per-rank partial sums, one all-reduce, identical covariance
                B = 5
                cov = t.comp_xi_covariance(n_bootstrap=B, mask=mask, seed=99)
                r = np.random.default_rng(99)
                xis = []
                for _ in range(B):
                    ind = r.integers(0, n - 1, size=n)
                    xis.append(po.comp_2pcf(X[ind], y[ind], y_err[ind], mn, mx, nb, aniso)[0][mask])
                d = np.array(xis) - np.mean(xis, axis=0)
                np.testing.assert_allclose(cov, d.T @ d / (B - 1.0), rtol=0, atol=1e-12)

        # ---- distributed forward-difference gradient of the likelihood search (opt-in) ----
        from treegp_b200 import log_likelihood as llmod
        import treegp_b200.log_likelihood  # noqa: F401
        import sys
        mod = sys.modules["treegp_b200.log_likelihood"]
        calls = []

        def fake_loglike(self, kernel):   # smooth stand-in for the device evaluation
            th = np.asarray(kernel.theta)
            calls.append(1)
            return -float(np.sum((th - np.array([0.3, -0.2])) ** 2 * np.array([1.0, 3.0])))

        mod.log_likelihood.log_likelihood = fake_loglike
        k0 = treegp.eval_kernel("1.0 * RBF(1.0)")
        serial = mod.log_likelihood(np.zeros((4, 1)), np.zeros(4), np.zeros(4))
        ks = serial.optimizer(k0)
        n_serial = len(calls)
        del calls[:]
        mod.DISTRIBUTED_FD = True
        par = mod.log_likelihood(np.zeros((4, 1)), np.zeros(4), np.zeros(4))
        kp = par.optimizer(k0)
        mod.DISTRIBUTED_FD = False
        np.testing.assert_allclose(kp.theta, [0.3, -0.2], atol=1e-5)
        np.testing.assert_allclose(kp.theta, ks.theta, atol=1e-6)
        assert len(calls) < 0.75 * n_serial            # each rank evaluated only its share of the probes
        th = torch.tensor(kp.theta)
        both = [torch.zeros_like(th) for _ in range(world)]
        tdist.all_gather(both, th)
        assert torch.equal(both[0], both[1])           # identical iterates on every rank

        # ---- slabs and gather ----
        m = 1001
        lo, hi = dist.slab(m, rank, world)
        local = torch.arange(lo, hi, dtype=torch.float64) * 2.0
        whole = dist.gather_slabs(local, m)
        np.testing.assert_array_equal(whole.numpy(), np.arange(m) * 2.0)
        ret[rank] = "ok"
    except Exception as e:  # surfaced by the parent
        ret[rank] = "FAIL: %r" % (e,)
        raise
    finally:
        tdist.destroy_process_group()


def test_two_rank_gloo_pairbin_allreduce_and_gather():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_slab_partition_properties():
    from treegp_b200 import dist

    for n in (0, 1, 7, 1000, 1001):
        for world in (1, 2, 3, 8):
            parts = [dist.slab(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
