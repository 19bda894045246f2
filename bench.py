#!/usr/bin/env python
"""bench.py -- headline measurement of the treegp_b200 hot path.

BASELINE.json metric: "GP fit+predict wall-s at N=40k; 2PCF pairs/s at 1/2/4/8 B200".

One JSON line is printed (rank 0).  Its primary metric is the one that is defined at 1/2/4/8 GPUs:

  metric  = "2pcf_pairs_per_s": unordered pairs binned per second by the anisotropic TwoD pair-binning
            kernel on N = 1,000,000 synthetic points (configs[3]; nbins = 21, min_sep = 0, max_sep = the
            reference default: half the field diagonal, two_pcf.py:421 -> 84 % of all pairs land in a bin).
            A "step" is one complete pair count over all N(N-1)/2 pairs.  With --gpus N the pair tiles are
            dealt to the ranks and the bin sums all-reduced over NCCL: total work is fixed -> "strong".
  e2e     = the same count through the reference-facing API, treegp_b200.two_pcf(...).comp_2pcf(X, y,
            y_err) with HOST numpy inputs (H2D of the points and D2H of xi inside the timed region).
  gp      = the other half of the BASELINE metric, reported in the same line under "gp_fit_predict":
            wall seconds of GPInterpolation.initialize + solve(optimizer='anisotropic') + predict(1e6
            points) for a 2-D AnisotropicVonKarman field with N = 40,000 (configs[2]), with a breakdown
            and the Cholesky's FP64 tensor-pipe roofline.  Test points are sharded over the ranks.

`--impl reference` times the CPU restatement of the same pair binning (oracle/, OpenMP over all host
cores; TreeCorr itself is not installable here) on a bounded slab of rows of the same N = 1e6 problem.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--npoints", type=int, default=1_000_000, help="2PCF points (configs[3]: 1e6)")
    ap.add_argument("--gp-train", type=int, default=40_000, help="GP training points (configs[2]: 40k)")
    ap.add_argument("--gp-predict", type=int, default=1_000_000)
    ap.add_argument("--skip-gp", action="store_true", help="only the 2PCF part (used for ncu captures)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-other", action="store_true", help="skip the configs[1] / configs[4] side measurements")
    ap.add_argument("--cpu-rows", type=int, default=20000, help="rows of the pair matrix in the CPU sample")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------
FIELD_2PCF = 1000.0
NBINS = 21


def make_2pcf_inputs(n, seed=42):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-FIELD_2PCF / 2, FIELD_2PCF / 2, size=(n, 2))
    y = rng.normal(size=n)  # pair counts and timing do not depend on the field values
    y_err = np.zeros(n)
    max_sep = np.sqrt(2.0) * FIELD_2PCF / 2.0  # reference default: half the field diagonal
    return X, y, y_err, 0.0, max_sep


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self._stop_evt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        # one nvidia-smi process sampling every 50 ms (a fresh process per sample takes ~0.2 s to start, which is
        # as long as a whole timed step by now)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.strip().split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)
            if self._stop_evt.is_set():
                break

    def stop(self):
        self._stop_evt.set()
        if getattr(self, "proc", None) is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm
# ------------------------------------------------------------------------------------------------
def cpu_pairbin_sample(X, y, min_sep, max_sep, rows):
    """Time the oracle (C, OpenMP, all host cores) on rows [0, rows) of the pair matrix."""
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU baseline is meant to use all host cores
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 or "OMP_NUM_THREADS" not in os.environ:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
    from oracle import pairbin_oracle as po

    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        po.set_threads(os.cpu_count())  # the OpenMP runtime may already be initialised with 1 thread
    n = len(y)
    rows = min(rows, n)
    k = y - np.mean(y)
    t0 = time.perf_counter()
    po.pairbin(X[:, 0], X[:, 1], k, None, min_sep, max_sep, NBINS, "TwoD", rows=(0, rows))
    dt = time.perf_counter() - t0
    pairs = rows * n - rows * (rows + 1) // 2
    return pairs / dt, dt, pairs, po.num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    X, y, y_err, mn, mx = make_2pcf_inputs(args.npoints)
    rates = []
    for i in range(args.warmup + args.steps):
        rate, dt, pairs, cores = cpu_pairbin_sample(X, y, mn, mx, args.cpu_rows)
        if i >= args.warmup:
            rates.append((rate, dt))
    value = float(np.mean([r for r, _ in rates]))
    ms = float(np.mean([d for _, d in rates]) * 1e3)
    sample = "rows [0,%d) of the N=%d pair matrix = %.3g unordered pairs per step" % (args.cpu_rows, args.npoints, pairs)
    line = {
        "impl": "reference", "metric": "2pcf_pairs_per_s", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "anisotropic TwoD 2PCF pair binning, N=%d, nbins=21, min_sep=0, max_sep=half field "
                               "diagonal (configs[3])" % args.npoints},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "oracle/pairbin_oracle.c (brute force, OpenMP); TreeCorr is not installable"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as tdist

    import treegp_b200 as treegp
    from treegp_b200 import _cabi, backend, binning, dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        tdist.init_process_group("nccl", device_id=dev)
    _cabi.load()

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    hbm_peak, peak_src = load_peaks()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # ---------------- 2PCF: device-resident timing ----------------
    n = args.npoints
    X, y, y_err, mn, mx = make_2pcf_inputs(n)
    npairs_total = n * (n - 1) // 2
    px, py = backend.to_device(X[:, 0]), backend.to_device(X[:, 1])
    pk = backend.to_device(y - np.mean(y))
    # device-resident inputs are stored in Hilbert order (what two_pcf.comp_2pcf does on the way in)
    order = backend.hilbert_order(px, py)
    px, py, pk = px[order].contiguous(), py[order].contiguous(), pk[order].contiguous()
    off = backend.to_device(np.array([0, n]), torch.int64)
    edges = backend.to_device(binning.twod_thresholds(mx, NBINS))
    launches = 0

    def step_device():
        res = backend.pairbin(px, py, pk, None, off, n, _cabi.BIN_TWOD, edges, NBINS, mn, mx, rank=rank, nranks=world)
        if world > 1:
            dist.allreduce_bins(None, res[0], res[1], res[2])
        return res

    # per-pair mode first (block forms off: every pair of every in-range block goes through the compare /
    # masked-FMA loop): this is the kernel the FP64-issue roofline of 10 ops per pair applies to
    backend.set_option("pairbin_block_sums", 0)
    step_device()
    pp_ms = []
    for _ in range(2):
        flush_buf.fill_(1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res_pp = backend.pairbin(px, py, pk, None, off, n, _cabi.BIN_TWOD, edges, NBINS, mn, mx, rank=rank, nranks=world)
        e1.record()
        barrier()
        pp_ms.append(max_over_ranks(e0.elapsed_time(e1)))
    backend.set_option("pairbin_block_sums", 1)
    backend.pairbin_stats(reset=True)

    for _ in range(args.warmup):
        res = step_device()
    barrier()
    stats = backend.pairbin_stats(reset=True)   # paths taken by one rank's share during the warm-up steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    step_ms, kern_ms = [], []
    for _ in range(args.steps):
        flush_buf.fill_(1)
        barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        res = backend.pairbin(px, py, pk, None, off, n, _cabi.BIN_TWOD, edges, NBINS, mn, mx, rank=rank, nranks=world)
        e1.record()
        if world > 1:
            dist.allreduce_bins(None, res[0], res[1], res[2])
        e2.record()
        barrier()
        step_ms.append(max_over_ranks(e0.elapsed_time(e2)))
        kern_ms.append(max_over_ranks(e0.elapsed_time(e1)))
        launches += 2  # pairbin_boxes_kernel + pairbin_kernel (ours); the counter memset and torch fills are not counted
    clocks = sampler.stop() if rank == 0 else None
    counted = int(res[0].sum().item())
    total_ms = float(np.sum(step_ms))
    value = npairs_total * args.steps / (total_ms * 1e-3)

    # ---------------- 2PCF: end to end through the public API (host buffers) ----------------
    def pinned(a):
        """numpy view of a page-locked copy of `a` (the contract's "inputs from pinned host memory"): the API
        takes plain numpy arrays; the driver recognises the pages as pinned and copies them by DMA"""
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        v = t.numpy()      # keeps `t` alive as its base
        v[...] = a
        return v

    Xp, yp, yerrp = pinned(np.ascontiguousarray(X)), pinned(y), pinned(y_err)
    e2e_ms = []
    tp = treegp.two_pcf(Xp, yp, yerrp, mn, mx, nbins=NBINS, anisotropic=True)
    tp.group = dist.WORLD if world > 1 else False
    for i in range(1 + args.steps):
        barrier()
        t0 = time.perf_counter()
        xi, _, _, _ = tp.comp_2pcf(Xp, yp, yerrp)
        torch.cuda.synchronize()
        dt = max_over_ranks((time.perf_counter() - t0) * 1e3)
        if i > 0:
            e2e_ms.append(dt)
    e2e_value = npairs_total / (float(np.mean(e2e_ms)) * 1e-3)
    h2d = 3 * 8 * n + 8 * (NBINS + 1) + 16
    d2h = NBINS * NBINS * (8 + 8 + 8)

    # ---------------- FP64 peaks (roofline denominators), measured live ----------------
    dfma_tf = backend.microbench_fp64(0, 20000)
    dmma_tf = backend.microbench_fp64(1, 20000)
    # algorithmic work per unordered pair: 10 FP64 operations; peak in operations/s is
    # the measured DFMA rate / 2 (one FMA instruction = 2 flop = 1 operation slot)
    ops_per_pair = 10.0
    my_pairs = npairs_total / world
    peak_gops = dfma_tf * 1e3 / 2.0
    kern_s = float(np.mean(kern_ms)) * 1e-3
    pp_s = float(np.mean(pp_ms)) * 1e-3
    # pairs that the timed (default) kernel evaluated one by one, per launch on this rank; blocks whose pairs
    # provably share a bin are summed in closed form and cost no per-pair FP64 work
    st_tot = max(1, sum(stats.values()))
    frac_paths = {k: v / st_tot for k, v in stats.items()}
    evaluated = my_pairs * (1.0 - frac_paths["closed_form"] - frac_paths["one_axis_sorted"])
    ach_gops = ops_per_pair * evaluated / kern_s / 1e9
    ach_pp = ops_per_pair * my_pairs / pp_s / 1e9
    same_counts = bool(torch.equal(res_pp[0], res[0])) if world == 1 else None
    roofline = {"bound": "fp64_alu", "kernel": "pairbin_kernel<TwoD, unweighted, pair-by-pair>", "achieved": ach_pp,
                "peak": peak_gops, "unit": "Gop/s (FP64 instructions x lanes)", "frac": ach_pp / peak_gops,
                # dram__bytes_read + dram__bytes_write of one launch, from the ncu --set full capture summarised in
                # profiles/r1_pairbin_v3_N1M.ncu.txt (N = 1e6; algorithmic input 24 MB + 1 MB chunk boxes)
                "traffic": 25.28e6 if (n == 1_000_000 and world == 1) else None,
                "note": "neither HBM- nor tensor-bound: 24 N bytes in, N^2/2 pairs; peak = measured DFMA issue rate "
                        "(tgp_microbench_fp64); algorithmic work 10 FP64 ops per unordered pair (SURVEY 8d).  That "
                        "figure describes the pair-by-pair kernel (block forms off, every pair through the compare / "
                        "masked-FMA loop), which is what achieved / frac are measured on here, live, on the same data "
                        "(ms_per_launch).  The TIMED kernel (value, ms_per_step) books blocks of 32 x 32 pairs that "
                        "provably fall into one bin from pre-computed chunk sums, answers one-axis blocks by a rank "
                        "query on sorted chunks and takes the marginal sums of 2 x 2-window blocks from two such "
                        "queries: identical counts, see timed_kernel.  In the pair-by-pair kernel every pair's forward "
                        "bin bits are evaluated individually; the mirrored-entry cross-check is evaluated per pair "
                        "except in blocks whose bounding boxes prove it (one bin per axis, mirrored window = its "
                        "mirror image)",
                "ms_per_launch": pp_s * 1e3, "pairs_per_s": my_pairs * world / pp_s,
                "timed_kernel": {"kernel": "pairbin_kernel<TwoD, unweighted, block forms>", "ms_per_launch": kern_s * 1e3,
                                 "speedup_over_pair_by_pair": pp_s / kern_s, "path_fractions": frac_paths,
                                 "counts_identical_to_pair_by_pair": same_counts,
                                 "fp64_frac_of_peak_counting_only_pairs_evaluated_one_by_one": ach_gops / peak_gops},
                "dram_GBs_for_reference": 24.0 * n / kern_s / 1e9,
                "hbm_peak_GBs": hbm_peak, "hbm_peak_source": peak_src}

    line = {
        "metric": "2pcf_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "anisotropic TwoD 2PCF pair binning, N=%d, nbins=21, min_sep=0, max_sep=half field "
                               "diagonal (configs[3])" % n,
                   "pairs_per_step": npairs_total, "pairs_in_range_x2": counted, "sharding": "pair tiles over %d rank(s) + NCCL allreduce of bin sums" % world,
                   "l2": "256 MB buffer written between timed iterations (L2 flush)"},
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "treegp_b200.two_pcf(...).comp_2pcf(X, y, y_err) with host numpy inputs (page-locked)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "fp64_peaks_measured": {"dfma_tflops": dfma_tf, "dmma_tflops": dmma_tf},
    }

    # ---------------- GP fit + predict at N = 40k ----------------
    if not args.skip_gp:
        line["gp_fit_predict"] = run_gp(args, treegp, backend, dist, rank, world, dev, barrier, max_over_ranks,
                                        hbm_peak, dmma_tf)
        if rank == 0 and not args.skip_other:
            line["other_configs"] = run_other_configs(args, treegp, backend, dmma_tf)

    # ---------------- CPU baseline (rank 0, bounded sample) ----------------
    if rank == 0 and not args.skip_cpu:
        rate, dt, pairs, cores = cpu_pairbin_sample(X, y, mn, mx, args.cpu_rows)
        line["cpu_baseline"] = {"value": rate, "unit": "pairs/s", "cores": cores, "kind": "port",
                                "sample": "rows [0,%d) of the N=%d pair matrix = %.3g unordered pairs, %.1f s"
                                          % (args.cpu_rows, n, pairs, dt),
                                "note": "oracle/pairbin_oracle.c (brute force, OpenMP); TreeCorr is not installable"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        tdist.destroy_process_group()


def run_gp(args, treegp, backend, dist, rank, world, dev, barrier, max_over_ranks, hbm_peak, dmma_tf):
    """configs[2]: 2-D AnisotropicVonKarman, N train / M predict, optimizer='anisotropic'."""
    import torch
    from treegp_b200.kernels import lower_kernel
    from treegp_b200.two_pcf import get_correlation_length_matrix

    n, m = args.gp_train, args.gp_predict
    L = 160.0 * np.sqrt(n / 40000.0)
    size, g1, g2, sigma, noise = 1.5, 0.2, 0.2, 2.0, 0.01
    inv = np.linalg.inv(get_correlation_length_matrix(size, g1, g2))
    kstr = "%r**2 * AnisotropicVonKarman(invLam=array([[%.17g, %.17g], [%.17g, %.17g]]))" % (
        sigma, inv[0, 0], inv[0, 1], inv[1, 0], inv[1, 1])
    rng = np.random.default_rng(42)
    X = rng.uniform(-L / 2, L / 2, size=(n, 2))
    Xs = rng.uniform(-L / 2, L / 2, size=(m, 2))
    # exact GRF draw y = L z + noise with our own factorisation (data generation, not timed)
    kern = treegp.eval_kernel(kstr)
    desc = lower_kernel(kern, 2)
    ws = backend.kmat_sym(X, desc, diag_add=backend.to_device(np.full(n, 1e-8)), lower_only=True)
    info = backend.potrf(ws, n)
    assert int(info.item()) == 0
    z = torch.as_tensor(rng.normal(size=n), device=dev)
    y = torch.zeros(n, dtype=torch.float64, device=dev)
    blk = 4096
    for r0 in range(0, n, blk):  # y = tril(L) z, block rows (avoids a 12.8 GB tril copy)
        r1 = min(n, r0 + blk)
        rows = ws[r0:r1, :n]
        y[r0:r1] = torch.tril(rows, diagonal=r0) @ z
    y = y.cpu().numpy() + rng.normal(scale=noise, size=n)
    y_err = np.full(n, noise)
    del ws, z
    torch.cuda.empty_cache()

    lo, hi = dist.slab(m, rank, world)
    Xs_local = Xs[lo:hi]
    out = {}

    def one_run():
        t = {}
        barrier()
        t0 = time.perf_counter()
        gp = treegp.GPInterpolation(kernel=kstr, optimizer="anisotropic", normalize=True, nbins=21, min_sep=0.0,
                                    max_sep=1.0, p0=[1.0, 0.0, 0.0])
        gp.initialize(X, y, y_err=y_err)
        gp.solve()
        torch.cuda.synchronize()
        t["solve_anisotropic_s"] = time.perf_counter() - t0
        t1 = time.perf_counter()
        ypred = gp.predict(Xs_local)
        if world > 1:  # every rank ends up with all M predictions (all-gather of the slabs, 8 M bytes)
            ypred = dist.gather_slabs(torch.as_tensor(ypred, device=dev), m).cpu().numpy()
        torch.cuda.synchronize()
        t["predict_s"] = time.perf_counter() - t1
        t["total_s"] = time.perf_counter() - t0
        return gp, ypred, t

    one_run()  # warm-up (allocator, kernel load)
    runs = []
    for _ in range(2):
        gp, ypred, t = one_run()
        runs.append({k: max_over_ranks(v) for k, v in t.items()})
    best = min(runs, key=lambda r: r["total_s"])

    # kernel-level breakdown with CUDA events (device-resident, rank-local)
    def ev(fn, reps=2):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e-3)
        return float(np.mean(ts))

    desc = lower_kernel(gp.kernel, 2)
    Xd = backend.as_points(X)
    e2 = backend.to_device(y_err ** 2)
    ws = backend.alloc_matrix(n, n)
    t_k = ev(lambda: backend.kmat_sym(Xd, desc, e2, out=ws, lower_only=True))

    def build_and_factor():
        backend.kmat_sym(Xd, desc, e2, out=ws, lower_only=True)
        backend.potrf(ws, n)

    # the same build for an RBF metric, full square as kernel.__call__(X) returns it: the HBM-store-bound case
    desc_rbf = lower_kernel(treegp.eval_kernel(kstr.replace("AnisotropicVonKarman", "AnisotropicRBF")), 2)
    t_k_rbf_full = ev(lambda: backend.kmat_sym(Xd, desc_rbf, e2, out=ws, lower_only=False), reps=3)

    t_kf = ev(build_and_factor)
    t_chol = t_kf - t_k
    b = backend.to_device(y)
    t_solve = ev(lambda: backend.potrs_vec(ws, n, b.clone()))
    Xsd = backend.as_points(Xs_local)
    alpha = backend.potrs_vec(ws, n, b.clone())
    t_mean_full = ev(lambda: backend.predict_mean(Xsd, Xd, desc, alpha, truncate=False))
    t_mean = ev(lambda: backend.predict_mean(Xsd, Xd, desc, alpha))   # what predict() runs: truncated support
    mv = min(len(Xs_local), 8192)
    t_var = ev(lambda: backend.predict_var(Xsd[:mv], Xd, desc, ws), reps=1)
    flops = n ** 3 / 3.0
    out.update({
        "metric": "gp_fit_predict_wall_s", "value": best["total_s"], "unit": "s", "higher_is_better": False,
        "config": {"workload": "2D AnisotropicVonKarman GP, N=%d train / M=%d predict (configs[2]); "
                               "GPInterpolation.initialize + solve(optimizer='anisotropic', nbins=21, max_sep=1 as tests/test_hyp_search.py, 444 bootstraps) "
                               "+ predict(M), host numpy in/out; test points sharded over %d rank(s)" % (n, m, world)},
        "wall_breakdown_s": best,
        "kernel_breakdown_s": {"kmat_lower": t_k, "kmat_rbf_full": t_k_rbf_full, "potrf": t_chol, "potrs_vec": t_solve,
                               "predict_mean_local_M=%d" % len(Xs_local): t_mean,
                               "predict_mean_full_sum_local_M=%d" % len(Xs_local): t_mean_full,
                               "predict_var_diag_M=%d" % mv: t_var},
        "fitted_theta": [float(v) for v in gp.kernel.theta],
        "true_theta": [float(v) for v in kern.theta],
        "roofline_potrf": {"bound": "tensor", "kernel": "gemm_nt_sub_kernel (DMMA.8x8x4) inside tgp_potrf",
                           "achieved": flops / t_chol / 1e12, "peak": dmma_tf, "unit": "TFLOP/s",
                           "frac": flops / t_chol / 1e12 / dmma_tf,
                           "note": "FP64 tensor pipe (mma.sync m8n8k4 f64); peak measured live by tgp_microbench_fp64"},
        "roofline_kmat": {"bound": "hbm", "kernel": "kmat_sym_kernel<RBF>, full N x N", "achieved": 8.0 * n * n / t_k_rbf_full / 1e9,
                          "peak": hbm_peak, "unit": "GB/s", "frac": 8.0 * n * n / t_k_rbf_full / 1e9 / hbm_peak,
                          "note": "AnisotropicRBF, full square (what kernel.__call__(X) returns), algorithmic bytes 8 N^2 "
                                  "stored once; the von Karman build of this fit (kmat_lower above, 4 N^2 bytes) is "
                                  "FP64-ALU bound by the Bessel-K evaluation",
                          "von_karman_lower_GBs": 4.0 * n * n / t_k / 1e9},
        "predict_mean_kernel_evals_per_s": len(Xs_local) * n / t_mean_full,
        "predict_mean_note": "predict() uses tgp_predict_mean_trunc (Hilbert-sorted blocks, pairs with correlation "
                             "< 1e-40 skipped); kernel_evals_per_s is the untruncated kernel evaluating all M x N pairs",
    })
    return out


def run_other_configs(args, treegp, backend, dmma_tf):
    """Side measurements for the remaining BASELINE.json configs (rank 0, single GPU):
    configs[1]  2-D AnisotropicRBF GP, N = 10,000, optimizer='log-likelihood' (FP64 Cholesky + marginal likelihood
                inside scipy's L-BFGS-B loop): whole-fit time, evaluation count, time per evaluation;
    configs[4]  robust 2PCF fit ingredients at N = 200,000: one pair count + 100 batched bootstrap resamples."""
    import torch
    from treegp_b200.kernels import lower_kernel
    from treegp_b200.two_pcf import get_correlation_length_matrix

    out = {}
    rng = np.random.default_rng(7)
    # ---- configs[1] ----
    n = 10_000
    L = 80.0 * np.sqrt(n / 16000.0)
    inv = np.linalg.inv(get_correlation_length_matrix(0.5, 0.2, 0.2))
    kstr = "4.0 * AnisotropicRBF(invLam=array([[%.17g, %.17g], [%.17g, %.17g]]))" % (inv[0, 0], inv[0, 1], inv[1, 0], inv[1, 1])
    X = rng.uniform(-L / 2, L / 2, size=(n, 2))
    kern = treegp.eval_kernel(kstr)
    desc = lower_kernel(kern, 2)
    ws = backend.kmat_sym(X, desc, diag_add=backend.to_device(np.full(n, 1e-8)), lower_only=True)
    backend.potrf(ws, n)
    z = torch.as_tensor(rng.normal(size=n), device=ws.device)
    y = (torch.tril(ws[:, :n]) @ z).cpu().numpy() + rng.normal(scale=0.01, size=n)
    del ws
    y_err = np.full(n, 0.01)
    gp = treegp.GPInterpolation(kernel=kstr, optimizer="log-likelihood", normalize=True)
    gp.initialize(X, y, y_err=y_err)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gp.solve()
    torch.cuda.synchronize()
    t_fit = time.perf_counter() - t0
    nev = gp._optimizer.n_evaluations
    out["loglike_fit_N10k"] = {
        "workload": "2D AnisotropicRBF GP, N=10000, optimizer='log-likelihood' (configs[1])",
        "fit_wall_s": t_fit, "likelihood_evaluations": nev, "s_per_evaluation": t_fit / max(nev, 1),
        "potrf_tflops_per_evaluation_upper_bound": n ** 3 / 3.0 / (t_fit / max(nev, 1)) / 1e12,
        "fitted_theta": [float(v) for v in gp.kernel.theta], "true_theta": [float(v) for v in kern.theta],
        "logL": float(gp._optimizer._logL)}
    # ---- configs[4] ----
    n = 200_000
    Lf = 1000.0 * np.sqrt(n / 1e6)
    X = rng.uniform(-Lf / 2, Lf / 2, size=(n, 2))
    yv = rng.normal(size=n)
    tp = treegp.two_pcf(X, yv, np.zeros(n), 0.0, np.sqrt(2.0) * Lf / 2.0, nbins=21, anisotropic=True)
    tp.group = False
    tp.comp_2pcf(X, yv, np.zeros(n))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tp.comp_2pcf(X, yv, np.zeros(n))
    torch.cuda.synchronize()
    t_one = time.perf_counter() - t0
    B = 100
    t0 = time.perf_counter()
    cov = tp.comp_xi_covariance(n_bootstrap=B, mask=None, seed=610639139)
    torch.cuda.synchronize()
    t_boot = time.perf_counter() - t0
    pairs = n * (n - 1) / 2
    out["bootstrap_2pcf_N200k"] = {
        "workload": "anisotropic 2PCF, N=200000, nbins=21, default max_sep: 1 pair count + %d batched bootstrap "
                    "resamples (configs[4])" % B,
        "pair_count_wall_s": t_one, "pairs_per_s": pairs / t_one,
        "bootstrap_wall_s": t_boot, "resamples": B,
        "resample_pairs_per_s": B * pairs / t_boot,
        "note": "a resample keeps ~63% of the distinct points (weights = multiplicities): ~40% of the pairs per resample",
        "cov_shape": list(cov.shape)}
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
