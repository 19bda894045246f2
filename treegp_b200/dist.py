"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box,
gloo in the CPU test-suite).

Only two pieces of the hot path shard:
* pair binning -- every rank holds all points, runs of pair tiles are dealt round-robin by
  ``tgp_pairbin(tile_rank, tile_nranks)`` and the bin sums are combined by ONE all-reduce
  (a few KB to ~1 MB): counts are integers, so the result is bit-exact for any number of GPUs;
* predict -- test points are split into contiguous slabs, X / alpha (/ L) are replicated, no data-path
  collective; results are gathered.
The Cholesky itself does not shard at these sizes (12.8 GB at N = 40k fits one GPU): replicas only.
"""
import numpy as np
import torch
import torch.distributed as tdist


WORLD = "world"   # sentinel: shard over the default process group


def _pg(group):
    return None if group in (None, False, WORLD) else group


def rank_world(group=None):
    """(rank, world) of `group`.  False: explicitly single-process -> (0, 1).  None / WORLD: the default process
    group, or (0, 1) when torch.distributed is not initialised.  A CabiComm carries its own.  Otherwise a ProcessGroup."""
    if group is False:
        return 0, 1
    if isinstance(group, CabiComm):
        return group.rank, group.world
    if group in (None, WORLD) and not (tdist.is_available() and tdist.is_initialized()):
        return 0, 1
    return tdist.get_rank(_pg(group)), tdist.get_world_size(_pg(group))


def allreduce_bins(group, *tensors):
    """Sum the per-rank bin arrays in place (one collective per array)."""
    for t in tensors:
        if t is not None:
            tdist.all_reduce(t, op=tdist.ReduceOp.SUM, group=_pg(group))


def allreduce_packed_bins(group, packed):
    """ONE all-reduce for the packed bin buffer of backend.pairbin_packed: plane 0 (int64 pair counts stored as raw
    words) is first converted to FP64 values -- exact below 2^53 pairs per bin, and sums of integer-valued doubles
    below 2^53 are exact in any order, so the counts stay bit-identical for any number of ranks.  Returns the buffer
    with plane 0 holding the summed counts as FP64 VALUES."""
    packed[0] = packed[0].view(torch.int64).to(torch.float64)
    tdist.all_reduce(packed, op=tdist.ReduceOp.SUM, group=_pg(group))
    return packed


class CabiComm(object):
    """The all-reduce of the bin sums behind the C ABI (tgp_comm_* / tgp_allreduce_bins: NCCL bound by the library at run
    time), for callers that shard the pair tiles without torch.distributed collectives on the data path.  Construction
    is collective: `exchange(buf)` must return rank 0's 128-byte id on every rank -- by default a torch.distributed
    broadcast (plumbing only; any transport will do).  Usable as `two_pcf.group`."""

    def __init__(self, rank, world, exchange=None):
        import ctypes
        from . import _cabi

        self.rank, self.world = int(rank), int(world)
        lib = _cabi.load()
        ident = (ctypes.c_ubyte * 128)()
        if self.rank == 0:
            _cabi.check(lib.tgp_comm_unique_id(ident), "tgp_comm_unique_id")
        if self.world > 1:
            if exchange is None:
                t = torch.tensor(list(ident), dtype=torch.uint8, device="cuda")
                tdist.broadcast(t, src=0)
                raw = bytes(t.cpu().tolist())
            else:
                raw = exchange(bytes(ident))
            ident = (ctypes.c_ubyte * 128).from_buffer_copy(raw)
        self._handle = ctypes.c_void_p(0)
        _cabi.check(lib.tgp_comm_init_rank(ident, self.rank, self.world, ctypes.byref(self._handle)), "tgp_comm_init_rank")

    def allreduce_packed_bins(self, packed):
        """In-place sum of the packed bin buffer over the ranks; plane 0 holds int64 counts (raw words) before and after."""
        import ctypes
        from . import _cabi

        planes, per_plane = int(packed.shape[0]), int(packed[0].numel())
        _cabi.check(_cabi.load().tgp_allreduce_bins(self._handle, ctypes.c_void_p(packed.data_ptr()), planes, per_plane,
                                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                    "tgp_allreduce_bins")
        return packed

    def close(self):
        from . import _cabi

        if self._handle:
            _cabi.load().tgp_comm_destroy(self._handle)
            self._handle = None


def slab(n, rank, world):
    """Contiguous [begin, end) share of n items for `rank`; sizes differ by at most one."""
    base, extra = divmod(int(n), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_slabs(local, n, group=None):
    """All-gather the per-rank slabs of a length-n device vector back into one vector on every rank."""
    rank, world = rank_world(group)
    if world == 1:
        return local
    sizes = [slab(n, r, world)[1] - slab(n, r, world)[0] for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=local.dtype, device=local.device)
    buf[: local.numel()] = local
    outs = [torch.empty_like(buf) for _ in range(world)]
    tdist.all_gather(outs, buf, group=_pg(group))
    return torch.cat([o[:s] for o, s in zip(outs, sizes)])
