"""Ensemble sharding across the GPUs of one box: one process per GPU (torchrun), contiguous
global trajectory-id ranges per rank, NO data-path collective; the only exchange is one
all-reduce (sum) of the ensemble statistics buffers (batch count / shifted sum / shifted sum of squares
[3, 64, T, C] doubles and the histogram [T, H, B] uint64 — tens of KB).  Saved trajectories stay on the rank that produced
them.  Philox counters use the GLOBAL trajectory id, so the union of the shards is identical
to a single-device run whatever the number of ranks (reference analogue: none — the reference
is single-task).
"""
import numpy as np

from . import _lib
from .solve import Ensemble, solve as _solve


def shard_range(n_total, rank, world):
    """Contiguous id range [offset, offset + count) of `rank` out of `world`."""
    base, rem = divmod(int(n_total), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def _dist():
    import torch.distributed as dist
    return dist


def rank_world(group=None):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def allreduce_stats_host(res, group=None):
    """All-reduce the statistics of an EnsembleResult held in host memory (any backend)."""
    import torch
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return res
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    parts = [res.stat_count, res.stat_sum, res.stat_sumsq]
    if res.batch_count is not None:      # the shifted batch sums add up too: the pivot is the same on every rank
        parts += [res.batch_count, res.batch_sum, res.batch_sumsq]
    f = torch.from_numpy(np.concatenate([a.ravel() for a in parts])).to(dev)
    h = torch.from_numpy(res.hist.astype(np.int64).ravel()).to(dev)
    if f.numel():
        dist.all_reduce(f, group=group)
    if h.numel():
        dist.all_reduce(h, group=group)
    f = f.cpu().numpy()
    n = res.stat_count.size
    res.stat_count = f[:n].reshape(res.stat_count.shape)
    res.stat_sum = f[n:2 * n].reshape(res.stat_sum.shape)
    res.stat_sumsq = f[2 * n:3 * n].reshape(res.stat_sumsq.shape)
    if res.batch_count is not None:
        m = res.batch_count.size
        res.batch_count = f[3 * n:3 * n + m].reshape(res.batch_count.shape)
        res.batch_sum = f[3 * n + m:3 * n + 2 * m].reshape(res.batch_sum.shape)
        res.batch_sumsq = f[3 * n + 2 * m:].reshape(res.batch_sumsq.shape)
    res.hist = h.cpu().numpy().astype(np.uint64).reshape(res.hist.shape)
    return res


class _DevView:
    """__cuda_array_interface__ holder: zero-copy torch view of a library device buffer."""

    def __init__(self, ptr, nelem, typestr):
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": typestr, "data": (ptr, False), "version": 2}


def stats_tensors(ens, device):
    """(float64 tensor [3*64*T*C], int64 tensor [T*H*B]) aliasing the ensemble's device statistics."""
    import torch
    pf, nf, ph, nh = ens.stats_device()
    f = torch.as_tensor(_DevView(pf, nf, "<f8"), device=device) if nf else None
    h = torch.as_tensor(_DevView(ph, nh, "<i8"), device=device) if nh else None
    return f, h


def allreduce_stats_device(ens, device, group=None):
    """In-place NCCL all-reduce of the statistics buffers resident in HBM (no host round trip)."""
    dist = _dist()
    f, h = stats_tensors(ens, device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if f is not None:
            dist.all_reduce(f, group=group)
        if h is not None:
            dist.all_reduce(h, group=group)
    return f, h


def solve_sharded(problem, alg, *, trajectories, group=None, local_solve=None, xc0=None, xd0=None, replay=None, **kw):
    """solve() for an ensemble of `trajectories` spread over the ranks of the process group.

    Every rank returns its own shard's trajectories (EnsembleResult.shard = (offset, count)) and
    the GLOBAL statistics.  `local_solve` lets the tests run the host logic on CPU with another
    backend; by default it is the GPU solve of this package (device = LOCAL rank's GPU)."""
    rank, world = rank_world(group)
    off, cnt = shard_range(trajectories, rank, world)
    sl = slice(off, off + cnt)
    if xc0 is not None and np.ndim(xc0) == 2:
        xc0 = np.asarray(xc0)[sl]
    if xd0 is not None and np.ndim(xd0) == 2:
        xd0 = np.asarray(xd0)[sl]
    if replay is not None:
        replay = np.asarray(replay)[sl]
    fn = local_solve or _solve
    res = fn(problem, alg, trajectories=cnt, traj_id_offset=kw.pop("traj_id_offset", 0) + off, xc0=xc0, xd0=xd0,
             replay=replay, **kw)
    res.shard = (off, cnt)
    return allreduce_stats_host(res, group)
