"""Sharding of an ensemble over the GPUs of one box: one process per GPU.

Members are independent (magpy/model.py:204-207), so the path shards by member index with no
data-path exchange; the only collective is one all-reduce (sum, fp64) of the [S][4] ensemble
sums.  Member `i` keeps its seed and its Philox member index whatever the number of ranks, so
results do not depend on the GPU count (up to fp64 summation order)."""
import numpy as np


def shard_bounds(n_members, world_size, rank):
    """Contiguous range [lo, hi) of member indices owned by `rank`: ceil(R/G) per rank."""
    if not 0 <= rank < world_size:
        raise ValueError('rank %d outside world of size %d' % (rank, world_size))
    per = -(-int(n_members) // int(world_size))
    lo = min(rank * per, n_members)
    return lo, min(lo + per, n_members)


def allreduce_sums(sums, group=None):
    """Sum the [S][4] ensemble sums over all ranks (no-op without an initialised process group)."""
    try:
        import torch
        import torch.distributed as dist
    except ImportError:      # single process, torch absent
        return sums
    if not (dist.is_available() and dist.is_initialized()):
        return sums
    on_gpu = dist.get_backend(group) == 'nccl'
    t = torch.from_numpy(np.ascontiguousarray(sums))
    if on_gpu:
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    sums[...] = t.cpu().numpy()
    return sums
