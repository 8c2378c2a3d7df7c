"""Sharding of an ensemble over the GPUs of one box: one process per GPU.

Members are independent (magpy/model.py:204-207), so the path shards by member index with no
data-path exchange; the only collective is one all-reduce (sum, fp64) of the [S][4] ensemble
sums, issued by the native library through NCCL (`core.Comm`, include/magpy_b200.h).  Member `i`
keeps its seed and its Philox member index whatever the number of ranks, so results do not depend
on the GPU count (up to fp64 summation order)."""
from . import core

_WORLD = None


def shard_bounds(n_members, world_size, rank):
    """Contiguous range [lo, hi) of member indices owned by `rank`: ceil(R/G) per rank."""
    if not 0 <= rank < world_size:
        raise ValueError('rank %d outside world of size %d' % (rank, world_size))
    per = -(-int(n_members) // int(world_size))
    lo = min(rank * per, n_members)
    return lo, min(lo + per, n_members)


def world_comm(device=-1):
    """The process-wide communicator built from the launcher's environment (RANK, WORLD_SIZE, LOCAL_RANK,
    MASTER_ADDR, MASTER_PORT — what torchrun sets); created on first use.  device < 0 = LOCAL_RANK."""
    global _WORLD
    if _WORLD is None:
        _WORLD = core.Comm.from_env(device)
    return _WORLD


def resolve_comm(shard, comm, device):
    """The communicator a sharded call must use: the one given, else the environment's; never a silent no-op — a
    sharded run whose sums are not reduced would report partial sums against the global member count."""
    rank, world = shard
    if comm is not None:
        return comm
    if world == 1:
        return None
    try:
        comm = world_comm(device)
    except (ConnectionError, RuntimeError) as exc:
        raise RuntimeError('EnsembleModel.simulate(shard=(%d, %d)) needs a communicator to sum the ensemble over the '
                           'ranks: pass comm=, or launch one process per GPU with RANK / WORLD_SIZE / MASTER_ADDR / '
                           'MASTER_PORT set (%s)' % (rank, world, exc)) from exc
    if (comm.rank, comm.world_size) != (rank, world):
        raise ValueError('shard=(%d, %d) does not match the communicator (rank %d of %d)' %
                         (rank, world, comm.rank, comm.world_size))
    return comm
