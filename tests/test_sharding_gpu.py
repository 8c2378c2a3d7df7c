"""GPU tests of the multi-GPU plumbing: global Philox member indices under grouping and sharding, the library's
communicator (world size 1 on any box; two NCCL ranks when the box has two GPUs)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope='module')
def core():
    import magpy_b200.core as core
    return core


class SumCollector:
    """Host stand-in communicator for one-process tests: records the local sums, reduces nothing."""
    def __init__(self, rank, world):
        self.rank, self.world_size, self.local = rank, world, None

    def allreduce_sums(self, sums):
        self.local = sums.copy()
        return sums


def test_member_index_is_the_philox_member(core):
    args = ([7e-9], [1e5], [[0, 0, 1.0]], [[1.0, 0, 0]], [[0, 0, 0.0]], 4e5, 0.1, 330.0, False, True, False, 1e-13, 2e-11, 5)
    seeds = np.arange(1, 65) * 3
    a = core.simulate_ensemble(*args, seeds, stream_offset=1000)
    b = core.simulate_ensemble(*args, seeds, member_index=np.arange(64) + 1000)
    assert np.array_equal(a['trajectories'], b['trajectories'])
    perm = np.random.default_rng(0).permutation(64)
    c = core.simulate_ensemble(*args, seeds[perm], member_index=perm + 1000)
    assert np.array_equal(c['trajectories'], a['trajectories'][perm])
    with pytest.raises(ValueError):
        core.simulate_ensemble(*args, seeds, member_index=np.arange(64, dtype=np.uint64) + (1 << 32))


@pytest.mark.parametrize('N', [1, 2, 8])
def test_grouped_and_sharded_run_equals_the_unsharded_run(core, N):
    """with several parameter groups the Philox member index used to depend on the grouping and the shard.
    Now every member keeps its global index: per-member results of a sharded, grouped run are bit-identical to the
    unsharded run, and the shard sums add up to the full sums."""
    import magpy_b200 as mp
    rng = np.random.default_rng(3)
    loc = np.cumsum(np.full((N, 3), 2.5e-8), axis=0)
    base = mp.Model(np.full(N, 7e-9), np.full(N, 1e5), np.tile([0, 0, 1.0], (N, 1)), np.tile([1.0, 0, 0], (N, 1)), loc,
                    4e5, 0.1, 330.0, field_shape='sine', field_frequency=5e9, field_amplitude=1e4)
    R = 24
    damping = [0.1, 0.07, 0.1, 0.13] * 6          # not a fast per-member parameter for N > 1: one plan per value
    ens = mp.EnsembleModel(R, base, damping=damping)
    kw = dict(implicit_solve=False, return_trajectories=True)
    full = ens.simulate(4e-11, 1e-13, 9, 17, **kw)
    assert len(full.stats) == (3 if N > 1 else 1)
    total = np.zeros_like(full._sums)
    for rank in range(3):
        comm = SumCollector(rank, 3)
        part = ens.simulate(4e-11, 1e-13, 9, 17, shard=(rank, 3), comm=comm, **kw)
        lo, hi = mp.sharding.shard_bounds(R, 3, rank)
        assert np.array_equal(part.final_state_array(), full.final_state_array()[lo:hi])
        assert np.array_equal(part._traj, full._traj[lo:hi])
        total += comm.local
    assert np.allclose(total, full._sums, rtol=1e-12, atol=1e-6)


def test_sharded_run_without_communicator_raises(core, monkeypatch):
    import magpy_b200 as mp
    from magpy_b200 import sharding
    monkeypatch.delenv('RANK', raising=False)
    monkeypatch.delenv('WORLD_SIZE', raising=False)
    monkeypatch.setattr(sharding, '_WORLD', None)
    base = mp.Model([7e-9], [1e5], [[0, 0, 1.0]], [[1.0, 0, 0]], [[0, 0, 0.0]], 4e5, 0.1, 330.0)
    ens = mp.EnsembleModel(8, base)
    with pytest.raises(RuntimeError, match='communicator'):
        ens.simulate(1e-11, 1e-13, 5, 1, implicit_solve=False, shard=(0, 2))
    # world size 1 needs none
    ens.simulate(1e-11, 1e-13, 5, 1, implicit_solve=False, shard=(0, 1))


def test_comm_world_size_one(core):
    comm = core.Comm(None, 0, 1, 0)
    v = np.array([1.0, 2.0, 3.0])
    assert np.array_equal(comm.allreduce(v.copy()), v) and np.array_equal(comm.allreduce(v.copy(), 'max'), v)
    comm.barrier()
    args = ([7e-9], [1e5], [[0, 0, 1.0]], [[1.0, 0, 0]], [[0, 0, 0.0]], 4e5, 0.1, 330.0, False, True, False, 1e-13, 2e-11, 5)
    seeds = np.arange(64)
    a = core.simulate_ensemble(*args, seeds)
    b = core.simulate_ensemble(*args, seeds, comm=comm)
    assert np.array_equal(a['sums'], b['sums'])
    with pytest.raises(ValueError):
        core.Comm(b'short', 0, 2, 0)


NCCL_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, {root!r})
import magpy_b200 as mp
from magpy_b200 import core, sharding
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
comm = sharding.world_comm()                       # RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* -> NCCL communicator
assert (comm.rank, comm.world_size) == (rank, world)
v = comm.allreduce(np.array([rank + 1.0, 10.0 * rank]))
assert np.array_equal(v, [world * (world + 1) / 2, 10.0 * world * (world - 1) / 2]), v
assert comm.allreduce(np.array([float(rank)]), 'max')[0] == world - 1
base = mp.Model([7e-9], [1e5], [[0, 0, 1.0]], [[1.0, 0, 0]], [[0, 0, 0.0]], 4e5, 0.1, 330.0, field_shape='sine',
                field_frequency=5e9, field_amplitude=1e4)
R = 1001
ens = mp.EnsembleModel(R, base)
part = ens.simulate(4e-11, 1e-13, 9, 5, implicit_solve=False, shard=(rank, world), device=rank)     # device all-reduce
full = ens.simulate(4e-11, 1e-13, 9, 5, implicit_solve=False, device=rank)
lo, hi = sharding.shard_bounds(R, world, rank)
assert part.stats[0]['kernel'] == 'heun_single'
assert np.array_equal(part.final_state_array(), full.final_state_array()[lo:hi])
assert np.allclose(part._sums, full._sums, rtol=1e-12, atol=1e-6)
assert np.allclose(part.ensemble_magnetisation(), full.ensemble_magnetisation(), rtol=1e-12, atol=1e-9)
# several parameter groups: local plans, ONE host-staged all-reduce of the accumulated sums
ens2 = mp.EnsembleModel(R, base, field_frequency=[5e9, 7e9] * 500 + [5e9])
p2 = ens2.simulate(4e-11, 1e-13, 9, 5, implicit_solve=False, shard=(rank, world), device=rank)
f2 = ens2.simulate(4e-11, 1e-13, 9, 5, implicit_solve=False, device=rank)
assert np.array_equal(p2.final_state_array(), f2.final_state_array()[lo:hi])
assert np.allclose(p2._sums, f2._sums, rtol=1e-12, atol=1e-6)
comm.barrier()
print('rank', rank, 'ok')
'''


def test_two_nccl_ranks_sharded_ensemble(core, tmp_path):
    """world size 2, one process per GPU, the library's own NCCL communicator built from torchrun's environment."""
    if core.device_count() < 2:
        pytest.skip('needs two GPUs (run with gpurun --gpus 2)')
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / 'worker.py'
    script.write_text(NCCL_WORKER.format(root=ROOT))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1',
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
