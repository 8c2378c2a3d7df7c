"""Run BASELINE.json configs 1, 2, 4 and 5 through the public API and report throughput (particle-steps/s).

    python scripts/run_configs.py                      # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/run_configs.py

Under torchrun every rank integrates its contiguous slice of the members (magpy_b200.sharding) and the
ensemble sums are all-reduced once (ncclAllReduce issued by libmagpy_b200).  Times are the library's CUDA-event device times, max over ranks.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

rank = int(os.environ.get('RANK', '0'))
local_rank = int(os.environ.get('LOCAL_RANK', '0'))
world = int(os.environ.get('WORLD_SIZE', '1'))
import magpy_b200 as mp  # noqa: E402
from magpy_b200 import core, geometry  # noqa: E402

comm = core.Comm.from_env(local_rank) if world > 1 else None   # the library's own NCCL communicator


ONLY = [a for a in sys.argv[1:] if a.startswith('C')]   # e.g. `run_configs.py C4` runs config 4 alone


def run(name, model, R, end_time, time_step, S, implicit, traj=False, **kw):
    if ONLY and name.split()[0] not in ONLY:
        return
    ens = mp.EnsembleModel(R, model)
    shard = (rank, world) if world > 1 else None
    out = None
    for _ in range(2):   # first pass warms the context / memory pool
        t0 = time.perf_counter()
        out = ens.simulate(end_time, time_step, S, 1001, implicit_solve=implicit, device=local_rank, shard=shard,
                           comm=comm, return_trajectories=traj, **kw)
        wall = time.perf_counter() - t0
    st = out.stats[0]
    ms = st['device_ms']
    if comm is not None:
        ms, wall = comm.allreduce(np.array([ms, wall]), 'max')
    N = len(np.atleast_1d(model.radius))
    total = R * N * st['steps_per_member']
    if rank == 0:
        mz = out.ensemble_magnetisation('z') / (N * model.magnetisation)
        line = {'config': name, 'n_gpus': world, 'members': R, 'particles': N, 'steps': st['steps_per_member'],
                'device_ms': ms, 'wall_ms': 1e3 * wall, 'particle_steps_per_s_device': total / (ms * 1e-3),
                'particle_steps_per_s_e2e': total / wall,
                'newton_iterations_per_step': st['newton_iterations'] / max(1, st['particle_steps'] / N),
                'newton_failures': st['newton_failures'], 'mz_first_last': [float(mz[0]), float(mz[-1])]}
        print(json.dumps(line), flush=True)


# config 1: 1000 members of the 12 nm particle, Heun dt = 1e-14 s to 1e-9 s, 1000 samples, per-member trajectories kept
c1 = mp.Model([12e-9], [4e4], [[0, 0, 1.0]], [[1.0, 0, 0]], [[0, 0, 0]], 4e5, 0.1, 300.0)
run('C1 single Heun 1000 x 1e5 steps, trajectories', c1, 1000, 1e-9, 1e-14, 1000, False, traj=True)

# config 2: two dipolar-coupled 7 nm particles 9 nm apart (two-particle-equilibrium notebook), implicit, 10k members
dimer = mp.Model([7e-9, 7e-9], [1e5, 1e5], [[0, 0, 1.0]] * 2, [[0, 0, 1.0]] * 2, [[0, 0, 0], [0, 0, 9e-9]], 4e5, 0.1, 330.0)
run('C2 dimer implicit 10k x 1000 steps', dimer, 10000, 1e-9, 1e-12, 500, True)

# config 4: 64-particle random-geometry clusters, all-pairs dipolar, Heun dt = 1e-14 s, 1e4 steps, 100k members
N = 64
axes = geometry.uniform_random_axes(N, rng=4)
cluster = mp.Model(np.full(N, 12e-9), np.full(N, 4e4), axes, axes.copy(),
                   geometry.random_cluster_coordinates(N, 3e-8, rng=4), 4e5, 0.1, 300.0)
run('C4 64-particle clusters Heun 100k x 10000 steps', cluster, 100000, 1e-10, 1e-14, 101, False)

# config 5 (one point of the sweep): 1M single-particle members, implicit midpoint, dt = 1e-12 s
single = mp.Model([12e-9], [4e4], [[0, 0, 1.0]], [[1.0, 0, 0]], [[0, 0, 0]], 4e5, 0.1, 300.0)
run('C5 single implicit 1M x 1000 steps', single, 1000000, 1e-9, 1e-12, 101, True)

if comm is not None:
    comm.barrier()
