"""End-to-end breakdown of the small BASELINE configs (C1: 1000 x 1e5 Heun steps with trajectories; C2: 10k dimers,
implicit): host wall clock of the public API call against the library's own stage timers."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.environ.get('MB_ROOT', '.'))
import magpy_b200 as mp
from magpy_b200 import core

def show(name, ens, *a, **kw):
    for it in range(4):
        t0 = time.perf_counter()
        res = ens.simulate(*a, **kw)
        wall = 1e3 * (time.perf_counter() - t0)
        st = res.stats[0]
        if it:
            print('%s pass %d: wall %.2f ms | lib total %.2f = setup %.2f + run %.2f + fetch %.2f (+%.2f release) | device %.2f '
                  'integrate %.2f ms | h2d %d d2h %d B | kernel %s' % (
                      name, it, wall, st['host_total_ms'], st['host_setup_ms'], st['host_run_ms'], st['host_fetch_ms'],
                      st['host_total_ms'] - st['host_setup_ms'] - st['host_run_ms'] - st['host_fetch_ms'], st['device_ms'],
                      st['integrate_ms'], st['h2d_bytes'], st['d2h_bytes'], st['kernel']), flush=True)

c1 = mp.Model([12e-9], [4e4], [[0, 0, 1.0]], [[1.0, 0, 0]], [[0, 0, 0]], 4e5, 0.1, 300.0)
show('C1', mp.EnsembleModel(1000, c1), 1e-9, 1e-14, 1000, 1001, implicit_solve=False, return_trajectories=True)
show('C1 no traj', mp.EnsembleModel(1000, c1), 1e-9, 1e-14, 1000, 1001, implicit_solve=False, return_trajectories=False)
dimer = mp.Model([7e-9, 7e-9], [1e5, 1e5], [[0, 0, 1.0]] * 2, [[0, 0, 1.0]] * 2, [[0, 0, 0], [0, 0, 9e-9]], 4e5, 0.1, 330.0)
show('C2', mp.EnsembleModel(10000, dimer), 1e-9, 1e-12, 500, 1001, implicit_solve=True)
big = mp.EnsembleModel(1 << 20, c1)
show('1Mi x 101 samples traj (2.5 GB)', big, 1e-10, 1e-12, 101, 1001, implicit_solve=False, return_trajectories=True)
