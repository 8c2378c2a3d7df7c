"""Where does the end-to-end time of EnsembleModel.simulate go? (host wall clock vs device ms)"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import magpy_b200 as mp
from magpy_b200 import core

R = 1_000_000
base = mp.Model([12e-9], [4e4], [[0, 0, 1.0]], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0,
                field_shape='sine', field_frequency=3e5, field_amplitude=2e4)
ens = mp.EnsembleModel(R, base)
for it in range(3):
    t0 = time.perf_counter()
    seeds = ens._member_seeds(1001)
    t1 = time.perf_counter()
    res = ens.simulate(1e-7, 1e-12, 101, 1001, implicit_solve=False, return_trajectories=False)
    t2 = time.perf_counter()
    st = res.stats[0]
    print('pass %d: seeds %.1f ms, simulate %.1f ms (device %.1f ms, integrate %.1f ms, h2d %d B, d2h %d B)' %
          (it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), st['device_ms'], st['integrate_ms'], st['h2d_bytes'], st['d2h_bytes']), flush=True)
# raw C-ABI call with prepared arrays
seeds = ens._member_seeds(1001)
for it in range(2):
    t0 = time.perf_counter()
    out = core.simulate_ensemble([12e-9], [4e4], [[0, 0, 1.0]], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, False, True, False,
                                 1e-12, 1e-7, 101, seeds, 'sine', 2e4, 3e5, return_trajectories=False)
    t1 = time.perf_counter()
    print('core.simulate_ensemble %.1f ms (device %.1f)' % (1e3 * (t1 - t0), out['stats']['device_ms']), flush=True)
