"""Persistent K1 kernel with fewer resident CTAs per SM (MAGPY_B200_K1_BAL_CTAS): the youngest CTAs of an SM are starved
by the oldest-first warp scheduler and hold their blocks for milliseconds — does leaving them out shorten the tail?"""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get('MB_ROOT', '.'))
import magpy_b200.core as core

def run(R, steps, ctas):
    os.environ['MAGPY_B200_K1_BALANCE'] = '1'
    os.environ['MAGPY_B200_K1_MIN_BLOCKS'] = '1'
    os.environ['MAGPY_B200_K1_BAL_CTAS'] = str(ctas)
    seeds = np.random.default_rng(0).integers(0, 2**31 - 1, R)
    plan = core.EnsemblePlan([12e-9], [4e4], [[0, 0, 1.0]], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, False, True, False,
                             1e-12, 1e-12 * steps, 101, seeds, field_shape='sine', field_amplitude=2e4, field_frequency=3e5,
                             gauss='f32p', return_trajectories=False)
    for i in range(2):
        plan.run(); st = plan.sync()
    return st['integrate_ms']

for R in (125000, 250000, 1000000):
    print('R=%d, ms per 100,000 steps with 3 / 4 / 5 / 6 CTAs per SM:' % R, ['%.2f' % run(R, 100000, c) for c in (3, 4, 5, 6)], flush=True)
