"""Resident CTAs per SM of the persistent K1 kernel after the 37-instruction rework, over the strong-scaling shard sizes."""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get('MB_ROOT', '.'))
import magpy_b200.core as core

def run(R, steps, ctas):
    os.environ['MAGPY_B200_K1_BAL_CTAS'] = ctas
    seeds = np.random.default_rng(0).integers(0, 2**31 - 1, R)
    plan = core.EnsemblePlan([12e-9], [4e4], [[0, 0, 1.0]], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, False, True,
                             False, 1e-12, 1e-12 * steps, 101, seeds, field_shape='sine', field_amplitude=2e4,
                             field_frequency=3e5, gauss='f32p', return_trajectories=False)
    for i in range(2):
        plan.run(); st = plan.sync()
    return st['integrate_ms'], st['kernel_variant']

for R in (1000000, 500000, 250000, 125000):
    row = []
    for c in ('3', '4', '5', '6'):
        ms, v = run(R, 100000, c)
        row.append('%s CTAs/SM: %7.2f ms (%.4e)' % (c, ms, R * 1e5 / (ms * 1e-3)))
    print('R=%8d  ' % R + '   '.join(row), flush=True)
