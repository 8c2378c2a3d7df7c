"""K1 register-allocation variants over shard sizes (strong scaling of C3 on one GPU's share of the members)."""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get('MB_ROOT', '.'))
import magpy_b200.core as core

def run(R, steps, renorm=False, axis=(0, 0, 1.0)):
    seeds = np.random.default_rng(0).integers(0, 2**31 - 1, R)
    plan = core.EnsemblePlan([12e-9], [4e4], [list(axis)], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, renorm, True,
                             False, 1e-12, 1e-12 * steps, 21, seeds, field_shape='sine', field_amplitude=2e4,
                             field_frequency=3e5, gauss='f32p', return_trajectories=False)
    for i in range(2):
        plan.run(); st = plan.sync()
    return st['particle_steps'] / (st['integrate_ms'] * 1e-3), st['integrate_ms'], st['kernel_variant']

base = None
for R in (1000000, 500000, 250000, 125000, 62500, 132608, 151552, 113664, 300000):
    row = []
    for v in ('1', '7', '8', ''):
        if v:
            os.environ['MAGPY_B200_K1_MIN_BLOCKS'] = v
        else:
            os.environ.pop('MAGPY_B200_K1_MIN_BLOCKS', None)
        rate, ms, var = run(R, 20000)
        row.append('%s: %.4e (%.2f ms, v%d)' % (v or 'auto', rate, ms, var))
    print('R=%8d  ' % R + '   '.join(row), flush=True)
for renorm, axis in ((True, (0, 0, 1.0)), (False, (0.6, 0, 0.8))):
    for v in ('1', '7', '8'):
        os.environ['MAGPY_B200_K1_MIN_BLOCKS'] = v
        print('renorm', renorm, 'axis', axis, 'v', v, ['%.4e' % run(R, 10000, renorm, axis)[0] for R in (1000000, 125000)], flush=True)
