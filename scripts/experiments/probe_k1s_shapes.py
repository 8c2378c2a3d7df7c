import os, sys
import numpy as np
sys.path.insert(0, '.')
sys.path.insert(0, 'scripts')
import magpy_b200.core as core
def run(R, steps, split, field='constant', axis=(0, 0, 1.0), renorm=False, dt=1e-14):
    os.environ['MAGPY_B200_K1_SPLIT'] = split
    seeds = np.arange(R) + 3
    plan = core.EnsemblePlan([12e-9], [4e4], [list(axis)], [[1.0, 0, 0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, renorm, True,
                             False, dt, dt * steps, 1000, seeds, field_shape=field, field_amplitude=2e4,
                             field_frequency=3e5, gauss='f32p', return_trajectories=True)
    for i in range(2):
        plan.run(); st = plan.sync()
    return st['integrate_ms'], st['kernel_variant']
for R in (4736, 9472):
    for field, axis, renorm in (('constant', (0,0,1.0), True), ('sine', (0.6,0,0.8), False), ('sine', (0.6,0,0.8), True), ('constant', (0.6,0,0.8), False)):
        a = run(R, 50000, '0', field, axis, renorm); b = run(R, 50000, '1', field, axis, renorm)
        print('R=%5d %-8s axis=%s renorm=%d: fused (variant %d) %.3f ms   split %.3f ms' % (R, field, axis, renorm, a[1], a[0], b[0]), flush=True)
