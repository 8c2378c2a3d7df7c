import os, sys
import numpy as np
sys.path.insert(0, '.')
import magpy_b200.core as core
def run(R, steps, env, field='constant'):
    for k in ('MAGPY_B200_K1_SPLIT', 'MAGPY_B200_K1_SPLIT_PRODUCERS'): os.environ.pop(k, None)
    os.environ.update(env)
    seeds = np.arange(R) + 3
    plan = core.EnsemblePlan([12e-9], [4e4], [[0, 0, 1.0]], [[1.0, 0, 0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, False, True,
                             False, 1e-14, 1e-14 * steps, 1000, seeds, field_shape=field, field_amplitude=2e4,
                             field_frequency=3e5, gauss='f32p', return_trajectories=True)
    for i in range(2):
        plan.run(); st = plan.sync()
    return '%.3f ms (variant %d)' % (st['integrate_ms'], st['kernel_variant'])
for R in (1000, 4736, 6000, 9472, 12000, 14208, 18944):
    print('R=%6d' % R, ' fused:', run(R, 100000, {'MAGPY_B200_K1_SPLIT': '0'}), ' K1s 3 producers:', run(R, 100000, {'MAGPY_B200_K1_SPLIT': '1', 'MAGPY_B200_K1_SPLIT_PRODUCERS': '3'}),
          ' K1s 1 producer:', run(R, 100000, {'MAGPY_B200_K1_SPLIT': '1', 'MAGPY_B200_K1_SPLIT_PRODUCERS': '1'}), ' default:', run(R, 100000, {}), flush=True)
print('sine R=9472 fused:', run(9472, 100000, {'MAGPY_B200_K1_SPLIT': '0'}, 'sine'), ' default:', run(9472, 100000, {}, 'sine'))
