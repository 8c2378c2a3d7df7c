"""K1 with renorm=True (Taylor-series renormalisation) and the general-axis kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get('MB_ROOT', '.'))
import magpy_b200.core as core
def run(R, steps, renorm, axis=(0, 0, 1.0)):
    seeds = np.random.default_rng(0).integers(0, 2**31 - 1, R)
    plan = core.EnsemblePlan([12e-9], [4e4], [list(axis)], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, renorm, True,
                             False, 1e-12, 1e-12 * steps, 21, seeds, field_shape='sine', field_amplitude=2e4,
                             field_frequency=3e5, gauss='f32p', return_trajectories=False)
    for i in range(2):
        plan.run(); st = plan.sync()
    out = plan.fetch()
    print('renorm', renorm, 'axis', axis, '%.4e particle-steps/s' % (st['particle_steps'] / (st['integrate_ms'] * 1e-3)),
          '|m|-1 max', np.abs(np.linalg.norm(out['final'][:, 0], axis=1) / 4e5 - 1).max(), flush=True)
run(1000000, 20000, False)
run(1000000, 20000, True)
run(1000000, 20000, True, (0.6, 0, 0.8))
run(1000000, 20000, False, (0.6, 0, 0.8))
