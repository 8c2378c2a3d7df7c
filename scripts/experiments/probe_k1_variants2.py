"""K1 source variants (scripts/build_variant.sh) at the bench shape, plain launch vs the persistent kernel with 4 / 5 / 6 CTAs per SM."""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get('MB_ROOT', '.'))
import magpy_b200.core as core

def run(R, steps, env):
    for k in ('MAGPY_B200_K1_BALANCE', 'MAGPY_B200_K1_BAL_CTAS'):
        os.environ.pop(k, None)
    os.environ.update(env)
    seeds = np.random.default_rng(0).integers(0, 2**31 - 1, R)
    plan = core.EnsemblePlan([12e-9], [4e4], [[0, 0, 1.0]], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, False, True,
                             False, 1e-12, 1e-12 * steps, 101, seeds, field_shape='sine', field_amplitude=2e4,
                             field_frequency=3e5, gauss='f32p', return_trajectories=False)
    for i in range(2):
        plan.run(); st = plan.sync()
    cyc = st['integrate_ms'] * 1e-3 * 1.965e9 / steps / (R / 32 / 592)
    return '%7.2f ms (%5.1f cyc)' % (st['integrate_ms'], cyc)

R, steps = 1000000, 50000
print('%-20s' % os.environ.get('MB_ROOT', '.'), ' plain:', run(R, steps, {'MAGPY_B200_K1_BALANCE': '0'}),
      ' '.join(' bal%s: %s' % (c, run(R, steps, {'MAGPY_B200_K1_BAL_CTAS': c})) for c in ('4', '5', '6')), flush=True)
