"""BASELINE config 1 (1000 x 1e5 Heun steps, dt = 1e-14 s) and small ensembles in general: heun_single's free-register
variant against its latency variant (applied-field table entries fetched one step pair ahead).  The first version of this
probe compared a producer / consumer split kernel (scripts/experiments/heun_single_split.cu): see
profiles/r02_probe_c1_split_kernel.log — slower, because a step is bound by the integrator's own dependent chain."""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get('MB_ROOT', '.'))
import magpy_b200.core as core

def run(R, steps, split, field='constant', axis=(0, 0, 1.0), renorm=False, dt=1e-14):
    os.environ['MAGPY_B200_K1_MIN_BLOCKS'] = split
    seeds = np.arange(R) + 3
    plan = core.EnsemblePlan([12e-9], [4e4], [list(axis)], [[1.0, 0, 0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, renorm, True,
                             False, dt, dt * steps, 1000, seeds, field_shape=field, field_amplitude=2e4,
                             field_frequency=3e5, gauss='f32p', return_trajectories=True)
    for i in range(2):
        plan.run(); st = plan.sync()
    print('R=%6d steps=%d %-9s axis=%s renorm=%d variant=%s: %-18s %.3f ms  (%.1f cycles/step at 1965 MHz)  %.3e particle-steps/s' % (
        R, steps, field, axis, renorm, split, st['kernel'], st['integrate_ms'], st['integrate_ms'] * 1e-3 * 1.965e9 / steps,
        st['particle_steps'] / (st['integrate_ms'] * 1e-3)), flush=True)

run(1000, 100000, '1')
for R in (1000, 9472, 18944, 37888, 75776):
    for variant in ('1', '100'):
        run(R, 100000, variant, 'sine')
for variant in ('1', '100'):
    run(1000, 100000, variant, 'sine', (0.6, 0, 0.8), True)
