"""BASELINE config 1 and other small single-particle Heun ensembles: K1s (heun_single_split.cu: integrator warp + generator
warp per 32 members) against the fused kernel (MAGPY_B200_K1_SPLIT=0)."""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get('MB_ROOT', '.'))
import magpy_b200.core as core

def run(R, steps, split, field='constant', axis=(0, 0, 1.0), renorm=False, dt=1e-14):
    os.environ['MAGPY_B200_K1_SPLIT'] = split
    seeds = np.arange(R) + 3
    plan = core.EnsemblePlan([12e-9], [4e4], [list(axis)], [[1.0, 0, 0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, renorm, True,
                             False, dt, dt * steps, 1000, seeds, field_shape=field, field_amplitude=2e4,
                             field_frequency=3e5, gauss='f32p', return_trajectories=True)
    for i in range(2):
        plan.run(); st = plan.sync()
    print('R=%6d steps=%d %-9s axis=%s renorm=%d split=%s: variant %3d  %8.3f ms  (%.1f cycles/step at 1965 MHz)  %.3e particle-steps/s' % (
        R, steps, field, axis, renorm, split, st['kernel_variant'], st['integrate_ms'], st['integrate_ms'] * 1e-3 * 1.965e9 / steps,
        st['particle_steps'] / (st['integrate_ms'] * 1e-3)), flush=True)

for R in (1000, 4736, 9472, 18944):
    for split in ('0', '1'):
        run(R, 100000, split)
for split in ('0', '1'):
    run(1000, 100000, split, 'sine')
    run(1000, 100000, split, 'sine', (0.6, 0, 0.8), True)
