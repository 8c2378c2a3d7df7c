"""Scratch GPU probe: Heun single-particle throughput (bench workload shape, shorter)."""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get('MB_ROOT', '.'))
import magpy_b200.core as core
print(core.__file__)

def run(R, steps, axis=(0, 0, 1.0), field='sine', gauss='f32', S=21, dt=1e-12, renorm=False):
    seeds = np.random.default_rng(0).integers(0, 2**31 - 1, R)
    plan = core.EnsemblePlan([12e-9], [4e4], [list(axis)], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, renorm, True,
                             False, dt, dt * steps, S, seeds, field_shape=field, field_amplitude=2e4,
                             field_frequency=3e5, gauss=gauss, return_trajectories=False)
    for i in range(2):
        plan.run(); st = plan.sync()
    out = plan.fetch()
    ps = st['particle_steps'] / (st['integrate_ms'] * 1e-3)
    print(f'R={R} steps={steps} axis={axis} field={field} gauss={gauss} renorm={renorm}: {st["integrate_ms"]:.2f} ms, '
          f'{ps:.4e} particle-steps/s, <mz>={out["sums"][-1,2]/R/4e5:.5f}', flush=True)

for g in (sys.argv[1:] or ['f32p', 'f32']):
    run(1000000, 20000, gauss=g)
    run(1000000, 20000, axis=(0.6, 0, 0.8), gauss=g)
    run(1000000, 20000, field='constant', gauss=g)
    run(1000000, 10000, renorm=True, gauss=g)
    run(148 * 2048 * 3, 20000, gauss=g)
