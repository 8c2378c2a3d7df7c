"""Heun step in 37 fp64 instructions (half-unit form, -DMB_HEUN37) against the shipped 40-instruction form: the bench shape
(1M members), an eighth of it, and BASELINE config 1 (latency bound: the new form's dependent chain is one DFMA longer)."""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get('MB_ROOT', '.'))
import magpy_b200.core as core
print(core.__file__)

def run(R, steps, field='sine', axis=(0, 0, 1.0), renorm=False, dt=1e-12, S=101):
    seeds = np.random.default_rng(0).integers(0, 2**31 - 1, R)
    plan = core.EnsemblePlan([12e-9], [4e4], [list(axis)], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, renorm, True,
                             False, dt, dt * steps, S, seeds, field_shape=field, field_amplitude=2e4,
                             field_frequency=3e5, gauss='f32p', return_trajectories=False)
    for i in range(2):
        plan.run(); st = plan.sync()
    out = plan.fetch()
    print('R=%7d steps=%d %-8s axis=%s renorm=%d: %-28s %9.3f ms  %.4e particle-steps/s  <mz>=%.6f' % (
        R, steps, field, axis, renorm, st['kernel'], st['integrate_ms'], st['particle_steps'] / (st['integrate_ms'] * 1e-3),
        out['sums'][-1, 2] / R / 4e5), flush=True)

run(1000000, 100000)
run(1000000, 50000, axis=(0.6, 0, 0.8))
run(1000000, 50000, renorm=True)
run(125000, 100000)
run(1000, 100000, 'constant', dt=1e-14, S=1000)
run(1000, 100000, 'sine', dt=1e-14, S=1000)
