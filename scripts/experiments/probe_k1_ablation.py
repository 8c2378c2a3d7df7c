"""Timing ablation of K1 (WRONG numerics on purpose, timing only): which of the generator's instruction groups the step
time of the bench kernel is made of.  Variants built with scripts/build_variant.sh (-DMB_ABL_*)."""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get('MB_ROOT', '.'))
import magpy_b200.core as core

def run(R, steps):
    seeds = np.random.default_rng(0).integers(0, 2**31 - 1, R)
    plan = core.EnsemblePlan([12e-9], [4e4], [[0, 0, 1.0]], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, False, True,
                             False, 1e-12, 1e-12 * steps, 101, seeds, field_shape='sine', field_amplitude=2e4,
                             field_frequency=3e5, gauss='f32p', return_trajectories=False)
    for i in range(2):
        plan.run(); st = plan.sync()
    warps_per_subpartition = R / 32 / 592
    cyc = st['integrate_ms'] * 1e-3 * 1.965e9 / steps / warps_per_subpartition
    print('%-22s R=%7d: %8.3f ms  %.4e particle-steps/s  %.1f cycles per warp-step per sub-partition' % (
        os.environ.get('MB_ROOT', '.'), R, st['integrate_ms'], st['particle_steps'] / (st['integrate_ms'] * 1e-3), cyc), flush=True)

run(1000000, 50000)
run(1000, 50000)
