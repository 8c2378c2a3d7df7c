"""One short launch of a chosen kernel for `ncu --set full` (see profiles/README.md)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import magpy_b200.core as core

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1
implicit = len(sys.argv) > 2 and sys.argv[2] == 'implicit'
R = int(sys.argv[3]) if len(sys.argv) > 3 else (1_000_000 if N == 1 else 1 << 14)
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
dt = 1e-12
radius = np.full(N, 12e-9); K = np.full(N, 4e4)
axis = np.tile([0, 0, 1.0], (N, 1)); m0 = np.tile([0, 0, 1.0], (N, 1))
loc = np.cumsum(np.full((N, 3), 3e-8), axis=0)
seeds = np.random.default_rng(0).integers(0, 2**31 - 1, R)
plan = core.EnsemblePlan(radius, K, axis, m0, loc, 4e5, 0.1, 300.0, False, True, implicit, dt, dt * steps, 21, seeds,
                         field_shape='sine', field_amplitude=2e4, field_frequency=3e5, return_trajectories=False)
plan.run()
st = plan.sync()
print(st['particle_steps'] / st['integrate_ms'] / 1e-3, 'particle-steps/s')
