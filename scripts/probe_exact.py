"""GPU probe: implicit midpoint, reference quasi-Newton vs exact-Jacobian Newton (implicit_newton='exact')."""
import sys
import numpy as np
sys.path.insert(0, '.')
import magpy_b200.core as core


def run(N, R, steps, newton, dt=1e-12, axis_z=True):
    rng = np.random.default_rng(1)
    radius = np.full(N, 12e-9 if N == 1 else 7e-9); K = np.full(N, 4e4 if N == 1 else 1e5)
    axis = np.tile([0, 0, 1.0], (N, 1)) if axis_z else rng.normal(size=(N, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    m0 = np.tile([1.0, 0, 0] if N == 1 else [0, 0, 1.0], (N, 1))
    loc = np.cumsum(np.full((N, 3), [0, 0, 9e-9]), axis=0)
    seeds = rng.integers(0, 2**31 - 1, R)
    plan = core.EnsemblePlan(radius, K, axis, m0, loc, 4e5, 0.1, 300.0 if N == 1 else 330.0, False, True, True, dt, dt * steps,
                             101, seeds, return_trajectories=False, implicit_newton=newton)
    for _ in range(2):
        plan.run(); st = plan.sync()
    ps = st['particle_steps'] / (st['integrate_ms'] * 1e-3)
    it = st['newton_iterations'] / max(1, st['particle_steps'] / N)
    mz = plan.fetch()['sums'][-1, 2] / R / 4e5 / N
    print(f'N={N} R={R:8d} steps={steps} {newton:9s} axis_z={axis_z}: {st["integrate_ms"]:8.2f} ms  {ps:.3e} particle-steps/s  '
          f'{it:.2f} it/step  <mz>={mz:.5f}', flush=True)


for newton in ('reference', 'exact'):
    run(1, 1_000_000, 1000, newton)                 # config 5 point
    run(1, 1_000_000, 1000, newton, axis_z=False)
    run(2, 10_000, 1000, newton)                    # config 2
    run(2, 1 << 18, 1000, newton)
    run(4, 1 << 17, 1000, newton)
