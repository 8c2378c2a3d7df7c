"""Scratch GPU probe: FP64 peak and raw throughput of the integration kernels."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import magpy_b200.core as core

print('devices', core.device_count())
print('fp64 peak TFLOP/s, max MHz', core.fp64_peak())

def run(R, steps, N=1, implicit=False, field='sine', gauss='f32', S=101, dt=5e-12, inter=True):
    rng = np.random.default_rng(0)
    radius = np.full(N, 12e-9); K = np.full(N, 4e4)
    axis = np.tile([0, 0, 1.0], (N, 1)); m0 = np.tile([1.0, 0, 0], (N, 1))
    loc = np.cumsum(np.full((N, 3), 3e-8), axis=0)
    seeds = rng.integers(0, 2**31 - 1, R)
    plan = core.EnsemblePlan(radius, K, axis, m0, loc, 4e5, 0.1, 300.0, False, inter, implicit, dt, dt * steps, S, seeds,
                             field_shape=field, field_amplitude=2e4, field_frequency=3e5, gauss=gauss,
                             return_trajectories=False)
    for i in range(2):
        plan.run(); st = plan.sync()
    out = plan.fetch()
    ps = st['particle_steps'] / (st['integrate_ms'] * 1e-3)
    print(f'R={R} N={N} steps={st["steps_per_member"]} implicit={implicit} field={field} gauss={gauss}: '
          f'{st["integrate_ms"]:.2f} ms integrate, {st["device_ms"]:.2f} ms device, {ps:.3e} particle-steps/s, '
          f'launches={st["kernel_launches"]} newton_it/step={st["newton_iterations"]/max(1,st["particle_steps"]/N):.2f}',
          flush=True)
    return out

out = run(1 << 20, 20000)
print('mean mz/Ms last', out['sums'][-1, 2] / (1 << 20) / 4e5)
run(1 << 20, 20000, field='constant')
run(1 << 20, 5000, gauss='f64')
run(1000000, 20000)
run(1 << 18, 2000, implicit=True)
run(1 << 16, 2000, N=2)
run(1 << 14, 1000, N=8)
run(1 << 12, 500, N=64)
run(1 << 14, 500, N=2, implicit=True)
