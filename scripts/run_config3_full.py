"""BASELINE config 3 at full size through the public API: 1,000,000 single-particle members under a 300 kHz /
20 kA/m sinusoidal field, Heun, dt = 1e-12 s, TWO field periods (6.67e6 steps per member, 6.67e12 particle-steps),
1000 samples per period; reports the throughput, the final-cycle hysteresis energy and the SAR, and checks the
size-independent properties a hysteresis run offers (periodic steady state, odd symmetry of the loop)."""
import json, sys, time
import numpy as np
sys.path.insert(0, '.')
import magpy_b200 as mp

f, H0 = 3e5, 2e4
radius = float(sys.argv[1]) if len(sys.argv) > 1 else 12e-9
R = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
base = mp.Model([radius], [4e4], [[0, 0, 1.0]], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0,
                field_shape='sine', field_frequency=f, field_amplitude=H0)
ens = mp.EnsembleModel(R, base)
t0 = time.perf_counter()
res = ens.simulate(end_time=2 / f, time_step=1e-12, max_samples=2001, random_state=1001, renorm=True,
                   implicit_solve=False, return_trajectories=False, devices='all')
wall = time.perf_counter() - t0
st = res.stats[0]
mz = res.ensemble_magnetisation() / 4e5
E = res.final_cycle_energy_dissipated(f)
half = 500                                  # samples per half period
loop2 = mz[1000:2001]
line = {
    'config': 'C3 full: %d x 1 particle r=%.1f nm, sine 300 kHz / 20 kA/m, Heun dt=1e-12, 2 periods, renorm' % (R, radius * 1e9),
    'steps_per_member': st['steps_per_member'], 'particle_steps': st['particle_steps'], 'wall_s': wall,
    'device_ms': st['device_ms'], 'particle_steps_per_s_e2e': st['particle_steps'] / wall,
    'kernel_launches': st['kernel_launches'],
    'final_cycle_energy_J_per_m3': E, 'SAR_W_per_kg': res.specific_absorption_rate(f),
    'mz_min_max_last_cycle': [float(loop2.min()), float(loop2.max())],
    'period_to_period_max_diff': float(np.abs(mz[1000:2000] - mz[0:1000])[200:].max()),
    'odd_symmetry_max_residual': float(np.abs(loop2[:half] + loop2[half:2 * half]).max()),
    'max_stderr': float((res.ensemble_magnetisation_stderr() / 4e5).max()),
}
print(json.dumps(line), flush=True)
