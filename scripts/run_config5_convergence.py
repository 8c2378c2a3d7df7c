"""BASELINE config 5: strong / weak convergence sweep of the implicit midpoint rule (and Heun) over 1,000,000
realisations on COMMON Brownian paths — coarse increments are sums of fine ones (test/convergence/task5.cpp:150-158),
formed inside the kernel from the packed Philox stream (noise_coarsen_log2), analysed like
docs/source/notebooks/convergence.ipynb cells 29-38 (Cauchy differences between consecutive step sizes).

    dt = 1e-15 s * 2^L, L = 1..10  (2e-15 ... 1.024e-12 s), horizon T = 2.048e-11 s (20 steps at the coarsest dt)

The C1 particle (12 nm, K = 4e4, Ms = 4e5, alpha = 0.1, T = 300 K), m0 tilted off the axis so that the weak error of
<m_z> does not vanish by symmetry.  Prints one JSON line per scheme; the third line repeats the implicit sweep with the opt-in
exact-Jacobian Newton (implicit_newton='exact'): same Brownian paths, same scheme, ~2 iterations per step."""
import json, sys, time
import numpy as np
sys.path.insert(0, '.')
import magpy_b200.core as core

radius, K, Ms, alpha, T = [12e-9], [4e4], 4e5, 0.1, 300.0
axis, m0, loc = [[0, 0, 1.0]], [[0.6, 0, 0.8]], [[0, 0, 0.0]]


def sweep(R, dt_fine, levels, n_coarsest, implicit, newton='reference'):
    t0 = time.perf_counter()
    seeds = np.random.default_rng(5).integers(0, 2 ** 31 - 1, R)
    horizon = dt_fine * (1 << levels[-1]) * n_coarsest
    finals, iters, steps, dev_ms = [], [], 0, 0.0
    for L in levels:
        dt = dt_fine * (1 << L)
        out = core.simulate_ensemble(radius, K, axis, m0, loc, Ms, alpha, T, False, True, implicit, dt,
                                     horizon * (1 + 1e-9), 2, seeds, implicit_tol=1e-9, return_trajectories=False,
                                     noise_coarsen_log2=L, implicit_newton=newton)
        st = out['stats']
        assert st['steps_per_member'] == n_coarsest << (levels[-1] - L), (st['steps_per_member'], L)
        steps += st['particle_steps']; dev_ms += st['device_ms']
        iters.append(st['newton_iterations'] / st['particle_steps'])
        finals.append(out['final'][:, 0, :] / Ms)
    n = len(levels)
    strong = [float(np.linalg.norm(finals[i + 1] - finals[i], axis=1).mean()) for i in range(n - 1)]
    dz = [finals[i + 1][:, 2] - finals[i][:, 2] for i in range(n - 1)]
    weak = [float(abs(d.mean())) for d in dz]
    weak_se = [float(d.std() / np.sqrt(R)) for d in dz]
    x = np.arange(n - 1)
    sig = [int(i) for i in x if weak[i] > 3 * weak_se[i]]
    print(json.dumps({
        'scheme': ('implicit midpoint' + (' (exact-Jacobian Newton, opt-in)' if newton == 'exact' else '')) if implicit else 'Heun', 'realisations': R, 'horizon_s': horizon,
        'dt_s': [dt_fine * (1 << L) for L in levels],
        'strong_cauchy_diff': strong, 'strong_order': float(np.polyfit(x, np.log2(strong), 1)[0]),
        'strong_local_slopes': [float(v) for v in np.diff(np.log2(strong))],
        'weak_cauchy_diff_mz': weak, 'weak_se': weak_se, 'significant_levels': sig,
        'weak_order_over_significant_levels':
            float(np.polyfit(np.array(sig), np.log2(np.array(weak)[sig]), 1)[0]) if len(sig) >= 3 else None,
        'newton_iterations_per_step': iters, 'particle_steps': steps, 'device_ms': dev_ms,
        'wall_s': time.perf_counter() - t0}), flush=True)


if __name__ == '__main__':
    R = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
    for implicit, newton in ((True, 'reference'), (False, 'reference'), (True, 'exact')):
        sweep(R, 1e-15, list(range(1, 11)), 20, implicit, newton)
