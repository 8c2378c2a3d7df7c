"""GPU probe: trajectory write-out path (sampled points stored by the kernel as [S][n][R], transposed to the
API layout [R][N][3][S] on fetch).  Reports the transpose kernel's effective HBM bandwidth."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import magpy_b200.core as core

R, S, steps = 1 << 20, 101, 2020
seeds = np.arange(R)
plan = core.EnsemblePlan([12e-9], [4e4], [[0, 0, 1.0]], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, False, True, False,
                         1e-12, 1e-12 * steps, S, seeds, return_trajectories=True)
for _ in range(2):
    plan.run(); st = plan.sync()
print('integrate %.2f ms for %d steps incl. %d samples x %.1f MB coalesced stores' % (st['integrate_ms'], steps, S, R * 24 / 1e6))
t0 = time.perf_counter(); out = plan.fetch(); t1 = time.perf_counter()
print('fetch (transpose + D2H of %.2f GB into pageable numpy) %.1f ms' % (out['trajectories'].nbytes / 1e9, 1e3 * (t1 - t0)))
t0 = time.perf_counter(); out = plan.fetch(); t1 = time.perf_counter()
print('fetch again %.1f ms' % (1e3 * (t1 - t0)))
