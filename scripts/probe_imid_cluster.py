"""Implicit-midpoint cluster kernels over N: matrix-product kernel (cluster_mma_imid.cu) vs scalar kernel (cluster.cu)."""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get('MB_ROOT', '.'))
import magpy_b200.core as core
from magpy_b200 import geometry

def run(N, R, steps, kernel, newton='reference'):
    os.environ['MAGPY_B200_CLUSTER_KERNEL'] = kernel
    axes = geometry.uniform_random_axes(N, rng=4)
    loc = geometry.random_cluster_coordinates(N, 3e-8, rng=4)
    seeds = np.arange(R)
    plan = core.EnsemblePlan(np.full(N, 12e-9), np.full(N, 4e4), axes, axes.copy(), loc, 4e5, 0.1, 300.0, False, True, True,
                             1e-12, 1e-12 * steps, 6, seeds, return_trajectories=False, implicit_newton=newton)
    for i in range(2):
        plan.run(); st = plan.sync()
    out = plan.fetch()
    it = st['newton_iterations'] / max(1, st['particle_steps'] / N)
    print('N=%3d R=%6d %-5s %-9s %-18s %8.2f ms  %.3e particle-steps/s  %.1f it/step  fails %d  <mz>=%.6f' % (
        N, R, kernel, newton, st['kernel'], st['integrate_ms'], st['particle_steps'] / (st['integrate_ms'] * 1e-3), it,
        st['newton_failures'], out['sums'][-1, 2] / R / N / 4e5), flush=True)

for N, R, steps in ((8, 65536, 100), (16, 32768, 100), (24, 16384, 50), (32, 16384, 50), (40, 16384, 40), (64, 9472, 40), (96, 4736, 20),
                    (128, 4736, 10)):
    for kernel in ('mma', 'simt'):
        if kernel == 'simt' and N > 40:
            continue
        run(N, R, steps, kernel)
run(16, 32768, 100, 'mma', 'exact')
run(64, 9472, 40, 'mma', 'exact')
