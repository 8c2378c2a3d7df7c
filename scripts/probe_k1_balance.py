"""K1 as a persistent task-queue kernel (heun_single_balanced.cu) against the plain launch over shard sizes: the strong-
scaling shares of BASELINE config 3 (1M members over 1, 2, 4, 8 GPUs) and sizes around whole waves."""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get('MB_ROOT', '.'))
import magpy_b200.core as core

def run(R, steps, balance, renorm=False, axis=(0, 0, 1.0)):
    os.environ['MAGPY_B200_K1_BALANCE'] = balance
    seeds = np.random.default_rng(0).integers(0, 2**31 - 1, R)
    plan = core.EnsemblePlan([12e-9], [4e4], [list(axis)], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, renorm, True,
                             False, 1e-12, 1e-12 * steps, 101, seeds, field_shape='sine', field_amplitude=2e4,
                             field_frequency=3e5, gauss='f32p', return_trajectories=False)
    for i in range(2):
        plan.run(); st = plan.sync()
    return st['particle_steps'] / (st['integrate_ms'] * 1e-3), st['integrate_ms'], st['kernel_variant']

base = None
for R in (1000000, 500000, 250000, 125000, 132608, 151552, 200000, 113664, 100000, 80000):
    row = []
    for b in ('0', '1'):
        rate, ms, var = run(R, 100000, b)
        if base is None:
            base = rate
        row.append('balance=%s: %.4e (%.2f ms, variant %d, %.3f of the 1M plain rate)' % (b, rate, ms, var, rate / base))
    print('R=%8d  ' % R + '   '.join(row), flush=True)
for renorm, axis in ((True, (0, 0, 1.0)), (False, (0.6, 0, 0.8))):
    for b in ('0', '1'):
        print('renorm', renorm, 'axis', axis, 'balance', b, ['%.4e' % run(R, 50000, b, renorm, axis)[0] for R in (1000000, 125000)], flush=True)
