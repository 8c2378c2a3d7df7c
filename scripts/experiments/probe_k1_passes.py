"""Where do the ~5 ms go that a 125,000-member launch of K1 loses against its share of the 1M run?  Step counts and
segment counts varied at fixed ensemble size (integrate ms; second of two passes)."""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get('MB_ROOT', '.'))
import magpy_b200.core as core

def run(R, steps, balance, segs=None, S=101, field='sine'):
    os.environ['MAGPY_B200_K1_BALANCE'] = balance
    os.environ['MAGPY_B200_K1_MIN_BLOCKS'] = '1'
    if segs:
        os.environ['MAGPY_B200_K1_SEGMENTS'] = str(segs)
    else:
        os.environ.pop('MAGPY_B200_K1_SEGMENTS', None)
    seeds = np.random.default_rng(0).integers(0, 2**31 - 1, R)
    plan = core.EnsemblePlan([12e-9], [4e4], [[0, 0, 1.0]], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, False, True, False,
                             1e-12, 1e-12 * steps, S, seeds, field_shape=field, field_amplitude=2e4, field_frequency=3e5,
                             gauss='f32p', return_trajectories=False)
    for i in range(2):
        plan.run(); st = plan.sync()
    return st['integrate_ms'], st['kernel_variant']

print('plain kernel, one full wave (113,664 members) and 1M, ms per 1000 steps:')
for R in (113664, 1000000):
    print('  R=%d' % R, ['%d steps: %.4f' % (n, run(R, n, '0')[0] / n * 1000) for n in (10000, 20000, 50000, 100000, 200000)], flush=True)
print('plain kernel, constant field (no table), one full wave:')
print('  R=113664', ['%d steps: %.4f' % (n, run(113664, n, '0', field='constant')[0] / n * 1000) for n in (20000, 100000)], flush=True)
print('balanced kernel, 125,000 members, ms per 1000 steps:')
print('  ', ['%d steps: %.4f' % (n, run(125000, n, '1')[0] / n * 1000) for n in (20000, 50000, 100000, 200000, 400000)], flush=True)
print('balanced kernel, 125,000 members x 100,000 steps, segments:')
print('  ', ['%d segments: %.2f ms' % (sg, run(125000, 100000, '1', sg)[0]) for sg in (8, 16, 32, 64, 128, 256)], flush=True)
print('start-up stagger per co-resident CTA (MAGPY_B200_K1_STAGGER, ns): plain one wave | balanced 125k, ms per 100,000 steps:')
for ns in (0, 500, 1500, 4000):
    os.environ['MAGPY_B200_K1_STAGGER'] = str(ns)
    print('   %5d ns: %.2f | %.2f' % (ns, run(113664, 100000, '0')[0], run(125000, 100000, '1', 128)[0]), flush=True)
os.environ['MAGPY_B200_K1_STAGGER'] = '0'
print('balanced kernel, 1M members x 100,000 steps, segments:')
print('  ', ['%d segments: %.2f ms' % (sg, run(1000000, 100000, '1', sg)[0]) for sg in (16, 64, 256)], flush=True)
