"""Per-task timeline of the persistent K1 kernel (MAGPY_B200_K1_TRACE): task durations and SM idle time, 125,000 members
against 1,000,000."""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get('MB_ROOT', '.'))
os.environ['MAGPY_B200_K1_TRACE'] = '/tmp/k1_trace.bin'
os.environ['MAGPY_B200_K1_BALANCE'] = '1'
os.environ['MAGPY_B200_K1_MIN_BLOCKS'] = '1'
import magpy_b200.core as core

for R, stagger in ((125000, 0), (250000, 0), (1000000, 0)):
    seeds = np.random.default_rng(0).integers(0, 2**31 - 1, R)
    plan = core.EnsemblePlan([12e-9], [4e4], [[0, 0, 1.0]], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, False, True, False,
                             1e-12, 1e-7, 101, seeds, field_shape='sine', field_amplitude=2e4, field_frequency=3e5,
                             gauss='f32p', return_trajectories=False)
    for i in range(2):
        plan.run(); st = plan.sync()
    tr = np.fromfile('/tmp/k1_trace.bin', dtype=np.uint64).reshape(-1, 4)
    tr = tr[tr[:, 1] > 0]
    t0 = tr[:, 0].min()
    beg, end, sm = (tr[:, 0] - t0) * 1e-6, (tr[:, 1] - t0) * 1e-6, tr[:, 2] & np.uint64(0xFFFFFFFF)
    cta = (tr[:, 2] >> np.uint64(32)).astype(int)
    dur = end - beg
    steps = (tr[:, 3] & np.uint64(0xFFFFFFFF)).astype(float)
    print('   slice length (steps): p5 %d  p50 %d  p95 %d; steps per CTA-ms: fast quartile %.0f, slow quartile %.0f' % (
        *np.percentile(steps, [5, 50, 95]), np.percentile(steps / dur, 75), np.percentile(steps / dur, 25)))
    print('R=%d stagger=%d ns: integrate %.2f ms, %d tasks, span %.2f ms; task duration ms: mean %.4f  p5 %.4f  p50 %.4f  p95 %.4f  max %.4f' % (
        R, stagger, st['integrate_ms'], len(tr), end.max(), dur.mean(), *np.percentile(dur, [5, 50, 95]), dur.max()))
    # busy CTA-slots over time: sum of durations / (span * 888)
    print('   slot utilisation %.3f (sum of task time / (888 slots x span)); per-SM tasks min %d max %d' % (
        dur.sum() / (end.max() * 888), np.bincount(sm.astype(int)).min(), np.bincount(sm.astype(int)).max()))
    # duration by quarter of the run
    q = np.minimum((beg / end.max() * 4).astype(int), 3)
    print('   mean task duration by quarter of the run:', ['%.4f' % dur[q == i].mean() for i in range(4)])
    # is slowness a property of the physical CTA (its slot on the SM)?
    per_cta = np.array([dur[cta == c].mean() for c in range(cta.max() + 1)])
    n_cta = np.bincount(cta)
    order = np.argsort(per_cta)
    print('   per physical CTA: mean task duration min %.3f  p25 %.3f  p50 %.3f  p75 %.3f  p90 %.3f  max %.3f ms; tasks per CTA min %d max %d' % (
        per_cta.min(), *np.percentile(per_cta, [25, 50, 75, 90]), per_cta.max(), n_cta.min(), n_cta.max()))
    slow = per_cta > 2 * np.median(per_cta)
    print('   CTAs with mean duration > 2x median: %d of %d; their blockIdx // 148 histogram: %s; slow-CTA share of all task time %.3f' % (
        slow.sum(), len(per_cta), np.bincount(np.nonzero(slow)[0] // 148, minlength=6).tolist(), dur[slow[cta]].sum() / dur.sum()))
    first_sm = {}
    for c in range(len(per_cta)):
        first_sm.setdefault(int(sm[cta == c][0]), []).append(per_cta[c])
    spread = np.array([max(v) / min(v) for v in first_sm.values()])
    print('   per SM: slowest / fastest CTA mean duration: median %.2f  max %.2f' % (np.median(spread), spread.max()))
    # how many tasks run concurrently on an SM, sampled
    ts = np.linspace(0.05, 0.95, 10) * end.max()
    conc = [np.sum((beg <= t) & (end > t)) for t in ts]
    print('   tasks in flight at 10 instants:', conc, flush=True)
