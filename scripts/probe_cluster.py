"""GPU probe: throughput of the cluster kernels (Heun and implicit midpoint) over N."""
import sys
import numpy as np
sys.path.insert(0, '.')
import magpy_b200.core as core


def geometry(N, spacing=3e-8, seed=0):
    """Random points in a ball with a minimum separation (BASELINE config 4 shape)."""
    rng = np.random.default_rng(seed)
    pts = []
    radius = spacing * (N ** (1 / 3)) * 0.9 + spacing
    while len(pts) < N:
        p = rng.uniform(-radius, radius, 3)
        if np.linalg.norm(p) > radius:
            continue
        if all(np.linalg.norm(p - q) >= spacing for q in pts):
            pts.append(p)
    return np.array(pts)


def run(N, R, steps, implicit=False, gauss='f32p', dt=1e-12, field='sine', S=11, renorm=False):
    rng = np.random.default_rng(1)
    radius = np.full(N, 12e-9); K = np.full(N, 4e4)
    axis = rng.normal(size=(N, 3)); axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    m0 = np.tile([0, 0, 1.0], (N, 1))
    loc = geometry(N) if N > 1 else np.zeros((1, 3))
    seeds = rng.integers(0, 2**31 - 1, R)
    plan = core.EnsemblePlan(radius, K, axis, m0, loc, 4e5, 0.1, 300.0, renorm, True, implicit, dt, dt * steps, S, seeds,
                             field_shape=field, field_amplitude=2e4, field_frequency=3e5, gauss=gauss,
                             return_trajectories=False)
    for _ in range(2):
        plan.run(); st = plan.sync()
    ps = st['particle_steps'] / (st['integrate_ms'] * 1e-3)
    walg = 98 + 36 * (N - 1)
    it = st['newton_iterations'] / max(1, st['particle_steps'] / N)
    print(f'N={N:3d} R={R:7d} steps={steps:6d} {"imid" if implicit else "heun"}{" renorm" if renorm else ""} {gauss}: {st["integrate_ms"]:9.2f} ms  '
          f'{ps:.3e} particle-steps/s  {ps / N:.3e} cluster-steps/s' +
          (f'  {ps * walg / 1e12:6.2f} TFLOP/s by W_alg' if not implicit else f'  {it:.2f} Newton it/step'), flush=True)


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if which in ('all', 'heun'):
        run(1, 1 << 20, 10000)
        run(2, 1 << 19, 10000)
        run(3, 1 << 18, 10000)
        run(4, 1 << 18, 10000)
        run(5, 1 << 17, 4000)
        run(8, 1 << 17, 4000)
        run(16, 1 << 16, 2000)
        run(32, 1 << 15, 1000)
        run(64, 12500, 1000)
        run(64, 148 * 32 * 2, 1000)
        run(128, 148 * 32, 400)
    if which == 'published':
        # the reference's best published figure: 10,000 x 2 particles, Heun, dipolar on, renorm,
        # 10,000 steps of 1e-13 s: 6.7e6 particle-steps/s on 8 processes
        run(2, 10000, 10000, dt=1e-13, field='constant', renorm=True)
        run(2, 1 << 19, 10000, dt=1e-13, field='constant', renorm=True)
        run(2, 1 << 19, 10000, dt=1e-13, field='constant', renorm=False)
    if which in ('all', 'imid'):
        run(1, 1 << 19, 1000, implicit=True)
        run(2, 1 << 18, 1000, implicit=True)
        run(3, 1 << 17, 1000, implicit=True)
        run(4, 1 << 17, 1000, implicit=True)
        run(5, 1 << 16, 500, implicit=True)
        run(8, 1 << 16, 500, implicit=True)
        run(16, 1 << 14, 300, implicit=True)
