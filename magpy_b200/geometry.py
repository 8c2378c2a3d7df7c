"""Host-side input generation for the cluster configurations.

The reference offers straight chains (`magpy/geometry/coordinates.py:70-102`), Arkus clusters of at most
eight particles (`magpy/geometry/arkus.py`) and random unit vectors (`magpy/initial_conditions.py:4-34`);
nothing there yields the "64-particle random-geometry cluster" of BASELINE config 4, so
:func:`random_cluster_coordinates` is added next to same-named equivalents of the two helpers that config
needs.  Everything here is plain numpy on the host: geometry is an input of the integration path, not part
of it.  The reference draws from numpy's global legacy generator (`np.random.seed`); these functions accept
a seed or a `numpy.random.Generator` instead and fall back to the global generator when given none, so
seeded reference scripts keep producing the same numbers for `uniform_random_axes`.
"""
import numpy as np


def _uniform(rng, n=None):
    return np.random.rand(*(() if n is None else (n,))) if rng is None else rng.random(n)


def _generator(seed_or_rng):
    if seed_or_rng is None or isinstance(seed_or_rng, np.random.Generator):
        return seed_or_rng
    return np.random.default_rng(seed_or_rng)


def random_point_on_unit_sphere(rng=None):
    """Uniformly distributed point on the unit sphere (magpy/initial_conditions.py:4-16: same two
    draws, same formula)."""
    rng = _generator(rng)
    theta = 2.0 * np.pi * _uniform(rng)
    phi = np.arccos(1 - 2.0 * _uniform(rng))
    return np.array([np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi)])


def uniform_random_axes(N, rng=None):
    """`N` unit vectors uniform on the sphere, shape (N, 3) (magpy/initial_conditions.py:19-34)."""
    rng = _generator(rng)
    return np.array([random_point_on_unit_sphere(rng) for _ in range(N)])


def chain_coordinates(n_particles, R, direction=(0, 0, 1)):
    """Particles on a straight, regularly spaced chain starting at the origin
    (magpy/geometry/coordinates.py:70-102): only the direction of `direction` matters."""
    d = np.asarray(direction, dtype=np.float64)
    return R * np.arange(n_particles)[:, None] * (d / np.linalg.norm(d))[None, :]


def random_cluster_coordinates(n_particles, min_separation, packing=0.3, rng=None, max_tries=200000):
    """Random non-overlapping cluster: `n_particles` points uniform in a ball, no two closer than
    `min_separation` (use twice the particle radius plus any coating).

    The ball radius is chosen so that spheres of diameter `min_separation` fill the fraction `packing` of
    its volume (random sequential addition jams near 0.38, so keep `packing` <= 0.35).  Returns an
    (n_particles, 3) array centred on the centroid.
    """
    if not 0 < packing <= 0.35:
        raise ValueError('packing must be in (0, 0.35]')
    rng = _generator(rng)
    if rng is None:
        rng = np.random.default_rng(np.random.randint(2 ** 31 - 1))
    ball = 0.5 * min_separation * (n_particles / packing) ** (1.0 / 3.0)
    pts = np.empty((n_particles, 3))
    n = 0
    for _ in range(max_tries):
        p = rng.uniform(-ball, ball, 3)
        if p @ p > ball * ball:
            continue
        if n and np.min(np.sum((pts[:n] - p) ** 2, axis=1)) < min_separation ** 2:
            continue
        pts[n] = p
        n += 1
        if n == n_particles:
            return pts - pts.mean(axis=0)
    raise RuntimeError('could not place %d particles; lower `packing`' % n_particles)
