"""GPU parity of the discrete-orientation model: `magpy_b200_simulate_dom` through
magpy_b200.core / DOModel against the CPU oracle and the golden vectors of the compiled reference.

Bar: the kernel executes the reference's operation sequence without FMA contraction, so the only differences are the
last-bit differences of exp / pow / sin between CUDA and glibc; the adaptive step controller turns those into slightly
different step sequences, each within the integrator's own tolerance (`time_step`).  Measured agreement is ~1e-12; the
bar written here is 1e-9 absolute on mz = p_0 - p_1 (|mz| <= 1)."""
import os
import sys

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
TOL = 1e-9

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
import make_golden  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_dom.npz'))


@pytest.fixture(scope='module')
def orc():
    return ol.load_oracle()


@pytest.fixture(scope='module')
def mp():
    import magpy_b200
    return magpy_b200


@pytest.mark.parametrize('name', sorted(make_golden.DOM_CASES))
def test_domodel_matches_golden_reference_and_oracle(orc, mp, name):
    kw = dict(make_golden.DOM_CASES[name])
    model = mp.DOModel(kw['radius'], kw['anisotropy'], kw['p0'], kw['Ms'], kw['alpha'], kw['T'],
                       field_shape=kw.get('field_shape', 'constant'), field_frequency=kw.get('f', 0.0),
                       field_amplitude=kw.get('H0', 0.0), field_n_components=kw.get('n_components', 1))
    res = model.simulate(kw['t_end'], kw['dt'], kw['S'])
    assert np.array_equal(res.time, GOLD[name + '/time'])
    assert np.allclose(res.field, GOLD[name + '/field'], rtol=1e-13, atol=1e-9)
    mz = res.z[0]
    assert np.abs(mz - GOLD[name + '/mz']).max() <= TOL
    assert np.abs(mz - ol.dom_simulate(orc, **kw)[2]).max() <= TOL
    # the reference's result dict: one particle, x and y identically zero, magnetisation('z') is the unitless p0 - p1
    assert res.N == 1 and not res.x[0].any() and not res.y[0].any()
    assert np.array_equal(res.magnetisation('z'), mz)


def test_dom_batch_matches_oracle_item_by_item(orc, mp):
    """A size / anisotropy distribution in one call (one thread per item) against the oracle run item by item; ragged
    batch (not a multiple of the CTA size), per-item initial probabilities."""
    rng = np.random.default_rng(5)
    n = 150
    radius = rng.uniform(5e-9, 8e-9, n)
    K = rng.uniform(3e4, 5e4, n)
    p0 = rng.dirichlet([1, 1], n)
    model = mp.DOModel(6e-9, 4e4, [1, 0], 4e5, 0.1, 300.0, field_shape='sine', field_frequency=3e5, field_amplitude=1.5e4)
    out = model.simulate_batch(radius, K, 4e-6, 1e-10, 60, initial_probabilities=p0)
    assert out['mz'].shape == (n, 60) and out['field'].shape == (n, 60) and out['steps'].shape == (n,)
    worst = 0.0
    for i in range(n):
        t, fl, mz = ol.dom_simulate(orc, radius[i], K[i], p0[i], 4e5, 0.1, 300.0, 1e-10, 4e-6, 60, 'sine', 1.5e4, 3e5)
        worst = max(worst, np.abs(out['mz'][i] - mz).max())
        assert np.allclose(out['field'][i], fl, rtol=1e-13, atol=1e-9)
    assert worst <= TOL, worst
    assert np.all(out['steps'] >= 59)
    # idempotence / batching invariance: an item's result does not depend on its neighbours
    sub = model.simulate_batch(radius[40:47], K[40:47], 4e-6, 1e-10, 60, initial_probabilities=p0[40:47])
    assert np.array_equal(sub['mz'], out['mz'][40:47])


def test_dom_relaxation_time_is_the_neel_brown_rate(mp):
    """Size-independent property at full scale: in zero field mz(t) = mz(0) exp(-2 r t) with r the Neel-Brown rate of
    lib/dom.cpp:44-56 — for every item of a 100k-particle size distribution."""
    n = 100_000
    radius = np.linspace(5e-9, 7e-9, n)
    K, Ms, alpha, T = 4e4, 4e5, 0.1, 300.0
    model = mp.DOModel(6e-9, K, [1, 0], Ms, alpha, T)
    t_end = 2e-7
    out = model.simulate_batch(radius, K, t_end, 1e-10, 11)
    V = 4 / 3 * np.pi * radius ** 3
    sigma = K * V / mp.get_KB() / T
    taun = V * Ms * (1 + alpha ** 2) / 2 / mp.get_gamma() / alpha / mp.get_KB() / T
    rate = 1 / (taun * np.sqrt(np.pi) / sigma ** 1.5) * np.exp(-sigma)
    want = np.exp(-2 * rate[:, None] * out['time'][None, :])
    # the RK45 error is ~1e-10 per step; what is left is the driver's first-order hold between steps of up to
    # end_time / 1000: (2e-10)^2 / 8 * (2 r)^2 <= 1.6e-5 for the fastest item (2 r = 5.6e7 / s)
    assert np.abs(out['mz'] - want).max() < 2e-5


def test_dom_argument_errors(mp):
    model = mp.DOModel(6e-9, 4e4, [1, 0], 4e5, 0.1, 300.0, field_shape='triangle')
    with pytest.raises(KeyError):
        model.simulate(1e-6, 1e-10, 10)
    with pytest.raises(ValueError):
        mp.DOModel(6e-9, 4e4, [1, 0], 4e5, 0.1, 300.0).simulate(1e-6, 1e-10, 1)
