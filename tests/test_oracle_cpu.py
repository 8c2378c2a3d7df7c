"""Pins the CPU oracle (oracle/sllg_oracle.c): against the known answers the reference's own
unit tests hold (tests/golden/reference_kat.json, restated from test/tests.cpp), against golden
trajectories produced by the compiled reference (tests/golden/reference_trajectories.npz), and
— where oracle/_ref has been built — against the compiled reference live."""
import ctypes as C
import json
import os

import sys

import numpy as np
import pytest

import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, 'golden', 'reference_kat.json')))
GOLD = np.load(os.path.join(HERE, 'golden', 'reference_trajectories.npz'))
GOLD_CASES = sorted({k.split('/')[0] for k in GOLD.files})

P = ol._p
SDE = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_void_p)
SDEJ = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                   C.POINTER(C.c_double), C.c_double, C.c_double, C.c_void_p)


@pytest.fixture(scope='module')
def orc():
    return ol.load_oracle()


def arr(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def test_constants_digit_for_digit():
    src = open(os.path.join(os.path.dirname(HERE), 'oracle', 'sllg_oracle.c')).read()
    # include/constants.hpp:10-12
    for lit in ('1.38064852e-23', '1.25663706e-6', '1.76086e11'):
        assert lit in src


def test_rng_known_answer(orc):
    k = KAT['rng_mt_norm']
    got = ol.mt_normal(orc, k['seed'], 5, k['std'])
    assert np.allclose(got, k['expect'], rtol=4e-16, atol=0)      # EXPECT_DOUBLE_EQ = 4 ulp


def test_drift_diffusion_known_answers(orc):
    k = KAT['llg_drift']
    out = np.zeros(3)
    orc.orc_drift(P(out), P(arr(k['state'])), C.c_double(k['alpha']), P(arr(k['field'])))
    assert out.tolist() == k['expect']
    k = KAT['llg_diffusion']
    out = np.zeros(9)
    orc.orc_diffusion(P(out), P(arr(k['state'])), C.c_double(k['sr']), C.c_double(k['alpha']))
    assert out.tolist() == k['expect']


def test_drift_jacobian_known_answer(orc):
    k = KAT['llg_drift_jacobian']
    out = np.zeros(9)
    orc.orc_drift_jacobian(P(out), P(arr(k['state'])), C.c_double(k['alpha']), P(arr(k['heff'])), P(arr(k['heff_jac'])))
    assert np.allclose(out, k['expect'], rtol=1e-13, atol=1e-10)


def test_diffusion_jacobian_is_the_references_table(orc):
    """Finite differences of the diffusion matrix reproduce every entry of the table except the
    two the reference tabulates differently (lib/llg.cpp:130,155), which the oracle keeps."""
    m = arr([0.3, -0.5, 0.8]); sr, al = 0.7, 0.2
    T = np.zeros(27)
    orc.orc_diffusion_jacobian(P(T), P(m), C.c_double(sr), C.c_double(al))
    fd = np.zeros(27)
    for z in range(3):
        d = np.zeros(3); d[z] = 1e-6
        bp = np.zeros(9); bm = np.zeros(9)
        orc.orc_diffusion(P(bp), P(arr(m + d)), C.c_double(sr), C.c_double(al))
        orc.orc_diffusion(P(bm), P(arr(m - d)), C.c_double(sr), C.c_double(al))
        fd[z::3] = (bp - bm) / 2e-6
    bad = np.where(np.abs(T - fd) > 1e-8)[0].tolist()
    assert bad == [4, 25]
    assert T[4] == -al * sr * m[2] and T[25] == 2 * al * sr * m[2]


def test_field_waveforms(orc):
    f = orc.orc_field_value
    assert f(C.c_int(2), C.c_double(3.3), C.c_double(0.7), C.c_double(9.0)) == 0.7
    assert f(C.c_int(0), C.c_double(0.3), C.c_double(2.0), C.c_double(0.5)) == 2.0 * np.sin(2 * np.pi * 0.5 * 0.3)
    # lib/field.cpp:51-54: +h on even half periods, -h on odd ones
    for t, want in ((0.1, 1), (0.6, -1), (1.2, 1), (1.7, -1)):
        assert f(C.c_int(1), C.c_double(t), C.c_double(1.5), C.c_double(1.0)) == 1.5 * want


def test_dipolar_known_answers(orc):
    k = KAT['dipolar_two_particles']
    MU0 = 1.25663706e-6
    field = np.zeros(6)
    dists = arr([0, 0, 0, -0.5, 0, np.sqrt(3.) / 2., 0.5, 0, -np.sqrt(3.) / 2., 0, 0, 0])
    cubes = arr([0, 1, 1, 0])
    orc.orc_multi_add_dipolar(P(field), C.c_double(1. / np.sqrt(MU0)), C.c_double(k['k_av']), P(arr(k['v_red'])),
                              P(arr(k['mag'])), P(dists), P(cubes), C.c_int(2))
    assert np.allclose(field, k['expect'], rtol=1e-15, atol=1e-17)
    # prefactor (test/tests.cpp:798-803) through a unit pair
    field = np.zeros(6)
    orc.orc_multi_add_dipolar(P(field), C.c_double(1. / np.sqrt(MU0)), C.c_double(0.5), P(arr([1, 1])),
                              P(arr([0, 0, 0, 0, 0, 1])), P(arr([0, 0, 0, 1, 0, 0, -1, 0, 0, 0, 0, 0])),
                              P(arr([0, 1, 1, 0])), C.c_int(2))
    assert np.isclose(-field[2], KAT['dipolar_prefactor']['expect'], rtol=1e-15)
    # closed form of two identical aligned particles (test/tests.cpp:646-720)
    r, R, k_av, ms = 2.5e-9, 2.917e-9, 6.33e4, 400e3
    v = 4. / 3 * np.pi * r ** 3
    Rr = R / v ** (1. / 3)
    mag = arr([1, 0, 0, 1, 0, 0])
    field = np.zeros(6)
    orc.orc_multi_add_dipolar(P(field), C.c_double(ms), C.c_double(k_av), P(arr([1, 1])), P(mag),
                              P(arr([0, 0, 0, 0, 0, 1, 0, 0, -1, 0, 0, 0])), P(arr([0, Rr ** 3, Rr ** 3, 0])), C.c_int(2))
    Hk = 2 * k_av / MU0 / ms
    pref = v * ms / (4 * np.pi * R ** 3) / Hk
    assert np.allclose(field, [-pref, 0, 0, -pref, 0, 0], atol=1e-14)


def test_heun_step_known_answer(orc):
    k = KAT['heun_multiplicative']

    @SDE
    def sde(a, B, x, t, ctx):
        a[0] = x[0] * x[1] * t; a[1] = 3 * x[0]
        B[0] = t; B[1] = x[0]; B[2] = x[1]; B[3] = x[1]; B[4] = 4.0; B[5] = x[0] * x[1]
    dw = arr(k['dw']) / np.sqrt(k['dt'])
    nxt = np.zeros(2); work = np.zeros(2 * 2 + 2 * 6)
    orc.orc_heun_step(P(nxt), P(arr(k['x0'])), P(dw), sde, None, C.c_int(2), C.c_int(3), C.c_double(k['t']),
                      C.c_double(k['dt']), P(work))
    assert np.allclose(nxt, k['expect'], rtol=1e-15)


def test_heun_driver_ou_pathwise(orc):
    k = KAT['heun_driver_ou']
    th, mu, sg, n, dt = k['theta'], k['mu'], k['sigma'], k['n_steps'], k['dt']

    @SDE
    def sde(a, B, x, t, ctx):
        a[0] = th * (mu - x[0]); B[0] = sg
    w = ol.mt_normal(orc, k['seed'], n)
    states = np.zeros(n + 1)
    orc.orc_driver_heun(P(states), P(arr([k['x0']])), P(w), sde, None, C.c_size_t(n), C.c_int(1), C.c_int(1),
                        C.c_double(dt))
    true = k['x0']
    for i in range(1, n + 1):
        true = (true * np.exp(-th * dt) + mu * (1 - np.exp(-th * dt))
                + sg * np.sqrt((1 - np.exp(-2 * th * dt)) / (2 * th)) * w[i - 1])
        assert abs(true - states[i]) < k['tol']


def _stiff_sde(a_, b_):
    @SDEJ
    def sde(a, B, ad, Bd, x, ta, tb, ctx):
        a[0] = a_ * (x[1] - x[0]) - 0.5 * b_ * b_ * x[0]
        a[1] = a_ * (x[0] - x[1]) - 0.5 * b_ * b_ * x[1]
        ad[0] = -a_ - 0.5 * b_ * b_; ad[1] = a_; ad[2] = a_; ad[3] = -a_ - 0.5 * b_ * b_
        B[0] = b_ * x[0]; B[1] = b_ * x[1]
        Bd[0] = b_; Bd[1] = 0; Bd[2] = 0; Bd[3] = b_
    return sde


def test_implicit_midpoint_step_kloeden_platen(orc):
    k = KAT['implicit_midpoint_2d_stiff']
    sde = _stiff_sde(k['a'], k['b'])
    n, w = 2, 1
    work = np.zeros(w + n + n * w + n * n + n * w * n + 2 * n + n * n)
    ipiv = np.zeros(n, dtype=np.int32); x = np.zeros(n); it = C.c_long(0)
    rc = orc.orc_implicit_midpoint_step(P(x), P(arr(k['x0'])), P(arr(k['dw'])), sde, None, C.c_int(n), C.c_int(w),
                                        C.c_double(k['t']), C.c_double(k['dt']), C.c_double(k['eps']),
                                        C.c_long(k['max_iter']), P(work), P(ipiv), C.byref(it))
    assert rc == 0 and it.value >= 1
    assert np.allclose(x, k['expect'], atol=k['tol'])


def test_implicit_driver_stiff_2d(orc):
    k = KAT['implicit_driver_stiff_2d']
    a_, b_, N, dt = k['a'], k['b'], k['N'], k['dt']
    sde = _stiff_sde(a_, b_)
    dw = ol.mt_normal(orc, k['seed'], N)
    Wt = np.concatenate([[0], np.cumsum(dw[:-1])])
    x = np.zeros(2 * N)
    rc = orc.orc_driver_implicit(P(x), P(arr([1.0, 2.0])), P(dw), sde, None, C.c_size_t(N - 1), C.c_int(2), C.c_int(1),
                                 C.c_double(0.0), C.c_double(dt), C.c_double(k['eps']), C.c_long(k['max_iter']))
    assert rc == 0
    for i in range(1, N):
        rp = -0.5 * b_ * b_ * i * dt + b_ * Wt[i] * np.sqrt(dt)
        rm = (-2 * a_ - 0.5 * b_ * b_) * i * dt + b_ * Wt[i] * np.sqrt(dt)
        on, off = np.exp(rp) + np.exp(rm), np.exp(rp) - np.exp(rm)
        assert abs(0.5 * (on * 1.0 + off * 2.0) - x[2 * i]) < k['tol']
        assert abs(0.5 * (off * 1.0 + on * 2.0) - x[2 * i + 1]) < k['tol']


def test_dgesv_pivoting_and_singular(orc):
    A = arr([[1e-3, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 10.0]])
    b = arr([1.0, 2.0, 3.0])
    want = np.linalg.solve(A, b)
    ipiv = np.zeros(3, dtype=np.int32)
    Ac, bc = A.copy(), b.copy()
    assert orc.orc_dgesv(C.c_int(3), P(Ac), P(ipiv), P(bc)) == 0
    assert np.allclose(bc, want, rtol=1e-13)
    assert ipiv.tolist() == [3, 3, 3]      # row partial pivoting, 1-based like LAPACK
    S = arr([[1.0, 2.0], [2.0, 4.0]])
    assert orc.orc_dgesv(C.c_int(2), P(S), P(np.zeros(2, dtype=np.int32)), P(arr([1.0, 1.0]))) == 2   # test/tests.cpp:314-344


def test_task4_deterministic_llg_closed_form(orc):
    """test/convergence/task4.cpp:33-169: T=0, no anisotropy, constant field H along z, m0 = x:
    m(t) = (sech(aHt)cos(Ht), sech(aHt)sin(Ht), tanh(aHt)) (task4.cpp:94-99).  Heun over the
    oracle's drift reproduces it with a global error that scales as dt^2."""
    H, alpha = 1.0, 0.1
    h = arr([0, 0, H])

    @SDE
    def sde(a, B, x, t, ctx):
        orc.orc_drift(a, x, C.c_double(alpha), P(h))
        for i in range(9):
            B[i] = 0.0
    errs = []
    for dt in (0.1, 0.05):
        n = int(round(10.0 / dt))
        states = np.zeros((n + 1, 3))
        orc.orc_driver_heun(P(states), P(arr([1.0, 0, 0])), P(np.zeros(3 * n)), sde, None, C.c_size_t(n), C.c_int(3),
                            C.c_int(3), C.c_double(dt))
        t = np.arange(n + 1) * dt
        exact = np.stack([np.cos(H * t) / np.cosh(alpha * H * t), np.sin(H * t) / np.cosh(alpha * H * t),
                          np.tanh(alpha * H * t)], axis=1)
        errs.append(np.abs(states - exact).max())
    assert errs[0] < 2e-2 and 3.5 < errs[0] / errs[1] < 4.5


def test_deterministic_relaxation_through_full_dynamics(orc):
    """Through the SI entry point (reduced anisotropy is always 1): T=0, no applied field, axis z:
    tan(theta(t)) = tan(theta0) exp(-alpha t) in reduced time, so m_z(t) = (1 + tan^2(theta0) e^{-2 alpha t})^-1/2."""
    th0, errs = np.pi / 3, []
    for dt_red in (0.02, 0.01):
        c = ol.make_case(N=1, T=0.0, alpha=0.1, S=11, axis=[[0, 0, 1.0]], m0=[[np.sin(th0), 0, np.cos(th0)]])
        tf = ol.reduced_scalars(orc, c)['time_factor']
        c['dt'] = dt_red / tf
        c['t_end'] = 4.0 / tf
        t, fl, m, _, _ = ol.oracle_simulate(orc, c, seed=1)
        cum = ol.schedule(orc, dt_red, 4.0, c.S)
        ts = np.maximum(cum.astype(float) - 1, 0) * dt_red       # zero-order hold: state after cum[k]-1 steps
        exact = 1.0 / np.sqrt(1 + np.tan(th0) ** 2 * np.exp(-2 * 0.1 * ts))
        errs.append(np.abs(m[0, 2] / c.Ms - exact).max())
    assert errs[0] < 5e-5 and 3.5 < errs[0] / errs[1] < 4.5


@pytest.mark.parametrize('name', GOLD_CASES)
def test_oracle_matches_golden_reference_trajectories(orc, name):
    g = {k.split('/')[1]: GOLD[k] for k in GOLD.files if k.startswith(name + '/')}
    Ms, alpha, T, eps, dt, t_end, H0, f = g['scalars']
    N, S, renorm, inter, impl, shape, seed, n_steps = [int(v) for v in g['flags']]
    shape_name = {v: k for k, v in ol.FIELD.items()}[shape]
    c = ol.make_case(N=N, radius=g['radius'], anisotropy=g['anisotropy'], axis=g['axis'], m0=g['m0'],
                     location=g['location'], Ms=Ms, alpha=alpha, T=T, renorm=bool(renorm), interactions=bool(inter),
                     implicit=bool(impl), eps=eps, dt=dt, t_end=t_end, S=S, field_shape=shape_name, H0=H0, f=f)
    assert ol.steps_executed(orc, c) == n_steps
    # the restated mt19937_64 + polar stream is the reference's, bit for bit
    assert np.array_equal(ol.mt_normal(orc, seed, n_steps * 3 * N).reshape(n_steps, 3 * N), g['dw'])
    t, fl, m, _, fails = ol.oracle_simulate(orc, c, seed=seed)
    assert fails == 0
    assert np.array_equal(t, g['time'])
    assert np.allclose(fl, g['field'], rtol=1e-15, atol=0)
    tol = 1e-12 if impl or renorm else 1e-14      # dgesv/dnrm2 are third-party (OpenBLAS) in the reference
    assert np.abs(m - g['m']).max() / Ms <= tol
    # injecting the same stream through the RngArray-style hook gives the same path
    t2, fl2, m2, _, _ = ol.oracle_simulate(orc, c, seed=0, dW=g['dw'])
    assert np.array_equal(m, m2)


def test_oracle_matches_compiled_reference_live(orc):
    ref = ol.load_reference()
    if ref is None:
        pytest.skip('oracle/_ref not built (needs /root/reference)')
    rng = np.random.default_rng(11)
    for N, impl, shape, inter, renorm in [(1, 0, 'constant', 1, 0), (1, 1, 'sine', 1, 0), (2, 0, 'square', 1, 1),
                                          (4, 1, 'constant', 1, 0), (6, 0, 'sine', 1, 0), (3, 1, 'sine', 0, 1)]:
        c = ol.make_case(N=N, radius=7e-9 * (1 + 0.2 * rng.random(N)), anisotropy=1e5 * (1 + 0.2 * rng.random(N)),
                         dt=1e-13 if not impl else 1e-12, t_end=1e-10, S=37, implicit=bool(impl),
                         interactions=bool(inter), renorm=bool(renorm), field_shape=shape, H0=2e4, f=5e9, rng=rng)
        t1, f1, m1 = ol.reference_simulate(ref, c, 777)
        t2, f2, m2, _, _ = ol.oracle_simulate(orc, c, seed=777)
        assert np.array_equal(t1, t2) and np.allclose(f1, f2, rtol=1e-15, atol=0)
        assert np.abs(m1 - m2).max() / c.Ms <= (1e-12 if impl or renorm else 1e-14)
    a = np.zeros(1000); b = ol.mt_normal(orc, 31337, 1000)
    ref.ref_rng_normal(C.c_ulong(31337), C.c_double(1.0), C.c_size_t(1000), P(a))
    assert np.array_equal(a, b)


def test_dom_transition_matrix_known_answer(orc):
    k = KAT['dom_transition_matrix']
    W = ol.dom_transition_matrix(orc, k['k'], k['v'], k['T'], k['h'], k['ms'], k['alpha'])
    assert np.allclose(W, k['expect'], rtol=4e-15, atol=0)   # EXPECT_DOUBLE_EQ = 4 ulp


def test_philox_random123_known_answers(orc):
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert tuple(ol.philox(orc, ctr, key)) == want


def test_packed_gaussian_stream_layout_against_an_independent_restatement(orc):
    """The production noise mode cuts one Philox4x32-10 block into three (radius, angle) pairs (magpy_b200/csrc/rng.cuh:
    philox_gauss6_f32; restated in oracle/sllg_oracle.c, mode 2).  The kernel and the oracle are compared draw by draw on the
    GPU; here the oracle's restatement is checked against a third, independent one written from the documented layout —
    each 64-bit half {w1:w0}, {w3:w2} is [23-bit radius field | 18-bit angle | 23 bits], radius uniforms are the midpoints
    of a 22-bit grid, 2^18 equidistant directions — so that oracle and kernel cannot drift together unnoticed."""
    rng = np.random.default_rng(5)
    for _ in range(200):
        seed = int(rng.integers(0, 2**62)); member = int(rng.integers(0, 2**32)); particle = int(rng.integers(0, 200))
        step = int(rng.integers(1, 2**33))
        w = ol.philox(orc, [((step - 1) >> 1) & 0xffffffff, member, seed & 0xffffffff, seed >> 32],
                      [particle | (2 << 24), 0xB2005EED])
        h0 = (w[1] << 32) | w[0]; h1 = (w[3] << 32) | w[2]
        fields = [(w[0] & 0x7fffff, (h0 >> 23) & 0x3ffff), (w[1] >> 9, (h1 >> 23) & 0x3ffff), (w[2] & 0x7fffff, w[3] >> 14)]
        g = []
        for r23, a18 in fields:
            u = np.float32(1.0 - ((r23 >> 1) + 0.5) * 2.0 ** -22)                 # u = 2 - f: the midpoints of a 22-bit grid on (0, 1)
            r = np.sqrt(np.float32(-2.0) * np.log(u, dtype=np.float32), dtype=np.float32)
            a = np.float32(2 * np.pi * (1 + a18 * 2.0 ** -18))
            g += [float(r * np.cos(a, dtype=np.float32)), float(r * np.sin(a, dtype=np.float32))]
        want = g[3:] if (step - 1) & 1 else g[:3]
        got = ol.philox_gauss3(orc, seed, member, particle, step, 2)
        assert np.allclose(got, want, rtol=0, atol=3e-6), (got, want)
    # the largest radius the stream can produce: the last cell, u = 2^-23 -> sqrt(2 ln 2^23) = 5.65
    assert abs(np.sqrt(-2 * np.log(0.5 * 2.0 ** -22)) - 5.6468) < 1e-3


def test_schedule_semantics(orc):
    # zero-order hold: sample k stores the state after cum[k]-1 steps; finer sampling than stepping repeats
    cum = ol.schedule(orc, 0.3, 1.0, 11)
    t = 0.0
    lit = [0]
    step = 0
    for k in range(1, 11):
        while t <= k * (1.0 / 10):
            step += 1
            t = step * 0.3
        lit.append(step)
    assert cum.tolist() == lit
    assert np.all(np.diff(cum.astype(np.int64)) >= 0) and len(set(cum.tolist())) < 11


# ---- discrete-orientation model ---------------------------------------------------------
DOM_GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_dom.npz'))


def _dom_cases():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    import make_golden
    return make_golden.DOM_CASES


@pytest.mark.parametrize('name', sorted(_dom_cases()))
def test_dom_oracle_matches_golden_reference(orc, name):
    """The restated master equation + adaptive RK45 + first-order-hold driver against vectors produced by the
    compiled reference (simulation::dom_ensemble_dynamics, tests/golden/make_golden.py).  Same operation order and no
    FMA contraction on either side: the step sequences coincide and the samples agree to the last bits."""
    kw = _dom_cases()[name]
    t, fl, mz = ol.dom_simulate(orc, **kw)
    assert np.array_equal(t, DOM_GOLD[name + '/time'])
    assert np.allclose(fl, DOM_GOLD[name + '/field'], rtol=1e-15, atol=0)
    assert np.abs(mz - DOM_GOLD[name + '/mz']).max() <= 1e-14
    # physics: probabilities stay normalised -> |mz| <= 1; zero field relaxes towards 0 from p = (1, 0)
    assert np.abs(mz).max() <= 1 + 1e-12
    if name == 'dom_relax_6nm':
        assert mz[0] == 1.0 and 0 < mz[-1] < 0.05 and np.all(np.diff(mz) < 0)


def test_dom_oracle_matches_compiled_reference_live(orc):
    ref = ol.load_reference()
    if ref is None:
        pytest.skip('oracle/_ref not built (needs /root/reference)')
    rng = np.random.default_rng(11)
    for _ in range(6):
        kw = dict(radius=float(rng.uniform(5e-9, 8e-9)), anisotropy=float(rng.uniform(2e4, 6e4)),
                  p0=[float(x) for x in rng.dirichlet([1, 1])], Ms=4e5, alpha=float(rng.uniform(0.05, 0.5)), T=300.0,
                  dt=1e-10, t_end=float(rng.uniform(1e-6, 5e-6)), S=int(rng.integers(20, 120)),
                  field_shape=str(rng.choice(['constant', 'sine', 'square', 'square_f'])), H0=float(rng.uniform(0, 2.5e4)),
                  f=float(rng.uniform(1e5, 1e6)), n_components=int(rng.integers(1, 9)))
        a = ol.dom_simulate(orc, **kw)
        b = ol.dom_simulate(ref, reference=True, **kw)
        assert np.array_equal(a[0], b[0]) and np.allclose(a[1], b[1], rtol=1e-15, atol=0)
        assert np.abs(a[2] - b[2]).max() <= 1e-13


def test_load_results_reads_files_written_by_the_reference(tmp_path):
    """`magpy_b200.results.load_results` reads what the reference's own
    `simulation::save_results` (lib/simulation.cpp:38-63, include/io.hpp:12-27) writes — the committed golden files
    (tests/golden/reference_saved.*, generator make_golden_saved_results.py) and, when the compiled reference is
    present, a file it writes live; and `save_results` writes byte-identical files back."""
    import sys
    from magpy_b200.results import load_results, save_results
    golden = os.path.join(HERE, 'golden')
    sys.path.insert(0, golden)
    import make_golden_saved_results as mg
    orc = ol.load_oracle()
    c = ol.make_case(**mg.CASE)
    t, fl, m, _, _ = ol.oracle_simulate(orc, c, seed=mg.SEED)
    prefixes = [os.path.join(golden, 'reference_saved')]
    ref = ol.load_reference()
    if ref is not None:
        ref.ref_simulate_and_save.restype = C.c_int
        mg.write(str(tmp_path / 'live'), ref)
        prefixes.append(str(tmp_path / 'live'))
    for prefix in prefixes:
        res = load_results(prefix)
        assert res.N == 1 and len(res.time) == c.S
        # bit-for-bit the trajectory of particle 1 of the reference run (the oracle is pinned to it bitwise)
        assert np.array_equal(res.time, t) and np.array_equal(res.field, fl)
        assert np.array_equal(res.x[0], m[mg.PARTICLE, 0]) and np.array_equal(res.y[0], m[mg.PARTICLE, 1])
        assert np.array_equal(res.z[0], m[mg.PARTICLE, 2])
        save_results(str(tmp_path / 'again'), res)
        for suffix in ('mx', 'my', 'mz', 'field', 'time'):
            assert open('%s.%s' % (prefix, suffix), 'rb').read() == open(str(tmp_path / 'again') + '.' + suffix, 'rb').read()
