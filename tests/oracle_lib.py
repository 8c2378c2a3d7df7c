"""ctypes bindings of the CPU oracle (oracle/libsllg_oracle.so) and, when it has been built,
of the compiled reference (oracle/_ref/libmagpy_ref.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, 'oracle', 'libsllg_oracle.so')
REF_SO = os.path.join(ROOT, 'oracle', '_ref', 'libmagpy_ref.so')

KB = 1.38064852e-23   # include/constants.hpp:10
FIELD = {'sine': 0, 'square': 1, 'constant': 2}


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def load_oracle():
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle')])
    lib = C.CDLL(ORACLE_SO)
    lib.orc_simulate.restype = C.c_long
    lib.orc_field_value.restype = C.c_double
    lib.orc_nrm2.restype = C.c_double
    lib.orc_steps_executed.restype = C.c_uint64
    lib.orc_dgesv.restype = C.c_int
    lib.orc_implicit_midpoint_step.restype = C.c_int
    lib.orc_driver_implicit.restype = C.c_int
    lib.orc_dom_simulate.restype = C.c_uint64
    return lib


def load_reference():
    """The compiled reference, or None when oracle/_ref has not been built (e.g. no /root/reference)."""
    if not os.path.exists(REF_SO):
        return None
    try:
        lib = C.CDLL(REF_SO)
    except OSError:
        return None
    lib.ref_simulate.restype = C.c_int
    lib.ref_ensemble.restype = C.c_double
    lib.ref_field_sinusoidal.restype = C.c_double
    lib.ref_field_square.restype = C.c_double
    return lib


class Case(dict):
    """Keyword bundle describing one cluster + run, in the reference's SI arguments."""
    def __getattr__(self, k):
        return self[k]


def make_case(N=1, radius=12e-9, anisotropy=4e4, axis=None, m0=None, location=None, Ms=4e5, alpha=0.1, T=300.0,
              renorm=False, interactions=True, implicit=False, eps=1e-9, dt=1e-14, t_end=1e-10, S=50,
              field_shape='constant', H0=0.0, f=0.0, rng=None):
    rng = rng or np.random.default_rng(0)

    def unit(v):
        return v / np.linalg.norm(v, axis=-1, keepdims=True)
    radius = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, dtype=np.float64), (N,)))
    anisotropy = np.ascontiguousarray(np.broadcast_to(np.asarray(anisotropy, dtype=np.float64), (N,)))
    axis = np.ascontiguousarray(unit(rng.normal(size=(N, 3))) if axis is None else np.asarray(axis, dtype=np.float64).reshape(N, 3))
    m0 = np.ascontiguousarray(unit(rng.normal(size=(N, 3))) if m0 is None else np.asarray(m0, dtype=np.float64).reshape(N, 3))
    if location is None:
        location = np.cumsum(np.abs(rng.normal(size=(N, 3))) * 1e-8 + 2.0 * radius.max(), axis=0)
    location = np.ascontiguousarray(np.asarray(location, dtype=np.float64).reshape(N, 3))
    return Case(N=N, radius=radius, anisotropy=anisotropy, axis=axis, m0=m0, location=location, Ms=Ms, alpha=alpha,
                T=T, renorm=renorm, interactions=interactions, implicit=implicit, eps=eps, dt=dt, t_end=t_end, S=S,
                field_shape=field_shape, H0=H0, f=f)


def oracle_simulate(lib, c, seed=1001, dW=None, axis=None, m0=None):
    """Run the oracle; dW (steps, 3N) injects the noise, else the reference MT stream of `seed`.
    Returns time[S], field[S], m[N,3,S] (A/m), (newton total, max, last-step count), failures."""
    N, S = c.N, c.S
    t = np.zeros(S); fl = np.zeros(S); m = np.zeros((N, 3, S)); it = np.zeros(3, dtype=np.int64)
    axis = c.axis if axis is None else np.ascontiguousarray(axis, dtype=np.float64)
    m0 = c.m0 if m0 is None else np.ascontiguousarray(m0, dtype=np.float64)
    if dW is not None:
        dW = np.ascontiguousarray(dW, dtype=np.float64)
        dwp, dwl = _p(dW), dW.size
    else:
        dwp, dwl = None, 0
    fails = lib.orc_simulate(
        C.c_int(N), _p(c.radius), _p(c.anisotropy), _p(axis), _p(m0), _p(c.location), C.c_double(c.Ms),
        C.c_double(c.alpha), C.c_double(c.T), C.c_int(int(c.renorm)), C.c_int(int(c.interactions)),
        C.c_int(int(c.implicit)), C.c_double(c.eps), C.c_double(c.dt), C.c_double(c.t_end), C.c_size_t(S),
        C.c_uint64(seed), C.c_int(FIELD[c.field_shape]), C.c_double(c.H0), C.c_double(c.f), dwp, C.c_size_t(dwl),
        _p(t), _p(fl), _p(m), _p(it))
    return t, fl, m, (int(it[0]), int(it[1]), int(it[2])), int(fails)


def reference_simulate(lib, c, seed=1001):
    N, S = c.N, c.S
    t = np.zeros(S); fl = np.zeros(S); m = np.zeros((N, 3, S))
    rc = lib.ref_simulate(
        C.c_size_t(N), _p(c.radius), _p(c.anisotropy), _p(c.axis), _p(c.m0), _p(c.location), C.c_double(c.Ms),
        C.c_double(c.alpha), C.c_double(c.T), C.c_int(int(c.renorm)), C.c_int(int(c.interactions)),
        C.c_int(int(c.implicit)), C.c_double(c.eps), C.c_double(c.dt), C.c_double(c.t_end), C.c_size_t(S),
        C.c_long(seed), C.c_int(FIELD[c.field_shape]), C.c_double(c.H0), C.c_double(c.f), _p(t), _p(fl), _p(m))
    assert rc == 0
    return t, fl, m


def mt_normal(lib, seed, n, std=1.0):
    out = np.zeros(n)
    lib.orc_rng_normal(C.c_uint64(seed), C.c_double(std), C.c_size_t(n), _p(out))
    return out


def reduced_scalars(lib, c):
    N = c.N
    k = np.zeros(N); v = np.zeros(N); s = np.zeros(N); ru = np.zeros(N * N * 3); rc = np.zeros(N * N); sc = np.zeros(9)
    lib.orc_reduce_units(C.c_int(N), _p(c.radius), _p(c.anisotropy), _p(c.location), C.c_double(c.Ms),
                         C.c_double(c.alpha), C.c_double(c.T), C.c_double(c.dt), C.c_double(c.t_end),
                         C.c_double(c.H0), C.c_double(c.f), _p(k), _p(v), _p(s), _p(ru), _p(rc), _p(sc))
    names = ('V_av', 'K_av', 'H_k', 'time_factor', 'dt_red', 'T_red', 'h0', 'f_red', 'dipolar_prefactor')
    out = dict(zip(names, sc.tolist()))
    out.update(k_red=k, v_red=v, sigma=s, runit=ru.reshape(N, N, 3), rcube=rc.reshape(N, N))
    return out


def steps_executed(lib, c):
    sc = reduced_scalars(lib, c)
    return int(lib.orc_steps_executed(C.c_double(sc['dt_red']), C.c_double(sc['T_red']), C.c_size_t(c.S)))


def schedule(lib, dt_red, T_red, S):
    cum = np.zeros(S, dtype=np.uint64)
    lib.orc_schedule(C.c_double(dt_red), C.c_double(T_red), C.c_size_t(S), _p(cum))
    return cum


def philox(lib, ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib.orc_philox4x32_10(c, k, o)
    return list(o)


def philox_gauss3(lib, seed, member, particle, step, mode):
    o = (C.c_double * 3)()
    lib.orc_philox_gauss3(C.c_uint64(seed), C.c_uint32(member), C.c_uint32(particle), C.c_uint64(step),
                          C.c_int(mode), o)
    return list(o)


def oracle_ensemble(lib, c, seeds, axis=None, m0=None):
    """OpenMP ensemble of oracle runs (reference MT noise per seed): sums [S,4] and final [R,N,3] in A/m."""
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    R = len(seeds)
    sums = np.zeros((c.S, 4)); final = np.zeros((R, c.N, 3))
    axis = c.axis if axis is None else np.ascontiguousarray(axis, dtype=np.float64)
    m0 = c.m0 if m0 is None else np.ascontiguousarray(m0, dtype=np.float64)
    lib.orc_ensemble.restype = C.c_double
    lib.orc_ensemble(C.c_size_t(R), _p(seeds), C.c_int(c.N), _p(c.radius), _p(c.anisotropy), _p(axis),
                     C.c_size_t(3 * c.N if axis.ndim == 3 else 0), _p(m0), C.c_size_t(3 * c.N if m0.ndim == 3 else 0),
                     _p(c.location), C.c_double(c.Ms), C.c_double(c.alpha), C.c_double(c.T), C.c_int(int(c.renorm)),
                     C.c_int(int(c.interactions)), C.c_int(int(c.implicit)), C.c_double(c.eps), C.c_double(c.dt),
                     C.c_double(c.t_end), C.c_size_t(c.S), C.c_int(FIELD[c.field_shape]), C.c_double(c.H0),
                     C.c_double(c.f), _p(sums), _p(final))
    return sums, final


def dom_transition_matrix(lib, k, v, T, h, ms, alpha):
    """2x2 Neel-Brown transition matrix (lib/dom.cpp:33-59), row-major [W00, W01, W10, W11]."""
    W = (C.c_double * 4)()
    lib.orc_dom_transition_matrix(W, C.c_double(k), C.c_double(v), C.c_double(T), C.c_double(h), C.c_double(ms),
                                  C.c_double(alpha))
    return list(W)


DOM_FIELD = {'sine': 0, 'square': 1, 'constant': 2, 'square_f': 3}


def dom_simulate(lib, radius, anisotropy, p0, Ms, alpha, T, dt, t_end, S, field_shape='constant', H0=0.0, f=0.0,
                 n_components=1, reference=False):
    """Discrete-orientation model of one particle: the oracle's restatement (`orc_dom_simulate`) or, with
    reference=True, the compiled reference (`ref_dom_simulate` -> simulation::dom_ensemble_dynamics), with the SI ->
    reduced conversion of magpy/core.pyx:224-225.  Returns (time, field [A/m], mz = p0 - p1)."""
    V = 4. / 3 * np.pi * radius ** 3
    H_k = 2.0 * anisotropy / 1.25663706e-6 / Ms
    t = np.zeros(S); fl = np.zeros(S); mz = np.zeros(S)
    p0 = np.ascontiguousarray(p0, dtype=np.float64)
    fn = lib.ref_dom_simulate if reference else lib.orc_dom_simulate
    fn(C.c_double(V), C.c_double(anisotropy), C.c_double(T), C.c_double(Ms), C.c_double(alpha),
       C.c_int(DOM_FIELD[field_shape]), C.c_double(H0 / H_k), C.c_double(f), C.c_size_t(n_components), _p(p0),
       C.c_double(dt), C.c_double(t_end), C.c_size_t(S), _p(t), _p(fl), _p(mz))
    return t, fl * H_k, mz
