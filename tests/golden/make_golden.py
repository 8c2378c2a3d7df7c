"""Generate golden trajectories from the COMPILED REFERENCE (oracle/_ref/libmagpy_ref.so, built
from /root/reference by oracle/ref_build/Makefile).  Run in the build container:

    make -C oracle/ref_build && python tests/golden/make_golden.py

Writes tests/golden/reference_trajectories.npz: for each case the SI inputs, the seed, the
reference's noise stream (`RngMtNorm(seed, 1.0)`, lib/rng.cpp:14-24) and the trajectories
`simulation::full_dynamics` returned (lib/simulation.cpp:476-624).  The .npz travels to the GPU
box; /root/reference does not.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

CASES = {
    # BASELINE config 1 (shortened to 5000 steps; the full 1e5-step run is compared live against the oracle)
    'c1_heun_single': dict(N=1, dt=1e-14, t_end=5e-11, S=101, axis=[[0, 0, 1.0]], m0=[[1.0, 0, 0]], seed=1001),
    'heun_single_sine': dict(N=1, dt=1e-13, t_end=2e-10, S=81, field_shape='sine', H0=2e4, f=3e9, seed=7,
                             axis=[[0.6, 0, 0.8]], m0=[[0, 0, 1.0]]),
    'heun_single_square_renorm': dict(N=1, dt=1e-13, t_end=2e-10, S=50, field_shape='square', H0=3e4, f=8e9,
                                      renorm=True, seed=99),
    'imid_single': dict(N=1, dt=1e-12, t_end=3e-10, S=61, implicit=True, seed=1001, axis=[[0, 0, 1.0]],
                        m0=[[1.0, 0, 0]]),
    'imid_single_sine_loose': dict(N=1, dt=1e-12, t_end=2e-10, S=41, implicit=True, eps=1e-5, field_shape='sine',
                                   H0=2e4, f=2e9, seed=5),
    # BASELINE config 2 geometry: two 7 nm particles 9 nm apart along z (two-particle-equilibrium notebook)
    'c2_imid_dimer': dict(N=2, radius=7e-9, anisotropy=1e5, T=330.0, dt=1e-12, t_end=2e-10, S=51, implicit=True,
                          axis=[[0, 0, 1.0], [0, 0, 1.0]], m0=[[0, 0, 1.0], [0, 0, 1.0]],
                          location=[[0, 0, 0], [0, 0, 9e-9]], seed=1234),
    'heun_dimer_renorm': dict(N=2, radius=7e-9, anisotropy=1e5, T=330.0, dt=1e-13, t_end=1e-10, S=41, renorm=True,
                              location=[[0, 0, 0], [0, 0, 9e-9]], seed=4321),
    'heun_cluster5_mixed': dict(N=5, radius=[6e-9, 7e-9, 8e-9, 7e-9, 6.5e-9],
                                anisotropy=[1e5, 0.9e5, 1.1e5, 1e5, 1.2e5], T=310.0, dt=1e-13, t_end=5e-11, S=26,
                                field_shape='sine', H0=1e4, f=1e10, seed=31),
    'imid_cluster3_nointeract': dict(N=3, radius=7e-9, anisotropy=1e5, T=330.0, dt=1e-12, t_end=1e-10, S=21,
                                     implicit=True, interactions=False, seed=8),
    'imid_cluster3': dict(N=3, radius=[7e-9, 6e-9, 8e-9], anisotropy=[1e5, 1.2e5, 0.8e5], T=330.0, dt=1e-12,
                          t_end=1e-10, S=21, implicit=True, seed=9),
}


# Discrete-orientation model (simulation::dom_ensemble_dynamics through ref_dom_simulate), SI arguments of
# magpy.DOModel.  The 6 nm particle (sigma = 8.7) relaxes / switches on these time scales; 12 nm is blocked.
DOM_CASES = {
    'dom_relax_6nm': dict(radius=6e-9, anisotropy=4e4, p0=[1.0, 0.0], Ms=4e5, alpha=0.1, T=300.0, dt=1e-10, t_end=2e-6, S=200),
    'dom_sine_6nm': dict(radius=6e-9, anisotropy=4e4, p0=[0.5, 0.5], Ms=4e5, alpha=0.1, T=300.0, dt=1e-10, t_end=1e-5, S=300,
                         field_shape='sine', H0=2e4, f=3e5),
    'dom_square_7nm': dict(radius=7e-9, anisotropy=3e4, p0=[0.9, 0.1], Ms=4.46e5, alpha=0.05, T=310.0, dt=1e-9, t_end=2e-5,
                           S=150, field_shape='square', H0=1.5e4, f=1e5),
    'dom_squaref_6nm': dict(radius=6e-9, anisotropy=4e4, p0=[1.0, 0.0], Ms=4e5, alpha=0.1, T=300.0, dt=1e-10, t_end=1e-5,
                            S=250, field_shape='square_f', H0=2e4, f=2e5, n_components=7),
    'dom_blocked_12nm': dict(radius=12e-9, anisotropy=4e4, p0=[1.0, 0.0], Ms=4e5, alpha=0.1, T=300.0, dt=1e-9, t_end=1e-5,
                             S=50, field_shape='sine', H0=2e4, f=3e5),
}


def main_dom(ref):
    out = {}
    for name, kw in DOM_CASES.items():
        t, fl, mz = ol.dom_simulate(ref, reference=True, **kw)
        out[name + '/time'] = t
        out[name + '/field'] = fl
        out[name + '/mz'] = mz
        print('%-20s mz[-1] = %.12g' % (name, mz[-1]))
    np.savez_compressed(os.path.join(HERE, 'reference_dom.npz'), **out)


def main():
    ref = ol.load_reference()
    if ref is None:
        raise SystemExit('oracle/_ref/libmagpy_ref.so missing: run `make -C oracle/ref_build` first')
    orc = ol.load_oracle()
    out = {}
    for name, kw in CASES.items():
        kw = dict(kw)
        seed = kw.pop('seed')
        c = ol.make_case(rng=np.random.default_rng(len(name)), **kw)
        t, fl, m = ol.reference_simulate(ref, c, seed)
        n_steps = ol.steps_executed(orc, c)
        dw = np.zeros(n_steps * 3 * c.N)
        import ctypes as C
        ref.ref_rng_normal(C.c_ulong(seed), C.c_double(1.0), C.c_size_t(dw.size), dw.ctypes.data_as(C.c_void_p))
        for k in ('radius', 'anisotropy', 'axis', 'm0', 'location'):
            out[name + '/' + k] = c[k]
        out[name + '/scalars'] = np.array([c.Ms, c.alpha, c.T, c.eps, c.dt, c.t_end, c.H0, c.f])
        out[name + '/flags'] = np.array([c.N, c.S, int(c.renorm), int(c.interactions), int(c.implicit),
                                         ol.FIELD[c.field_shape], seed, n_steps], dtype=np.int64)
        out[name + '/time'] = t
        out[name + '/field'] = fl
        out[name + '/m'] = m
        out[name + '/dw'] = dw.reshape(n_steps, 3 * c.N).astype(np.float64)
        print('%-28s N=%d steps=%d  m[...,-1]=%s' % (name, c.N, n_steps, m[0, :, -1]))
    np.savez_compressed(os.path.join(HERE, 'reference_trajectories.npz'), **out)
    main_dom(ref)


if __name__ == '__main__':
    main()
