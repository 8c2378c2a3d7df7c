"""Write tests/golden/reference_saved.{mx,my,mz,field,time}: files produced by the REFERENCE's own
`simulation::save_results` (lib/simulation.cpp:38-63) after `simulation::full_dynamics`, through the compiled
reference (oracle/_ref/libmagpy_ref.so; `make -C oracle/ref_build`).  The persistence test reads them with
magpy_b200.results.load_results.  Run in the build container (needs /root/reference to have been compiled)."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

CASE = dict(N=2, radius=7e-9, anisotropy=1e5, dt=1e-13, t_end=5e-11, S=12, implicit=False, T=330.0, field_shape='sine',
            H0=1e4, f=1e10, location=[[0, 0, 0], [0, 0, 9e-9]])
SEED, PARTICLE = 4242, 1


def write(prefix, lib=None):
    lib = lib or ol.load_reference()
    c = ol.make_case(**CASE)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.ref_simulate_and_save(
        C.c_size_t(c.N), P(c.radius), P(c.anisotropy), P(c.axis), P(c.m0), P(c.location), C.c_double(c.Ms),
        C.c_double(c.alpha), C.c_double(c.T), C.c_int(0), C.c_int(1), C.c_int(0), C.c_double(c.eps), C.c_double(c.dt),
        C.c_double(c.t_end), C.c_size_t(c.S), C.c_long(SEED), C.c_int(ol.FIELD[c.field_shape]), C.c_double(c.H0),
        C.c_double(c.f), C.c_size_t(PARTICLE), prefix.encode())
    assert rc == 0
    return c


if __name__ == '__main__':
    write(os.path.join(HERE, 'reference_saved'))
    print('wrote', [f for f in sorted(os.listdir(HERE)) if f.startswith('reference_saved')])
