"""BASELINE.md section 3, baseline (i) with the REAL reference Python layer: builds the reference's own Cython module
(magpy/core.pyx) against oracle/_ref/libmagpy_ref.so into a scratch directory, imports the reference package from
/root/reference with stand-ins for its missing optional imports (toolz, matplotlib, transforms3d, scipy.integrate.trapz)
and times `magpy.EnsembleModel.simulate(n_jobs=nproc)` (joblib process pool, warm pool, first call discarded) on the
bench workload.  Only runs where /root/reference exists (the build container); nothing in tests/ or bench.py uses it.
bench.py's `cpu_baseline.joblib` drives the same compiled reference through the same kind of pool and can travel.

    python scripts/reference_native_baseline.py [n_members] > profiles/r02_reference_native_joblib.log
"""
import os
import subprocess
import sys
import sysconfig
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
SCRATCH = '/tmp/magpy_ref_native'


def build_reference_core():
    os.makedirs(os.path.join(SCRATCH, 'magpy'), exist_ok=True)
    ext = sysconfig.get_config_var('EXT_SUFFIX')
    out = os.path.join(SCRATCH, 'magpy', 'core' + ext)
    if os.path.exists(out):
        return
    csrc = os.path.join(SCRATCH, 'core.cpp')
    subprocess.check_call([sys.executable, '-m', 'cython', '--cplus', '-3', os.path.join(REF, 'magpy', 'core.pyx'), '-o', csrc])
    inc = sysconfig.get_paths()['include']
    libdir = os.path.join(ROOT, 'oracle', '_ref')
    subprocess.check_call(['g++', '-O2', '-fPIC', '-shared', '-w', '--std=c++11', '-I' + inc, '-I' + np.get_include(),
                           '-I' + os.path.join(REF, 'include'), '-I' + os.path.join(ROOT, 'oracle', 'ref_build', 'shim'),
                           '-include', 'stdexcept', '-include', 'cstdio', csrc, '-o', out, '-L' + libdir, '-lmagpy_ref',
                           '-Wl,-rpath,' + libdir])
    # the package's Python files are imported from where they lie: a namespace of symlinks, no copies
    for name in os.listdir(os.path.join(REF, 'magpy')):
        if name.endswith('.py') or os.path.isdir(os.path.join(REF, 'magpy', name)):
            dst = os.path.join(SCRATCH, 'magpy', name)
            if not os.path.lexists(dst):
                os.symlink(os.path.join(REF, 'magpy', name), dst)


def stub_optional_imports():
    """Stand-ins for the reference's optional imports that this image lacks, as real files in the scratch directory so
    that the joblib worker processes (fresh interpreters) find them too (PYTHONPATH + sitecustomize)."""
    def write(path, text):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, 'w') as fh:
            fh.write(text)
    write(os.path.join(SCRATCH, 'toolz', '__init__.py'), 'from . import dicttoolz\n')
    write(os.path.join(SCRATCH, 'toolz', 'dicttoolz.py'),
          'def merge(*ds):\n    return {k: v for d in ds for k, v in d.items()}\n')
    write(os.path.join(SCRATCH, 'matplotlib', '__init__.py'), '')
    write(os.path.join(SCRATCH, 'matplotlib', 'pyplot.py'), '')
    write(os.path.join(SCRATCH, 'transforms3d', '__init__.py'), 'from . import quaternions\n')
    write(os.path.join(SCRATCH, 'transforms3d', 'quaternions.py'), '')
    write(os.path.join(SCRATCH, 'sitecustomize.py'),
          'import scipy.integrate\nif not hasattr(scipy.integrate, "trapz"):\n    scipy.integrate.trapz = scipy.integrate.trapezoid\n')
    os.environ['PYTHONPATH'] = SCRATCH + os.pathsep + os.environ.get('PYTHONPATH', '')
    sys.path.insert(0, SCRATCH)
    import scipy.integrate
    if not hasattr(scipy.integrate, 'trapz'):
        scipy.integrate.trapz = scipy.integrate.trapezoid


def main():
    if not os.path.isdir(REF):
        raise SystemExit('needs the reference tree at ' + REF)
    build_reference_core()
    stub_optional_imports()
    import magpy
    cores = len(os.sched_getaffinity(0))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8 * cores
    base = magpy.Model(radius=[12e-9], anisotropy=[4e4], anisotropy_axis=[np.array([0., 0., 1.])],
                       magnetisation_direction=[np.array([0., 0., 1.])], location=[np.array([0., 0., 0.])],
                       magnetisation=4e5, damping=0.1, temperature=300., field_shape='sine', field_frequency=3e5,
                       field_amplitude=2e4)
    ens = magpy.EnsembleModel(n, base)
    kw = dict(end_time=1e-7, time_step=1e-12, max_samples=101, random_state=1001, n_jobs=cores, implicit_solve=False,
              interactions=True, renorm=False)
    ens.simulate(**kw)                       # warm pool
    t0 = time.perf_counter()
    res = ens.simulate(**kw)
    el = time.perf_counter() - t0
    steps = 100001
    print('reference-native magpy.EnsembleModel.simulate(n_jobs=%d): %d members x %d Heun steps in %.2f s = %.3e '
          'particle-steps/s (mean mz/Ms at end %.5f)' % (cores, n, steps, el, n * steps / el,
                                                         res.ensemble_magnetisation()[-1] / 4e5))


if __name__ == '__main__':
    main()
