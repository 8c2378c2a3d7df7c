"""In-tree build of libmagpy_b200.so (nvcc, sm_100a) and the Cython layer magpy_b200.core.

Run as ``python -m magpy_b200._build`` or through ``__graft_entry__.build()``.  Everything
is written next to the sources so the binaries travel with a snapshot of the repository.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _run(cmd, cwd=None):
    print('+', ' '.join(cmd), flush=True)
    subprocess.check_call(cmd, cwd=cwd)


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_library():
    _run(['make', '-C', os.path.join(HERE, 'csrc')])
    return os.path.join(HERE, 'libmagpy_b200.so')


def build_cython():
    import numpy
    pyx = os.path.join(HERE, 'core.pyx')
    csrc = os.path.join(HERE, 'core.c')
    ext = sysconfig.get_config_var('EXT_SUFFIX')
    out = os.path.join(HERE, 'core' + ext)
    hdr = os.path.join(ROOT, 'include', 'magpy_b200.h')
    if not _stale(out, [pyx, hdr]):
        return out
    _run([sys.executable, '-m', 'cython', '-3', pyx, '-o', csrc])
    inc = sysconfig.get_paths()['include']
    _run(['gcc', '-O2', '-fPIC', '-shared', '-w', '-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION',
          '-I' + inc, '-I' + numpy.get_include(), '-I' + os.path.join(ROOT, 'include'),
          csrc, '-o', out, '-L' + HERE, '-lmagpy_b200', '-Wl,-rpath,$ORIGIN'])
    return out


def build_all():
    build_library()
    build_cython()


if __name__ == '__main__':
    build_all()
