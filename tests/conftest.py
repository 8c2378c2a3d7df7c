import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    """The binaries are git-ignored: on a fresh checkout build them once (nvcc cross-compiles without a GPU)."""
    import glob
    have = (os.path.exists(os.path.join(ROOT, 'magpy_b200', 'libmagpy_b200.so'))
            and glob.glob(os.path.join(ROOT, 'magpy_b200', 'core.*.so'))
            and os.path.exists(os.path.join(ROOT, 'oracle', 'libsllg_oracle.so')))
    if not have:
        import __graft_entry__
        __graft_entry__.build()


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    _ensure_built()


def _cuda_devices():
    try:
        import magpy_b200.core as core
        return core.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
