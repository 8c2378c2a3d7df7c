"""magpy_b200 — B200-native ensemble stochastic Landau-Lifshitz-Gilbert integration behind
the API of owlas/magpy's ``magpy.Model`` / ``magpy.EnsembleModel`` / ``magpy.core``.

Only the sLLG time-integration path is implemented; the compute runs in
hand-written sm_100a CUDA kernels behind a C ABI (include/magpy_b200.h).  Importing
``magpy_b200.core`` fails loudly if the native library has not been built
(``python -m magpy_b200._build``); there is no CPU fallback.
"""
from . import core
from . import results
from . import sharding
from . import geometry
from .core import simulate, simulate_ensemble, simulate_dom, simulate_dom_batch, get_KB, get_mu0, get_gamma
from .model import Model, EnsembleModel, DOModel
from .results import Results, EnsembleResults

__all__ = ['core', 'results', 'sharding', 'geometry', 'simulate', 'simulate_ensemble', 'get_KB', 'get_mu0', 'get_gamma',
           'Model', 'EnsembleModel', 'Results', 'EnsembleResults']
