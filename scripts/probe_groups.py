"""GPU probe: an ensemble split into many parameter groups (field-amplitude sweep), groups as concurrent plans vs one at a time."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import magpy_b200 as mp
from magpy_b200 import model as model_mod

groups, per = 64, 1000
amps = np.repeat(np.linspace(5e3, 2.5e4, groups), per)
base = mp.Model([12e-9], [4e4], [[0, 0, 1.0]], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, field_shape='sine',
                field_frequency=3e5, field_amplitude=1e4)
ens = mp.EnsembleModel(groups * per, base, field_amplitude=list(amps))
for conc in (32, 1, 32):
    model_mod._MAX_CONCURRENT_PLANS = conc
    t0 = time.perf_counter()
    res = ens.simulate(1e-8, 1e-12, 101, 7, implicit_solve=False, return_trajectories=False)
    print(f'{groups} groups x {per} members x 10000 Heun steps, {conc:2d} plans in flight: {1e3 * (time.perf_counter() - t0):8.1f} ms wall', flush=True)
