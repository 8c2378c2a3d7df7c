"""GPU probe: discrete-orientation model batch throughput (one thread per particle) next to the compiled reference
(simulation::dom_ensemble_dynamics, one particle per call on one host core)."""
import sys
import time
import numpy as np
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import magpy_b200 as mp
import oracle_lib as ol

model = mp.DOModel(6e-9, 4e4, [1, 0], 4e5, 0.1, 300.0, field_shape='sine', field_frequency=3e5, field_amplitude=2e4)
t_end, dt, S = 1e-5, 1e-10, 101   # three field periods; >= 1000 adaptive steps per particle (step cap end_time / 1000)
for n in (1, 1000, 100_000, 1_000_000):
    radius = np.linspace(5e-9, 8e-9, n) if n > 1 else np.array([6e-9])
    model.simulate_batch(radius[:1], 4e4, t_end, dt, S)
    t0 = time.perf_counter()
    out = model.simulate_batch(radius, 4e4, t_end, dt, S)
    el = time.perf_counter() - t0
    steps = int(out['steps'].sum())
    print(f'GPU n={n:8d}: {el*1e3:9.2f} ms wall  {steps/el:.3e} RK45 steps/s  {n/el:.3e} particles/s  (mean {steps/n:.0f} steps per particle)', flush=True)
lib = ol.load_reference()
kind = 'reference'
if lib is None:
    lib, kind = ol.load_oracle(), 'oracle port'
radius = np.linspace(5e-9, 8e-9, 200)
t0 = time.perf_counter()
for r in radius:
    ol.dom_simulate(lib, float(r), 4e4, [1, 0], 4e5, 0.1, 300.0, dt, t_end, S, 'sine', 2e4, 3e5, reference=(kind == 'reference'))
el = time.perf_counter() - t0
print(f'CPU {kind}, 1 core, n=200: {el*1e3:9.2f} ms  {200/el:.3e} particles/s')
