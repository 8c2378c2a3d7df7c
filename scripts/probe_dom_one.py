"""One launch of the discrete-orientation kernel for `ncu --set full` (see profiles/README.md)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import magpy_b200 as mp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
model = mp.DOModel(6e-9, 4e4, [1, 0], 4e5, 0.1, 300.0, field_shape='sine', field_frequency=3e5, field_amplitude=2e4)
out = model.simulate_batch(np.linspace(5e-9, 8e-9, n), 4e4, 1e-6, 1e-10, 21)
print(int(out['steps'].sum()), 'RK45 steps')
