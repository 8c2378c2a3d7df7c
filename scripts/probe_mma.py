"""GPU probe: Heun cluster kernels, scalar (cluster.cu) vs DMMA (cluster_mma.cu), over N."""
import os
import sys
sys.path.insert(0, '.')
sys.path.insert(0, 'scripts')
from probe_cluster import run

cases = [(8, 1 << 17, 4000), (16, 1 << 16, 2000), (24, 1 << 15, 1000), (32, 1 << 15, 1000), (40, 20000, 1000), (48, 20000, 1000),
         (56, 9472, 1000), (64, 9472, 1000), (64, 12500, 1000), (72, 4736, 400), (96, 4736, 400), (128, 4736, 400), (128, 2368, 400)]
if len(sys.argv) > 1:
    want = [int(x) for x in sys.argv[1].split(',')]
    cases = [c for c in cases if c[0] in want]
for N, R, steps in cases:
    for kern in ('simt', 'mma'):
        os.environ['MAGPY_B200_CLUSTER_KERNEL'] = kern
        print(kern, end=' ')
        run(N, R, steps)
