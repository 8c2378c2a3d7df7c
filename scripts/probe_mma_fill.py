"""GPU probe: where does the DMMA cluster kernel (groups of 8 particles, padded) beat the scalar one?"""
import os
import sys
sys.path.insert(0, '.')
sys.path.insert(0, 'scripts')
from probe_cluster import run

for N in (9, 10, 12, 14, 17, 20, 25, 28, 33, 36, 41, 49, 57):
    R = max(4096, (1 << 20) // N // 2)
    steps = max(200, 40000 // N)
    for kern in ('simt', 'mma'):
        os.environ['MAGPY_B200_CLUSTER_KERNEL'] = kern
        print(kern, end=' ')
        run(N, R, steps)
