"""GPU probe: BASELINE config 2 (10k implicit dimers) with one thread per cluster vs one lane per particle."""
import os
import sys
sys.path.insert(0, '.')
sys.path.insert(0, 'scripts')
from probe_cluster import run

for R in (1000, 10000, 40000, 1 << 18):
    for N in (2, 4):
        for kern in ('thread', 'split', 'warps'):
            os.environ['MAGPY_B200_SMALL_KERNEL'] = kern
            print(kern, end=' ')
            run(N, R, 1000, implicit=True)
