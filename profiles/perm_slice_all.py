import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, eqtlbma_b200, numpy as np
from eqtlbma_b200.synth import make_dataset, make_grid
pds = make_dataset(**dict(bench.PERM_WORKLOAD, gridL=make_grid("general")[:10]))
peng = eqtlbma_b200.Engine(pds, analysis="join", bfs="all")
pp = int(peng.pair_offsets()[-1])
for i in range(2):
    ms = peng.run_permutations_device_only(20, 1859, pbf="all", wrtsize=10)
    print('all ms', ms, pp*20/ms/1e3, 'M pair-perms/s')
