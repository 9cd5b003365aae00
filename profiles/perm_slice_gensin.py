import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, eqtlbma_b200, numpy as np
from eqtlbma_b200.synth import make_dataset
pds = make_dataset(**bench.PERM_WORKLOAD)
peng = eqtlbma_b200.Engine(pds, analysis="join", bfs="sin")
pp = int(peng.pair_offsets()[-1])
for i in range(3):
    ms = peng.run_permutations_device_only(bench.PERM_NPERM, 1859, pbf="gen-sin", wrtsize=10)
    print('gen-sin ms', ms, pp*bench.PERM_NPERM/ms/1e3, 'M pair-perms/s')
ms = peng.run_permutations_device_only(bench.PERM_NPERM, 1859, pbf="gen", wrtsize=10); print('gen ms', ms)
