import os, sys
sys.path.insert(0, '.')
import bench, eqtlbma_b200
lib = eqtlbma_b200.load_library()
ds, _ = bench.make_shard(0, 1, lib, None)
ds = bench.pinned_copy(ds)
dfx = bench.fixed_point_copy(ds)
eng = eqtlbma_b200.Engine(dfx, analysis="join", bfs="sin")
for _ in range(6):
    eng.run_device_only(raw=True)
