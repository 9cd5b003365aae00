"""A/B of the true-pass contraction on the c2 bench workload: genotypes resident as doubles (eqb_set_genotypes) vs as the u16
numerators of the fixed-point transport (eqb_set_genotypes_fixed: phase A of fast_pair_warp_kernel reads 4x fewer bytes);
results must be bit-identical.  usage (GPU box): python profiles/r2_x16_ab.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, eqtlbma_b200
lib = eqtlbma_b200.load_library()
ds, _ = bench.make_shard(0, 1, lib, None)
ds = bench.pinned_copy(ds)
dfx = bench.fixed_point_copy(ds)
res = {}
for name, d in (("f64", ds), ("u16", dfx)):
    eng = eqtlbma_b200.Engine(d, analysis="join", bfs="sin")
    for _ in range(5):
        eng.run_device_only(raw=True)
    ms = [eng.run_device_only(raw=True) for _ in range(10)]
    k = [eng.last_pair_kernel_ms()]
    for _ in range(9):
        eng.run_device_only(raw=True); k.append(eng.last_pair_kernel_ms())
    res[name] = eng.run(raw=True)
    pairs = int(eng.pair_offsets()[-1])
    print(f"{name}: step {np.mean(ms):.4f} ms, pair kernel {np.mean(k):.4f} ms, {pairs / np.mean(ms) / 1e3:.1f} M pairs/s")
    eng.close()
same = all(np.array_equal(getattr(res["f64"], f), getattr(res["u16"], f), equal_nan=True) for f in ("n", "sstats", "abf_gen", "abf_cfg", "abf_w"))
print("bit-identical:", same)
