"""u16-resident c2 kernel at 2 / 3 / 4 CTAs per SM (EQB_FASTW_MINB_X16: 128 / 80 / 64 registers; variants built by hand into
csrc/build/variants/lib_x3.so, lib_x4.so).  usage (GPU box): python profiles/r2_x16_occupancy.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, eqtlbma_b200


class AnyEngine(eqtlbma_b200.Engine):
    def __init__(self, lib, prefix, ds, **kw):
        eqtlbma_b200._capi.Engine.__init__(self, lib, prefix, ds, **kw)


lib0 = eqtlbma_b200.load_library()
ds, _ = bench.make_shard(0, 1, lib0, None)
ds = bench.pinned_copy(ds)
dfx = bench.fixed_point_copy(ds)
for name in ["default", "x3", "x4", "default"]:
    path = "eqtlbma_b200/libeqtlbma_b200.so" if name == "default" else f"eqtlbma_b200/csrc/build/variants/lib_{name}.so"
    if not os.path.exists(path):
        continue
    lib = ctypes.CDLL(os.path.abspath(path))
    f = lib.eqb_last_pair_kernel_ms
    f.restype = ctypes.c_float
    eng = AnyEngine(lib, "eqb_", dfx, analysis="join", bfs="sin")
    for _ in range(5):
        eng.run_device_only(raw=True)
    ms, k = [], []
    for _ in range(10):
        ms.append(eng.run_device_only(raw=True))
        k.append(float(f(eng.ctx)))
    pairs = int(eng.pair_offsets()[-1])
    print(f"{name:8s} step {np.mean(ms):.4f} ms, pair kernel {np.mean(k):.4f} ms, {pairs / np.mean(ms) / 1e3:.1f} M pairs/s")
    eng.close()
