"""A/B of fast_pair_all_kernel build variants (FA_U, FA_MINB, FA_CLAMP; built by hand into csrc/build/variants/lib_<name>.so) on
the c3 slice.  usage (GPU box): python profiles/r2_fast_all_variants.py"""
import os, sys, ctypes
sys.path.insert(0, '.')
import numpy as np
import eqtlbma_b200


class AnyEngine(eqtlbma_b200.Engine):
    """Engine on an explicitly chosen build of the library"""
    def __init__(self, lib, prefix, ds, **kw):
        eqtlbma_b200._capi.Engine.__init__(self, lib, prefix, ds, **kw)


from eqtlbma_b200.synth import make_dataset, make_grid
ds = make_dataset(seed=3, n_subgroups=9, n_inds=450, n_genes=32, snps_per_gene=5000, ragged=True, ragged_min_frac=0.34,
                  radius=10000, gene_spacing=20001, far_snp=False, n_chr=2, gridL=make_grid("general")[:10])
ref = None
for name in ["default", "u1", "u4", "noclamp", "u4b3", "u1noclamp", "default"]:
    path = "eqtlbma_b200/libeqtlbma_b200.so" if name == "default" else f"eqtlbma_b200/csrc/build/variants/lib_{name}.so"
    lib = ctypes.CDLL(os.path.abspath(path))
    eng = AnyEngine(lib, "eqb_", ds, analysis="join", bfs="all")
    pairs = int(eng.pair_offsets()[-1])
    f = lib.eqb_last_pair_kernel_ms; f.restype = ctypes.c_float
    out = []
    for raw in (False, True):
        for _ in range(3):
            ms = eng.run_device_only(raw=raw)
        out.append((ms, float(f(eng.ctx))))
    r = eng.run(0, 2, raw=False)
    if ref is None: ref = r.abf_w.copy()
    ok = np.allclose(r.abf_w, ref, rtol=0, atol=1e-10, equal_nan=True)
    print(f"{name:10s} no-raw {pairs/out[0][0]/1e3:6.2f} M/s (kern {out[0][1]:.3f} ms)  raw {pairs/out[1][0]/1e3:6.2f} M/s (kern {out[1][1]:.3f} ms) same={ok}")
    eng.close()
