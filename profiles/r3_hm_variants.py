"""A/B of hm_estep_kernel build variants (HM_MIN_CTAS, HM_STAGES_SMALL; hm.cu alone built by hand into
csrc/build/variants/lib_hm_<name>.so) on the bench's hm workload.  usage (GPU box): python profiles/r3_hm_variants.py"""
import ctypes
import os
import sys

sys.path.insert(0, ".")
import numpy as np

import eqtlbma_b200
import eqtlbma_b200.hm as hmmod
from eqtlbma_b200.hm_synth import make_hm_dataset

ds = make_hm_dataset(seed=1861, n_genes=10000, snps_lo=50, snps_hi=150, n_subgroups=3, grid=10, round_text=False)
gw, cp = np.full(ds.grid, 1.0 / ds.grid), np.full(ds.dim, 1.0 / ds.dim)
alg = 8.0 * ds.n_pairs * ds.dim * ds.grid
names = ["default"] + sorted(f[7:-3] for f in os.listdir("eqtlbma_b200/csrc/build/variants") if f.startswith("lib_hm_")) + ["default"]
ref = None
for name in names:
    path = "eqtlbma_b200/libeqtlbma_b200.so" if name == "default" else f"eqtlbma_b200/csrc/build/variants/lib_hm_{name}.so"
    eqtlbma_b200._lib = ctypes.CDLL(os.path.abspath(path))
    hm = hmmod.HmEngine(ds.dim, ds.grid)
    hm.append(ds.B, ds.gene_off)
    hm.finalize()
    hm.estep_device_only(gw, cp, reps=3)
    ms = min(hm.estep_device_only(gw, cp, reps=20) for _ in range(3))
    lik = hm.loglik(0.4, gw, cp)
    ref = lik if ref is None else ref
    print(f"{name:10s} {ms:.4f} ms  {alg / ms / 1e6:7.1f} GB/s  loglik rel diff {abs(lik - ref) / abs(ref):.2e}")
    hm.close()
