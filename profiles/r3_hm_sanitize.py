"""Small run of every hm kernel (several grid sizes / dims, posterior pass) for compute-sanitizer:
compute-sanitizer --tool memcheck|racecheck|synccheck python profiles/r3_hm_sanitize.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
from eqtlbma_b200.hm import HmEngine, HmFit
from eqtlbma_b200.hm_synth import make_hm_dataset

for shape in (dict(n_genes=40, snps_lo=1, snps_hi=60, n_subgroups=3, grid=10), dict(n_genes=20, snps_lo=1, snps_hi=9, n_subgroups=6, grid=7),
              dict(n_genes=15, snps_lo=30, snps_hi=300, n_subgroups=2, grid=25)):
    ds = make_hm_dataset(seed=5, round_text=False, **shape)
    hm = HmEngine(ds.dim, ds.grid)
    hm.append(ds.B, ds.gene_off)
    hm.finalize()
    fit = hm.em(HmFit(0.5, np.full(ds.grid, 1.0 / ds.grid), np.full(ds.dim, 1.0 / ds.dim)), thresh=0.05, stepmax=3.0, maxit=6)
    post = hm.posteriors(fit)
    print(shape, "loglik", fit.loglik, "launches", hm.launch_count, float(post["gene_post"].mean()))
    hm.close()
