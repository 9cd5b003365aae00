"""c3 slice (9 ragged tissues of 450, ~5000 cis SNPs per gene, --bfs all: 511 configurations x 10 grid points): true pass with and
without the raw-ABF emission.  usage (GPU box): python profiles/gtex_slice.py [n_genes]"""
import os, sys, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, eqtlbma_b200
from eqtlbma_b200.synth import make_dataset, make_grid
n_genes = int(sys.argv[1]) if len(sys.argv) > 1 else 48
ds = make_dataset(seed=3, n_subgroups=9, n_inds=450, n_genes=n_genes, snps_per_gene=5000, ragged=True, ragged_min_frac=0.34,
                  radius=10000, gene_spacing=20001, far_snp=False, n_chr=2, gridL=make_grid("general")[:10])
eng = eqtlbma_b200.Engine(ds, analysis="join", bfs="all")
pairs = int(eng.pair_offsets()[-1]); print('pairs', pairs, 'fast genes', eng.fast_gene_count(), 'configs', eng.n_configs)
raw_bytes = pairs * (3 * eng.L + eng.n_configs * eng.K) * 8
for raw in (False, True):
    for i in range(3):
        ms = eng.run_device_only(raw=raw)
    kms = eng.last_pair_kernel_ms()
    print('c3 slice bfs all raw=%s: step %.2f ms (pair kernel %.2f ms) -> %.2f M pairs/s%s' % (
        raw, ms, kms, pairs / ms / 1e3, (', raw emission %.2f TB/s' % (raw_bytes / (kms * 1e-3) / 1e12)) if raw else ''))
