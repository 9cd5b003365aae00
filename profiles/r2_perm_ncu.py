"""One permutation run on the c4-shaped slice for ncu captures.  usage: python profiles/r2_perm_ncu.py <pbf> <nperm> [n_genes]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eqtlbma_b200
from eqtlbma_b200.synth import make_dataset, make_grid

pbf, npm = sys.argv[1], int(sys.argv[2])
n_genes = int(sys.argv[3]) if len(sys.argv) > 3 else 8
ds = make_dataset(seed=3, n_subgroups=9, n_inds=450, n_genes=n_genes, snps_per_gene=5000, ragged=True, ragged_min_frac=0.34,
                  radius=10000, gene_spacing=20001, far_snp=False, n_chr=2, gridL=make_grid("general")[:10])
eng = eqtlbma_b200.Engine(ds, analysis="join", bfs="all" if pbf == "all" else "sin")
print(pbf, npm, eng.run_permutations_device_only(npm, 1859, pbf=pbf, wrtsize=10), "ms")
