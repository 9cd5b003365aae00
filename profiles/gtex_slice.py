import os, sys, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, eqtlbma_b200
from eqtlbma_b200.synth import make_dataset, make_grid
t=time.time()
ds = make_dataset(seed=3, n_subgroups=9, n_inds=450, n_genes=48, snps_per_gene=5000, ragged=True, ragged_min_frac=0.34,
                  radius=10000, gene_spacing=20001, far_snp=False, n_chr=2, gridL=make_grid("general")[:10])
print('gen', time.time()-t, ds.n_snps)
eng = eqtlbma_b200.Engine(ds, analysis="join", bfs="all")
pairs = int(eng.pair_offsets()[-1]); print('pairs', pairs, 'fast genes', eng.fast_gene_count(), 'configs', eng.n_configs)
for raw in (False, True):
    for i in range(3):
        ms = eng.run_device_only(raw=raw)
    print('c3 slice bfs all raw=%s: %.2f ms -> %.2f M pairs/s' % (raw, ms, pairs/ms/1e3))
for pbf, npm in (("gen", 200), ("gen-sin", 200), ("all", 20)):
    for i in range(2):
        ms = eng.run_permutations_device_only(npm, 1859, pbf=pbf, wrtsize=10)
    print('c4 slice pbf=%s nperm=%d: %.1f ms -> %.2f M pair-perms/s' % (pbf, npm, ms, pairs*npm/ms/1e3))
