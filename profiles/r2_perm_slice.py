"""Permutation throughput on a c4-shaped slice (9 ragged tissues of 450 individuals, ~5000 cis SNPs per gene) with the
per-kernel device times of the batched-GEMM path, the FP64 pipe peaks and the GEMM kernel's own throughput.
usage (GPU box): python profiles/r2_perm_slice.py [n_genes] [nperm_gen] [nperm_all]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eqtlbma_b200
from eqtlbma_b200.synth import make_dataset, make_grid

n_genes = int(sys.argv[1]) if len(sys.argv) > 1 else 16
np_gen = int(sys.argv[2]) if len(sys.argv) > 2 else 511
np_all = int(sys.argv[3]) if len(sys.argv) > 3 else 127
pk = eqtlbma_b200.measure_fp64_peaks()
print("fp64 peaks:", json.dumps(pk))
for shape in ((5000, 512, 464), (5000, 3456, 464), (1000, 1024, 304)):
    print("gemm selftest", shape, json.dumps(eqtlbma_b200.selftest_perm_gemm(*shape)))
t = time.time()
ds = make_dataset(seed=3, n_subgroups=9, n_inds=450, n_genes=n_genes, snps_per_gene=5000, ragged=True, ragged_min_frac=0.34,
                  radius=10000, gene_spacing=20001, far_snp=False, n_chr=2, gridL=make_grid("general")[:10])
print("dataset %.1f s, snps %d" % (time.time() - t, ds.n_snps))
eng = eqtlbma_b200.Engine(ds, analysis="join", bfs="all")
pairs = int(eng.pair_offsets()[-1])
for pbf, npm in (("gen", np_gen), ("gen-sin", np_gen), ("all", np_all)):
    eng.set_perm_timing(False)
    for i in range(2):
        ms = eng.run_permutations_device_only(npm, 1859, pbf=pbf, wrtsize=10)
    eng.set_perm_timing(True)   # per-kernel events (one synchronisation per column batch)
    eng.run_permutations_device_only(npm, 1859, pbf=pbf, wrtsize=10)
    tm = eng.last_perm_timing()
    tm["pbf"], tm["nperm"], tm["pairs"], tm["ms"] = pbf, npm, pairs, ms
    tm["M_pair_perms_per_s"] = pairs * (npm + 1) / ms / 1e3
    tm["gemm_tflops_issued"] = tm["gemm_flops"] / (tm["gemm_ms"] * 1e-3) / 1e12 if tm["gemm_ms"] else None
    tm["gemm_tflops_useful"] = tm["gemm_useful_flops"] / (tm["gemm_ms"] * 1e-3) / 1e12 if tm["gemm_ms"] else None
    print(json.dumps(tm))
