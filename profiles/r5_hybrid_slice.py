"""--error hybrid slice: 3 ragged subgroups of up to 450 individuals, 200 genes x ~50 cis SNPs, true pass (--bfs sin and
--bfs all) and a permutation pass on the device, with the unmodified reference binary (oracle/_ref, one core: its true pass
is single-threaded) on the first genes of the same workload beside it.
usage (GPU box): python profiles/r5_hybrid_slice.py [n_genes] [ref_genes]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, eqtlbma_b200
from eqtlbma_b200.synth import make_dataset
import bench

n_genes = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ref_genes = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ds = make_dataset(seed=11, n_subgroups=3, n_inds=450, n_genes=n_genes, snps_per_gene=50, ragged=True, ragged_min_frac=0.6,
                  radius=1000, gene_spacing=2001, far_snp=False, n_chr=2)
out = {"workload": f"hybrid slice: 3 ragged subgroups <= 450 individuals, {n_genes} genes x ~50 cis SNPs"}
for bfs in ("sin", "all"):
    eng = eqtlbma_b200.Engine(ds, analysis="join", bfs=bfs, error="hybrid", fiterr=0.5)
    pairs = int(eng.pair_offsets()[-1])
    for i in range(4):
        ms = eng.run_device_only(raw=True)
    out[f"true_pass_{bfs}"] = {"pairs": pairs, "ms": ms, "pairs_per_s": pairs / ms * 1e3}
    if bfs == "sin":
        nperm = 100
        for i in range(2):
            pms = eng.run_permutations_device_only(nperm, 7, pbf="gen-sin", wrtsize=n_genes)
        out["perm_gen-sin"] = {"pair_perms": pairs * nperm, "ms": pms, "pair_perms_per_s": pairs * nperm / pms * 1e3}
    eng.close()
    if os.environ.get("HYBRID_NO_REF"):
        continue
    sub = bench.subset_dataset(ds, ref_genes)
    r = bench.time_reference_binary(sub, ["--analys", "join", "--bfs", bfs, "--error", "hybrid", "--fiterr", "0.5", "--outss", "--outw"])
    if r:
        out[f"reference_{bfs}"] = {"pairs": r["pairs"], "seconds": r["seconds"], "pairs_per_s": r["pairs"] / r["seconds"], "cores": 1}
        out[f"ratio_{bfs}"] = out[f"true_pass_{bfs}"]["pairs_per_s"] / out[f"reference_{bfs}"]["pairs_per_s"]
print(json.dumps(out))
