#!/bin/bash
# c3 / c4 shape at scale on N GPUs of one box (default 8): every rank takes 32 genes x ~5,000 cis SNPs (S = 9 ragged tissues of
# 450 individuals) -- true pass --bfs all, then 10^4 permutations with --pbf all and --pbf gen-sin -- next to the c2 bench
# lines; one dataset per rank for the c4 block, ONE partitioned dataset for c2.  Output: gpurun_out/r2_c4_scale_n$N.json
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 5 --warmup 3 --perm-genes 32 --perm-nperm 10000 \
  > gpurun_out/r2_c4_scale_n$N.json 2> gpurun_out/r2_c4_scale_n$N.err
tail -c 400 gpurun_out/r2_c4_scale_n$N.err
python - <<P
import json
d = json.loads(open("gpurun_out/r2_c4_scale_n$N.json").read().strip().splitlines()[-1])
p = d["perm"]
print("N", d["n_gpus"], "c2 value %.1f M pairs/s, e2e %.1f M" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6))
print("c3 true pass --bfs all:", {k: round(v["pairs_per_s"] / 1e6, 1) for k, v in p["true_pass_bfs_all"].items() if isinstance(v, dict)}, "M pairs/s")
print("c4 nperm", p["nperm"], {k: round(v["permuted_pairs_per_s"] / 1e6, 1) for k, v in p["runs"].items()}, "M permuted pairs/s")
P
