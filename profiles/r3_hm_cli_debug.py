"""eqtlbma_b200/eqtlbma_hm with 1 / 4 / 16 loader threads against the reference binary on the first 1000 genes of the bench's
hm workload (same file)."""
import gzip, os, subprocess, sys, tempfile
sys.path.insert(0, ".")
from eqtlbma_b200.hm_synth import make_hm_dataset
ds = make_hm_dataset(seed=1861, n_genes=10000, snps_lo=50, snps_hi=150, n_subgroups=3, grid=10, round_text=False)
tmp = tempfile.mkdtemp()
f = os.path.join(tmp, "s_l10abfs_raw.txt.gz")
ds.write_raw_file(f, 0, 1000)
base = ["--data", f, "--nsubgrp", "3", "--dim", "7", "--ngrid", "10", "-v", "1"]
subprocess.run(["oracle/_ref/eqtlbma_hm_ref"] + base + ["--out", os.path.join(tmp, "ref.gz"), "--thread", "16"], capture_output=True)
ref = gzip.open(os.path.join(tmp, "ref.gz"), "rt").read()
for th in (1, 4, 16):
    r = subprocess.run(["eqtlbma_b200/eqtlbma_hm"] + base + ["--out", os.path.join(tmp, "o%d.gz" % th), "--thread", str(th)], capture_output=True, text=True)
    got = gzip.open(os.path.join(tmp, "o%d.gz" % th), "rt").read()
    its = [l for l in r.stdout.splitlines() if l.startswith("iter ")]
    print("threads", th, "rc", r.returncode, "same as reference:", got == ref, "iteration lines", len(its))
    if got != ref:
        print(r.stderr[-300:])
        print(its[0][:150]); print(its[-1][:150])
        for a, b in zip(got.splitlines(), ref.splitlines()):
            if a != b: print("  ours", a, "| ref", b)
