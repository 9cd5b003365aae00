#!/usr/bin/env python3
"""TEST INFRASTRUCTURE.  Generates tests/golden/hm/*: for each scenario of HM_SCENARIOS the synthetic raw-ABF
dataset of eqtlbma_b200/hm_synth.py is written as `_l10abfs_raw.txt.gz` files, the UNMODIFIED reference eqtlbma_hm
(oracle/_ref/eqtlbma_hm_ref_dump, built by oracle/Makefile from /root/reference/src/eqtlbma_hm.cpp + hm_methods.cpp)
is run on them, and its full-precision dump plus its own text output are stored.  The datasets themselves are NOT
stored: the tests regenerate them from the seed and check the digest recorded in <name>.json.

Run from the repo root (needs /root/reference only through the prebuilt oracle/_ref binaries):
    python oracle/make_golden_hm.py
"""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from hm_scenarios import HM_SCENARIOS, build_dataset, ref_cmdline  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "eqtlbma_hm_ref_dump")
OUT = os.path.join(ROOT, "tests", "golden", "hm")


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, sc in HM_SCENARIOS.items():
        ds = build_dataset(sc)
        tmp = tempfile.mkdtemp(prefix="hm_" + name + "_")
        try:
            nfiles = sc.get("files", 1)
            per = (ds.n_genes + nfiles - 1) // nfiles
            for i in range(nfiles):
                ds.write_raw_file(os.path.join(tmp, "in_%d_l10abfs_raw.txt.gz" % i), i * per, min(ds.n_genes, (i + 1) * per))
            init = None
            if "init" in sc:
                init = os.path.join(tmp, "init.txt")
                with open(init, "w") as f:
                    f.write(sc["init"])
            cmd = [REF] + ref_cmdline(sc, ds, os.path.join(tmp, "in_*_l10abfs_raw.txt.gz"), os.path.join(tmp, "out_hm.txt.gz"), init)
            env = dict(os.environ, EQTLBMA_HM_DUMP=os.path.join(tmp, "dump.txt"))
            r = subprocess.run(cmd, env=env, cwd=tmp, capture_output=True, text=True)
            if r.returncode != 0:
                print(r.stdout[-2000:], r.stderr[-2000:])
                raise SystemExit("reference failed on " + name)
            with open(os.path.join(tmp, "dump.txt"), "rb") as f, gzip.GzipFile(os.path.join(OUT, name + ".dump.gz"), "wb", mtime=0) as g:
                g.write(f.read())
            shutil.copy(os.path.join(tmp, "out_hm.txt.gz"), os.path.join(OUT, name + ".out_hm.txt.gz"))
            iters = [ln for ln in r.stdout.splitlines() if ln.startswith("iter ")]
            with open(os.path.join(OUT, name + ".json"), "w") as f:
                json.dump({"digest": ds.digest(), "cmd": [os.path.basename(c) if os.sep in c else c for c in cmd[1:]],
                           "n_iter_lines": len(iters), "last_iter_line": iters[-1] if iters else ""}, f, indent=1)
            print(name, "ok", ds.n_genes, "genes", ds.n_pairs, "pairs", len(iters), "iteration lines")
        finally:
            shutil.rmtree(tmp, ignore_errors=True)


def chain():
    """The bf -> hm chain of the reference's functional test (tests/test_hm.bash:117-131): the reference eqtlbma_hm on the
    `_l10abfs_raw.txt.gz` file the reference eqtlbma_bf wrote for a scenario of tests/scenarios.py (stored as text in
    tests/golden/<scenario>.text.json.gz) -> tests/golden/hm/chain_<scenario>.out_hm.txt.gz."""
    from hm_scenarios import CHAIN_SCENARIOS, chain_cmdline
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "eqtlbma_hm_ref")
    for name in CHAIN_SCENARIOS:
        gold = json.loads(gzip.open(os.path.join(ROOT, "tests", "golden", name + ".text.json.gz"), "rt").read())
        tmp = tempfile.mkdtemp(prefix="hmchain_")
        try:
            raw = os.path.join(tmp, "ref_bf_l10abfs_raw.txt.gz")
            with gzip.open(raw, "wt") as f:
                f.write(gold["l10abfs_raw.txt.gz"])
            out = os.path.join(tmp, "out_hm.txt.gz")
            r = subprocess.run([ref_bin] + chain_cmdline(name, raw, out), capture_output=True, text=True, cwd=tmp)
            if r.returncode != 0:
                print(r.stdout[-2000:], r.stderr[-2000:])
                raise SystemExit("reference failed on chain " + name)
            shutil.copy(out, os.path.join(OUT, "chain_" + name + ".out_hm.txt.gz"))
            print("chain", name, "ok", len([ln for ln in r.stdout.splitlines() if ln.startswith("iter ")]), "iteration lines")
        finally:
            shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
    chain()
