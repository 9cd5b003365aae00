"""Golden fixtures of the --inss path (Bayes factors from summary-statistics files): the UNMODIFIED reference
(oracle/_ref/eqtlbma_bf_ref) first writes `_sumstats_<subgroup>.txt.gz` with --outss on a seeded scenario, then reads them
back with --inss; inputs (sumstats files, grids) and outputs (l10abfs_raw, l10abfs_avg-grids) go to tests/golden/inss/<name>/.
usage: python oracle/make_golden_inss.py   (build container only: needs /root/reference compiled under oracle/_ref)"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from scenarios import SCENARIOS, build_dataset  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "eqtlbma_bf_ref")
CASES = {"basic_all": ("basic_all_perm", "all"), "absent_sin": ("absent_genes_nan", "sin"), "ragged5_all": ("ragged5", "all"),
         "basic_gen": ("basic_all_perm", "gen")}

for name, (scen, bfs) in CASES.items():
    ds = build_dataset(SCENARIOS[scen])
    tmp = tempfile.mkdtemp(prefix="eqb_inss_")
    out = os.path.join(ROOT, "tests", "golden", "inss", name)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    ds.write_files(tmp)
    subprocess.check_call([REF] + ds.ref_args(tmp, os.path.join(tmp, "o1")) + ["--analys", "join", "--bfs", "sin", "--outss", "-v", "0"])
    lst = os.path.join(tmp, "list_sstats.txt")
    with open(lst, "w") as fh:
        for f in sorted(os.listdir(tmp)):
            if f.startswith("o1_sumstats_"):
                sg = f[len("o1_sumstats_"):-len(".txt.gz")]
                shutil.copy(os.path.join(tmp, f), os.path.join(out, f"sumstats_{sg}.txt.gz"))
                fh.write(f"{sg}\t{os.path.join(tmp, f)}\n")
    grids = ["--gridL", os.path.join(tmp, "grid_phi2_oma2_general.txt.gz")]
    if bfs != "gen":
        grids += ["--gridS", os.path.join(tmp, "grid_phi2_oma2_with-configs.txt.gz")]
    for g in grids[1::2]:
        shutil.copy(g, out)
    subprocess.check_call([REF, "--inss", lst, "--out", os.path.join(tmp, "o2"), "--analys", "join", "--bfs", bfs, "--outw", "-v", "0"] + grids)
    for f in ("o2_l10abfs_raw.txt.gz", "o2_l10abfs_avg-grids.txt.gz"):
        shutil.copy(os.path.join(tmp, f), os.path.join(out, "expected_" + f[3:]))
    with open(os.path.join(out, "bfs.txt"), "w") as fh:
        fh.write(bfs + "\n")
    shutil.rmtree(tmp, ignore_errors=True)
    print(name, sorted(os.listdir(out)))
