"""--inss (Bayes factors from summary-statistics files, data_loader.cpp:1212-1343 + gene.cpp:293-311): the drop-in binary
against the UNMODIFIED reference's outputs on the same sumstats files (fixtures: oracle/make_golden_inss.py)."""
import gzip
import os
import subprocess

import pytest

from test_cli_dropin import cells_match

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "eqtlbma_b200", "eqtlbma_bf")
GOLD = os.path.join(ROOT, "tests", "golden", "inss")


@pytest.mark.parametrize("name", sorted(os.listdir(GOLD)) if os.path.isdir(GOLD) else [])
@pytest.mark.parametrize("threads", [1, 3])
def test_inss_outputs_match_reference_text(tmp_path, name, threads):
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "eqtlbma_b200", "host")])
    d = os.path.join(GOLD, name)
    bfs = open(os.path.join(d, "bfs.txt")).read().strip()
    lst = tmp_path / "list_sstats.txt"
    with open(lst, "w") as fh:
        for f in sorted(os.listdir(d)):
            if f.startswith("sumstats_"):
                fh.write(f"{f[len('sumstats_'):-len('.txt.gz')]}\t{os.path.join(d, f)}\n")
    out = str(tmp_path / "obs")
    cmd = [EXE, "--inss", str(lst), "--out", out, "--analys", "join", "--bfs", bfs, "--outw", "-v", "0", "--thread", str(threads),
           "--gridL", os.path.join(d, "grid_phi2_oma2_general.txt.gz")]
    if bfs != "gen":
        cmd += ["--gridS", os.path.join(d, "grid_phi2_oma2_with-configs.txt.gz")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    n_cells = n_exact = 0
    for fn in ("l10abfs_raw.txt.gz", "l10abfs_avg-grids.txt.gz"):
        exp = gzip.open(os.path.join(d, "expected_" + fn), "rt").read().splitlines()
        got = gzip.open(out + "_" + fn, "rt").read().splitlines()
        assert len(exp) == len(got), (fn, len(exp), len(got))
        for ln, (e, g) in enumerate(zip(exp, got)):
            et, gt = e.split("\t"), g.split("\t")
            assert len(et) == len(gt), (fn, ln, e, g)
            for ci, (a, b) in enumerate(zip(et, gt)):
                ok, exact = cells_match(a, b)
                assert ok, (fn, ln, ci, a, b)
                n_cells += 1
                n_exact += exact
    assert n_exact >= 0.99 * n_cells, (n_exact, n_cells)
