"""Command-line behaviour that needs no device (option validation and input errors happen before the first CUDA call):
messages follow eqtlbma_bf.cpp:463-692 and data_loader.cpp:1280-1292."""
import gzip
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "eqtlbma_b200", "eqtlbma_bf")
GOLD = os.path.join(ROOT, "tests", "golden", "inss", "basic_gen")


def _exe():
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "eqtlbma_b200", "host")])
    return EXE


def _run(args):
    return subprocess.run([_exe()] + args, capture_output=True, text=True, timeout=120)


def test_help_states_limits():
    r = _run(["--help"])
    assert r.returncode == 0
    assert "2048 samples" in r.stdout and "--gpus" in r.stdout and "--shard" in r.stdout


def test_inss_requires_join_and_uvlr(tmp_path):
    lst = tmp_path / "l.txt"
    lst.write_text("s1\t%s\n" % os.path.join(GOLD, "sumstats_s1.txt.gz"))
    grid = os.path.join(GOLD, "grid_phi2_oma2_general.txt.gz")
    r = _run(["--inss", str(lst), "--out", str(tmp_path / "o"), "--analys", "sep", "--gridL", grid])
    assert r.returncode != 0 and "--inss requires --analys join" in r.stderr
    r = _run(["--inss", str(lst), "--out", str(tmp_path / "o"), "--analys", "join", "--error", "mvlr", "--gridL", grid])
    assert r.returncode != 0 and "--inss requires --error uvlr" in r.stderr
    r = _run(["--inss", str(tmp_path / "absent.txt"), "--out", str(tmp_path / "o"), "--analys", "join", "--gridL", grid])
    assert r.returncode != 0 and "can't find" in r.stderr


def test_inss_missing_column_is_reported(tmp_path):
    bad = tmp_path / "bad.txt.gz"
    with gzip.open(bad, "wt") as fh:
        fh.write("gene\tsnp\tn\tbetahat.geno\tsebetahat.geno\ngene1\tsnp1\t100\t0.1\t0.05\n")
    lst = tmp_path / "l.txt"
    lst.write_text("s1\t%s\n" % bad)
    r = _run(["--inss", str(lst), "--out", str(tmp_path / "o"), "--analys", "join", "--bfs", "gen", "-v", "0",
              "--gridL", os.path.join(GOLD, "grid_phi2_oma2_general.txt.gz")])
    assert r.returncode != 0 and "missing sigmahat in header" in r.stderr


def test_missing_compulsory_options():
    r = _run(["--out", "x", "--analys", "join"])
    assert r.returncode != 0 and "missing compulsory option --geno" in r.stderr


def test_error_model_validation(tmp_path):
    """--error uvlr|mvlr|hybrid are the reference's three models (eqtlbma_bf.cpp:617-622); anything else is rejected, and
    --inss stays restricted to uvlr."""
    lst = tmp_path / "l.txt"
    lst.write_text("s1\t%s\n" % os.path.join(GOLD, "sumstats_s1.txt.gz"))
    grid = os.path.join(GOLD, "grid_phi2_oma2_general.txt.gz")
    r = _run(["--geno", grid, "--scoord", grid, "--exp", grid, "--gcoord", grid, "--out", str(tmp_path / "o"), "--analys", "join",
              "--bfs", "gen", "--gridL", grid, "--error", "bogus"])
    assert r.returncode != 0 and "--error bogus is not valid" in r.stderr
    r = _run(["--inss", str(lst), "--out", str(tmp_path / "o"), "--analys", "join", "--error", "hybrid", "--gridL", grid])
    assert r.returncode != 0 and "is not valid" not in r.stderr and "--inss requires --error uvlr" in r.stderr
    assert "hybrid" in _run(["--help"]).stdout
