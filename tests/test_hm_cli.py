"""The drop-in front-end eqtlbma_b200/eqtlbma_hm against the reference's own output files (tests/golden/hm/*.out_hm.txt.gz,
written by the unmodified reference eqtlbma_hm on the same input files): same lines, same text cells, numeric cells equal
up to the last of the 4 printed digits.  CPU: option checks, the loader, and the loud failure without a device."""
import gzip
import os
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hm_scenarios import GOLDEN_HM, HM_SCENARIOS, build_dataset, ref_cmdline  # noqa: E402

BIN = os.path.join(ROOT, "eqtlbma_b200", "eqtlbma_hm")


def _write_inputs(sc, ds, tmp):
    nfiles = sc.get("files", 1)
    per = (ds.n_genes + nfiles - 1) // nfiles
    for i in range(nfiles):
        ds.write_raw_file(os.path.join(tmp, "in_%d_l10abfs_raw.txt.gz" % i), i * per, min(ds.n_genes, (i + 1) * per))
    init = None
    if "init" in sc:
        init = os.path.join(tmp, "init.txt")
        with open(init, "w") as f:
            f.write(sc["init"])
    return init


def _cells_match(a, b):
    if a == b:
        return True
    try:
        fa, fb = float(a), float(b)
    except ValueError:
        return False
    if fa != fa or fb != fb:
        return (fa != fa) and (fb != fb)
    return abs(fa - fb) <= 2.1e-4 * max(abs(fa), abs(fb)) + 1e-300


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(HM_SCENARIOS))
def test_cli_output_matches_reference_file(cuda_lib, name):
    sc = HM_SCENARIOS[name]
    ds = build_dataset(sc)
    with tempfile.TemporaryDirectory() as tmp:
        init = _write_inputs(sc, ds, tmp)
        out = os.path.join(tmp, "out_hm.txt.gz")
        cmd = [BIN] + ref_cmdline(sc, ds, os.path.join(tmp, "in_*_l10abfs_raw.txt.gz"), out, init) + ["--thread", "4"]
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp)
        assert r.returncode == 0, r.stderr[-2000:]
        got = gzip.open(out, "rt").read().splitlines()
    ref = gzip.open(os.path.join(GOLDEN_HM, name + ".out_hm.txt.gz"), "rt").read().splitlines()
    assert len(got) == len(ref)
    exact = total = 0
    for lg, lr in zip(got, ref):
        cg, cr = lg.split("\t"), lr.split("\t")
        assert len(cg) == len(cr), (lg, lr)
        for a, b in zip(cg, cr):
            assert _cells_match(a, b), (lg, lr)
            exact += a == b
            total += 1
    assert exact >= 0.99 * total, f"only {exact}/{total} cells are byte-identical"
    # the progress lines of the EM are the reference's (show_state_EM)
    assert sum(ln.startswith("iter ") for ln in r.stdout.splitlines()) > 0


@pytest.mark.gpu
def test_cli_ci_only_mode(cuda_lib):
    """--ci: intervals around the estimates of a parameter file (the `#` lines of a previous output with the `#` stripped,
    tests/test_hm.bash:131 of the reference); same intervals as the --getci run of the same scenario."""
    sc = HM_SCENARIOS["getci"]
    ds = build_dataset(sc)
    ref = gzip.open(os.path.join(GOLDEN_HM, "getci.out_hm.txt.gz"), "rt").read().splitlines()
    with tempfile.TemporaryDirectory() as tmp:
        _write_inputs(sc, ds, tmp)
        cif = os.path.join(tmp, "for_ci.txt")
        with open(cif, "w") as f:
            f.write("\n".join(ln[1:] for ln in ref if ln.startswith("#")) + "\n")
        out = os.path.join(tmp, "ci.txt.gz")
        cmd = [BIN, "--data", os.path.join(tmp, "in_*_l10abfs_raw.txt.gz"), "--nsubgrp", "2", "--dim", str(ds.dim), "--ngrid",
               str(ds.grid), "--out", out, "--getci", "--ci", cif]
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp)
        assert r.returncode == 0, r.stderr[-2000:]
        got = gzip.open(out, "rt").read().splitlines()
    assert all(ln.startswith("#") for ln in got) and len(got) == 2 + ds.dim + ds.grid
    for lg, lr in zip(got[1:], [ln for ln in ref if ln.startswith("#")][1:]):
        cg, cr = lg.split("\t"), lr.split("\t")
        assert cg[0] == cr[0]
        # estimates are read back from 4 printed digits: the intervals move by at most a tick or two
        for a, b in zip(cg[1:4], cr[1:4]):
            assert abs(float(a) - float(b)) <= 2.5e-3, (lg, lr)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["basic_all_perm", "basic_all_trick1"])
def test_bf_to_hm_chain_matches_reference_chain(cuda_lib, tmp_path, name):
    """The reference's functional test of the pair of programs (tests/test_hm.bash:117-131): eqtlbma_bf --bfs all writes
    `_l10abfs_raw.txt.gz`, eqtlbma_hm --getbf --getci fits the model on it.  Both drop-in front-ends in a row against the
    reference's two programs in a row (golden: oracle/make_golden_hm.py chain()).  The raw files agree up to the last
    printed digit of some cells, so the 4-digit estimates may differ in their last digit too."""
    from hm_scenarios import chain_cmdline
    from scenarios import SCENARIOS, build_dataset as build_bf_dataset, cli_extra, ref_flags
    sc = SCENARIOS[name]
    ds = build_bf_dataset(sc)
    d = str(tmp_path / "in")
    ds.write_files(d)
    out = str(tmp_path / "obs")
    bf = os.path.join(ROOT, "eqtlbma_b200", "eqtlbma_bf")
    r = subprocess.run([bf] + ds.ref_args(d, out) + ref_flags(sc) + cli_extra(sc, ds, d) + ["-v", "0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    hm_out = str(tmp_path / "obs_hm.txt.gz")
    r = subprocess.run([BIN] + chain_cmdline(name, out + "_l10abfs_raw.txt.gz", hm_out), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    got = gzip.open(hm_out, "rt").read().splitlines()
    ref = gzip.open(os.path.join(GOLDEN_HM, "chain_" + name + ".out_hm.txt.gz"), "rt").read().splitlines()
    assert len(got) == len(ref)
    for lg, lr in zip(got, ref):
        cg, cr = lg.split("\t"), lr.split("\t")
        assert len(cg) == len(cr), (lg, lr)
        for a, b in zip(cg, cr):
            if a == b:
                continue
            fa, fb = float(a), float(b)
            # estimates and Bayes factors to 1e-3 relative; interval ends may move by one 0.001 tick
            assert abs(fa - fb) <= 1e-3 * max(abs(fa), abs(fb)) + 1.001e-3, (lg, lr)


def test_cli_option_checks_and_loud_failure_without_device():
    assert os.path.exists(BIN), "eqtlbma_b200/eqtlbma_hm is not built (python -c 'import __graft_entry__ as g; g.build()')"
    r = subprocess.run([BIN, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--msl" in r.stdout and "--getci" in r.stdout
    r = subprocess.run([BIN, "--nsubgrp", "3"], capture_output=True, text=True)
    assert r.returncode != 0 and "--data" in r.stderr
    r = subprocess.run([BIN, "--data", "x", "--nsubgrp", "3", "--dim", "7", "--ngrid", "10", "--out", "o.gz", "--model", "types"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "types" in r.stderr
    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run([BIN, "--data", os.path.join(tmp, "none*"), "--nsubgrp", "3", "--dim", "7", "--ngrid", "10", "--out",
                            os.path.join(tmp, "o.gz")], capture_output=True, text=True)
        assert r.returncode != 0 and "no input file" in r.stderr
        bad = os.path.join(tmp, "bad_l10abfs_raw.txt.gz")
        with gzip.open(bad, "wt") as f:
            f.write("snp\tgene\tconfig\tl10abf.grid1\n")
        r = subprocess.run([BIN, "--data", bad, "--nsubgrp", "3", "--dim", "7", "--ngrid", "1", "--out", os.path.join(tmp, "o.gz")],
                           capture_output=True, text=True)
        assert r.returncode != 0 and "wrong header" in r.stderr
        import torch
        if not torch.cuda.is_available():
            sc = HM_SCENARIOS["classic_bf"]
            ds = build_dataset(sc)
            _write_inputs(sc, ds, tmp)
            r = subprocess.run([BIN] + ref_cmdline(sc, ds, os.path.join(tmp, "in_*_l10abfs_raw.txt.gz"), os.path.join(tmp, "o.gz"), None),
                               capture_output=True, text=True)
            assert r.returncode != 0 and "no CUDA device" in r.stderr and "finish loading" not in r.stderr
