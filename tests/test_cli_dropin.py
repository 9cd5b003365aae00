"""Drop-in check of the C++ front-end eqtlbma_b200/eqtlbma_bf (GPU): same command line and input
files as the reference, gzipped text outputs compared with the reference's own outputs
(tests/golden/<scenario>.text.json.gz, produced by oracle/_ref/eqtlbma_bf_ref_dump).
Text cells must be identical; a numeric cell may differ in its last printed digit (the files carry
7 significant digits; the underlying doubles agree to 1e-9 / 1e-8, see test_cuda_parity.py).
`med.perm.l10abf` is excluded: the reference reads past its vector there (gene.cpp:713-714)."""
import gzip
import json
import os
import subprocess

import pytest

from scenarios import CLI_SCENARIOS, SCENARIOS, build_dataset, cli_extra, ref_flags
from test_oracle_vs_reference import GOLD

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "eqtlbma_b200", "eqtlbma_bf")
NOT_YET = set()
DEGENERATE = {"monomorphic"}
DECODERS = {"vcf_input", "impute_input"}  # run by test_genotype_decoders.py


def cells_match(a, b):
    if a == b:
        return True, True
    try:
        x, y = float(a), float(b)
    except ValueError:
        return False, False
    if x != x and y != y:  # nan vs -nan
        return True, False
    return abs(x - y) <= 2e-6 * max(abs(x), abs(y)) + 1e-300, False


@pytest.mark.parametrize("name", sorted(set(SCENARIOS) - NOT_YET - DEGENERATE - DECODERS))
def test_cli_outputs_match_reference_text(tmp_path, name):
    _run_and_compare(tmp_path, name, threads=1)


@pytest.mark.parametrize("name", ["ragged5", "covariates", "sep_permsep2_trick2"])
def test_cli_parallel_encoder_is_byte_identical_after_gunzip(tmp_path, name):
    """--thread N drives the multi-threaded output encoder (one gzip member per thread slice)."""
    _run_and_compare(tmp_path, name, threads=4)


def _run_and_compare(tmp_path, name, threads):
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "eqtlbma_b200", "host")])
    sc = SCENARIOS.get(name) or CLI_SCENARIOS[name]
    ds = build_dataset(sc)
    d = str(tmp_path / "in")
    ds.write_files(d)
    out = str(tmp_path / "obs")
    cmd = [EXE] + ds.ref_args(d, out) + ref_flags(sc) + cli_extra(sc, ds, d) + ["-v", "0", "--thread", str(threads)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    gold = json.loads(gzip.open(os.path.join(GOLD, name + ".text.json.gz"), "rt").read())
    assert gold, "no golden text"
    n_cells = n_exact = 0
    for fn, exp_txt in gold.items():
        path = out + "_" + fn
        assert os.path.exists(path), f"missing output {fn}"
        got_txt = gzip.open(path, "rt").read()
        exp_lines, got_lines = exp_txt.splitlines(), got_txt.splitlines()
        assert len(exp_lines) == len(got_lines), (fn, len(exp_lines), len(got_lines))
        med_col = None
        for ln, (e, g) in enumerate(zip(exp_lines, got_lines)):
            et, gt = e.split("\t"), g.split("\t")
            assert len(et) == len(gt), (fn, ln, e, g)
            if "med.perm.l10abf" in et:
                med_col = et.index("med.perm.l10abf")
            for ci, (a, b) in enumerate(zip(et, gt)):
                if med_col is not None and ci == med_col and ln > 1:
                    continue
                ok, exact = cells_match(a, b)
                assert ok, (fn, ln, ci, a, b)
                n_cells += 1
                n_exact += exact
    assert n_exact >= 0.995 * n_cells, (n_exact, n_cells)


def _read_outputs(prefix):
    out = {}
    d, base = os.path.dirname(prefix), os.path.basename(prefix)
    for fn in sorted(os.listdir(d)):
        if fn.startswith(base + "_") and fn.endswith(".txt.gz"):
            out[fn[len(base) + 1:]] = gzip.open(os.path.join(d, fn), "rt").read()
    return out


@pytest.mark.parametrize("name", ["basic_all_perm", "sep_permsep2_trick2", "ragged5"])
def test_cli_shards_concatenate_to_the_unsharded_outputs(tmp_path, name):
    """--shard k/N (contiguous cost-balanced ranges of whole write-groups, eqb_partition_by_cost): the shards' outputs,
    concatenated in shard order without the later headers (the `zcat | sed 1d` merge of the reference's recipe,
    doc/manual_eqtlbma.texi:1211-1218), are identical to the outputs of one run -- permutation p-values included, since
    the generator is re-seeded per write-group.  --gpus N does the same through the launcher (one process per shard,
    gzip members concatenated byte-wise)."""
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "eqtlbma_b200", "host")])
    sc = SCENARIOS[name]
    ds = build_dataset(sc)
    d = str(tmp_path / "in")
    ds.write_files(d)

    def run(out, extra):
        cmd = [EXE] + ds.ref_args(d, out) + ref_flags(sc) + cli_extra(sc, ds, d) + ["-v", "0"] + extra
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        return _read_outputs(out)

    full = run(str(tmp_path / "full"), [])
    assert full
    n = 3
    parts = [run(str(tmp_path / f"part{k}"), ["--shard", f"{k}/{n}"]) for k in range(n)]
    for fn, txt in full.items():
        n_head = 2 if "PermPvals" in fn else 1  # the permutation files carry a comment line above the column names
        merged = parts[0][fn]
        for k in range(1, n):
            merged += "".join(parts[k][fn].splitlines(keepends=True)[n_head:])
        assert merged == txt, fn
    launched = run(str(tmp_path / "launched"), ["--gpus", "3"])
    assert launched == full
    assert not [f for f in os.listdir(tmp_path) if ".shard" in f]  # the launcher removed its intermediate files
