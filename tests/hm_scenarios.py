"""Scenarios of the hierarchical-model EM (eqtlbma_hm, `--model configs`) shared by oracle/make_golden_hm.py
(which runs the unmodified reference on them) and the parity tests.  Each scenario = arguments of
eqtlbma_b200.hm_synth.make_hm_dataset + the eqtlbma_hm options (/root/reference/src/eqtlbma_hm.cpp:1797-2051)."""
from __future__ import annotations

import gzip
import os

import numpy as np

from eqtlbma_b200.hm_synth import make_hm_dataset

GOLDEN_HM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hm")

HM_SCENARIOS = {
    # classical EM, default initialisation, posteriors and Bayes factors
    "classic_bf": dict(data=dict(seed=11, n_genes=150), getbf=True),
    # SQUAREM (--msl 3)
    "squarem_bf": dict(data=dict(seed=12, n_genes=200, snps_hi=12), msl=3.0, getbf=True, thresh=0.01),
    # --pi0 fixed on the command line
    "fixed_pi0": dict(data=dict(seed=13, n_genes=120), pi0=0.6, getbf=True),
    # initialisation file with the grid weights fixed
    "init_fixed_grid": dict(data=dict(seed=14, n_genes=100, grid=4), getbf=True,
                            init="param\tvalue\tfixed\npi0\t0.4\tFALSE\ngrid.1\t0.1\tTRUE\ngrid.2\t0.2\tTRUE\n"
                                 "grid.3\t0.3\tTRUE\ngrid.4\t0.4\tTRUE\n"
                                 + "".join("config.%d\t%g\tFALSE\n" % (k + 1, v) for k, v in
                                           enumerate([0.05, 0.05, 0.1, 0.1, 0.2, 0.2, 0.3]))),
    # subset of configurations (--configs, --dim 4), three input files
    "configs_subset": dict(data=dict(seed=15, n_genes=90), configs="1|2|3|1-2-3", dim=4, files=3, getbf=True),
    # BMAlite-style: only the `gen` rows (--keepgen --configs gen --dim 1)
    "keepgen_dim1": dict(data=dict(seed=16, n_genes=80, with_gen=True), configs="gen", dim=1, keepgen=True, getbf=True),
    # iteration cap
    "maxit5": dict(data=dict(seed=17, n_genes=100, strength=2.0), maxit=5, thresh=1e-6),
    # profile-likelihood confidence intervals (single-threaded and slow in the reference: small case)
    "getci": dict(data=dict(seed=18, n_genes=60, n_subgroups=2, grid=3), getci=True, getbf=True),
    # random initialisation (--seed: gsl_rng_mt19937 + gsl_ran_exponential, eqtlbma_hm.cpp:529-579)
    "rand_seed": dict(data=dict(seed=20, n_genes=100), seed=1859, getbf=True),
    # SQUAREM with 4 subgroups (15 configurations) and strong effects
    "squarem_s4": dict(data=dict(seed=19, n_genes=120, n_subgroups=4, grid=6, strength=2.5), msl=4.0, getbf=True),
}


# bf -> hm chain (the reference's tests/test_hm.bash): scenarios of tests/scenarios.py whose raw ABFs are all finite
CHAIN_SCENARIOS = {"basic_all_perm": dict(nsubgrp=3, dim=7, ngrid=10), "basic_all_trick1": dict(nsubgrp=3, dim=7, ngrid=10)}


def chain_cmdline(name, raw_file, out):
    c = CHAIN_SCENARIOS[name]
    return ["--data", raw_file, "--nsubgrp", str(c["nsubgrp"]), "--dim", str(c["dim"]), "--ngrid", str(c["ngrid"]),
            "--out", out, "--getbf", "--getci", "-v", "1"]


def build_dataset(sc):
    return make_hm_dataset(**sc["data"])


def kept_configs(sc, ds):
    """Indices (into ds.cfg_names, or -1 for the `gen` row) of the rows the loader keeps, in file order
    (eqtlbma_hm.cpp:315-332)."""
    keep = sc.get("configs")
    keep = keep.split("|") if keep else None
    rows = []
    if ds.gen is not None and sc.get("keepgen"):
        if keep is None or "gen" in keep:
            rows.append(-1)
    for k, nm in enumerate(ds.cfg_names):
        if keep is None or nm in keep:
            rows.append(k)
    return rows


def model_arrays(sc, ds):
    """B [pairs][dim][grid] as the reference's loader would hold it, and the configuration names."""
    rows = kept_configs(sc, ds)
    parts, names = [], []
    for k in rows:
        if k < 0:
            parts.append(ds.gen[:, 0, :])
            names.append("gen")
        else:
            parts.append(ds.B[:, k, :])
            names.append(ds.cfg_names[k])
    return np.ascontiguousarray(np.stack(parts, axis=1)), names


def ref_cmdline(sc, ds, pattern, out, init_path):
    dim = sc.get("dim", ds.dim)
    cmd = ["--data", pattern, "--nsubgrp", str(ds.n_subgroups), "--dim", str(dim), "--ngrid", str(ds.grid),
           "--out", out, "-v", "1", "--thresh", repr(sc.get("thresh", 0.05))]
    if "msl" in sc:
        cmd += ["--msl", repr(sc["msl"])]
    if "maxit" in sc:
        cmd += ["--maxit", str(sc["maxit"])]
    if "pi0" in sc:
        cmd += ["--pi0", repr(sc["pi0"])]
    if "configs" in sc:
        cmd += ["--configs", sc["configs"]]
    if sc.get("keepgen"):
        cmd += ["--keepgen"]
    if sc.get("getbf"):
        cmd += ["--getbf"]
    if sc.get("getci"):
        cmd += ["--getci"]
    if "seed" in sc:
        cmd += ["--seed", str(sc["seed"])]
    if init_path:
        cmd += ["--init", init_path]
    return cmd


def initial_params(sc, dim, grid):
    """Controller::init_params (eqtlbma_hm.cpp:452-613) for the scenarios above: (pi0, grid_wts, config_prior, fixed)."""
    fixed = dict(pi0=False, grid=False, configs=False)
    pi0 = 0.5
    gw = np.full(grid, 1.0 / grid)
    cp = np.full(dim, 1.0 / dim)
    if "pi0" in sc:
        pi0 = sc["pi0"]
        fixed["pi0"] = True
    if "seed" in sc:
        # gsl_rng_mt19937 seeded by gsl_rng_set = numpy's MT19937 seeded with the same integer; gsl_rng_uniform = raw 32 bits / 2^32
        rs = np.random.RandomState(sc["seed"])
        u = lambda: float(rs.randint(0, 2 ** 32, dtype=np.uint64)) / 4294967296.0  # noqa: E731
        if not fixed["pi0"]:
            pi0 = u()
        gw = np.array([-np.log1p(-u()) for _ in range(grid)])
        gw = gw / gw.sum()
        if dim > 1:
            cp = np.array([-np.log1p(-u()) for _ in range(dim)])
            cp = cp / cp.sum()
    if "init" in sc:
        ig = ic = 0
        for ln in sc["init"].splitlines():
            t = ln.split("\t")
            if ln.startswith("#") or (t[0] == "param" and t[1] == "value"):
                continue
            fx = len(t) == 3 and t[2] in ("TRUE", "true")
            if "pi0" in t[0]:
                pi0 = float(t[1])
                fixed["pi0"] |= fx
            elif "grid" in t[0]:
                gw[ig] = float(t[1])
                ig += 1
                fixed["grid"] |= fx
            elif "config" in t[0]:
                cp[ic] = float(t[1])
                ic += 1
                fixed["configs"] |= fx
    return pi0, gw, cp, fixed


def load_hm_dump(name):
    """Parse tests/golden/hm/<name>.dump.gz (written by oracle/ref_hm_dump_main.cpp)."""
    out = {"genes": [], "snps": [], "postcfg": []}
    with gzip.open(os.path.join(GOLDEN_HM, name + ".dump.gz"), "rt") as f:
        for ln in f:
            t = ln.rstrip("\n").split("\t")
            tag = t[0]
            if tag == "SHAPE":
                out["shape"] = tuple(int(x) for x in t[1:])
            elif tag == "LOGLIK":
                out["loglik"] = float(t[1])
            elif tag == "PI0":
                out["pi0"] = [float(x) for x in t[1:]]
            elif tag in ("CONFIG", "CONFIG_LEFT", "CONFIG_RIGHT", "GRID", "GRID_LEFT", "GRID_RIGHT"):
                out[tag.lower()] = np.array([float(x) for x in t[1:]])
            elif tag == "NAMES":
                out["names"] = t[1:]
            elif tag == "GENE":
                out["genes"].append((t[1], int(t[2]), float(t[3]), float(t[4])))
            elif tag == "SNP":
                out["snps"].append((t[1], float(t[2]), float(t[3]), [float(x) for x in t[4:]]))
            elif tag == "POSTCFG":
                out["postcfg"].append([float(x) for x in t[1:]])
    return out
