"""Seeded parity scenarios shared by the oracle-vs-reference (CPU) and CUDA-vs-oracle (GPU) tests.

Each scenario = synthetic-data parameters (eqtlbma_b200.synth.make_dataset) + eqtlbma_bf options.
They cover the shapes the reference's own functional tests exercise (tests/test_basic.bash,
test_with-covariates.bash, test_genes-absent-in-some-subgroups*.bash, test_mvlr.bash) and the paths
those tests never touch (permutations, --analys sep, --bfs gen|sin, --qnorm, --fiterr 0.5, ragged
individuals under uvlr, more than 3 subgroups)."""
from __future__ import annotations

import hashlib

import numpy as np

from eqtlbma_b200.synth import make_dataset

BASE = dict(seed=1859, n_subgroups=3, n_inds=200, n_genes=10, snps_per_gene=2)

SCENARIOS = {
    # test_basic.bash shape: join, all 7 configs, wrtsize 3 (+ permutations the reference never tests)
    "basic_all_perm": dict(data=dict(BASE), analysis="join", bfs="all", wrtsize=3,
                           perm=dict(nperm=100, pbf="all", seed=1859)),
    "basic_sin_permgensin_maxbf": dict(data=dict(BASE, seed=7), analysis="join", bfs="sin", wrtsize=4,
                                       perm=dict(nperm=60, pbf="gen-sin", seed=42, maxbf=True)),
    "basic_gen_permgen_trick2": dict(data=dict(BASE, seed=11, n_genes=12, null_frac=0.7), analysis="join",
                                     bfs="gen", wrtsize=5,
                                     perm=dict(nperm=200, pbf="gen", seed=3, trick=2, tricut=10)),
    "basic_all_trick1": dict(data=dict(BASE, seed=12, n_genes=12, null_frac=0.7), analysis="join", bfs="all",
                             wrtsize=5, perm=dict(nperm=150, pbf="gen", seed=5, trick=1, tricut=5)),
    # test_with-covariates.bash shape
    "covariates": dict(data=dict(BASE, seed=21, n_cov=3, dosage=True), analysis="join", bfs="sin", wrtsize=10,
                       perm=dict(nperm=40, pbf="gen", seed=9)),
    # genes absent in some subgroups (+ individual NaNs): test_genes-absent-in-some-subgroups*.bash
    "absent_genes_nan": dict(data=dict(BASE, seed=31, absent_gene_frac=0.3, nan_frac=0.01), analysis="join",
                             bfs="all", wrtsize=3, perm=dict(nperm=50, pbf="all", seed=77)),
    # ragged individuals under uvlr, 5 subgroups, SNPs missing from some genotype files
    "ragged5": dict(data=dict(BASE, seed=41, n_subgroups=5, n_inds=120, ragged=True, snps_per_gene=4,
                              separate_geno_files=True, missing_geno_frac=0.15),
                    analysis="join", bfs="all", wrtsize=4, perm=dict(nperm=50, pbf="all", seed=1859)),
    # separate analysis, both permutation flavours
    "sep_permsep1": dict(data=dict(BASE, seed=51, n_cov=2, ragged=True), analysis="sep", bfs="gen", wrtsize=3,
                         perm=dict(nperm=80, permsep=1, seed=13)),
    "sep_permsep2_trick2": dict(data=dict(BASE, seed=52, absent_gene_frac=0.2), analysis="sep", bfs="gen",
                                wrtsize=4, perm=dict(nperm=80, permsep=2, seed=14, trick=2, tricut=3)),
    # --trick 1 with --permsep 2: the generator is re-seeded per subgroup and a gene's stopping point is its own in
    # every subgroup (eqtlbma_bf.cpp:806-822, gene.cpp:380-450)
    "sep_permsep2_trick1": dict(data=dict(BASE, seed=53, n_genes=12, null_frac=0.7, absent_gene_frac=0.1), analysis="sep",
                                bfs="gen", wrtsize=4, perm=dict(nperm=120, permsep=2, seed=15, trick=1, tricut=3)),
    # c4 slices at the permutation count of the GTEx target (10^4): 9 ragged tissues, exceedance counts exact
    "c4_slice_gensin_10k_trick2": dict(data=dict(seed=1901, n_subgroups=9, n_inds=60, n_genes=3, snps_per_gene=5, ragged=True,
                                                 ragged_min_frac=0.5, null_frac=0.5), analysis="join", bfs="sin",
                                       wrtsize=10, perm=dict(nperm=10000, pbf="gen-sin", seed=1859, trick=2, tricut=10)),
    "c4_slice_all_10k": dict(data=dict(seed=1902, n_subgroups=9, n_inds=60, n_genes=2, snps_per_gene=3, ragged=True,
                                       ragged_min_frac=0.5, null_frac=0.5), analysis="join", bfs="all", wrtsize=10,
                             perm=dict(nperm=10000, pbf="all", seed=1859)),
    # almost deterministic eQTLs: Student tail below the double range -> Phi^-1(0) = -inf -> NaN / infinite standardised
    # statistics -> ABF 0 through the reference's own guards
    "huge_t": dict(data=dict(BASE, seed=91, n_genes=8, perfect_frac=0.5), analysis="join", bfs="all", wrtsize=4,
                   perm=dict(nperm=20, pbf="all", seed=4)),
    # --qnorm
    "qnorm": dict(data=dict(BASE, seed=61, n_inds=60, ragged=True), analysis="join", bfs="sin", wrtsize=10,
                  qnorm=True, perm=dict(nperm=30, pbf="gen-sin", seed=21)),
    # MVLR (test_mvlr.bash: --fiterr 0.0) and the default --fiterr 0.5 the reference never pins
    "mvlr_fit0": dict(data=dict(BASE, seed=71, n_inds=100, n_genes=6), analysis="join", bfs="all", wrtsize=3,
                      error="mvlr", fiterr=0.0, perm=dict(nperm=20, pbf="all", seed=5)),
    "mvlr_fit05_cov": dict(data=dict(BASE, seed=72, n_inds=100, n_genes=6, n_cov=2), analysis="join", bfs="all",
                           wrtsize=3, error="mvlr", fiterr=0.5, perm=dict(nperm=20, pbf="gen-sin", seed=6)),
    # --error hybrid (gene_snp_pair.cpp:760-1423; test_common-uniq-inds.bash: --fiterr 0.0): individuals common to /
    # unique to each pair of subgroups, expression levels missing per gene, and the default --fiterr 0.5 with covariates
    "hybrid_fit0": dict(data=dict(BASE, seed=73, n_inds=100, n_genes=6, ragged=True, ragged_min_frac=0.6, nan_frac=0.03),
                        analysis="join", bfs="all", wrtsize=3, error="hybrid", fiterr=0.0,
                        perm=dict(nperm=20, pbf="all", seed=7)),
    "hybrid_fit05_cov": dict(data=dict(BASE, seed=74, n_inds=100, n_genes=6, n_cov=2, ragged=True, ragged_min_frac=0.6,
                                       pad_names=True, n_subgroups=4), analysis="join", bfs="sin", wrtsize=3,
                             error="hybrid", fiterr=0.5, perm=dict(nperm=20, pbf="gen-sin", seed=9)),
    # hybrid with --qnorm (diagonals only: the off-diagonal blocks read the raw expression levels), genes absent from
    # some subgroups (skipped: eqtlbma_bf.cpp:747-762), 2 subgroups, --maxbf permutations on the general ABF
    "hybrid_qnorm_maxbf": dict(data=dict(BASE, seed=75, n_subgroups=2, n_inds=80, n_genes=8, snps_per_gene=3, ragged=True,
                                         ragged_min_frac=0.7, absent_gene_frac=0.2), analysis="join", bfs="gen", wrtsize=4,
                               qnorm=True, error="hybrid", fiterr=0.3, perm=dict(nperm=30, pbf="gen", seed=11, maxbf=True)),
    # degenerate inputs: monomorphic SNPs (rank-deficient designs)
    "monomorphic": dict(data=dict(BASE, seed=81, monomorphic_frac=0.3, snps_per_gene=4), analysis="join",
                        bfs="all", wrtsize=3),
    # genotype decoders other than the custom dose matrix (data_loader.cpp:570-1010, snp.cpp:118-185): no --scoord,
    # coordinates come from the genotype files themselves
    "vcf_input": dict(data=dict(BASE, seed=101, snps_per_gene=3, geno_format="vcf", separate_geno_files=True,
                                missing_geno_frac=0.1), analysis="join", bfs="all", wrtsize=4,
                      perm=dict(nperm=30, pbf="all", seed=8)),
    "impute_input": dict(data=dict(BASE, seed=102, snps_per_gene=3, n_cov=2, geno_format="impute"), analysis="join",
                         bfs="sin", wrtsize=10),
    # TSS+TES anchor
    "anchor_tss_tes": dict(data=dict(BASE, seed=91, anchor="TSS+TES", radius=150, snps_per_gene=3),
                           analysis="join", bfs="gen", wrtsize=10),
}

# Front-end filters (command line only: they change what the loader hands to the hot path, so these have the
# reference's text outputs as golden but no engine-level dump test): --sbgrp, --maf, --snp
# (--outm is declared by the reference's option table but never handled: eqtlbma_bf.cpp:275 falls through to the help text)
CLI_SCENARIOS = {
    "cli_sbgrp_maf": dict(data=dict(BASE, seed=111, snps_per_gene=4, maf=0.25), analysis="join", bfs="sin",
                          wrtsize=10, extra=["--sbgrp", "s1+s3", "--maf", "0.24"]),
    "cli_snp_list": dict(data=dict(BASE, seed=112, snps_per_gene=4), analysis="sep", bfs="gen", wrtsize=10,
                         snp_list_every=2),
}


def cli_extra(sc, ds, d):
    """Extra command-line flags of a CLI-only scenario; writes the --snp list next to the input files."""
    f = list(sc.get("extra", []))
    k = sc.get("snp_list_every")
    if k:
        path = f"{d}/snps_to_keep.txt"
        with open(path, "w") as fh:
            fh.write("\n".join(ds.snp_names[::k]) + "\n")
        f += ["--snp", path]
    return f


def build_dataset(sc):
    return make_dataset(**sc["data"])


def dataset_digest(ds) -> str:
    h = hashlib.sha256()
    for G in ds.genos:
        h.update(np.ascontiguousarray(np.nan_to_num(G, nan=-9.0)).tobytes())
    for sg in ds.subgroups:
        h.update(np.ascontiguousarray(np.nan_to_num(sg.Y, nan=-9.0)).tobytes())
        h.update(np.ascontiguousarray(sg.C).tobytes())
        h.update(sg.all2exp.tobytes())
        h.update(sg.gene_has_exp.tobytes())
        h.update(sg.snp_has_geno.tobytes())
    h.update(ds.snp_pos.tobytes())
    h.update(ds.gene_start.tobytes())
    return h.hexdigest()[:16]


def ref_flags(sc):
    """eqtlbma_bf command-line flags of a scenario (beyond the input files)."""
    f = ["--analys", sc["analysis"], "--bfs", sc["bfs"], "--wrtsize", str(sc.get("wrtsize", 10)), "--outss", "--outw"]
    if sc.get("error", "uvlr") != "uvlr":
        f += ["--error", sc["error"], "--fiterr", repr(float(sc.get("fiterr", 0.5)))]
    if sc.get("qnorm"):
        f += ["--qnorm"]
    p = sc.get("perm")
    if p:
        f += ["--nperm", str(p["nperm"]), "--seed", str(p["seed"])]
        if p.get("trick"):
            f += ["--trick", str(p["trick"]), "--tricut", str(p.get("tricut", 10))]
        if sc["analysis"] == "join":
            f += ["--pbf", p["pbf"]]
            if p.get("maxbf"):
                f += ["--maxbf"]
        else:
            f += ["--permsep", str(p["permsep"])]
    return f


def engine_kwargs(sc):
    # the reference holds --fiterr in a float (eqtlbma_bf.cpp:245,398): 0.3 on the command line is 0.300000011920929 in
    # every formula; the engines take the double the front-end hands them (eqtlbma_bf_main.cpp keeps a float as well)
    return dict(analysis=sc["analysis"], bfs=sc["bfs"], error=sc.get("error", "uvlr"),
                fiterr=float(np.float32(sc.get("fiterr", 0.5))), qnorm=bool(sc.get("qnorm")))


def perm_kwargs(sc):
    p = sc.get("perm")
    if not p:
        return None
    return dict(nperm=p["nperm"], seed=p["seed"], trick=p.get("trick", 0), tricut=p.get("tricut", 10),
                permsep=p.get("permsep", 0), pbf=p.get("pbf", "none"), maxbf=p.get("maxbf", False),
                wrtsize=sc.get("wrtsize", 10))
