"""Synthetic eqtlbma_bf inputs (simul_flutre_et_al / functional_tests.R style).

Builds, from one seed, (1) the in-memory column layouts the C-ABI consumes (what the C++ host
produces after parsing) and (2) the reference's own input files (App. C of SURVEY.md: list files,
MatrixEQTL-like dose / expression / covariate matrices, BED coordinates, grids), so that the very
same data can be pushed through the reference binary, the CPU oracle and the CUDA library.

Shapes follow /root/reference/tests/functional_tests.R:239-436 (S subgroups, HWE genotypes at
MAF 0.3, ES-model effects) and src/simul_flutre_et_al.cpp:545-743; orderings follow the
reference loader: subgroups sorted by id, samples = sorted union, genes in
byte-wise name order, SNPs per chromosome in position order, covariates in name order.
Every number is rounded to the text precision first, so files and arrays hold identical doubles.
"""
from __future__ import annotations

import gzip
import os
from dataclasses import dataclass, field

import numpy as np


def make_grid(kind: str = "general") -> np.ndarray:
    """(phi2, oma2) grid of functional_tests.R:180-201 (getGrid): 25 points (general) or 10 (small)."""
    aes = [0.1 ** 2, 0.2 ** 2, 0.4 ** 2, 0.8 ** 2, 1.6 ** 2]
    homs = [0.0, 0.25, 0.5, 0.75, 1.0] if kind == "general" else [0.75, 1.0]
    rows = []
    for a in aes:
        for h in homs:
            rows.append((a * (1 - h), a * h))
    return np.array(rows, dtype=np.float64)


@dataclass
class Subgroup:
    name: str
    geno_id: int
    all2geno: np.ndarray  # int32 [N_all], -1 = absent
    snp_has_geno: np.ndarray  # uint8 [M]
    Y: np.ndarray  # float64 [n_genes, n_exp_cols], NaN = missing value
    all2exp: np.ndarray  # int32 [N_all]
    gene_has_exp: np.ndarray  # uint8 [n_genes]
    C: np.ndarray  # float64 [Q, n_cov_cols]
    all2cov: np.ndarray  # int32 [N_all]
    exp_samples: list = field(default_factory=list)
    cov_names: list = field(default_factory=list)


@dataclass
class Dataset:
    samples: list  # sorted union of sample names
    subgroups: list  # list[Subgroup], sorted by name
    genos: list  # list of float64 [M, n_geno_cols] (SNP-major), one per distinct genotype file
    geno_samples: list  # per geno matrix: sample names in file order
    snp_names: list  # global SNP order: chromosomes in name order, position order inside
    snp_chr: np.ndarray  # int32 [M] index into chr_names
    snp_pos: np.ndarray  # int64 [M]
    snp_bed_start: np.ndarray
    gene_names: list  # byte-wise name order
    gene_chr: np.ndarray  # int32 [n_genes]
    gene_start: np.ndarray  # int64, 1-based (BED start + 1)
    gene_end: np.ndarray  # int64
    chr_names: list
    gridL: np.ndarray
    gridS: np.ndarray
    anchor: str = "TSS"
    radius: int = 100000
    maf: np.ndarray | None = None  # [n_genos][M] folded MAF as the reference computes it
    geno_format: str = "custom"  # "custom" (dose matrix + --scoord BED), "vcf" or "impute" (hard calls only)

    @property
    def n_all(self):
        return len(self.samples)

    @property
    def n_snps(self):
        return len(self.snp_names)

    @property
    def n_genes(self):
        return len(self.gene_names)

    def cis_windows(self):
        """[begin,end) SNP index range per gene, restating Snp::IsInCis (snp.cpp:274-297) +
        Gene::SetCisSnps (gene.cpp:140-157) with integer arithmetic."""
        beg = np.zeros(self.n_genes, dtype=np.int64)
        end = np.zeros(self.n_genes, dtype=np.int64)
        for c in np.unique(self.snp_chr):
            idx = np.nonzero(self.snp_chr == c)[0]
            lo_chr, hi_chr = int(idx[0]), int(idx[-1]) + 1
            pos = self.snp_pos[lo_chr:hi_chr]
            gs = np.nonzero(self.gene_chr == c)[0]
            if len(gs) == 0:
                continue
            start, endc = self.gene_start[gs], self.gene_end[gs]
            lo = np.where(start >= self.radius, start - self.radius, 0)
            hi = (start if self.anchor == "TSS" else endc) + self.radius
            b = np.searchsorted(pos, lo, side="left")
            e = np.searchsorted(pos, hi, side="right")
            beg[gs], end[gs] = lo_chr + b, lo_chr + np.maximum(b, e)
        return beg, end

    # ------------------------------------------------------------------ files
    def write_files(self, d: str, with_covariates: bool | None = None):
        os.makedirs(d, exist_ok=True)

        def w(path, lines, gz=None):
            gz = path.endswith(".gz") if gz is None else gz
            op = gzip.open if gz else open
            with op(os.path.join(d, path), "wt") as fh:
                fh.write("\n".join(lines) + "\n")

        def fmt(v):
            return "NA" if np.isnan(v) else repr(float(v))

        # coordinates (BED): genes chr start(0-based) end name score strand; SNPs chr start end name
        w("gene_coords.bed.gz", [
            f"{self.chr_names[self.gene_chr[g]]}\t{int(self.gene_start[g]) - 1}\t{int(self.gene_end[g])}\t{self.gene_names[g]}\t1000\t+"
            for g in range(self.n_genes)])
        w("snp_coords.bed.gz", [
            f"{self.chr_names[self.snp_chr[m]]}\t{int(self.snp_pos[m]) - 1}\t{int(self.snp_pos[m])}\t{self.snp_names[m]}"
            for m in range(self.n_snps)])
        for gi, G in enumerate(self.genos):
            if self.geno_format == "vcf":
                # Snp::AddSubgroupFromVcfLine (snp.cpp:118-150): dosage = number of "1" alleles of GT
                lines = ["##fileformat=VCFv4.1", "##source=eqtlbma_b200.synth",
                         "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(self.geno_samples[gi])]
                gts = (("0|0", "0|1", "1|1"), ("0/0", "1/0", "1/1"))
                for m in range(self.n_snps):
                    if np.all(np.isnan(G[m])):
                        continue
                    extra = m % 2  # every other line carries a second FORMAT field before GT
                    cells = [("7:" if extra else "") + gts[(m + i) % 2][int(v)] for i, v in enumerate(G[m])]
                    lines.append(f"{self.chr_names[self.snp_chr[m]]}\t{int(self.snp_pos[m])}\t{self.snp_names[m]}\tA\tG\t.\tPASS\t.\t"
                                 + ("DP:GT" if extra else "GT") + "\t" + "\t".join(cells))
                w(f"genotypes_{gi}.txt.gz", lines)
                continue
            if self.geno_format == "impute":
                # Snp::AddSubgroupFromImputeLine (snp.cpp:152-185): dosage = P(a1a2) + 2 P(a2a2)
                lines = ["chr name coord a1 a2 " + " ".join(f"{s}_a1a1 {s}_a1a2 {s}_a2a2" for s in self.geno_samples[gi])]
                trip = ("1 0 0", "0 1 0", "0 0 1")
                for m in range(self.n_snps):
                    if np.all(np.isnan(G[m])):
                        continue
                    lines.append(f"{self.chr_names[self.snp_chr[m]]} {self.snp_names[m]} {int(self.snp_pos[m])} A G "
                                 + " ".join(trip[int(v)] for v in G[m]))
                w(f"genotypes_{gi}.txt.gz", lines)
                continue
            lines = ["id\t" + "\t".join(self.geno_samples[gi])]
            for m in range(self.n_snps):
                if np.all(np.isnan(G[m])):
                    continue  # SNP absent from this genotype file
                lines.append(self.snp_names[m] + "\t" + "\t".join(fmt(v) for v in G[m]))
            w(f"genotypes_{gi}.txt.gz", lines)
        w("list_genotypes.txt", [f"{sg.name} {d}/genotypes_{sg.geno_id}.txt.gz" for sg in self.subgroups])
        for sg in self.subgroups:
            lines = ["id\t" + "\t".join(sg.exp_samples)]
            for g in range(self.n_genes):
                if sg.gene_has_exp[g]:
                    lines.append(self.gene_names[g] + "\t" + "\t".join(fmt(v) for v in sg.Y[g]))
            w(f"phenotypes_{sg.name}.txt.gz", lines)
        w("list_phenotypes.txt", [f"{sg.name} {d}/phenotypes_{sg.name}.txt.gz" for sg in self.subgroups])
        has_cov = any(sg.C.shape[0] > 0 for sg in self.subgroups)
        if has_cov:
            cl = []
            for sg in self.subgroups:
                if sg.C.shape[0] == 0:
                    continue
                cov_samples = [None] * sg.C.shape[1]
                for i, c in enumerate(sg.all2cov):
                    if c >= 0:
                        cov_samples[c] = self.samples[i]
                lines = ["id\t" + "\t".join(cov_samples)]
                for q, nm in enumerate(sg.cov_names):
                    lines.append(nm + "\t" + "\t".join(fmt(v) for v in sg.C[q]))
                w(f"covariates_{sg.name}.txt.gz", lines)
                cl.append(f"{sg.name} {d}/covariates_{sg.name}.txt.gz")
            w("list_covariates.txt", cl)
        w("grid_phi2_oma2_general.txt.gz", [f"{repr(float(a))}\t{repr(float(b))}" for a, b in self.gridL])
        w("grid_phi2_oma2_with-configs.txt.gz", [f"{repr(float(a))}\t{repr(float(b))}" for a, b in self.gridS])
        return d

    def ref_args(self, d: str, out_prefix: str):
        """Command-line arguments of the reference eqtlbma_bf for the files of write_files()."""
        a = ["--geno", f"{d}/list_genotypes.txt"]
        if self.geno_format == "custom":
            a += ["--scoord", f"{d}/snp_coords.bed.gz"]
        a += ["--exp", f"{d}/list_phenotypes.txt", "--gcoord", f"{d}/gene_coords.bed.gz",
             "--anchor", self.anchor, "--cis", str(self.radius), "--out", out_prefix,
             "--gridL", f"{d}/grid_phi2_oma2_general.txt.gz",
             "--gridS", f"{d}/grid_phi2_oma2_with-configs.txt.gz"]
        if any(sg.C.shape[0] > 0 for sg in self.subgroups):
            a += ["--covar", f"{d}/list_covariates.txt"]
        return a


def _round(a, nd=6):
    return np.round(np.asarray(a, dtype=np.float64), nd)


def make_dataset(seed=1859, n_subgroups=3, n_inds=200, n_genes=10, snps_per_gene=2, n_chr=2,
                 n_cov=0, cov_per_subgroup=False, ragged=False, ragged_min_frac=0.4, absent_gene_frac=0.0, nan_frac=0.0,
                 dosage=False, maf=0.3, gridL=None, gridS=None, radius=None, anchor="TSS",
                 null_frac=0.3, separate_geno_files=False, missing_geno_frac=0.0,
                 pad_names=False, monomorphic_frac=0.0, gene_spacing=1000, far_snp=True,
                 geno_format="custom", layout_only=False, perfect_frac=0.0) -> Dataset:
    """Generate a dataset. One SNP stream per chromosome at uniform spacing; each gene's +-radius
    TSS window holds ~snps_per_gene SNPs; expression y = mu_s + b_s*g + N(0,1) with ES-model
    effects from the first cis SNP of the gene (simul_flutre_et_al.cpp:682-743)."""
    rng = np.random.default_rng(seed)
    S = n_subgroups
    gridL = make_grid("general") if gridL is None else np.asarray(gridL, dtype=np.float64)
    gridS = make_grid("small") if gridS is None else np.asarray(gridS, dtype=np.float64)
    width = len(str(max(n_inds, n_genes, n_genes * snps_per_gene + 8)))

    def nm(prefix, i):
        return f"{prefix}{i:0{width}d}" if pad_names else f"{prefix}{i}"

    ind_names = [nm("ind", i + 1) for i in range(n_inds)]
    samples = sorted(ind_names)
    # (zero-padded with pad_names: the byte-wise chromosome order of the loader then equals the numeric one, so genes that
    # are contiguous in name order have contiguous genotype rows)
    chr_names_unsorted = [f"chr{c + 1:0{len(str(n_chr))}d}" if pad_names else f"chr{c + 1}" for c in range(n_chr)]
    chr_names = sorted(chr_names_unsorted)

    # genes: equally spread over chromosomes, spacing 1000 bp, length 200
    spacing, glen = gene_spacing, 200
    if radius is None:
        radius = 100  # window 2*radius+1
    snp_step = max(1, (2 * radius + 1) // max(1, snps_per_gene))
    genes = []
    per_chr = int(np.ceil(n_genes / n_chr))
    for g in range(n_genes):
        c = g // per_chr
        k = g % per_chr
        start1 = 1000 + k * spacing + 1  # 1-based start
        genes.append((nm("gene", g + 1), chr_names_unsorted[c], start1, start1 + glen))
    # SNPs: uniform ladder over each chromosome's gene span (+ one SNP far away with no gene)
    snps = []
    sid = 0
    for c in range(n_chr):
        gs = [g for g in genes if g[1] == chr_names_unsorted[c]]
        if not gs:
            continue
        lo = max(1, min(g[2] for g in gs) - radius)
        hi = max(g[2] for g in gs) + radius
        pos = lo + snp_step // 2
        while pos <= hi:
            sid += 1
            snps.append((nm("snp", sid), chr_names_unsorted[c], pos))
            pos += snp_step
        if far_snp:
            sid += 1
            snps.append((nm("snp", sid), chr_names_unsorted[c], hi + 50 * radius + 7))
    # orderings of the reference loader
    genes.sort(key=lambda t: t[0].encode())
    snps.sort(key=lambda t: (chr_names.index(t[1]), t[2], t[0].encode()))
    M = len(snps)
    snp_names = [t[0] for t in snps]
    snp_chr = np.array([chr_names.index(t[1]) for t in snps], dtype=np.int32)
    snp_pos = np.array([t[2] for t in snps], dtype=np.int64)
    gene_names = [t[0] for t in genes]
    gene_chr = np.array([chr_names.index(t[1]) for t in genes], dtype=np.int32)
    gene_start = np.array([t[2] for t in genes], dtype=np.int64)
    gene_end = np.array([t[3] for t in genes], dtype=np.int64)

    if layout_only:
        # coordinates only (cis-window sizes for a cost partition): no genotypes, no expression levels
        return Dataset(samples=samples, subgroups=[], genos=[], geno_samples=[], snp_names=snp_names, snp_chr=snp_chr,
                       snp_pos=snp_pos, snp_bed_start=snp_pos - 1, gene_names=gene_names, gene_chr=gene_chr,
                       gene_start=gene_start, gene_end=gene_end, chr_names=chr_names, gridL=gridL, gridS=gridS,
                       anchor=anchor, radius=radius, geno_format=geno_format)
    # genotypes (file column order = ind1..indN, i.e. NOT the sorted order)
    p = np.array([(1 - maf) ** 2, 2 * maf * (1 - maf), maf ** 2])
    G = rng.choice(3, size=(M, n_inds), p=p).astype(np.float64)
    if dosage:
        G = np.clip(G + rng.normal(0, 0.08, size=G.shape), 0.0, 2.0)
        G = _round(G, 3)
    if monomorphic_frac > 0:
        mono = rng.random(M) < monomorphic_frac
        G[mono] = 0.0
    n_files = S if separate_geno_files else 1
    genos, geno_samples = [], []
    for gi in range(n_files):
        Gi = G.copy()
        if missing_geno_frac > 0 and gi > 0:
            drop = rng.random(M) < missing_geno_frac
            Gi[drop] = np.nan  # SNP absent from this file
        genos.append(Gi)
        geno_samples.append(list(ind_names))
    maf_arr = np.zeros((n_files, M))
    for gi in range(n_files):
        with np.errstate(invalid="ignore"):
            f = genos[gi].sum(axis=1) / (2 * n_inds)
        maf_arr[gi] = np.where(f <= 0.5, f, 1 - f)

    ds = Dataset(samples=samples, subgroups=[], genos=genos, geno_samples=geno_samples,
                 snp_names=snp_names, snp_chr=snp_chr, snp_pos=snp_pos,
                 snp_bed_start=snp_pos - 1, gene_names=gene_names, gene_chr=gene_chr,
                 gene_start=gene_start, gene_end=gene_end, chr_names=chr_names, gridL=gridL,
                 gridS=gridS, anchor=anchor, radius=radius, maf=maf_arr, geno_format=geno_format)
    beg, end = ds.cis_windows()

    # effects
    pos_in_sorted = {n: i for i, n in enumerate(samples)}
    ind_pos = {n: i for i, n in enumerate(ind_names)}
    all2geno = np.array([ind_pos[s] for s in samples], dtype=np.int32)
    mus = rng.normal(4, 2, size=S)
    cov_names = sorted([f"cov{q + 1}" for q in range(max(0, n_cov - 1))] + (["sex"] if n_cov > 0 else []))
    Cfull = np.zeros((n_cov, n_inds))
    for q, cn in enumerate(cov_names):
        Cfull[q] = rng.integers(0, 2, n_inds) if cn == "sex" else _round(rng.normal(0, 1, n_inds), 5)
    Cshared = Cfull
    for s in range(S):
        name = f"s{s + 1}"
        gi = s if separate_geno_files else 0
        if cov_per_subgroup and n_cov > 0 and s > 0:
            # tissue-specific covariates (PEER-like factors); "sex" stays an attribute of the individual
            Cfull = Cshared.copy()
            for q, cn in enumerate(cov_names):
                if cn != "sex":
                    Cfull[q] = _round(rng.normal(0, 1, n_inds), 5)
        if ragged:
            n_s = int(rng.integers(int(ragged_min_frac * n_inds), n_inds + 1))
            cols = np.sort(rng.choice(n_inds, size=n_s, replace=False))
        else:
            cols = np.arange(n_inds)
        exp_samples = [ind_names[c] for c in cols]
        Y = np.zeros((len(genes), len(cols)))
        for g in range(len(genes)):
            y = mus[s] + rng.normal(0, 1, n_inds)
            if end[g] > beg[g] and rng.random() > null_frac:
                m = int(beg[g])
                pve = rng.uniform(0.1, 0.4)
                tot = pve / ((1 - pve) * 2 * maf * (1 - maf))
                het = rng.uniform(0, 0.2)
                bbar = rng.normal(0, np.sqrt(tot * (1 - het)))
                b = rng.normal(bbar, np.sqrt(tot * het))
                y = y + b * np.nan_to_num(G[m])
            if n_cov > 0:
                y = y + 0.5 * Cfull[0]
            if perfect_frac > 0 and end[g] > beg[g] and rng.random() < perfect_frac:
                # almost deterministic eQTL: |t| of several hundreds, the Student tail underflows
                y = mus[s] + np.nan_to_num(G[int(beg[g])]) + rng.normal(0, 0.004, n_inds)
            Y[g] = y[cols]
        Y = _round(Y, 6)
        if nan_frac > 0:
            Y[rng.random(Y.shape) < nan_frac] = np.nan
        gene_has_exp = np.ones(len(genes), dtype=np.uint8)
        if absent_gene_frac > 0 and s > 0:
            gene_has_exp[rng.random(len(genes)) < absent_gene_frac] = 0
        all2exp = np.full(len(samples), -1, dtype=np.int32)
        for j, nmx in enumerate(exp_samples):
            all2exp[pos_in_sorted[nmx]] = j
        snp_has = (~np.all(np.isnan(genos[gi]), axis=1)).astype(np.uint8)
        if n_cov > 0:
            # covariates for ALL genotyped individuals, file order = ind order
            C = Cfull.copy()
            all2cov = all2geno.copy()
        else:
            C = np.zeros((0, 0))
            all2cov = np.full(len(samples), -1, dtype=np.int32)
        ds.subgroups.append(Subgroup(name=name, geno_id=gi, all2geno=all2geno.copy(),
                                     snp_has_geno=snp_has, Y=Y, all2exp=all2exp,
                                     gene_has_exp=gene_has_exp, C=C, all2cov=all2cov,
                                     exp_samples=exp_samples, cov_names=list(cov_names)))
    return ds
