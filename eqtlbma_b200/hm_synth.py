"""Synthetic inputs of the hierarchical model (eqtlbma_hm): raw log10 Bayes factors per
(gene, SNP, configuration, grid point), in memory as `[pairs][dim][grid]` doubles with gene offsets
(what the C ABI of include/eqtlbma_hm_b200.h takes) and as the `_l10abfs_raw.txt.gz` text of
eqtlbma_bf (header `gene snp config l10abf.grid1 ...`, `%.6e` cells; /root/reference/src/eqtlbma_bf.cpp:1083-1229)
that the reference's loader reads (eqtlbma_hm.cpp:287-371).  Every value is rounded to the text
precision first, so file and array hold identical doubles.

The values come from a small generative model with the structure the EM estimates: a fraction
pi0 of null genes, one causal SNP per non-null gene with a true configuration and grid point,
per-subgroup z-scores, and per-configuration Bayes factors that multiply over the active subgroups."""
from __future__ import annotations

import gzip
import hashlib
import itertools
from dataclasses import dataclass

import numpy as np


def config_names(n_subgroups: int, singletons_only: bool = False) -> list:
    """Configuration names in the order eqtlbma_bf writes them: by size, lexicographic (gsl_combination order,
    gene_snp_pair.cpp:469-485): "1","2","3","1-2","1-3","2-3","1-2-3"."""
    out = []
    for k in range(1, n_subgroups + 1):
        for comb in itertools.combinations(range(1, n_subgroups + 1), k):
            out.append("-".join(str(c) for c in comb))
        if singletons_only:
            break
    return out


@dataclass
class HmDataset:
    gene_names: list
    snp_names: list  # per pair
    gene_off: np.ndarray  # int64 [n_genes + 1]
    cfg_names: list  # [dim]
    B: np.ndarray  # float64 [pairs, dim, grid]
    gen: np.ndarray | None  # float64 [pairs, 3, grid]: "gen", "gen-fix", "gen-maxh" rows of the file (or None)
    n_subgroups: int

    @property
    def dim(self):
        return self.B.shape[1]

    @property
    def grid(self):
        return self.B.shape[2]

    @property
    def n_pairs(self):
        return self.B.shape[0]

    @property
    def n_genes(self):
        return len(self.gene_names)

    def digest(self) -> str:
        h = hashlib.sha256()
        h.update(np.ascontiguousarray(self.B).tobytes())
        h.update(np.ascontiguousarray(self.gene_off).tobytes())
        h.update("|".join(self.gene_names + self.snp_names + self.cfg_names).encode())
        return h.hexdigest()[:16]

    def write_raw_file(self, path: str, gene_lo: int = 0, gene_hi: int | None = None, header: bool = True):
        """One `_l10abfs_raw.txt.gz` file holding genes [gene_lo, gene_hi)."""
        gene_hi = self.n_genes if gene_hi is None else gene_hi
        G = self.grid
        with gzip.open(path, "wt") as f:
            if header:
                f.write("gene\tsnp\tconfig" + "".join(f"\tl10abf.grid{i + 1}" for i in range(G)) + "\n")
            for g in range(gene_lo, gene_hi):
                for p in range(int(self.gene_off[g]), int(self.gene_off[g + 1])):
                    pre = f"{self.gene_names[g]}\t{self.snp_names[p]}\t"
                    if self.gen is not None:
                        for j, nm in enumerate(("gen", "gen-fix", "gen-maxh")):
                            f.write(pre + nm + "".join("\t%.6e" % v for v in self.gen[p, j]) + "\n")
                    for k, nm in enumerate(self.cfg_names):
                        f.write(pre + nm + "".join("\t%.6e" % v for v in self.B[p, k]) + "\n")


def _round_text(a: np.ndarray) -> np.ndarray:
    flat = np.array([float("%.6e" % v) for v in a.ravel()], dtype=np.float64)
    return flat.reshape(a.shape)


def make_hm_dataset(seed: int = 1859, n_genes: int = 150, snps_lo: int = 3, snps_hi: int = 9, n_subgroups: int = 3,
                    grid: int = 10, pi0: float = 0.3, n_ind: int = 100, with_gen: bool = False, strength: float = 1.0,
                    singletons_only: bool = False, round_text: bool = True) -> HmDataset:
    rs = np.random.RandomState(seed)
    S = n_subgroups
    names = config_names(S, singletons_only)
    masks = np.zeros((len(names), S), dtype=bool)
    for k, nm in enumerate(names):
        for t in nm.split("-"):
            masks[k, int(t) - 1] = True
    dim = len(names)
    # prior variances of the grid (geometric, like makeGrid) scaled by the sample size
    W = n_ind * 0.01 * 4.0 ** (np.arange(grid) % 5) * (1.0 + 0.5 * (np.arange(grid) // 5))
    true_cfg_p = rs.dirichlet(np.ones(dim) * 0.7)
    true_grid_p = rs.dirichlet(np.ones(grid))
    m = rs.randint(snps_lo, snps_hi + 1, size=n_genes)
    gene_off = np.concatenate([[0], np.cumsum(m)]).astype(np.int64)
    n_pairs = int(gene_off[-1])
    z = rs.normal(size=(n_pairs, S))
    for g in range(n_genes):
        if rs.uniform() < pi0:
            continue
        p = int(gene_off[g]) + rs.randint(m[g])
        k = rs.choice(dim, p=true_cfg_p)
        l = rs.choice(grid, p=true_grid_p)
        eff = rs.normal(scale=np.sqrt(W[l]) * strength, size=S)
        z[p, masks[k]] += eff[masks[k]]
    # per-subgroup log10 BF at every grid point, summed over the active subgroups of a configuration
    shrink = W / (1.0 + W)  # [grid]
    per = (-0.5 * np.log1p(W)[None, None, :] + 0.5 * (z ** 2)[:, :, None] * shrink[None, None, :]) / np.log(10.0)  # [pairs,S,grid]
    B = np.einsum("ks,psl->pkl", masks.astype(np.float64), per)
    gen = None
    if with_gen:
        full = per.sum(axis=1)  # consistent configuration
        gen = np.stack([full, full * 0.9, full * 1.05], axis=1)
    if round_text:
        B = _round_text(B)
        if gen is not None:
            gen = _round_text(gen)
    gene_names = ["gene%04d" % (g + 1) for g in range(n_genes)]
    snp_names = []
    for g in range(n_genes):
        snp_names += ["snp%d_%d" % (g + 1, j + 1) for j in range(m[g])]
    return HmDataset(gene_names, snp_names, gene_off, names, np.ascontiguousarray(B), gen, S)
