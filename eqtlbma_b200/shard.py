"""Gene sharding over the GPUs of one box: contiguous ranges of whole write-groups,
balanced on cis-window cost; no collective on the data path, a final host gather of the per-gene /
per-pair results in gene order (replaces scripts/eqtlbma_bf_parallel.bash:248-345, one OS process per gene batch).

  partition()       eqb_partition_by_cost of the C ABI
  slice_dataset()   the inputs one shard needs: its genes and the genotype rows their cis windows touch
  concat_datasets() several gene-contiguous pieces (same individuals, grids, covariates) as one dataset
"""
from __future__ import annotations

import copy
import ctypes

import numpy as np


def gene_costs(cis_begin, cis_end, nperm=0):
    """cost model: cis SNPs x (1 + permutations) pair evaluations per gene"""
    return ((np.asarray(cis_end) - np.asarray(cis_begin)).astype(np.int64) * (1 + int(nperm))).astype(np.int64)


def partition(lib: ctypes.CDLL, costs, wrtsize: int, n_shards: int) -> np.ndarray:
    costs = np.ascontiguousarray(costs, dtype=np.int64)
    out = np.zeros(n_shards + 1, dtype=np.int64)
    f = lib.eqb_partition_by_cost
    f.restype = ctypes.c_int
    rc = f(costs.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(costs)), ctypes.c_int64(wrtsize),
           ctypes.c_int32(n_shards), out.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError("eqb_partition_by_cost failed")
    return out


def slice_dataset(ds, g_lo: int, g_hi: int):
    """Genes [g_lo, g_hi) of ds and the SNP rows of their cis windows (genes are in name order, their windows are
    index ranges of the position-sorted SNP table, so both slices are contiguous)."""
    beg, end = ds.cis_windows()
    keep = np.arange(g_lo, g_hi)
    has = end[keep] > beg[keep]
    m_lo = int(beg[keep][has].min()) if has.any() else 0
    m_hi = int(end[keep][has].max()) if has.any() else 0
    sub = copy.copy(ds)
    sub.genos = [G[m_lo:m_hi] for G in ds.genos]
    sub.snp_names = ds.snp_names[m_lo:m_hi]
    sub.snp_chr = ds.snp_chr[m_lo:m_hi]
    sub.snp_pos = ds.snp_pos[m_lo:m_hi]
    sub.snp_bed_start = ds.snp_bed_start[m_lo:m_hi]
    sub.gene_names = ds.gene_names[g_lo:g_hi]
    sub.gene_chr = ds.gene_chr[g_lo:g_hi]
    sub.gene_start = ds.gene_start[g_lo:g_hi]
    sub.gene_end = ds.gene_end[g_lo:g_hi]
    if ds.maf is not None:
        sub.maf = ds.maf[:, m_lo:m_hi]
    sub.subgroups = []
    for sg in ds.subgroups:
        s2 = copy.copy(sg)
        s2.Y = sg.Y[g_lo:g_hi]
        s2.gene_has_exp = sg.gene_has_exp[g_lo:g_hi]
        s2.snp_has_geno = sg.snp_has_geno[m_lo:m_hi]
        sub.subgroups.append(s2)
    return sub


def concat_datasets(parts, tags=None):
    """Gene-contiguous pieces with the same individuals / sample maps / covariates / grids as ONE dataset: genes and
    SNPs are appended in order, the chromosomes of piece k are renamed <tag_k>_<chr> so that they stay distinct."""
    if len(parts) == 1 and tags is None:
        return parts[0]
    tags = tags or [f"b{k}" for k in range(len(parts))]
    out = copy.copy(parts[0])
    out.chr_names, out.snp_names, out.gene_names = [], [], []
    snp_chr, gene_chr = [], []
    for k, p in enumerate(parts):
        off = len(out.chr_names)
        out.chr_names += [f"{tags[k]}_{c}" for c in p.chr_names]
        out.snp_names += [f"{tags[k]}_{n}" for n in p.snp_names]
        out.gene_names += [f"{tags[k]}_{n}" for n in p.gene_names]
        snp_chr.append(np.asarray(p.snp_chr) + off)
        gene_chr.append(np.asarray(p.gene_chr) + off)
    out.snp_chr = np.concatenate(snp_chr).astype(np.int32)
    out.gene_chr = np.concatenate(gene_chr).astype(np.int32)
    out.snp_pos = np.concatenate([p.snp_pos for p in parts])
    out.snp_bed_start = np.concatenate([p.snp_bed_start for p in parts])
    out.gene_start = np.concatenate([p.gene_start for p in parts])
    out.gene_end = np.concatenate([p.gene_end for p in parts])
    out.genos = [np.concatenate([p.genos[i] for p in parts], axis=0) for i in range(len(parts[0].genos))]
    out.maf = None if parts[0].maf is None else np.concatenate([p.maf for p in parts], axis=1)
    out.subgroups = []
    for si, sg in enumerate(parts[0].subgroups):
        s2 = copy.copy(sg)
        s2.Y = np.concatenate([p.subgroups[si].Y for p in parts], axis=0)
        s2.gene_has_exp = np.concatenate([p.subgroups[si].gene_has_exp for p in parts])
        s2.snp_has_geno = np.concatenate([p.subgroups[si].snp_has_geno for p in parts])
        out.subgroups.append(s2)
    return out
