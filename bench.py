#!/usr/bin/env python
"""bench.py -- throughput of the eqtlbma_bf hot path on B200: cis gene-SNP pair BFs/sec and permuted pairs/sec.

Workload (BASELINE.json configs[1], "c2"): 3 subgroups x 300 individuals, 11 covariates (subgroup-specific values),
dosage genotypes shared by the subgroups, 5000 genes x ~50 cis SNPs per GPU, `--analys join --bfs sin`.
ONE dataset of world x 5000 genes is laid out in blocks of 5000 genes whose cis windows hold 36..64 SNPs (skewed window
sizes); the genes are cut into `world` contiguous shards of whole write-groups balanced on cis-window cost by
eqb_partition_by_cost (the C ABI's partitioner), rank k builds and runs shard k (weak scaling: 250k pairs per GPU), no
collective on the data path; the shards' results are what a final host gather concatenates in shard order
(scripts/eqtlbma_bf_parallel.bash:248-345 of the reference).  A "step" = one pass of the hot path over the rank's shard.

  value        pairs/s with inputs resident in HBM (CUDA events around the kernels, on the library's stream)
  e2e          pairs/s through the C ABI with HOST buffers: eqb_create + H2D of genotypes / expression / covariates +
               eqb_finalize + eqb_run + D2H of every result the reference's writers need in join mode (sample sizes,
               summary statistics, raw per-grid-point ABFs, grid-averaged ABFs), every step; `e2e_avg_only` is the same
               without the raw ABFs (a caller that only needs the averaged ABFs: 5x fewer bytes down)
  perm         BASELINE.json's second metric, permuted pairs/s, on the c4 shape (9 ragged tissues of 450 individuals,
               ~5000 cis SNPs per gene, 2047 permutations, --pbf all and gen-sin) with an FP64 roofline against the
               DFMA / DMMA peaks measured in the same run, and the reference's permutation loop (--thread nproc) on a slice
  parity       the first genes of the bench workload are diffed against the reference's own full-precision results
  --impl reference   times the reference's own CPU eqtlbma_bf (oracle/_ref, built from the unmodified sources against the
               GSL shim) the way the reference scales on a multi-core host: one single-threaded process per gene batch.
"""
from __future__ import annotations

import argparse
import ctypes
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# (zero-padded names: the byte-wise gene order of the reference = the positional order, so a contiguous gene range has a
# contiguous genotype-row range -- what a shard uploads)
C2 = dict(n_subgroups=3, n_inds=300, n_genes=5000, n_cov=11, cov_per_subgroup=True, dosage=True,
          radius=100, gene_spacing=201, far_snp=False, n_chr=22, pad_names=True)
BLOCK_SPG = [50, 36, 64, 44, 58, 40, 62, 46]  # cis SNPs per gene in block b (mean 50): skewed window sizes
WRTSIZE = 10
WORKLOAD_DESC = ("c2: S=3 x N=300, Q=11 covariates (subgroup-specific values), dosage genotypes, 5000 genes x ~50 cis SNPs per GPU "
                 "(one dataset of n_gpus blocks with 36..64 SNPs per window, cut by eqb_partition_by_cost), "
                 "--analys join --bfs sin, gridL 25 / gridS 10, no permutations")
C4 = dict(n_subgroups=9, n_inds=450, n_genes=8, snps_per_gene=5000, ragged=True, ragged_min_frac=0.34, radius=10000,
          gene_spacing=20001, far_snp=False, n_chr=2)
C4_NPERM = 2047
C4_DESC = "c4 shape: S=9 ragged tissues (150-450 of 450 individuals), 8 genes x ~4400-5000 cis SNPs per GPU, gridL 10 / gridS 10"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)  # never let nvidia-smi run into the next timed region
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ workload
def canonical_covariates(n_cov, n_inds, n_subgroups):
    """Covariates belong to the individuals, not to a block of genes: one fixed set for the whole dataset."""
    rng = np.random.default_rng(4242)
    names = sorted([f"cov{q + 1}" for q in range(n_cov - 1)] + ["sex"])
    sex = rng.integers(0, 2, n_inds).astype(np.float64)
    out = []
    for _ in range(n_subgroups):
        C = np.round(rng.normal(0, 1, (n_cov, n_inds)), 5)
        C[names.index("sex")] = sex
        out.append(C)
    return out


def block_kwargs(b, n_genes=None):
    kw = dict(C2, seed=1859 + b, snps_per_gene=BLOCK_SPG[b % len(BLOCK_SPG)])
    if n_genes:
        kw["n_genes"] = n_genes
    return kw


def make_shard(rank, world, lib, genes_per_block=None):
    """The global dataset is block-major (block b = genes [b * G, (b + 1) * G) of the gene order, on its own
    chromosomes); only the blocks that intersect the rank's shard are generated."""
    from eqtlbma_b200.shard import concat_datasets, partition, slice_dataset
    from eqtlbma_b200.synth import make_dataset
    G = genes_per_block or C2["n_genes"]
    costs = []
    for b in range(world):
        lay = make_dataset(layout_only=True, **block_kwargs(b, G))
        beg, end = lay.cis_windows()
        costs.append((end - beg).astype(np.int64))
    costs = np.concatenate(costs)
    sb = partition(lib, costs, WRTSIZE, world)
    lo, hi = int(sb[rank]), int(sb[rank + 1])
    covs = canonical_covariates(C2["n_cov"], C2["n_inds"], C2["n_subgroups"])
    parts, tags = [], []
    for b in range(lo // G, (max(hi, lo + 1) - 1) // G + 1):
        blk = make_dataset(**block_kwargs(b, G))
        for s, sg in enumerate(blk.subgroups):
            sg.C = covs[s]
        l0, l1 = max(lo, b * G) - b * G, min(hi, (b + 1) * G) - b * G
        parts.append(slice_dataset(blk, l0, l1) if (l0, l1) != (0, G) else blk)
        tags.append(f"b{b}")
    ds = concat_datasets(parts, tags) if world > 1 or len(parts) > 1 else parts[0]
    return ds, dict(shard_begin=[int(x) for x in sb], genes=[lo, hi], total_cost=int(costs.sum()),
                    shard_cost=int(costs[lo:hi].sum()))


def algorithmic_bytes(ds, eng, raw: bool):
    """8*N_s*M_g bytes of genotypes per (gene, subgroup) + 8*N_s per gene for y,
    6*8 B of summary statistics per (pair, subgroup), and the weighted (and raw) ABFs written."""
    S = len(ds.subgroups)
    mg = (eng.cis_end - eng.cis_begin).astype(np.float64)
    pairs = float(mg.sum())
    b = 0.0
    for sg in ds.subgroups:
        n_s = float((sg.all2exp >= 0).sum())
        b += 8.0 * n_s * pairs + 8.0 * n_s * ds.n_genes
    b += pairs * S * 48.0
    b += pairs * (5 + eng.n_configs) * 8.0
    if raw:
        b += pairs * (3 * eng.L + eng.n_configs * eng.K) * 8.0
    return b


def pinned_copy(ds):
    """Place the big host inputs in pinned memory (the C ABI copies asynchronously from it)."""
    import torch
    ds._pinned = []
    for i, G in enumerate(ds.genos):
        t = torch.from_numpy(np.ascontiguousarray(np.nan_to_num(G, nan=0.0))).pin_memory()
        ds.genos[i] = t.numpy()
        ds._pinned.append(t)
    for sg in ds.subgroups:
        t = torch.from_numpy(np.ascontiguousarray(sg.Y)).pin_memory()
        sg.Y = t.numpy()
        ds._pinned.append(t)
    ds._clean = True
    return ds


def fixed_point_copy(ds):
    """What the front-end's text parser hands to the ABI for a dosage file written with 3 decimals
    (eqtlbma_bf_main.cpp, eqb_set_genotypes_fixed): u16 numerators of 1000 in pinned memory.  Lossless:
    k / 1000 (IEEE division) is checked here to reproduce every double of the matrix bit for bit."""
    import copy
    import torch
    d2 = copy.copy(ds)
    d2.genos, d2._pinned, d2.geno_denoms = [], list(getattr(ds, "_pinned", [])), []
    for G in ds.genos:
        k = np.rint(G * 1000.0)
        if not (np.all(k >= 0) and np.all(k <= 65535) and np.array_equal(k / 1000.0, G)):
            return None  # not representable: the caller keeps the double matrix
        t = torch.from_numpy(k.astype(np.uint16)).pin_memory()
        d2.genos.append(t.numpy())
        d2._pinned.append(t)
        d2.geno_denoms.append(1000.0)
    return d2


def h2d_bytes(ds):
    b = sum(G.nbytes for G in ds.genos)
    for sg in ds.subgroups:
        b += sg.Y.nbytes + sg.C.nbytes + 3 * 4 * ds.n_all + ds.n_snps + ds.n_genes
    b += (ds.n_genes * 2 + ds.n_snps) * 8 + (ds.n_genes + ds.n_snps) * 4
    return int(b)


def result_digest(r):
    h = hashlib.sha256()
    for a in (r.n, r.sstats, r.abf_gen, r.abf_cfg, r.abf_w):
        if a is not None:
            h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


# ------------------------------------------------------------------------------------------------ permutation block
def perm_flop_model(S, n_all, Q, L, K, pbf):
    """contraction 2 * N_all * (Q + 3) flop per (SNP, subgroup, permutation) (ragged form: the kept rows
    change with every permutation, so the mask and squared-genotype columns are part of the product); BF: (6 * mean
    configuration size + 40) flop-equivalents per closed-form evaluation; the permutation statistic needs the L gen-row
    evaluations (size S), plus S*K singleton evaluations (gen-sin) or (2^S - 1) * K configuration evaluations (all)."""
    contraction = 2.0 * n_all * (Q + 3) * S
    if pbf == "gen":
        bf = L * (6.0 * S + 40.0)
    elif pbf == "gen-sin":
        bf = L * (6.0 * S + 40.0) + S * K * (6.0 + 40.0)
    else:
        C = 2 ** S - 1
        bf = C * K * (6.0 * (S * 2 ** (S - 1) / C) + 40.0)
    return contraction, bf


def run_perm_block(eqtlbma_b200, rank, local_rank, fp64, with_cpu, n_genes=None, nperm=None):
    from eqtlbma_b200.synth import make_dataset, make_grid
    C4_NPERM = nperm or globals()["C4_NPERM"]
    n_genes = n_genes or C4["n_genes"]
    pds = make_dataset(**dict(C4, n_genes=n_genes, seed=1860 + rank, gridL=make_grid("general")[:10]))
    out = {"workload": C4_DESC.replace("8 genes", f"{n_genes} genes"), "nperm": C4_NPERM, "runs": {}}
    # the c3 true pass on the same shape: --bfs all (511 configurations x 10 grid points), with and without the raw ABFs
    eng = eqtlbma_b200.Engine(pds, analysis="join", bfs="all", device=local_rank)
    pairs = int(eng.pair_offsets()[-1])
    tp = {"pairs": pairs, "configurations": int(eng.n_configs), "kernels": "fast_pair_warp_kernel (first pass) + fast_pair_all_kernel"}
    for raw in (False, True):
        for _ in range(3):
            ms = eng.run_device_only(raw=raw)
        kms = eng.last_pair_kernel_ms()
        key = "with_raw_abfs" if raw else "averaged_only"
        tp[key] = {"pairs_per_s": pairs / (ms * 1e-3), "ms_per_step": ms, "pair_kernels_ms": kms}
        if raw:
            raw_bytes = pairs * (3 * eng.L + eng.n_configs * eng.K) * 8
            tp[key]["raw_emission_gbs"] = raw_bytes / (kms * 1e-3) / 1e9
    eng.close()
    out["true_pass_bfs_all"] = tp
    for pbf in ("all", "gen-sin"):
        eng = eqtlbma_b200.Engine(pds, analysis="join", bfs="all" if pbf == "all" else "sin", device=local_rank)
        pairs = int(eng.pair_offsets()[-1])
        eng.set_perm_timing(False)
        eng.run_permutations_device_only(C4_NPERM, 1859, pbf=pbf, wrtsize=WRTSIZE)  # warm-up (also builds the shuffle tables)
        reps = 2 if pairs * C4_NPERM < 5e8 else 1
        ms = [eng.run_permutations_device_only(C4_NPERM, 1859, pbf=pbf, wrtsize=WRTSIZE) for _ in range(reps)]
        eng.set_perm_timing(True)
        eng.run_permutations_device_only(C4_NPERM, 1859, pbf=pbf, wrtsize=WRTSIZE)
        tm = eng.last_perm_timing()
        eng.close()
        ms = float(np.mean(ms))
        evals = pairs * (C4_NPERM + 1)  # the true data (identity) runs through the same kernels
        contraction, bf = perm_flop_model(9, pds.n_all, 0, 10, len(pds.gridS), pbf)
        r = {"permuted_pairs_per_s": evals / (ms * 1e-3), "pairs": pairs, "ms": ms,
             "kernel_ms": {k: tm[k] for k in ("prep_ms", "gemm_ms", "bf_ms", "merge_ms")},
             "flop_per_pair_perm": {"contraction": contraction, "bf_equiv": bf}}
        if fp64:
            gemm_tf = contraction * evals / (tm["gemm_ms"] * 1e-3) / 1e12 if tm["gemm_ms"] else None
            step_tf = (contraction + bf) * evals / (ms * 1e-3) / 1e12
            r["roofline"] = {"bound": "tensor", "unit": "TFLOP/s", "kernel": "perm_gemm_kernel (FP64 mma.sync DMMA, TMA-fed)",
                             "achieved": gemm_tf, "peak": fp64["dmma_tflops"], "frac": gemm_tf / fp64["dmma_tflops"] if gemm_tf else None,
                             "peak_kind": "measured in this run (eqb_measure_fp64_peaks)", "traffic": None,
                             "kernel_ms": tm["gemm_ms"], "issued_tflops": tm["gemm_flops"] / (tm["gemm_ms"] * 1e-3) / 1e12,
                             "step_achieved": step_tf, "step_frac": step_tf / fp64["dfma_tflops"],
                             "step_note": "contraction flop + BF flop-equivalents of SURVEY 8(d) over the whole permutation step "
                                          "(prep + GEMM + BF + merge) against the measured DFMA peak"}
        out["runs"][pbf] = r
    if with_cpu:
        out["cpu_baseline"] = cpu_baseline_perm(eqtlbma_b200, local_rank)
    return out


def run_hm_block(eqtlbma_b200, local_rank, hbm_gbs, with_cpu, n_genes=10000, snps=(50, 150), rank=0, world=1, dist=None):
    """the EM of the hierarchical model (eqtlbma_hm --model configs) on raw ABFs resident in HBM.
    Dominant kernel hm_estep_kernel: one streaming pass over B[pairs][7][10] per fixed-point iteration (algorithmic bytes =
    8 * pairs * dim * grid, read once), HBM-bound.  CPU arm: the unmodified reference eqtlbma_hm (oracle/_ref) with
    --thread = host cores on the first genes of the same data written as `_l10abfs_raw.txt.gz`, its own `EM ran for`
    clock (whole seconds) falling back to wall clock minus a --maxit 2 run of the same files.
    N > 1 (weak scaling): every rank holds its own 10,000 genes of ONE model; the sums over genes of every evaluation are
    all-gathered over NCCL (eqb_hm_set_collective: 2 + dim + grid doubles per rank and evaluation) -- the one real exchange
    step of this repo's paths."""
    import time
    from eqtlbma_b200.hm import HmEngine, HmFit
    from eqtlbma_b200.hm_synth import make_hm_dataset
    ds = make_hm_dataset(seed=1861 + rank, n_genes=n_genes, snps_lo=snps[0], snps_hi=snps[1], n_subgroups=3, grid=10, round_text=False)
    t0 = time.perf_counter()
    hm = HmEngine(ds.dim, ds.grid, device=local_rank)
    hm.append(ds.B, ds.gene_off)
    hm.finalize()
    t_load = time.perf_counter() - t0
    if world > 1:
        hm.set_collective()
    gw, cp = np.full(ds.grid, 1.0 / ds.grid), np.full(ds.dim, 1.0 / ds.dim)
    hm.estep_device_only(gw, cp, reps=3)
    ms = hm.estep_device_only(gw, cp, reps=20)
    alg = 8.0 * ds.n_pairs * ds.dim * ds.grid
    out = {"workload": f"eqtlbma_hm --model configs: {ds.n_genes} genes, {ds.n_pairs} pairs, dim {ds.dim}, grid {ds.grid} "
                       f"({alg / 1e6:.0f} MB of raw log10 ABFs, larger than L2)",
           "pairs": int(ds.n_pairs), "h2d_s": t_load,
           "roofline": {"bound": "hbm", "kernel": "hm_estep_kernel", "kernel_ms": ms, "algorithmic_bytes_per_launch": alg,
                        "achieved": alg / (ms * 1e-3) / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                        "frac": alg / (ms * 1e-3) / 1e9 / hbm_gbs, "traffic": None}}
    tot_pairs = float(ds.n_pairs)
    if world > 1:
        import torch
        c = torch.tensor([tot_pairs], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        tot_pairs = float(c[0])
        out["n_gpus"], out["pairs_all_ranks"] = world, tot_pairs
        out["collective"] = ("hm_xchg_kernel: %d doubles per rank and evaluation stored into every peer's buffer over NVLink (CUDA IPC "
                             "peer memory), epoch flags, combined in rank order on the device; NCCL only carries the IPC handles and "
                             "this block's timing reductions" % (2 + ds.dim + ds.grid))
    for label, msl in (("classic", 1.0), ("squarem", 3.0)):
        l0 = hm.launch_count
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        fit = hm.em(HmFit(0.5, gw, cp), thresh=0.05, stepmax=msl)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        n_lines = len([ln for ln in fit.log_lines if ln.startswith("iter ")])
        out[label] = {"em_s": dt, "likelihood_evaluations": n_lines, "pairs_iterations_per_s": tot_pairs * n_lines / dt,
                      "loglik": fit.loglik, "pi0": fit.pi0, "gpu_launches": hm.launch_count - l0}
    hm.close()
    if with_cpu:
        out["cpu_baseline"] = cpu_baseline_hm(ds)
    return out


def run_hybrid_block(eqtlbma_b200, local_rank, with_cpu, n_genes=200, ref_genes=6):
    """--error hybrid on a slice of 3 ragged subgroups of up to 450 individuals, ~50 cis SNPs
    per gene: device-timed true pass (--bfs sin) and permutation pass (--pbf gen-sin), inputs resident in HBM; the unmodified
    reference binary on the first genes of the same workload beside it (one core: its true pass is single-threaded)."""
    from eqtlbma_b200.synth import make_dataset
    ds = make_dataset(seed=11, n_subgroups=3, n_inds=450, n_genes=n_genes, snps_per_gene=50, ragged=True, ragged_min_frac=0.6,
                      radius=1000, gene_spacing=2001, far_snp=False, n_chr=2)
    eng = eqtlbma_b200.Engine(ds, analysis="join", bfs="sin", error="hybrid", fiterr=0.5, device=local_rank)
    pairs = int(eng.pair_offsets()[-1])
    l0 = eng.launch_count()
    ms = min(eng.run_device_only(raw=True) for _ in range(5))
    launches = (eng.launch_count() - l0) // 5
    nperm = 100
    pms = min(eng.run_permutations_device_only(nperm, 7, pbf="gen-sin", wrtsize=n_genes) for _ in range(2))
    eng.close()
    out = {"workload": f"--error hybrid --bfs sin: 3 ragged subgroups of <= 450 individuals, {n_genes} genes x ~50 cis SNPs "
                       f"({pairs} pairs); permutations: {nperm} per gene, --pbf gen-sin",
           "value": pairs / ms * 1e3, "unit": "pairs/s", "ms_per_pass": ms, "gpu_launches_per_pass": int(launches),
           "kernels": "hybrid_offdiag_kernel + hybrid_kernel", "timing": "device-timed, inputs resident in HBM, best of 5",
           "perm": {"value": pairs * nperm / pms * 1e3, "unit": "pair-permutations/s", "ms": pms}}
    if with_cpu:
        sub = subset_dataset(ds, ref_genes)
        r = time_reference_binary(sub, ["--analys", "join", "--bfs", "sin", "--error", "hybrid", "--fiterr", "0.5", "--outss", "--outw"])
        if r and r.get("pairs"):
            out["cpu_baseline"] = {"value": r["pairs"] / r["seconds"], "unit": "pairs/s", "cores": 1, "kind": "reference",
                                   "sample": f"first {ref_genes} genes ({r['pairs']} pairs) through oracle/_ref/eqtlbma_bf_ref, "
                                             f"association loop {r['seconds']:.2f} s"}
    return out


def cpu_baseline_hm(ds, sample_genes=1000):
    import re
    import shutil
    import tempfile
    import time
    ref = os.path.join(ROOT, "oracle", "_ref", "eqtlbma_hm_ref")
    if not os.path.exists(ref):
        return {"unavailable": "oracle/_ref/eqtlbma_hm_ref is not built"}
    cores = os.cpu_count() or 1
    tmp = tempfile.mkdtemp(prefix="hm_cpu_")
    try:
        # the text precision of the file is the reference's own input format; names are needed only here
        ds.write_raw_file(os.path.join(tmp, "s_l10abfs_raw.txt.gz"), 0, sample_genes)
        pairs = int(ds.gene_off[sample_genes])
        base = [ref, "--data", os.path.join(tmp, "s_l10abfs_raw.txt.gz"), "--nsubgrp", "3", "--dim", str(ds.dim), "--ngrid",
                str(ds.grid), "--out", os.path.join(tmp, "o.txt.gz"), "--thread", str(cores), "-v", "1"]
        t0 = time.perf_counter()
        r = subprocess.run(base, capture_output=True, text=True, cwd=tmp)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"unavailable": "reference eqtlbma_hm failed: " + r.stderr[-200:]}
        n_lines = len([ln for ln in r.stdout.splitlines() if ln.startswith("iter ")])
        t0 = time.perf_counter()
        short = list(base)
        short[short.index(os.path.join(tmp, "o.txt.gz"))] = os.path.join(tmp, "o_short.txt.gz")
        subprocess.run(short + ["--maxit", "2"], capture_output=True, text=True, cwd=tmp)
        wall1 = time.perf_counter() - t0
        r1_lines = 3  # iteration 0, one fixed point, the closing fixed point
        em_s = max(wall - wall1, 1e-3)
        # the drop-in front-end on the same file (wall clock, CUDA context creation included)
        cli = None
        ours = os.path.join(ROOT, "eqtlbma_b200", "eqtlbma_hm")
        if os.path.exists(ours):
            cmd = [ours] + base[1:]
            cmd[cmd.index(os.path.join(tmp, "o.txt.gz"))] = os.path.join(tmp, "o2.txt.gz")
            t0 = time.perf_counter()
            r2 = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp)
            wall2 = time.perf_counter() - t0
            if r2.returncode == 0:
                import gzip
                a = gzip.open(os.path.join(tmp, "o.txt.gz"), "rt").read().splitlines()
                b = gzip.open(os.path.join(tmp, "o2.txt.gz"), "rt").read().splitlines()
                em2 = [ln for ln in r2.stdout.splitlines() if ln.startswith("EM ran for")]
                worst = 0.0
                for la, lb in zip(a[1:], b[1:]):
                    for x, y in zip(la.split("\t")[1:2], lb.split("\t")[1:2]):
                        worst = max(worst, abs(float(x) - float(y)) / max(abs(float(x)), 1e-300))
                cli = {"ours_wall_s": wall2, "reference_wall_s": wall, "wall_ratio": wall / wall2, "same_parameter_lines": a == b,
                       "worst_rel_diff_of_printed_estimates": worst if len(a) == len(b) else None,
                       "ours_em": em2[0] if em2 else None,
                       "what": "eqtlbma_b200/eqtlbma_hm vs oracle/_ref/eqtlbma_hm_ref on the same `_l10abfs_raw.txt.gz` file, "
                               "classical EM, parameter lines of the two output files compared as text"}
            else:
                cli = {"error": r2.stderr[-300:]}
        return {"value": pairs * (n_lines - r1_lines) / em_s, "unit": "pairs x likelihood evaluations/s", "cores": cores,
                "kind": "reference", "sample": f"first {sample_genes} genes ({pairs} pairs), classical EM to --thresh 0.05, "
                f"{n_lines} likelihood evaluations, wall {wall:.2f} s minus {wall1:.2f} s of a --maxit 2 run (file parsing + 3 evaluations)",
                "wall_s": wall, "load_s": wall1, "cli_e2e": cli}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cpu_baseline_perm(eqtlbma_b200=None, device=0):
    """The reference's permutation loop (gene.cpp:598-717; OpenMP over SNPs, --thread = host cores) on a bounded slice
    of the c4 shape: run with permutations minus the same run without.  The same slices go through
    oracle/_ref/eqtlbma_bf_ref_dump and through the CUDA path: true statistic, number of permutations and p-value (i.e. the
    exceedance count) of every gene must agree ('parity' of each entry)."""
    from eqtlbma_b200.synth import make_dataset, make_grid
    threads = os.cpu_count() or 1
    res = {}
    for pbf, n_genes, spg, nperm in (("all", 2, 150, 40), ("gen-sin", 3, 400, 60)):
        ds = make_dataset(**dict(C4, seed=1861, n_genes=n_genes, snps_per_gene=spg, radius=1000, gene_spacing=2001,
                                 gridL=make_grid("general")[:10]))
        bfs = "all" if pbf == "all" else "sin"
        base = ["--analys", "join", "--bfs", bfs, "--outw"]
        pflags = ["--nperm", str(nperm), "--seed", "1859", "--pbf", pbf]
        t1 = time_reference_binary(ds, base + pflags, threads=threads, raw_wall=True)
        t0 = time_reference_binary(ds, base, threads=threads, raw_wall=True)
        if not t1 or not t0:
            return None
        dt = max(t1["t_full"] - t0["t_full"], 1e-9)
        res[pbf] = {"value": t1["pairs"] * nperm / dt, "unit": "permuted pairs/s", "cores": threads, "kind": "reference",
                    "sample": f"{n_genes} genes x ~{spg} cis SNPs of the c4 shape ({t1['pairs']} pairs), --nperm {nperm} --pbf {pbf} "
                              f"--thread {threads} through oracle/_ref/eqtlbma_bf_ref: {t1['t_full']:.1f} s with permutations - "
                              f"{t0['t_full']:.1f} s without"}
        exe = os.path.join(ROOT, "oracle", "_ref", "eqtlbma_bf_ref_dump")
        if eqtlbma_b200 is not None and os.path.exists(exe):
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from refdump import parse_dump
            tmp = tempfile.mkdtemp(prefix="eqb_ppar_")
            try:
                ds.write_files(tmp)
                dump = os.path.join(tmp, "dump.txt")
                cmd = [exe] + ds.ref_args(tmp, os.path.join(tmp, "obs")) + base + pflags + ["--thread", str(threads), "-v", "0"]
                r = subprocess.run(cmd, env=dict(os.environ, EQTLBMA_DUMP=dump), capture_output=True, text=True)
                d = parse_dump(dump) if r.returncode == 0 else None
            finally:
                shutil.rmtree(tmp, ignore_errors=True)
            if d:
                eng = eqtlbma_b200.Engine(ds, analysis="join", bfs=bfs, device=device)
                eng.run()
                pr = eng.run_permutations(nperm=nperm, seed=1859, pbf=pbf, wrtsize=10)
                eng.close()
                ok, worst = True, 0.0
                for g, name in enumerate(ds.gene_names):
                    e = d["permjoin"].get(name)
                    if e is None:
                        ok = ok and pr.nperm_done[g] == 0
                        continue
                    ok = ok and int(pr.nperm_done[g]) == e["nperm"] and int(pr.count[g]) == round(e["pval"] * (e["total"] + 1))
                    worst = max(worst, abs(float(pr.true_stat[g]) - e["true"]))
                res[pbf]["parity"] = {"ok": bool(ok and worst <= 1e-8), "genes": len(d["permjoin"]), "exceedance_counts_exact": bool(ok),
                                      "max_abs_err_true_statistic": worst,
                                      "against": "oracle/_ref/eqtlbma_bf_ref_dump on the same slice, seed and write-group size"}
    return res


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, world, local_rank):
    import torch
    import eqtlbma_b200
    from eqtlbma_b200.shard import partition, slice_dataset

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = eqtlbma_b200.load_library()
    ds, shard_info = make_shard(rank, world, lib, args.genes or None)
    ds = pinned_copy(ds)
    kw = dict(analysis="join", bfs="sin", device=local_rank)

    # ---- device-resident throughput ("value")
    eng = eqtlbma_b200.Engine(ds, **kw)
    pairs = int(eng.pair_offsets()[-1])
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_w = time.perf_counter()
    n_w = 0
    # at least --warmup untimed steps, and at least ~1.5 s of load so that clocks settle and the
    # nvidia-smi sampler sees the device under load
    while n_w < args.warmup or time.perf_counter() - t_w < 1.5:
        eng.run_device_only(raw=True)
        n_w += 1
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = eng.launch_count()
    ms_list, kms_list = [], []
    for _ in range(args.steps):
        ms_list.append(eng.run_device_only(raw=True))
        kms_list.append(eng.last_pair_kernel_ms())
    torch.cuda.synchronize()
    launches = eng.launch_count() - l0
    clocks = sampler.stop()
    ms_step = float(np.mean(ms_list))
    ms_kernel = float(np.mean(kms_list))
    alg_bytes = algorithmic_bytes(ds, eng, raw=True)

    # ---- end to end through the C ABI with host buffers ("e2e")
    ds_fx = fixed_point_copy(ds)
    ds_e2e = ds_fx if ds_fx is not None else ds
    bufs = {True: eng.alloc_results(raw=True, pinned=True), False: eng.alloc_results(raw=False, pinned=True)}

    def e2e_step(raw, d=None):
        e = eqtlbma_b200.Engine(ds_e2e if d is None else d, **kw)
        r_ = e.run(raw=raw, out=bufs[raw])
        e.close()
        return r_

    def time_e2e(raw, d=None):
        for _ in range(max(1, args.warmup)):
            r_ = e2e_step(raw, d)
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        tt = []
        for _ in range(args.steps):
            ts = time.perf_counter()
            e2e_step(raw, d)
            tt.append(time.perf_counter() - ts)
        torch.cuda.synchronize()
        nbytes = sum(int(a.nbytes) for a in (r_.n, r_.sstats, r_.abf_gen, r_.abf_cfg, r_.abf_w) if a is not None)
        # the GPU box's host is shared: single steps are occasionally 2-3x slower (PCIe / memory contention from other
        # tenants); the per-step median is the robust estimate, the mean is reported next to it
        return float(np.median(tt)), float(np.mean(tt)), nbytes

    if args.no_e2e:
        e2e_s = e2e_mean = float("nan")
        d2h = 0
        e2e_raw_s = e2e_f64_s = None
    else:
        e2e_s, e2e_mean, d2h = time_e2e(True)
        e2e_raw_s, _, d2h_raw = time_e2e(False)  # (without the raw ABFs)
        e2e_f64_s = time_e2e(True, ds)[0] if (ds_fx is not None and world == 1) else None
    full = eng.run(raw=True)  # results of the shard (digest, parity, sharding check)
    digest = result_digest(full)
    # the same device-resident step when the genotypes came through the fixed-point transport: the u16 numerators stay
    # resident and the contraction of fast_pair_warp_kernel reads them (4x fewer bytes; bit-identical results)
    fx_info = None
    if ds_fx is not None and world == 1:
        eng_fx = eqtlbma_b200.Engine(ds_fx, **kw)
        for _ in range(max(3, args.warmup)):
            eng_fx.run_device_only(raw=True)
        fms, fk = [], []
        for _ in range(args.steps):
            fms.append(eng_fx.run_device_only(raw=True))
            fk.append(eng_fx.last_pair_kernel_ms())
        same = result_digest(eng_fx.run(raw=True)) == digest
        eng_fx.close()
        fx_info = {"value": pairs / (float(np.mean(fms)) * 1e-3), "unit": "pairs/s", "ms_per_step": float(np.mean(fms)),
                   "kernel_ms": float(np.mean(fk)), "bit_identical_to_f64_resident": bool(same),
                   "note": "genotypes resident as u16 numerators (eqb_set_genotypes_fixed) next to the doubles; the contraction "
                           "reads 2 bytes per genotype, so the SURVEY 8(d) byte count of the headline roofline (8 bytes) does not apply"}

    # ---- results do not depend on the sharding: the rank's shard cut in two by the partitioner, each half as its own
    # dataset / context, concatenated == the shard's result, bit for bit
    shard_check = None
    if not args.no_check:
        costs = (eng.cis_end - eng.cis_begin).astype(np.int64)
        sb2 = partition(lib, costs, WRTSIZE, 2)
        halves = []
        for k in range(2):
            sub = slice_dataset(ds, int(sb2[k]), int(sb2[k + 1]))
            sub._clean = False
            e = eqtlbma_b200.Engine(sub, **kw)
            halves.append(e.run(raw=True))
            e.close()
        ok = all(np.array_equal(np.concatenate([getattr(h, f) for h in halves]), getattr(full, f), equal_nan=True)
                 for f in ("n", "sstats", "abf_gen", "abf_cfg", "abf_w"))
        shard_check = {"ok": bool(ok), "pairs": pairs, "cut": [int(x) for x in sb2],
                       "what": "shard cut in two by eqb_partition_by_cost, halves run as separate datasets, concatenation "
                               "bit-identical to the unsharded run"}

    # ---- FP64 peaks + permuted pairs/s on the c4 shape
    fp64 = eqtlbma_b200.measure_fp64_peaks(local_rank) if rank == 0 else None
    if dist:
        obj = [fp64]
        dist.broadcast_object_list(obj, src=0)
        fp64 = obj[0]
    perm_info = None
    if not args.no_perm:
        perm_info = run_perm_block(eqtlbma_b200, rank, local_rank, fp64, with_cpu=(world == 1 and not args.no_cpu),
                                   n_genes=args.perm_genes or None, nperm=args.perm_nperm or None)

    hm_info = None
    if not args.no_hm:
        try:
            hm_info = run_hm_block(eqtlbma_b200, local_rank, peaks()[0]["hbm_gbs"], with_cpu=(world == 1 and not args.no_cpu),
                                   rank=rank, world=world, dist=dist)
        except Exception as exc:  # the headline line must not depend on the widening block
            hm_info = {"error": repr(exc)[:300]}

    hybrid_info = None
    if rank == 0 and not args.no_hybrid:
        try:
            hybrid_info = run_hybrid_block(eqtlbma_b200, local_rank, with_cpu=(world == 1 and not args.no_cpu))
        except Exception as exc:  # the headline line must not depend on the widening block
            hybrid_info = {"error": repr(exc)[:300]}

    # ---- max over ranks, whole-job aggregate
    tot_pairs = pairs
    digests = [digest]
    shard_ok = [shard_check["ok"] if shard_check else None]
    if dist:
        t = torch.tensor([ms_step, e2e_s, e2e_raw_s or 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_s, e2e_raw_s = float(t[0]), float(t[1]), float(t[2])
        c = torch.tensor([float(pairs)], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        tot_pairs = float(c[0])
        if perm_info:
            for pbf, r in perm_info["runs"].items():
                pp = torch.tensor([r["permuted_pairs_per_s"]], device="cuda", dtype=torch.float64)
                dist.all_reduce(pp, op=dist.ReduceOp.SUM)
                r["permuted_pairs_per_s"] = float(pp[0])
            for key in ("averaged_only", "with_raw_abfs"):
                pp = torch.tensor([perm_info["true_pass_bfs_all"][key]["pairs_per_s"]], device="cuda", dtype=torch.float64)
                dist.all_reduce(pp, op=dist.ReduceOp.SUM)
                perm_info["true_pass_bfs_all"][key]["pairs_per_s"] = float(pp[0])
        gathered = [None] * world
        dist.all_gather_object(gathered, (digest, shard_ok[0], shard_info["genes"], pairs))
        digests = [g[0] for g in gathered]
        shard_ok = [g[1] for g in gathered]
        shard_info["genes_per_rank"] = [g[2] for g in gathered]
        shard_info["pairs_per_rank"] = [g[3] for g in gathered]
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    pk, pk_kind = peaks()
    # dominant kernel = fast_pair_warp_kernel (K2+K3): algorithmic bytes of SURVEY 8(d) / its own CUDA-event duration;
    # the whole step (K1b + K1c + fix-up + K2+K3) is reported next to it
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    achieved_step = alg_bytes / (ms_step * 1e-3) / 1e9
    traffic = None
    for tp in ("r2_traffic.json", "r1_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tp)
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("fast_pair_warp_kernel_dram_bytes_per_launch")
            break
    out = {
        "metric": "cis gene-SNP pair BFs/sec", "value": tot_pairs / (ms_step * 1e-3), "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC, "pairs_per_gpu": pairs, "genes_per_gpu": ds.n_genes,
                   "l2": "inputs (genotypes %.0f MB per GPU) larger than the 126 MB L2" % (ds.genos[0].nbytes / 1e6),
                   "sharding": "ONE dataset, genes cut into contiguous shards of whole write-groups by eqb_partition_by_cost "
                               "(cost = cis SNPs per gene), one process per GPU, no collective on the data path",
                   "shards": shard_info},
        "clocks": clocks,
        "e2e": {"value": tot_pairs / e2e_s if e2e_s == e2e_s else None, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes(ds_e2e),
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3, "stat": "median of per-step wall times, max over ranks",
                "mean_ms_per_step": e2e_mean * 1e3,
                "results": "sample sizes, summary statistics, raw and grid-averaged ABFs (everything the reference writes in join mode)",
                "genotype_transport": ("u16 numerators of 1000 (lossless for the 3-decimal dosage file; "
                                       "eqb_set_genotypes_fixed)" if ds_fx is not None else "f64")},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_kind": pk_kind,
                     "kernel": "fast_pair_warp_kernel", "kernel_ms": ms_kernel, "algorithmic_bytes_per_launch": alg_bytes,
                     "step_achieved": achieved_step, "step_frac": achieved_step / pk["hbm_gbs"],
                     "step_kernels": "prep_y + prep_x_dmma + fix-up + fast_pair_warp"},
        "result_digests": digests,
        "shard_check": {"ok": all(bool(x) for x in shard_ok) if shard_check else None,
                        "what": shard_check["what"] if shard_check else None, "per_rank": shard_ok},
        "fp64_peaks": fp64,
    }
    if e2e_raw_s:
        out["e2e_avg_only"] = {"value": tot_pairs / e2e_raw_s, "unit": "pairs/s", "ms_per_step": e2e_raw_s * 1e3,
                               "d2h_bytes_per_step": d2h_raw if not args.no_e2e else None,
                               "results": "as e2e without the raw per-grid-point ABFs"}
    if e2e_f64_s is not None:
        out["e2e_f64"] = {"value": pairs / e2e_f64_s, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes(ds),
                          "ms_per_step": e2e_f64_s * 1e3, "genotype_transport": "f64 (eqb_set_genotypes)"}
    if fx_info:
        out["value_u16_resident"] = fx_info
    if perm_info:
        out["perm"] = perm_info
    if hm_info:
        out["hm"] = hm_info
    if hybrid_info:
        out["hybrid"] = hybrid_info
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline_reference(ds, sample_genes=args.cpu_genes)
        par = parity_vs_reference(ds, eng, full, args.cpu_genes)
        if par:
            out.update(par)
        cli = cli_e2e(ds, args.cli_genes)
        if cli:
            out["cli_e2e"] = cli
    print(json.dumps(out))
    if dist:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ reference runs
def subset_dataset(ds, n_genes, first=0):
    """n genes of the workload starting at `first` and the SNPs of their windows (same shape per gene)."""
    from eqtlbma_b200.shard import slice_dataset
    return slice_dataset(ds, first, min(first + n_genes, ds.n_genes))


def time_reference_binary(sub, flags, threads=1, raw_wall=False, tmp=None):
    """Wall time of the association loop of the reference binary on `sub`: total wall of the run
    minus the wall of the same invocation restricted to a gene without cis SNPs (input loading)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "eqtlbma_bf_ref")
    if not os.path.exists(exe):
        return None
    own = tmp is None
    tmp = tmp or tempfile.mkdtemp(prefix="eqb_ref_")
    try:
        sub.write_files(tmp)
        base = [exe] + sub.ref_args(tmp, os.path.join(tmp, "obs")) + flags + ["--thread", str(threads), "-v", "1"]
        t0 = time.perf_counter()
        r = subprocess.run(base, capture_output=True, text=True)
        t_full = time.perf_counter() - t0
        if r.returncode != 0:
            return None
        pairs = None
        for line in r.stdout.splitlines():
            if line.startswith("nb of analyzed gene-SNP pairs:"):
                pairs = int(line.split(":")[1].split("(")[0])
        if raw_wall:
            return dict(pairs=pairs, t_full=t_full)
        # loading-only run: a gene far from every SNP
        import gzip
        with gzip.open(os.path.join(tmp, "gene_far.bed.gz"), "wt") as fh:
            fh.write(f"{sub.chr_names[sub.gene_chr[0]]}\t999999999\t1000000100\t{sub.gene_names[0]}\t1000\t+\n")
        load = [a if a != f"{tmp}/gene_coords.bed.gz" else f"{tmp}/gene_far.bed.gz" for a in base]
        t0 = time.perf_counter()
        subprocess.run(load, capture_output=True, text=True)
        t_load = time.perf_counter() - t0
        return dict(pairs=pairs, seconds=max(t_full - t_load, 1e-9), t_full=t_full, t_load=t_load)
    finally:
        if own:
            shutil.rmtree(tmp, ignore_errors=True)


REF_FLAGS = ["--analys", "join", "--bfs", "sin", "--outss", "--outw"]


class ReferenceParallel:
    """The reference's own recipe for a multi-core host (scripts/eqtlbma_bf_parallel.bash:248-345): one single-threaded
    eqtlbma_bf per gene batch, all at once.  run() returns aggregate pairs / seconds over the association loops."""

    def __init__(self, ds, genes_per_proc, n_proc):
        self.exe = os.path.join(ROOT, "oracle", "_ref", "eqtlbma_bf_ref")
        self.ok = os.path.exists(self.exe)
        self.tmps, self.cmds, self.n_proc, self.t_load = [], [], n_proc, None
        if not self.ok:
            return
        for k in range(n_proc):
            sub = subset_dataset(ds, genes_per_proc, first=(k * genes_per_proc) % max(1, ds.n_genes - genes_per_proc))
            tmp = tempfile.mkdtemp(prefix=f"eqb_refp{k}_")
            sub.write_files(tmp)
            self.tmps.append(tmp)
            self.cmds.append([self.exe] + sub.ref_args(tmp, os.path.join(tmp, "obs")) + REF_FLAGS + ["--thread", "1", "-v", "1"])
        # loading-only time of one batch (subtracted once: the batches load concurrently)
        single = time_reference_binary(subset_dataset(ds, genes_per_proc), REF_FLAGS, threads=1)
        self.ok = single is not None
        self.t_load = single["t_load"] if single else None

    def run(self):
        if not self.ok:
            return None
        t0 = time.perf_counter()
        procs = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for c in self.cmds]
        outs = [p.communicate()[0] for p in procs]
        wall = time.perf_counter() - t0
        pairs = 0
        for o in outs:
            for line in o.splitlines():
                if line.startswith("nb of analyzed gene-SNP pairs:"):
                    pairs += int(line.split(":")[1].split("(")[0])
        if not pairs:
            return None
        return dict(pairs=pairs, seconds=max(wall - self.t_load, 1e-9), wall=wall, t_load=self.t_load, n_proc=self.n_proc)

    def close(self):
        for t in self.tmps:
            shutil.rmtree(t, ignore_errors=True)


def reference_parallel(ds, genes_per_proc, n_proc):
    rp = ReferenceParallel(ds, genes_per_proc, n_proc)
    try:
        return rp.run()
    finally:
        rp.close()


def cpu_baseline_reference(ds, sample_genes=60):
    sub = subset_dataset(ds, sample_genes)
    r = time_reference_binary(sub, REF_FLAGS, threads=1)
    if r is None or not r["pairs"]:
        return cpu_baseline_port(ds, sample_genes)
    out = {"value": r["pairs"] / r["seconds"], "unit": "pairs/s", "cores": 1, "kind": "reference",
           "sample": f"first {len(sub.gene_names)} genes ({r['pairs']} pairs) of the same workload through "
                     f"oracle/_ref/eqtlbma_bf_ref (unmodified reference + GSL shim); association loop "
                     f"{r['seconds']:.2f} s (run {r['t_full']:.2f} s - loading {r['t_load']:.2f} s); the reference's "
                     f"non-permuted pass is single-threaded by design (gene.cpp:282-284)"}
    n_proc = os.cpu_count() or 1
    par = reference_parallel(ds, max(8, sample_genes // 4), n_proc)
    if par:
        out["best_cpu"] = {"value": par["pairs"] / par["seconds"], "unit": "pairs/s", "cores": n_proc,
                           "sample": f"{n_proc} concurrent single-threaded reference processes (the reference's own multi-core "
                                     f"recipe, scripts/eqtlbma_bf_parallel.bash), {par['pairs']} pairs in {par['seconds']:.2f} s "
                                     f"(wall {par['wall']:.2f} s - loading {par['t_load']:.2f} s)"}
    return out


def cli_e2e(ds, n_genes=500):
    """Command-line drop-in, wall clock: eqtlbma_b200/eqtlbma_bf and the reference binary on the SAME input files (text
    parsing, association pass, %.6e formatting and gzip output all inside both times)."""
    exe = os.path.join(ROOT, "eqtlbma_b200", "eqtlbma_bf")
    ref = os.path.join(ROOT, "oracle", "_ref", "eqtlbma_bf_ref")
    if not os.path.exists(exe):
        return None
    sub = subset_dataset(ds, n_genes)
    tmp = tempfile.mkdtemp(prefix="eqb_cli_")
    try:
        sub.write_files(tmp)
        in_bytes = sum(os.path.getsize(os.path.join(tmp, f)) for f in os.listdir(tmp))
        nthr = str(os.cpu_count() or 1)

        def run(binary, tag, threads, verbose="1"):
            cmd = [binary] + sub.ref_args(tmp, os.path.join(tmp, tag)) + REF_FLAGS + ["--thread", threads, "-v", verbose]
            t0 = time.perf_counter()
            r = subprocess.run(cmd, capture_output=True, text=True)
            wall = time.perf_counter() - t0
            if r.returncode != 0:
                return None
            pairs, phases = None, None
            for line in r.stdout.splitlines():
                if line.startswith("nb of analyzed gene-SNP pairs:"):
                    pairs = int(line.split(":")[1].split("(")[0])
                if line.startswith("phases (wall clock, s):"):
                    phases = {kv.split("=")[0]: float(kv.split("=")[1]) for kv in line.split(":", 1)[1].split()}
            out_ = {"pairs": pairs, "wall_s": wall, "pairs_per_s": (pairs or 0) / wall}
            if phases:
                import gzip
                text = sum(len(gzip.open(os.path.join(tmp, f), "rb").read()) for f in os.listdir(tmp)
                           if f.startswith(tag + "_") and f.endswith(".gz"))
                out_["phases_s"] = phases
                out_["output_encoder"] = {"text_bytes": text, "threads": int(threads),
                                          "MB_per_s": text / 1e6 / max(phases.get("write", 0.0), 1e-9),
                                          "what": "%.6e formatting + one gzip member per thread slice (parallel_emit), "
                                                  "the 'write' phase of the run"}
            return out_

        run(exe, "warm", nthr)  # first process on the device pays the driver / module load once
        ours = run(exe, "ours", nthr, verbose="2")
        out = {"workload": f"first {len(sub.gene_names)} genes of the bench workload written as the reference's input files "
                           f"({in_bytes / 1e6:.1f} MB gzipped), --analys join --bfs sin --outss --outw", "ours": ours}
        if os.path.exists(ref):
            theirs = run(ref, "ref", "1")
            out["reference"] = theirs
            if ours and theirs and ours["pairs"] == theirs["pairs"]:
                out["wall_ratio"] = theirs["wall_s"] / ours["wall_s"]
                import gzip
                same = True
                for fn in sorted(os.listdir(tmp)):
                    if fn.startswith("ours_") and fn.endswith(".gz"):
                        a = gzip.open(os.path.join(tmp, fn), "rt").read().splitlines()
                        b = gzip.open(os.path.join(tmp, "ref_" + fn[5:]), "rt").read().splitlines()
                        same = same and len(a) == len(b) and all(x.split("\t")[:2] == y.split("\t")[:2] for x, y in zip(a, b))
                out["same_rows"] = bool(same)
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cpu_baseline_port(ds, sample_genes=60):
    from eqtlbma_b200._capi import Engine as AnyEngine
    path = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    sub = subset_dataset(ds, sample_genes)
    ora = AnyEngine(ctypes.CDLL(path), "eqo_", sub, analysis="join", bfs="sin")
    t0 = time.perf_counter()
    r = ora.run()
    dt = time.perf_counter() - t0
    return {"value": r.n.shape[0] / dt, "unit": "pairs/s", "cores": 1, "kind": "port",
            "sample": f"first {sample_genes} genes ({r.n.shape[0]} pairs) through the oracle restatement, {dt:.2f} s"}


def parity_vs_reference(ds, eng, full, sample_genes):
    """At-scale parity inside the run: the first genes of the bench workload through the UNMODIFIED reference
    (oracle/_ref/eqtlbma_bf_ref_dump: every getter at %.17g) against the CUDA results of the same genes.
    Tolerances of BASELINE.json: pair set exact, sample sizes exact, summary statistics 1e-9 relative, log10 ABFs 1e-8."""
    exe = os.path.join(ROOT, "oracle", "_ref", "eqtlbma_bf_ref_dump")
    if not os.path.exists(exe):
        return None
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from refdump import parse_dump
    sub = subset_dataset(ds, sample_genes)
    tmp = tempfile.mkdtemp(prefix="eqb_par_")
    try:
        sub.write_files(tmp)
        dump = os.path.join(tmp, "dump.txt")
        cmd = [exe] + sub.ref_args(tmp, os.path.join(tmp, "obs")) + REF_FLAGS + ["-v", "0"]
        r = subprocess.run(cmd, env=dict(os.environ, EQTLBMA_DUMP=dump), capture_output=True, text=True)
        if r.returncode != 0:
            return None
        d = parse_dump(dump)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    S = len(ds.subgroups)
    got_pairs = [(ds.gene_names[g], ds.snp_names[m]) for g in range(len(sub.gene_names)) if full.gene_analyzed[g]
                 for m in range(eng.cis_begin[g], eng.cis_end[g])]
    exp_pairs = [(p["gene"], p["snp"]) for p in d["pairs"]]
    pair_set_ok = got_pairs == exp_pairs
    n_ok, pve_ok, ss_err, abf_err = True, True, 0.0, 0.0
    if pair_set_ok:
        for p, pr in enumerate(d["pairs"]):
            for s in range(S):
                v = pr["ss"].get(s)
                if v is None:
                    n_ok = n_ok and full.n[p, s] == 0
                    continue
                n_ok = n_ok and int(full.n[p, s]) == v[0]
                a, b = np.asarray(full.sstats[p, s, 1:]), np.asarray(v[2:])
                with np.errstate(invalid="ignore", divide="ignore"):
                    e = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
                ss_err = max(ss_err, float(np.nanmax(np.where(np.isnan(a) & np.isnan(b), 0.0, e))))
                # pve = 1 - rss/tss carries an absolute rounding error of ~1e-16 in the reference itself (cancellation
                # for null pairs): compared with an absolute floor, as in tests/test_oracle_vs_reference.py
                pve_ok = pve_ok and abs(float(full.sstats[p, s, 0]) - v[1]) <= 1e-12 + 1e-9 * abs(v[1])
            for j, nm in enumerate(["gen", "gen-fix", "gen-maxh"]):
                abf_err = max(abf_err, float(np.max(np.abs(np.asarray(full.abf_gen[p, j]) - np.asarray(pr["raw"][nm])))))
                abf_err = max(abf_err, abs(float(full.abf_w[p, j]) - pr["w"][nm]))
            for c in range(S):
                abf_err = max(abf_err, float(np.max(np.abs(np.asarray(full.abf_cfg[p, c]) - np.asarray(pr["raw"][str(c + 1)])))))
                abf_err = max(abf_err, abs(float(full.abf_w[p, 5 + c]) - pr["w"][str(c + 1)]))
            abf_err = max(abf_err, abs(float(full.abf_w[p, 3]) - pr["w"]["gen-sin"]))
    ok = bool(pair_set_ok and n_ok and pve_ok and ss_err <= 1e-9 and abf_err <= 1e-8)
    return {"parity_checked_pairs": len(exp_pairs) if pair_set_ok else 0,
            "parity": {"ok": ok, "pair_set_exact": bool(pair_set_ok), "sample_sizes_exact": bool(n_ok),
                       "max_rel_err_sumstats": ss_err, "max_abs_err_log10_abf": abf_err,
                       "against": f"oracle/_ref/eqtlbma_bf_ref_dump (unmodified reference, %.17g) on the first {len(sub.gene_names)} "
                                  f"genes of the bench workload; tolerances 1e-9 relative / 1e-8 absolute"}}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from eqtlbma_b200.synth import make_dataset
    n_proc = os.cpu_count() or 1
    gpp = max(args.cpu_genes // 4, 8)  # genes per process: a bounded sample of the same shape (same per-gene layout)
    ds = make_dataset(**dict(block_kwargs(0), n_genes=max(gpp * 4, 64)))
    covs = canonical_covariates(C2["n_cov"], C2["n_inds"], C2["n_subgroups"])
    for s, sg in enumerate(ds.subgroups):
        sg.C = covs[s]
    vals, sample, kind = [], "", "reference"
    rp = ReferenceParallel(ds, gpp, n_proc)
    for it in range(args.warmup + args.steps):
        r = rp.run()
        if r is None:
            kind = "port"
            b = cpu_baseline_port(ds, gpp)
            v, sample = b["value"], b["sample"]
        else:
            v = r["pairs"] / r["seconds"]
            sample = (f"{n_proc} concurrent single-threaded eqtlbma_bf_ref processes (the reference's multi-core recipe, "
                      f"scripts/eqtlbma_bf_parallel.bash; its non-permuted pass has no threads, gene.cpp:282-284), {gpp} genes "
                      f"of the c2 shape each: {r['pairs']} pairs per step, association loops {r['seconds']:.2f} s "
                      f"(wall {r['wall']:.2f} s - loading {r['t_load']:.2f} s)")
        if it >= args.warmup:
            vals.append(v)
    rp.close()
    v = float(np.mean(vals))
    pairs_step = float(gpp * n_proc * BLOCK_SPG[0])
    out = {"impl": "reference", "metric": "cis gene-SNP pair BFs/sec", "value": v, "unit": "pairs/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": pairs_step / v * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": WORKLOAD_DESC},
           "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": n_proc if kind == "reference" else 1,
                            "kind": kind, "sample": sample},
           "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genes", type=int, default=0, help="override the number of genes per block / GPU (debug)")
    ap.add_argument("--cpu-genes", type=int, default=60, help="genes in the bounded CPU sample")
    ap.add_argument("--cli-genes", type=int, default=500, help="genes in the command-line end-to-end comparison")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-perm", action="store_true")
    ap.add_argument("--no-hm", action="store_true", help="skip the hierarchical-model (eqtlbma_hm) block")
    ap.add_argument("--no-hybrid", action="store_true", help="skip the --error hybrid block")
    ap.add_argument("--perm-genes", type=int, default=0, help="genes per GPU of the c3/c4-shape block (default 8)")
    ap.add_argument("--perm-nperm", type=int, default=0, help="permutations of the c4-shape block (default 2047)")
    ap.add_argument("--no-check", action="store_true", help="skip the sharding-invariance check")
    ap.add_argument("--no-e2e", action="store_true", help="kernel A/B runs only: skip the end-to-end loop (e2e = null)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
