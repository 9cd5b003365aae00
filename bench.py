#!/usr/bin/env python
"""bench.py -- throughput of the eqtlbma_bf hot path (cis gene-SNP pair BFs/sec) on B200.

Workload at N=1: BASELINE.json configs[1] ("c2"): 3 subgroups x 300 individuals, 11 covariates,
dosage genotypes shared by the subgroups, 5000 genes x ~50 cis SNPs (250k pairs),
`--analys join --bfs sin`, no permutations.  A "step" = one pass of the hot path over that batch.

  value  pairs/s with inputs resident in HBM (CUDA events around the kernels, on the library's stream)
  e2e    pairs/s through the C ABI with HOST buffers: eqb_create + H2D of genotypes / expression /
         covariates + eqb_finalize + eqb_run + D2H of every result the writers need, every step
  perm   secondary metric (BASELINE.json: "permuted pairs/sec"): pair x permutation evaluations/s on
         a bounded c4-style slice (9 ragged subgroups, --pbf gen-sin)
  --impl reference   times the reference's own CPU eqtlbma_bf (oracle/_ref, built from the unmodified
         sources against the GSL shim) on a bounded sample of the same workload.

N>1 (torchrun): genes are independent, so every rank owns a gene shard of the same shape (weak
scaling), no data-path collective; timing = max over ranks.
"""
from __future__ import annotations

import argparse
import ctypes
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

WORKLOAD = dict(seed=1859, n_subgroups=3, n_inds=300, n_genes=5000, snps_per_gene=50, n_cov=11, cov_per_subgroup=True,
                dosage=True,
                radius=100, gene_spacing=201, far_snp=False, n_chr=22)
WORKLOAD_DESC = ("c2: S=3 x N=300, Q=11 covariates (subgroup-specific values), dosage genotypes, 5000 genes x ~50 cis SNPs, "
                 "--analys join --bfs sin, gridL 25 / gridS 10, no permutations")
PERM_WORKLOAD = dict(seed=1860, n_subgroups=9, n_inds=450, n_genes=40, snps_per_gene=200, ragged=True,
                     ragged_min_frac=0.34, radius=100, gene_spacing=201, far_snp=False, n_chr=2)
PERM_NPERM = 200


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
    for i, G in enumerate(ds.genos):
        t = torch.from_numpy(np.ascontiguousarray(np.nan_to_num(G, nan=0.0))).pin_memory()
        ds.genos[i] = t.numpy()
        ds._pinned = getattr(ds, "_pinned", []) + [t]
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


def run_ours(args, rank, world, local_rank):
    import torch
    import eqtlbma_b200
    from eqtlbma_b200.synth import make_dataset

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = dict(WORKLOAD, seed=WORKLOAD["seed"] + rank)  # one gene shard of the same shape per rank
    if args.genes:
        wl["n_genes"] = args.genes
    ds = pinned_copy(make_dataset(**wl))
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

    # ---- end to end through the C ABI with host buffers ("e2e"): context creation, H2D of every
    # input from pinned host memory, layout build, kernels, D2H of every result into pinned buffers
    out_buf = eng.alloc_results(raw=True, pinned=True)
    # host genotype buffers in the compact lossless transport format the front-end's parser produces for this
    # dosage file (3 decimals -> u16 numerators of 1000); the plain double matrix is timed next to it ("e2e_f64")
    ds_fx = fixed_point_copy(ds)
    ds_e2e = ds_fx if ds_fx is not None else ds

    def e2e_step(d=None):
        e = eqtlbma_b200.Engine(ds_e2e if d is None else d, **kw)
        r_ = e.run(raw=True, out=out_buf)
        e.close()
        return r_

    for _ in range(0 if args.no_e2e else max(1, args.warmup)):  # untimed warm-up steps of the end-to-end path as well
        r = e2e_step()
    if args.no_e2e:
        r = eng.run(raw=True, out=out_buf)
        args_steps_e2e = 0
    else:
        args_steps_e2e = args.steps
    d2h = int(r.n.nbytes + r.sstats.nbytes + r.abf_gen.nbytes + r.abf_cfg.nbytes + r.abf_w.nbytes)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    if args.verbose:
        from eqtlbma_b200._capi import Engine as _E
        _E.timing = {}
    t0 = time.perf_counter()
    step_times = []
    for _ in range(args_steps_e2e):
        ts = time.perf_counter()
        e2e_step()
        step_times.append(time.perf_counter() - ts)
    torch.cuda.synchronize()
    e2e_mean_s = (time.perf_counter() - t0) / args.steps
    if args.no_e2e:
        step_times = [float("nan")]
    # the GPU box's host is shared: single steps are occasionally 2-3x slower (PCIe / memory contention from other
    # tenants); the per-step median is the robust estimate, the mean is reported next to it
    e2e_s = float(np.median(step_times))
    e2e_f64_s = None
    if ds_fx is not None and not args.no_e2e:
        e2e_step(ds)
        tt = []
        for _ in range(args.steps):
            ts = time.perf_counter()
            e2e_step(ds)
            tt.append(time.perf_counter() - ts)
        e2e_f64_s = float(np.median(tt))
    if args.verbose:
        print("e2e per-step ms:", [round(t * 1e3, 2) for t in step_times], file=sys.stderr)
    if args.verbose:
        print("e2e host timing per step (ms):", {k: round(v * 1e3 / args.steps, 2) for k, v in _E.timing.items()},
              file=sys.stderr)
        _E.timing = None

    # ---- secondary metric: permuted pairs/s on a bounded c4-style slice
    perm_info = None
    if not args.no_perm:
        pds = make_dataset(**dict(PERM_WORKLOAD, seed=PERM_WORKLOAD["seed"] + rank))
        peng = eqtlbma_b200.Engine(pds, analysis="join", bfs="sin", device=local_rank)
        ppairs = int(peng.pair_offsets()[-1])
        peng.run_permutations_device_only(PERM_NPERM, 1859, pbf="gen-sin", wrtsize=10)
        pms = [peng.run_permutations_device_only(PERM_NPERM, 1859, pbf="gen-sin", wrtsize=10) for _ in range(2)]
        perm_info = {"permuted_pairs_per_s": ppairs * PERM_NPERM / (np.mean(pms) * 1e-3), "pairs": ppairs,
                     "nperm": PERM_NPERM, "ms": float(np.mean(pms)),
                     "workload": "c4 slice: S=9 ragged of 450, 40 genes x ~200 cis SNPs, --pbf gen-sin"}
        peng.close()

    # ---- max over ranks, whole-job aggregate
    tot_pairs = pairs
    if dist:
        t = torch.tensor([ms_step, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_s = float(t[0]), float(t[1])
        c = torch.tensor([float(pairs)], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        tot_pairs = float(c[0])
        if perm_info:
            pp = torch.tensor([perm_info["permuted_pairs_per_s"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(pp, op=dist.ReduceOp.SUM)
            perm_info["permuted_pairs_per_s"] = float(pp[0])
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
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("fast_pair_warp_kernel_dram_bytes_per_launch")
    out = {
        "metric": "cis gene-SNP pair BFs/sec", "value": tot_pairs / (ms_step * 1e-3), "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC, "pairs_per_gpu": pairs, "genes_per_gpu": ds.n_genes,
                   "l2": "inputs (genotypes %.0f MB per GPU) larger than the 126 MB L2" % (ds.genos[0].nbytes / 1e6),
                   "sharding": "genes sharded across ranks, no collective"},
        "clocks": clocks,
        "e2e": {"value": tot_pairs / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes(ds_e2e),
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3, "stat": "median of per-step wall times",
                "mean_ms_per_step": e2e_mean_s * 1e3,
                "genotype_transport": ("u16 numerators of 1000 (lossless for the 3-decimal dosage file; "
                                       "eqb_set_genotypes_fixed)" if ds_fx is not None else "f64")},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_kind": pk_kind,
                     "kernel": "fast_pair_warp_kernel", "kernel_ms": ms_kernel, "algorithmic_bytes_per_launch": alg_bytes,
                     "step_achieved": achieved_step, "step_frac": achieved_step / pk["hbm_gbs"],
                     "step_kernels": "prep_y + prep_x_dmma + fix-up + fast_pair_warp"},
    }
    if e2e_f64_s is not None and world == 1:
        out["e2e_f64"] = {"value": pairs / e2e_f64_s, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes(ds),
                          "ms_per_step": e2e_f64_s * 1e3, "genotype_transport": "f64 (eqb_set_genotypes)"}
    if perm_info:
        out["perm"] = perm_info
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline_reference(ds, sample_genes=args.cpu_genes)
    print(json.dumps(out))
    if dist:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def subset_dataset(ds, n_genes):
    """First n genes of the workload and the SNPs of their windows (same shape per gene)."""
    import copy
    beg, end = ds.cis_windows()
    keep_g = np.arange(min(n_genes, ds.n_genes))
    m_hi = int(end[keep_g].max())
    m_lo = int(beg[keep_g].min())
    sub = copy.copy(ds)
    sub.genos = [G[m_lo:m_hi] for G in ds.genos]
    sub.snp_names = ds.snp_names[m_lo:m_hi]
    sub.snp_chr = ds.snp_chr[m_lo:m_hi]
    sub.snp_pos = ds.snp_pos[m_lo:m_hi]
    sub.snp_bed_start = ds.snp_bed_start[m_lo:m_hi]
    sub.gene_names = [ds.gene_names[g] for g in keep_g]
    sub.gene_chr = ds.gene_chr[keep_g]
    sub.gene_start = ds.gene_start[keep_g]
    sub.gene_end = ds.gene_end[keep_g]
    sub.subgroups = []
    for sg in ds.subgroups:
        s2 = copy.copy(sg)
        s2.Y = sg.Y[keep_g]
        s2.gene_has_exp = sg.gene_has_exp[keep_g]
        s2.snp_has_geno = sg.snp_has_geno[m_lo:m_hi]
        sub.subgroups.append(s2)
    return sub


def time_reference_binary(sub, flags, threads=1):
    """Wall time of the association loop of the reference binary on `sub`: total wall of the run
    minus the wall of the same invocation restricted to a gene without cis SNPs (input loading)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "eqtlbma_bf_ref")
    if not os.path.exists(exe):
        return None
    tmp = tempfile.mkdtemp(prefix="eqb_ref_")
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
        shutil.rmtree(tmp, ignore_errors=True)


def cpu_baseline_reference(ds, sample_genes=60):
    sub = subset_dataset(ds, sample_genes)
    r = time_reference_binary(sub, ["--analys", "join", "--bfs", "sin", "--outss", "--outw"], threads=1)
    if r is None or not r["pairs"]:
        return cpu_baseline_port(ds, sample_genes)
    return {"value": r["pairs"] / r["seconds"], "unit": "pairs/s", "cores": 1, "kind": "reference",
            "sample": f"first {len(sub.gene_names)} genes ({r['pairs']} pairs) of the same workload through "
                      f"oracle/_ref/eqtlbma_bf_ref (unmodified reference + GSL shim); association loop "
                      f"{r['seconds']:.2f} s (run {r['t_full']:.2f} s - loading {r['t_load']:.2f} s); the reference's "
                      f"non-permuted pass is single-threaded by design (gene.cpp:282-284)"}


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


def run_reference(args, rank, world):
    if rank != 0:
        return
    from eqtlbma_b200.synth import make_dataset
    wl = dict(WORKLOAD)
    wl["n_genes"] = max(args.cpu_genes, 8)  # bounded sample of the same shape (same per-gene layout)
    ds = make_dataset(**wl)
    threads = os.cpu_count() or 1
    vals = []
    kind, sample = "reference", ""
    for it in range(args.warmup + args.steps):
        r = time_reference_binary(ds, ["--analys", "join", "--bfs", "sin", "--outss", "--outw"], threads=threads)
        if r is None or not r["pairs"]:
            kind = "port"
            b = cpu_baseline_port(ds, wl["n_genes"])
            v, sample = b["value"], b["sample"]
        else:
            v = r["pairs"] / r["seconds"]
            sample = (f"{wl['n_genes']} genes ({r['pairs']} pairs) of the c2 shape per step through oracle/_ref/eqtlbma_bf_ref, "
                      f"--thread {threads} (only permutation loops are threaded in the reference)")
        if it >= args.warmup:
            vals.append(v)
    v = float(np.mean(vals))
    pairs_step = float(wl["n_genes"] * wl["snps_per_gene"])
    out = {"impl": "reference", "metric": "cis gene-SNP pair BFs/sec", "value": v, "unit": "pairs/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": pairs_step / v * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": WORKLOAD_DESC},
           "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads if kind == "reference" else 1,
                            "kind": kind, "sample": sample},
           "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genes", type=int, default=0, help="override the number of genes per GPU (debug)")
    ap.add_argument("--cpu-genes", type=int, default=60, help="genes in the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-perm", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="kernel A/B runs only: skip the end-to-end loop (e2e = null)")
    ap.add_argument("--verbose", action="store_true")
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
