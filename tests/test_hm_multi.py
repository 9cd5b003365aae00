"""Multi-GPU hierarchical-model EM: genes sharded over ranks, the partial sums of every evaluation all-gathered and combined
in rank order (eqb_hm_set_collective / eqb_hm_combine_partials of include/eqtlbma_hm_b200.h).
CPU (world_size 2, gloo): the host-side combination against the numpy restatement, per-rank partials computed by the
restatement.  GPU (needs 2 devices; skipped otherwise): two ranks on two GPUs over NCCL against the single-GPU fit."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CPU_WORKER = r"""
import ctypes, os, sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "oracle"))
import torch, torch.distributed as dist
from hm_oracle import HmOracle, l10ws
from eqtlbma_b200.hm_synth import make_hm_dataset
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
lib = ctypes.CDLL(os.path.join(sys.argv[1], "eqtlbma_b200", "libeqtlbma_b200.so"))  # host-only entry point
ds = make_hm_dataset(seed=31, n_genes=301, snps_lo=1, snps_hi=20, round_text=False)
cut = [0, 117, ds.n_genes]
lo, hi = cut[rank], cut[rank + 1]
p0, p1 = int(ds.gene_off[lo]), int(ds.gene_off[hi])
o = HmOracle(ds.B[p0:p1], ds.gene_off[lo:hi + 1] - p0)       # compute stand-in on CPU for this rank's genes
rs = np.random.RandomState(2)
pi0, gw, cp = 0.37, rs.dirichlet(np.ones(ds.grid)), rs.dirichlet(np.ones(ds.dim))
lik = o.gene_lik(pi0, gw, cp)
ones = np.ones(o.G)
cg = o._over_snps(o.cfg_bf(gw)) - lik[:, None]
gg = o._over_snps(l10ws(o.B, cp[None, :, None], axis=1)) - lik[:, None]
# what one rank reads back: [sum lik, sum 10^(log10 pi0 - lik), log10 sum_g 10^cg[:, k] ..., log10 sum_g 10^gg[:, l] ...]
def lse(v):
    m = v.max(axis=0); return m + np.log10(np.sum(10.0 ** (v - m), axis=0))
mine = np.concatenate([[lik.sum(), np.sum(10.0 ** (np.log10(pi0) - lik))], lse(cg), lse(gg)])
parts = [torch.empty(len(mine), dtype=torch.float64) for _ in range(2)]
dist.all_gather(parts, torch.from_numpy(mine))
gathered = np.ascontiguousarray(torch.stack(parts).numpy())
out = np.zeros(len(mine))
f = lib.eqb_hm_combine_partials; f.restype = ctypes.c_int
assert f(gathered.ctypes.data_as(ctypes.c_void_p), 2, len(mine), out.ctypes.data_as(ctypes.c_void_p)) == 0
full = HmOracle(ds.B, ds.gene_off)
flik = full.gene_lik(pi0, gw, cp)
assert abs(out[0] - flik.sum()) <= 1e-12 * abs(flik.sum())
n_pi0, n_gw, n_cp = full.fixedpoint(pi0, gw, cp, dict(pi0=False, grid=False, configs=False))
assert abs(out[1] / ds.n_genes - n_pi0) <= 1e-13
t = out[2:2 + ds.dim] + np.log10(cp); got = 10.0 ** (t - lse(t[:, None])[0])
assert np.allclose(got, n_cp, rtol=1e-11, atol=0)
t = out[2 + ds.dim:] + np.log10(gw); got = 10.0 ** (t - lse(t[:, None])[0])
assert np.allclose(got, n_gw, rtol=1e-11, atol=0)
# NaN and empty (-inf) partials
g2 = np.array([[1.0, 2.0, -np.inf, np.nan, -np.inf], [3.0, 4.0, -np.inf, 1.0, 2.0]]); o2 = np.zeros(5)
assert f(g2.ctypes.data_as(ctypes.c_void_p), 2, 5, o2.ctypes.data_as(ctypes.c_void_p)) == 0
assert o2[0] == 4.0 and o2[1] == 6.0 and o2[2] == -np.inf and np.isnan(o2[3]) and o2[4] == 2.0
if rank == 0: print("HM_COMBINE_OK")
dist.destroy_process_group()
"""

GPU_WORKER = r"""
import os, sys, numpy as np
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from eqtlbma_b200.hm import HmEngine, HmFit
from eqtlbma_b200.hm_synth import make_hm_dataset
rank, world = int(sys.argv[3]), int(sys.argv[4])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=rank, world_size=world)
ds = make_hm_dataset(seed=41, n_genes=4001, snps_lo=1, snps_hi=40, round_text=False)
cut = np.linspace(0, ds.n_genes, world + 1).astype(int)
lo, hi = int(cut[rank]), int(cut[rank + 1])
p0, p1 = int(ds.gene_off[lo]), int(ds.gene_off[hi])
hm = HmEngine(ds.dim, ds.grid, device=rank)
hm.append(ds.B[p0:p1], ds.gene_off[lo:hi + 1] - p0)
hm.finalize()
gw, cp = np.full(ds.grid, 1.0 / ds.grid), np.full(ds.dim, 1.0 / ds.dim)
fits, by_mode = {}, {}
for native in (True, False):    # hm_xchg_kernel over peer memory, then the NCCL all-gather through the host callback
    hm.set_collective(native=native)
    for label, msl in (("classic", 1.0), ("squarem", 3.0)):
        fits[label] = hm.em(HmFit(0.5, gw, cp), thresh=0.01, stepmax=msl)
    by_mode[native] = dict(fits)
for label in fits:   # the two transports combine the same values in the same order (device exp10/log10 vs host pow/log10)
    a, b = by_mode[True][label], by_mode[False][label]
    assert abs(a.loglik - b.loglik) <= 1e-12 * abs(a.loglik) and len(a.log_lines) == len(b.log_lines)
    assert np.allclose(a.grid_wts, b.grid_wts, rtol=1e-10, atol=1e-15)
hm.set_collective(native=True)
fits = by_mode[True]
post = hm.posteriors(hm.em(HmFit(0.5, gw, cp), thresh=0.01))
# every rank holds bit-identical estimates (the combination runs in rank order on every rank)
mine = torch.tensor([fits["squarem"].loglik, fits["squarem"].pi0] + list(fits["squarem"].grid_wts), dtype=torch.float64, device="cuda")
both = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(both, mine)
assert all(torch.equal(both[0], t) for t in both)
if rank == 0:
    one = HmEngine(ds.dim, ds.grid, device=0)
    one.append(ds.B, ds.gene_off); one.finalize()
    for label, msl in (("classic", 1.0), ("squarem", 3.0)):
        ref, got = one.em(HmFit(0.5, gw, cp), thresh=0.01, stepmax=msl), fits[label]
        assert len(ref.log_lines) == len(got.log_lines), label
        assert abs(ref.loglik - got.loglik) <= 1e-11 * abs(ref.loglik), (label, ref.loglik, got.loglik)
        assert abs(ref.pi0 - got.pi0) <= 1e-10 and np.allclose(ref.grid_wts, got.grid_wts, rtol=1e-9, atol=1e-14)
        assert np.allclose(ref.config_prior, got.config_prior, rtol=1e-9, atol=1e-14)
    rpost = one.posteriors(one.em(HmFit(0.5, gw, cp), thresh=0.01))
    assert np.allclose(rpost["gene_post"][lo:hi], post["gene_post"], rtol=1e-9, atol=1e-14)
    assert np.allclose(rpost["snp_bf"][p0:p1], post["snp_bf"], rtol=0, atol=1e-10)
    print("HM_MULTI_GPU_OK")
dist.barrier()
dist.destroy_process_group()
"""


def _run(tmp_path, src, world, extra=()):
    script = tmp_path / "worker.py"
    script.write_text(src)
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)] + list(extra), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    return outs


def test_two_rank_combination_matches_restatement(tmp_path, cuda_lib):
    outs = _run(tmp_path, CPU_WORKER, 2)
    assert "HM_COMBINE_OK" in outs[0][0]


@pytest.mark.gpu
def test_two_gpu_em_matches_single_gpu(tmp_path, cuda_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    outs = _run(tmp_path, GPU_WORKER, 2, extra=("2",))
    assert "HM_MULTI_GPU_OK" in outs[0][0]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["configs_subset", "classic_bf", "getci"])
def test_cli_two_gpus_writes_the_reference_file(tmp_path, cuda_lib, name):
    """eqtlbma_hm --gpus 2: one process per GPU on contiguous shards of the files (configs_subset: 3 files) or of the genes
    (one file), hm_xchg_kernel between them, shard outputs merged byte-wise: the reference's output file."""
    import gzip
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from hm_scenarios import GOLDEN_HM, HM_SCENARIOS, build_dataset, ref_cmdline
    from test_hm_cli import BIN, _cells_match, _write_inputs
    sc = HM_SCENARIOS[name]
    ds = build_dataset(sc)
    tmp = str(tmp_path)
    init = _write_inputs(sc, ds, tmp)
    out = os.path.join(tmp, "out_hm.txt.gz")
    cmd = [BIN] + ref_cmdline(sc, ds, os.path.join(tmp, "in_*_l10abfs_raw.txt.gz"), out, init) + ["--gpus", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp)
    assert r.returncode == 0, r.stderr[-2000:]
    got = gzip.open(out, "rt").read().splitlines()
    ref = gzip.open(os.path.join(GOLDEN_HM, name + ".out_hm.txt.gz"), "rt").read().splitlines()
    assert len(got) == len(ref)
    for lg, lr in zip(got, ref):
        cg, cr = lg.split("\t"), lr.split("\t")
        assert len(cg) == len(cr), (lg, lr)
        assert all(_cells_match(a, b) for a, b in zip(cg, cr)), (lg, lr)
    assert not [f for f in os.listdir(tmp) if ".shard" in f]
