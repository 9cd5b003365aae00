"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/*.h declares,
fails loudly without a device, and the multi-GPU gene sharding (world_size 2, gloo) reproduces the
single-process result through the final host gather."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    inc = os.path.join(ROOT, "include")
    hdr = "".join(open(os.path.join(inc, f)).read() for f in sorted(os.listdir(inc)) if f.endswith(".h"))
    # function declarations only (eqb_hm_ctx, eqb_hm_fit ... are types)
    return sorted(set(re.findall(r"\b(eqb_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(cuda_lib):
    syms = declared_symbols()
    assert len(syms) >= 30 and "eqb_hm_em" in syms and "eqb_raw_abfs_device" in syms
    for s in syms:
        assert hasattr(cuda_lib, s), f"libeqtlbma_b200.so does not export {s}"


def test_no_cpu_fallback_without_device(cuda_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from eqtlbma_b200._capi import Config, ABI_VERSION
    cfg = Config(ABI_VERSION, 2, 10, 1, 5, 3, 0, 0, 0, 0, 0.5)
    ctx = ctypes.c_void_p()
    f = cuda_lib.eqb_create
    f.restype = ctypes.c_int
    rc = f(ctypes.byref(ctx), ctypes.byref(cfg))
    assert rc != 0, "eqb_create must fail without a CUDA device (no CPU fallback)"
    e = cuda_lib.eqb_last_error
    e.restype = ctypes.c_char_p
    assert b"no CUDA device" in e(ctx)
    d = cuda_lib.eqb_destroy
    d.restype = None
    d(ctx)


def test_partition_contiguous_whole_groups(cuda_lib):
    from eqtlbma_b200.shard import partition
    rng = np.random.default_rng(1)
    for G, w, n in [(103, 10, 8), (20, 3, 2), (7, 10, 4), (1000, 7, 8), (5, 1, 8)]:
        cost = rng.integers(0, 500, size=G)
        sb = partition(cuda_lib, cost, w, n)
        assert sb[0] == 0 and sb[-1] == G
        assert np.all(np.diff(sb) >= 0)
        assert all(b % w == 0 or b == G for b in sb)
        if G >= 10 * n * w:  # balanced when there are many groups
            per = [cost[sb[k]:sb[k + 1]].sum() for k in range(n)]
            assert max(per) <= 1.5 * cost.sum() / n + cost.max() * w


WORKER = r"""
import ctypes, os, sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch.distributed as dist
from eqtlbma_b200._capi import Engine
from eqtlbma_b200.shard import partition, gene_costs
from eqtlbma_b200.synth import make_dataset
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
ds = make_dataset(seed=5, n_subgroups=3, n_inds=60, n_genes=23, snps_per_gene=3, ragged=True)
ora = ctypes.CDLL(os.path.join(sys.argv[1], "oracle", "liboracle.so"))       # compute stand-in on CPU
lib = ctypes.CDLL(os.path.join(sys.argv[1], "eqtlbma_b200", "libeqtlbma_b200.so"))  # host-only entry point
eng = Engine(ora, "eqo_", ds, analysis="join", bfs="sin")
wrt, nperm = 4, 20
sb = partition(lib, gene_costs(eng.cis_begin, eng.cis_end, nperm), wrt, 2)
lo, hi = int(sb[rank]), int(sb[rank + 1])
res = eng.run(lo, hi)
perm = eng.run_permutations(nperm, 1859, lo=lo, hi=hi, pbf="gen-sin", wrtsize=wrt)
mine = dict(lo=lo, hi=hi, w=res.abf_w, n=res.n, count=perm.count, pval=perm.pval)
gathered = [None, None] if rank == 0 else None
dist.gather_object(mine, gathered, dst=0)      # the final host gather: no collective on the data path
if rank == 0:
    full = eng.run()
    fperm = eng.run_permutations(nperm, 1859, pbf="gen-sin", wrtsize=wrt)
    w = np.concatenate([g["w"] for g in gathered]); n = np.concatenate([g["n"] for g in gathered])
    cnt = np.concatenate([g["count"] for g in gathered]); pv = np.concatenate([g["pval"] for g in gathered])
    assert gathered[0]["hi"] == gathered[1]["lo"] and gathered[1]["hi"] == ds.n_genes
    assert np.array_equal(n, full.n) and np.allclose(w, full.abf_w, equal_nan=True, rtol=0, atol=0)
    assert np.array_equal(cnt, fperm.count) and np.allclose(pv, fperm.pval, equal_nan=True, rtol=0, atol=0)
    print("SHARDING_OK")
dist.destroy_process_group()
"""


def test_two_rank_sharding_matches_single_process(tmp_path, oracle_lib, cuda_lib):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "SHARDING_OK" in outs[0][0]
