"""GPU parity tests: the CUDA library, called through its C ABI, against (a) the golden dumps of the
UNMODIFIED reference and (b) the CPU oracle on the same seeded inputs.
Tolerances (BASELINE.json north_star): pair sets / sample sizes / exceedance counts bit-exact,
summary statistics 1e-9 relative, log10 ABFs 1e-8 absolute."""
import os

import numpy as np
import pytest

from eqtlbma_b200._capi import Engine as AnyEngine
from refdump import parse_dump
from scenarios import SCENARIOS, build_dataset, engine_kwargs, perm_kwargs
from test_oracle_vs_reference import GOLD, check_against_dump

pytestmark = pytest.mark.gpu

# paths of the reference that the CUDA library does not cover yet (reported as errors by the ABI)
NOT_YET = set()
# degenerate rank-deficient designs: documented tie, checked separately
DEGENERATE = {"monomorphic"}
# 10^4-permutation slices: checked against the reference dumps only (the CPU oracle needs minutes for them and is itself
# pinned to the same dumps by the CPU suite)
DUMP_ONLY = {"c4_slice_gensin_10k_trick2", "c4_slice_all_10k"}


@pytest.mark.parametrize("name", sorted(set(SCENARIOS) - NOT_YET - DEGENERATE))
def test_cuda_matches_reference_dump(cuda_lib, name):
    import eqtlbma_b200
    sc = SCENARIOS[name]
    ds = build_dataset(sc)
    dump = parse_dump(os.path.join(GOLD, name + ".dump.gz"))
    eng = eqtlbma_b200.Engine(ds, **engine_kwargs(sc))
    res = eng.run()
    pk = perm_kwargs(sc)
    perm = eng.run_permutations(**pk) if pk else None
    check_against_dump(eng, ds, sc, dump, res, perm)
    assert eng.launch_count() > 0
    eng.close()


@pytest.mark.parametrize("name", sorted(set(SCENARIOS) - NOT_YET - DUMP_ONLY))
def test_cuda_matches_oracle(cuda_lib, oracle_lib, name):
    import eqtlbma_b200
    sc = SCENARIOS[name]
    ds = build_dataset(sc)
    eng = eqtlbma_b200.Engine(ds, **engine_kwargs(sc))
    ora = AnyEngine(oracle_lib, "eqo_", ds, **engine_kwargs(sc))
    a, b = eng.run(), ora.run()
    assert np.array_equal(a.offsets, b.offsets)
    assert np.array_equal(a.gene_analyzed, b.gene_analyzed)
    assert np.array_equal(a.n, b.n)
    # pve = 1 - rss/tss is compared absolutely (it is exactly 0 up to rounding for null designs)
    assert np.allclose(a.sstats[..., 0], b.sstats[..., 0], rtol=1e-9, atol=1e-12, equal_nan=True)
    assert np.allclose(a.sstats[..., 1:], b.sstats[..., 1:], rtol=1e-9, atol=0, equal_nan=True)
    if sc["analysis"] == "join":
        assert np.allclose(a.abf_gen, b.abf_gen, rtol=0, atol=1e-8, equal_nan=True)
        assert np.allclose(a.abf_cfg, b.abf_cfg, rtol=0, atol=1e-8, equal_nan=True)
        assert np.allclose(a.abf_w, b.abf_w, rtol=0, atol=1e-8, equal_nan=True)
        # without the raw arrays (--bfs all: the linear-domain kernel with its log-domain fall-backs): same averages
        c = eng.run(raw=False)
        assert np.allclose(c.abf_w, b.abf_w, rtol=0, atol=1e-8, equal_nan=True)
        assert np.array_equal(np.isnan(c.abf_w), np.isnan(b.abf_w))
    pk = perm_kwargs(sc)
    if pk:
        pa, pb = eng.run_permutations(**pk), ora.run_permutations(**pk)
        assert np.array_equal(pa.count, pb.count)
        assert np.array_equal(pa.nperm_done, pb.nperm_done)
        assert np.allclose(pa.pval, pb.pval, rtol=1e-12, atol=0, equal_nan=True)
        tol = dict(rtol=0, atol=1e-8) if sc["analysis"] == "join" else dict(rtol=1e-9, atol=0)
        assert np.allclose(pa.true_stat, pb.true_stat, equal_nan=True, **tol)
        assert np.allclose(pa.perm_stats, pb.perm_stats, equal_nan=True, **tol)
    eng.close()
    ora.close()


# --error hybrid beyond the golden shapes, against the restatement (itself pinned to the reference dumps): 6 subgroups x
# 2000 individuals (the per-subgroup bases no longer fit in shared memory: global workspace, bounded grid) and a
# covariate design with unequal subgroups; true pass and a few permutations
HYBRID_EXTRA = {
    "wide_workspace": dict(data=dict(seed=801, n_subgroups=6, n_inds=2000, n_genes=3, snps_per_gene=2, ragged=True,
                                     ragged_min_frac=0.7, nan_frac=0.01), bfs="sin", fiterr=0.5,
                           perm=dict(nperm=3, seed=5, pbf="gen-sin", wrtsize=2)),
    "cov3_all": dict(data=dict(seed=802, n_subgroups=4, n_inds=300, n_genes=5, snps_per_gene=3, ragged=True, n_cov=3,
                               pad_names=True, dosage=True), bfs="all", fiterr=0.25,
                     perm=dict(nperm=6, seed=6, pbf="all", wrtsize=3)),
}


@pytest.mark.parametrize("name", sorted(HYBRID_EXTRA))
def test_hybrid_extra_shapes_match_oracle(cuda_lib, oracle_lib, name):
    import eqtlbma_b200
    from eqtlbma_b200.synth import make_dataset
    sc = HYBRID_EXTRA[name]
    ds = make_dataset(**sc["data"])
    kw = dict(analysis="join", bfs=sc["bfs"], error="hybrid", fiterr=sc["fiterr"])
    eng = eqtlbma_b200.Engine(ds, **kw)
    ora = AnyEngine(oracle_lib, "eqo_", ds, **kw)
    a, b = eng.run(), ora.run()
    assert np.array_equal(a.n, b.n)
    assert np.isfinite(b.abf_w[:, 0]).all()
    assert np.allclose(a.sstats[..., 0], b.sstats[..., 0], rtol=1e-9, atol=1e-12, equal_nan=True)
    assert np.allclose(a.sstats[..., 1:], b.sstats[..., 1:], rtol=1e-9, atol=0, equal_nan=True)
    for x, y in ((a.abf_gen, b.abf_gen), (a.abf_cfg, b.abf_cfg), (a.abf_w, b.abf_w)):
        assert np.allclose(x, y, rtol=0, atol=1e-8, equal_nan=True)
    pk = dict(trick=0, tricut=10, permsep=0, maxbf=False, **sc["perm"])
    pa, pb = eng.run_permutations(**pk), ora.run_permutations(**pk)
    assert np.array_equal(pa.count, pb.count)
    assert np.allclose(pa.true_stat, pb.true_stat, rtol=0, atol=1e-8, equal_nan=True)
    assert np.allclose(pa.perm_stats, pb.perm_stats, rtol=0, atol=1e-8, equal_nan=True)
    eng.close()
    ora.close()


def test_hybrid_reference_failure_modes_are_reported(cuda_lib, oracle_lib):
    """Inputs on which the reference stops or reads past its vectors: two subgroups without a common individual
    (gene_snp_pair.cpp:897-901, fatal) and covariate files that are not in the order of the sorted sample names (the
    off-diagonal designs index them by the all-sample index, :931-933).  Both are errors of the ABI, never silent values."""
    import eqtlbma_b200
    from eqtlbma_b200.synth import make_dataset
    kw = dict(analysis="join", bfs="gen", error="hybrid", fiterr=0.5)
    ds = make_dataset(seed=5, n_subgroups=2, n_inds=60, n_genes=3, snps_per_gene=2)
    half = len(ds.samples) // 2
    ds.subgroups[0].all2exp[half:] = -1
    ds.subgroups[1].all2exp[:half] = -1
    eng = eqtlbma_b200.Engine(ds, **kw)
    with pytest.raises(RuntimeError, match="no individuals in common"):
        eng.run()
    eng.close()
    ora = AnyEngine(oracle_lib, "eqo_", ds, **kw)
    with pytest.raises(RuntimeError, match="no individuals in common"):
        ora.run()
    ora.close()
    ds = make_dataset(seed=6, n_subgroups=2, n_inds=60, n_genes=3, snps_per_gene=2, n_cov=1)  # ind1, ind10, ind11, ...: not identity
    with pytest.raises(RuntimeError, match="order of the sorted sample names"):
        eqtlbma_b200.Engine(ds, **kw)


def test_cis_windows_match_linear_scan(cuda_lib, oracle_lib):
    """a1: the device binary search reproduces Gene::SetCisSnps / Snp::IsInCis bit-exactly,
    including the start < radius underflow guard and both anchors."""
    import eqtlbma_b200
    from eqtlbma_b200.synth import make_dataset
    for anchor, radius in [("TSS", 100), ("TSS+TES", 150), ("TSS", 5000), ("TSS+TES", 1)]:
        ds = make_dataset(seed=3, n_genes=40, snps_per_gene=5, n_inds=20, n_chr=3, anchor=anchor, radius=radius)
        ds.radius = radius
        eng = eqtlbma_b200.Engine(ds, analysis="sep", bfs="gen")
        ora = AnyEngine(oracle_lib, "eqo_", ds, analysis="sep", bfs="gen")
        assert np.array_equal(eng.cis_begin, ora.cis_begin)
        assert np.array_equal(eng.cis_end, ora.cis_end)
        eng.close()
        ora.close()
