"""Size-independent properties of the CUDA path at bench-scale shapes (GPU):
 * the split K1/K2+K3 fast path and the general fused kernel are two independent implementations of
   the same statistics: they must agree on every pair of a large workload;
 * results do not depend on how genes are sharded (gene ranges processed separately = all at once);
 * permutation statistics do not depend on the sharding either, and the batched-GEMM permutation path agrees
   with the CPU oracle when every permutation changes the kept rows (ragged individuals with covariates)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ds(**kw):
    from eqtlbma_b200.synth import make_dataset
    base = dict(seed=77, n_subgroups=3, n_inds=300, n_genes=400, snps_per_gene=50, n_cov=11, dosage=True,
                radius=100, gene_spacing=201, far_snp=False, n_chr=4)
    base.update(kw)
    return make_dataset(**base)


def test_fast_path_equals_general_path_at_scale(cuda_lib):
    import eqtlbma_b200
    ds = _ds()
    eng = eqtlbma_b200.Engine(ds, analysis="join", bfs="sin")
    assert eng.fast_gene_count() == ds.n_genes
    a = eng.run()
    # same data with ONE NaN per gene in the last subgroup: every gene leaves the fast path; the
    # other subgroups' statistics must be unchanged
    ds2 = _ds()
    for g in range(ds2.n_genes):
        ds2.subgroups[2].Y[g, g % 300] = np.nan
    eng2 = eqtlbma_b200.Engine(ds2, analysis="join", bfs="sin")
    assert eng2.fast_gene_count() == 0
    b = eng2.run()
    assert np.array_equal(a.n[:, :2], b.n[:, :2])
    assert np.allclose(a.sstats[:, :2, 1:], b.sstats[:, :2, 1:], rtol=1e-9, atol=0, equal_nan=True)
    assert np.allclose(a.sstats[:, :2, 0], b.sstats[:, :2, 0], rtol=1e-9, atol=1e-12, equal_nan=True)
    # singleton ABFs of the untouched subgroups (configs 0 and 1) agree to the 1e-8 budget
    assert np.allclose(a.abf_cfg[:, :2], b.abf_cfg[:, :2], rtol=0, atol=1e-8, equal_nan=True)
    assert np.allclose(a.abf_w[:, 5:7], b.abf_w[:, 5:7], rtol=0, atol=1e-8, equal_nan=True)


def test_results_independent_of_gene_sharding(cuda_lib):
    import eqtlbma_b200
    from eqtlbma_b200.shard import gene_costs, partition
    ds = _ds(n_genes=203, n_subgroups=4, n_cov=2, ragged=True)
    eng = eqtlbma_b200.Engine(ds, analysis="join", bfs="all")
    full = eng.run()
    pfull = eng.run_permutations(40, 1859, pbf="all", wrtsize=10)
    sb = partition(cuda_lib, gene_costs(eng.cis_begin, eng.cis_end, 40), 10, 3)
    parts = [eng.run(int(sb[k]), int(sb[k + 1])) for k in range(3)]
    pparts = [eng.run_permutations(40, 1859, lo=int(sb[k]), hi=int(sb[k + 1]), pbf="all", wrtsize=10) for k in range(3)]
    assert np.array_equal(np.concatenate([p.n for p in parts]), full.n)
    assert np.array_equal(np.concatenate([p.abf_w for p in parts]), full.abf_w, equal_nan=True)
    assert np.array_equal(np.concatenate([p.count for p in pparts]), pfull.count)
    assert np.array_equal(np.concatenate([p.perm_stats for p in pparts]), pfull.perm_stats, equal_nan=True)


def test_gemm_permutation_path_matches_oracle_with_covariates(cuda_lib, oracle_lib):
    """Ragged individuals WITH covariates: every permutation changes the kept rows of every subgroup, so the
    per-(gene, permutation) bases are rebuilt (perm_prep_kernel, general case) and x~'x~ comes from the Gram form of
    the batched GEMM; against the CPU oracle."""
    import eqtlbma_b200
    from eqtlbma_b200._capi import Engine as AnyEngine
    ds = _ds(n_genes=30, n_subgroups=5, n_inds=200, n_cov=3, ragged=True, snps_per_gene=30)
    eng = eqtlbma_b200.Engine(ds, analysis="join", bfs="sin")
    ora = AnyEngine(oracle_lib, "eqo_", ds, analysis="join", bfs="sin")
    a = eng.run_permutations(64, 7, pbf="gen-sin", wrtsize=7)
    b = ora.run_permutations(64, 7, pbf="gen-sin", wrtsize=7)
    assert eng.last_perm_timing()["path"] == 1
    assert np.array_equal(a.count, b.count)
    assert np.allclose(a.perm_stats, b.perm_stats, rtol=0, atol=1e-8, equal_nan=True)
    assert np.allclose(a.true_stat, b.true_stat, rtol=0, atol=1e-8, equal_nan=True)


def test_gtex_shape_slice_matches_oracle(cuda_lib, oracle_lib):
    """c3/c4 shape at reduced counts: 9 ragged subgroups, --bfs all (511 configurations), permutations
    with --pbf all and --pbf gen-sin, against the CPU oracle."""
    import eqtlbma_b200
    from eqtlbma_b200._capi import Engine as AnyEngine
    from eqtlbma_b200.synth import make_dataset, make_grid
    ds = make_dataset(seed=99, n_subgroups=9, n_inds=150, n_genes=8, snps_per_gene=6, ragged=True,
                      ragged_min_frac=0.34, gridL=make_grid("general")[:10], n_chr=2)
    eng = eqtlbma_b200.Engine(ds, analysis="join", bfs="all")
    ora = AnyEngine(oracle_lib, "eqo_", ds, analysis="join", bfs="all")
    a, b = eng.run(), ora.run()
    assert eng.n_configs == 511
    assert np.array_equal(a.n, b.n)
    assert np.allclose(a.sstats[..., 1:], b.sstats[..., 1:], rtol=1e-9, atol=0, equal_nan=True)
    assert np.allclose(a.abf_gen, b.abf_gen, rtol=0, atol=1e-8, equal_nan=True)
    assert np.allclose(a.abf_cfg, b.abf_cfg, rtol=0, atol=1e-8, equal_nan=True)
    assert np.allclose(a.abf_w, b.abf_w, rtol=0, atol=1e-8, equal_nan=True)
    for pbf in ("all", "gen-sin"):
        pa = eng.run_permutations(25, 1859, pbf=pbf, wrtsize=3)
        pb = ora.run_permutations(25, 1859, pbf=pbf, wrtsize=3)
        assert np.array_equal(pa.count, pb.count)
        assert np.allclose(pa.perm_stats, pb.perm_stats, rtol=0, atol=1e-8, equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("n_sub,n_inds,n_cov", [(3, 300, 11), (9, 120, 6), (9, 450, 2), (2, 700, 3), (2, 1800, 2)])
def test_subgroup_specific_covariates_match_oracle(cuda_lib, oracle_lib, n_sub, n_inds, n_cov):
    """c2 shape with tissue-specific covariate values (no duplicate-subgroup shortcut: one basis per
    subgroup, several DMMA column tiles per launch), --bfs sin, plus permutations, against the CPU oracle."""
    import eqtlbma_b200
    from eqtlbma_b200._capi import Engine as AnyEngine
    from eqtlbma_b200.synth import make_dataset
    ds = make_dataset(seed=7 + n_sub, n_subgroups=n_sub, n_inds=n_inds, n_genes=12, snps_per_gene=9, n_cov=n_cov,
                      cov_per_subgroup=True, dosage=True, ragged=(n_sub == 9), n_chr=2, radius=100, gene_spacing=201,
                      far_snp=False)
    eng = eqtlbma_b200.Engine(ds, analysis="join", bfs="sin")
    ora = AnyEngine(oracle_lib, "eqo_", ds, analysis="join", bfs="sin")
    assert eng.fast_gene_count() == ds.n_genes
    a, b = eng.run(), ora.run()
    assert np.array_equal(a.n, b.n)
    assert np.allclose(a.sstats[..., 1:], b.sstats[..., 1:], rtol=1e-9, atol=0, equal_nan=True)
    assert np.allclose(a.sstats[..., 0], b.sstats[..., 0], rtol=1e-9, atol=1e-12, equal_nan=True)
    assert np.allclose(a.abf_gen, b.abf_gen, rtol=0, atol=1e-8, equal_nan=True)
    assert np.allclose(a.abf_cfg, b.abf_cfg, rtol=0, atol=1e-8, equal_nan=True)
    assert np.allclose(a.abf_w, b.abf_w, rtol=0, atol=1e-8, equal_nan=True)
    pa = eng.run_permutations(20, 1859, pbf="gen-sin", wrtsize=5)
    pb = ora.run_permutations(20, 1859, pbf="gen-sin", wrtsize=5)
    assert np.array_equal(pa.count, pb.count)


@pytest.mark.gpu
@pytest.mark.parametrize("analysis,bfs", [("join", "sin"), ("join", "all"), ("sep", "gen")])
def test_upload_pipeline_matches_plain_path(cuda_lib, analysis, bfs, monkeypatch):
    """Enough SNPs for the chunked upload pipeline (genes finish in genotype-chunk order and their results
    are scattered into PINNED host arrays by a kernel): bit-identical to the plain path (pageable arrays,
    DMA copy, gene order), for the whole range and for a sub-range of genes."""
    import eqtlbma_b200
    from eqtlbma_b200.synth import make_dataset
    ds = make_dataset(seed=21, n_subgroups=3, n_inds=40, n_genes=700, snps_per_gene=14, n_cov=2, cov_per_subgroup=True,
                      dosage=True, n_chr=3, radius=100, gene_spacing=201, far_snp=False, absent_gene_frac=0.05)
    assert ds.n_snps >= 8192
    eng = eqtlbma_b200.Engine(ds, analysis=analysis, bfs=bfs)
    plain = eng.run()
    pinned = eng.run(out=eng.alloc_results(pinned=True))
    for a, b in ((plain.n, pinned.n), (plain.sstats, pinned.sstats), (plain.abf_gen, pinned.abf_gen),
                 (plain.abf_cfg, pinned.abf_cfg), (plain.abf_w, pinned.abf_w)):
        assert (a is None) == (b is None)
        if a is not None:
            assert np.array_equal(a, b, equal_nan=True)
    sub = eng.run(lo=100, hi=420, out=eng.alloc_results(100, 420, pinned=True))
    off = eng.pair_offsets()
    assert np.array_equal(sub.sstats, plain.sstats[off[100]:off[420]], equal_nan=True)
    if plain.abf_w is not None:
        assert np.array_equal(sub.abf_w, plain.abf_w[off[100]:off[420]], equal_nan=True)
    eng.close()
    # first run of a fresh context goes through the pipeline while the upload is still in flight
    eng2 = eqtlbma_b200.Engine(ds, analysis=analysis, bfs=bfs)
    first = eng2.run(out=eng2.alloc_results(pinned=True))
    assert np.array_equal(first.sstats, plain.sstats, equal_nan=True)
    if plain.abf_w is not None:
        assert np.array_equal(first.abf_w, plain.abf_w, equal_nan=True)
    eng2.close()


@pytest.mark.gpu
def test_table_math_against_cuda_library(cuda_lib):
    """rcp_n / log_tab16 / rsqrt_newton1 / exp_tab16 (perm_gemm.cuh: the permutation BF kernel's elementary functions)
    against 1/x, log, rsqrt, exp, exp10 of the CUDA math library over 3e6 pseudo-random arguments (wide range and around
    1) and the special values.  Budget: a log10 BF must be good to 1e-8 absolute, i.e. ~2e-8 relative on a linear-domain
    sum and ~2e-8 absolute on a natural logarithm; the forms below are three to four orders of magnitude inside it."""
    import ctypes
    out = (ctypes.c_double * 5)()
    f = cuda_lib.eqb_math_selftest
    f.restype = ctypes.c_int
    assert f(ctypes.c_int32(0), ctypes.c_int64(3_000_000), out) == 0
    rcp_rel, log_abs, rsqrt_rel, exp_rel, special = list(out)
    assert rcp_rel < 2e-12
    assert log_abs < 1e-12
    assert rsqrt_rel < 2e-12
    assert exp_rel < 1e-10
    assert special == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("dosage,dtype,denom", [(True, np.uint16, 1000.0), (False, np.uint8, 1.0)])
def test_fixed_point_genotype_transport_is_lossless(cuda_lib, dosage, dtype, denom):
    """eqb_set_genotypes_fixed (u8 hard calls / u16 numerators of 10^d for a d-decimal dosage file) must give
    bit-identical results to the double matrix: k / 10^d correctly rounded IS the parsed double."""
    import copy
    import eqtlbma_b200
    ds = _ds(n_genes=120, dosage=dosage)
    a = eqtlbma_b200.Engine(ds, analysis="join", bfs="sin").run(raw=True)
    ds2 = copy.copy(ds)
    ds2.genos = []
    for G in ds.genos:
        k = np.rint(G * denom)
        assert np.array_equal(k / denom, G)
        ds2.genos.append(k.astype(dtype))
    ds2.geno_denoms = [denom] * len(ds.genos)
    b = eqtlbma_b200.Engine(ds2, analysis="join", bfs="sin").run(raw=True)
    assert np.array_equal(a.n, b.n)
    for f in ("sstats", "abf_gen", "abf_cfg", "abf_w"):
        assert np.array_equal(getattr(a, f), getattr(b, f), equal_nan=True), f


@pytest.mark.gpu
def test_fixed_point_rejects_bad_arguments(cuda_lib):
    import ctypes as C
    import eqtlbma_b200
    ds = _ds(n_genes=8)
    eng = eqtlbma_b200.Engine(ds, analysis="join", bfs="gen")
    f = cuda_lib.eqb_set_genotypes_fixed
    f.restype = C.c_int
    G = np.zeros((ds.n_snps, ds.genos[0].shape[1]), dtype=np.uint8)
    # wrong element width, non-integer denominator, and a call after eqb_finalize() all fail with a message
    assert f(eng.ctx, 0, G.ctypes.data_as(C.c_void_p), 4, C.c_double(1.0), C.c_int64(G.shape[0]), G.shape[1]) != 0
    assert f(eng.ctx, 0, G.ctypes.data_as(C.c_void_p), 1, C.c_double(2.5), C.c_int64(G.shape[0]), G.shape[1]) != 0
    assert f(eng.ctx, 0, G.ctypes.data_as(C.c_void_p), 1, C.c_double(1.0), C.c_int64(G.shape[0]), G.shape[1]) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["s44_sin", "big_grid", "separate_files", "sep_analysis"])
def test_warp_kernel_variants_match_oracle(cuda_lib, oracle_lib, case):
    """fast_pair_warp_kernel instantiations the c2 bench does not reach: 44 subgroups (--bfs sin, the c5 shape: six
    groups of 8 subgroups in the DMMA contraction), a grid too large for the constant-bank tables (tables read from
    global memory), separate genotype files per subgroup (shuffle-reduced fallback contraction) and --analys sep."""
    import eqtlbma_b200
    from eqtlbma_b200._capi import Engine as AnyEngine
    from eqtlbma_b200.synth import make_dataset
    kw = dict(seed=91, n_subgroups=3, n_inds=150, n_genes=14, snps_per_gene=13, n_cov=2, dosage=True, n_chr=2, radius=100,
              gene_spacing=201, far_snp=False)
    analysis, bfs = "join", "sin"
    if case == "s44_sin":
        kw.update(n_subgroups=44, n_inds=90, n_cov=1, ragged=True, ragged_min_frac=0.5)
    elif case == "big_grid":
        rng = np.random.default_rng(5)
        gl = np.abs(rng.normal(0.3, 0.3, size=(70, 2))) + 1e-3   # 70 points: 3L = 210 > 192 entries
        gl[::7, 1] = 0.0
        gl[3::9, 0] = 0.0
        kw.update(gridL=gl, gridS=np.abs(rng.normal(0.3, 0.3, size=(40, 2))) + 1e-3)   # K = 40 > 32
    elif case == "separate_files":
        kw.update(separate_geno_files=True, missing_geno_frac=0.1)
    else:
        analysis, bfs = "sep", "gen"
    ds = make_dataset(**kw)
    eng = eqtlbma_b200.Engine(ds, analysis=analysis, bfs=bfs)
    ora = AnyEngine(oracle_lib, "eqo_", ds, analysis=analysis, bfs=bfs)
    assert eng.fast_gene_count() > 0
    a, b = eng.run(), ora.run()
    assert np.array_equal(a.n, b.n)
    assert np.allclose(a.sstats[..., 1:], b.sstats[..., 1:], rtol=1e-9, atol=0, equal_nan=True)
    assert np.allclose(a.sstats[..., 0], b.sstats[..., 0], rtol=1e-9, atol=1e-12, equal_nan=True)
    if analysis == "join":
        assert np.allclose(a.abf_gen, b.abf_gen, rtol=0, atol=1e-8, equal_nan=True)
        assert np.allclose(a.abf_cfg, b.abf_cfg, rtol=0, atol=1e-8, equal_nan=True)
        assert np.allclose(a.abf_w, b.abf_w, rtol=0, atol=1e-8, equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("n_sub,K,n_genes,spg,absent", [
    (9, 10, 30, 70, 0.0),   # c3 shape, 3 pairs per tile, > 592 tiles: every persistent CTA loops over several tiles
    (10, 16, 6, 9, 0.2),    # 4 + 3 + 3 subgroups, 1,023 configurations, 2 pairs per tile, genes absent from subgroups
    (6, 5, 10, 7, 0.2),     # 2 + 2 + 2, 4 pairs per tile (32 / 5 capped), ragged last tile
    (4, 11, 9, 5, 0.0),     # 2 + 1 + 1, 2 pairs per tile, one chunk of 15 configurations
    (2, 3, 7, 3, 0.3),      # 1 + 1 + 0
    (1, 10, 5, 4, 0.0),     # a single subgroup: one configuration
])
def test_bfs_all_kernel_shapes_match_oracle(cuda_lib, oracle_lib, n_sub, K, n_genes, spg, absent):
    """fast_pair_all_kernel (fast_all_kernel.cuh) over its shape space -- parts of the subset-sum tables, pairs per tile, chunk
    counts, tiles per CTA, absent subgroups -- against the CPU oracle: raw ABFs of every configuration, their grid averages,
    BMAlite and BMA (gene_snp_pair.cpp:504-602)."""
    import eqtlbma_b200
    from eqtlbma_b200._capi import Engine as AnyEngine
    from eqtlbma_b200.synth import make_dataset, make_grid
    rng = np.random.default_rng(5 + n_sub)
    gridS = np.column_stack([rng.choice([0.0, 0.01, 0.04, 0.16, 0.64], size=K), rng.choice([0.0, 0.0025, 0.04, 0.3, 2.56], size=K)])
    ds = make_dataset(seed=300 + n_sub, n_subgroups=n_sub, n_inds=90, n_genes=n_genes, snps_per_gene=spg, ragged=True,
                      ragged_min_frac=0.5, absent_gene_frac=absent, gridL=make_grid("general")[:7], gridS=gridS, n_chr=2,
                      radius=100, gene_spacing=201, far_snp=False)
    eng = eqtlbma_b200.Engine(ds, analysis="join", bfs="all")
    ora = AnyEngine(oracle_lib, "eqo_", ds, analysis="join", bfs="all")
    a, b = eng.run(), ora.run()
    assert eng.n_configs == 2 ** n_sub - 1
    assert np.array_equal(a.offsets, b.offsets) and np.array_equal(a.n, b.n)
    assert np.allclose(a.sstats[..., 1:], b.sstats[..., 1:], rtol=1e-9, atol=0, equal_nan=True)
    assert np.allclose(a.abf_gen, b.abf_gen, rtol=0, atol=1e-8, equal_nan=True)
    assert np.allclose(a.abf_cfg, b.abf_cfg, rtol=0, atol=1e-8, equal_nan=True)
    assert np.allclose(a.abf_w, b.abf_w, rtol=0, atol=1e-8, equal_nan=True)
    # the averaged-only call (no raw arrays) takes the linear-domain instantiation of the kernel: same averages
    c = eng.run(raw=False)
    assert np.allclose(c.abf_w, b.abf_w, rtol=0, atol=1e-8, equal_nan=True)
    assert np.allclose(c.abf_w, a.abf_w, rtol=0, atol=1e-9, equal_nan=True)
    eng.close()
    ora.close()
