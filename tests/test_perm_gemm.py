"""GPU tests of the batched-GEMM permutation path (perm_gemm.cu): the TMA / DMMA product kernel against a plain
device reference, and the whole path (prep -> GEMM -> BF -> merge) against the CPU oracle on the shapes the golden
scenarios do not reach: the GTEx shape (9 ragged subgroups, --pbf gen / gen-sin / all), complete individuals with
covariates (fixed basis, permutation-invariant x~'x~), several column batches (more than 512 permutations), separate
genotype files, --analys sep with --permsep 1 / 2, --maxbf.  Tolerances: exceedance counts exact, log10 ABF
statistics 1e-8 absolute, minimum p-values 1e-9 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_rows,n_cols,ldn", [(300, 200, 208), (1000, 129, 464), (128, 128, 16), (77, 5, 48)])
def test_gemm_kernel_matches_plain_product(cuda_lib, n_rows, n_cols, ldn):
    import eqtlbma_b200
    r = eqtlbma_b200.selftest_perm_gemm(n_rows, n_cols, ldn)
    assert r["tiles"] > 0
    assert r["worst_rel_err"] < 1e-14, r


def test_fp64_peaks_are_measurable(cuda_lib):
    import eqtlbma_b200
    pk = eqtlbma_b200.measure_fp64_peaks()
    assert pk["dfma_tflops"] > 1.0 and pk["dmma_tflops"] > 1.0, pk


def _cmp(eng, ora, nperm, join=True, **kw):
    a = eng.run_permutations(nperm, 1859, **kw)
    b = ora.run_permutations(nperm, 1859, **kw)
    assert eng.last_perm_timing()["path"] == 1   # the GEMM path ran (not the general fused kernel)
    assert np.array_equal(a.count, b.count)
    assert np.array_equal(a.nperm_done, b.nperm_done)
    tol = dict(rtol=0, atol=1e-8) if join else dict(rtol=1e-9, atol=0)
    assert np.allclose(a.true_stat, b.true_stat, equal_nan=True, **tol)
    assert np.allclose(a.perm_stats, b.perm_stats, equal_nan=True, **tol)
    assert np.allclose(a.pval, b.pval, rtol=1e-12, atol=0, equal_nan=True)


@pytest.mark.parametrize("pbf,maxbf", [("gen", False), ("gen-sin", False), ("all", False), ("all", True)])
def test_gtex_shape_permutations_match_oracle(cuda_lib, oracle_lib, pbf, maxbf):
    import eqtlbma_b200
    from eqtlbma_b200._capi import Engine as AnyEngine
    from eqtlbma_b200.synth import make_dataset, make_grid
    ds = make_dataset(seed=5, n_subgroups=9, n_inds=160, n_genes=4, snps_per_gene=140, ragged=True, ragged_min_frac=0.34,
                      gridL=make_grid("general")[:10], n_chr=2, radius=1000, gene_spacing=2001, far_snp=False)
    bfs = "all" if pbf == "all" else "sin"
    eng = eqtlbma_b200.Engine(ds, analysis="join", bfs=bfs)
    ora = AnyEngine(oracle_lib, "eqo_", ds, analysis="join", bfs=bfs)
    _cmp(eng, ora, 24, pbf=pbf, maxbf=maxbf, wrtsize=3)


def test_complete_individuals_with_covariates_and_many_column_batches(cuda_lib, oracle_lib):
    """c2 shape: every sample in every subgroup (fixed K1 basis, only y~ depends on the permutation), 600 permutations
    = two column batches, trick 2 stop rule applied on the full statistics."""
    import eqtlbma_b200
    from eqtlbma_b200._capi import Engine as AnyEngine
    from eqtlbma_b200.synth import make_dataset
    ds = make_dataset(seed=11, n_subgroups=3, n_inds=90, n_genes=9, snps_per_gene=7, n_cov=3, cov_per_subgroup=True,
                      dosage=True, n_chr=2, radius=100, gene_spacing=201, far_snp=False)
    eng = eqtlbma_b200.Engine(ds, analysis="join", bfs="sin")
    ora = AnyEngine(oracle_lib, "eqo_", ds, analysis="join", bfs="sin")
    _cmp(eng, ora, 600, pbf="gen-sin", wrtsize=4, trick=2, tricut=10)


@pytest.mark.parametrize("permsep", [1, 2])
def test_separate_analysis_and_separate_genotype_files(cuda_lib, oracle_lib, permsep):
    import eqtlbma_b200
    from eqtlbma_b200._capi import Engine as AnyEngine
    from eqtlbma_b200.synth import make_dataset
    ds = make_dataset(seed=13, n_subgroups=3, n_inds=80, n_genes=10, snps_per_gene=9, n_cov=1, ragged=True,
                      separate_geno_files=True, missing_geno_frac=0.1, n_chr=2, radius=100, gene_spacing=201, far_snp=False,
                      absent_gene_frac=0.1)
    eng = eqtlbma_b200.Engine(ds, analysis="sep", bfs="gen")
    ora = AnyEngine(oracle_lib, "eqo_", ds, analysis="sep", bfs="gen")
    _cmp(eng, ora, 50, join=False, permsep=permsep, wrtsize=4)


def test_nan_expression_values_change_the_kept_rows(cuda_lib, oracle_lib):
    """NaN expression values: the kept rows of a subgroup follow the permutation even when every sample is present."""
    import eqtlbma_b200
    from eqtlbma_b200._capi import Engine as AnyEngine
    from eqtlbma_b200.synth import make_dataset
    ds = make_dataset(seed=17, n_subgroups=3, n_inds=70, n_genes=8, snps_per_gene=6, n_cov=2, n_chr=2, radius=100,
                      gene_spacing=201, far_snp=False)
    rng = np.random.default_rng(3)
    for sg in ds.subgroups:
        for g in range(0, ds.n_genes, 2):
            sg.Y[g, rng.integers(0, sg.Y.shape[1], size=4)] = np.nan
    eng = eqtlbma_b200.Engine(ds, analysis="join", bfs="all")
    ora = AnyEngine(oracle_lib, "eqo_", ds, analysis="join", bfs="all")
    _cmp(eng, ora, 30, pbf="all", wrtsize=3)
