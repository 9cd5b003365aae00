"""GPU parity tests of the hierarchical-model EM (eqb_hm_* of include/eqtlbma_hm_b200.h) against
(a) the full-precision dumps of the UNMODIFIED reference eqtlbma_hm (tests/golden/hm/, oracle/make_golden_hm.py),
(b) the numpy restatement (oracle/hm_oracle.py) on larger seeded inputs.
Tolerances: log10 likelihood 1e-9 relative; weights 1e-8 relative (they are fixed points of a map iterated 10-25
times, differences of the summation order are not amplified but carried along); per-pair log10 BFs 1e-9 absolute;
posteriors 1e-8 relative; iteration counts and the profile-likelihood interval ticks exact."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from hm_oracle import HmOracle  # noqa: E402  (checker only)
from hm_scenarios import GOLDEN_HM, HM_SCENARIOS, build_dataset, initial_params, load_hm_dump, model_arrays  # noqa: E402

pytestmark = pytest.mark.gpu


def _engine(B, gene_off):
    from eqtlbma_b200.hm import HmEngine
    hm = HmEngine(B.shape[1], B.shape[2])
    hm.append(B, gene_off)
    hm.finalize()
    return hm


@pytest.mark.parametrize("name", sorted(HM_SCENARIOS))
def test_hm_matches_reference_dump(cuda_lib, name):
    from eqtlbma_b200.hm import HmFit
    sc = HM_SCENARIOS[name]
    ds = build_dataset(sc)
    B, names = model_arrays(sc, ds)
    ref = load_hm_dump(name)
    hm = _engine(B, ds.gene_off)
    assert (hm.n_genes, hm.n_pairs) == (ds.n_genes, ds.n_pairs)
    pi0, gw, cp, fixed = initial_params(sc, B.shape[1], B.shape[2])
    fit = hm.em(HmFit(pi0, gw, cp), thresh=sc.get("thresh", 0.05), maxit=sc.get("maxit"), stepmax=sc.get("msl", 1.0), fixed=fixed)
    with open(os.path.join(GOLDEN_HM, name + ".json")) as f:
        meta = json.load(f)
    iter_lines = [ln for ln in fit.log_lines if ln.startswith("iter ")]
    assert len(iter_lines) == meta["n_iter_lines"]
    # the reference's own last progress line (printed with 4 significant digits) is reproduced character by character
    # up to the last digit of each field
    ref_last, got_last = meta["last_iter_line"].split(), iter_lines[-1].split()
    assert len(ref_last) == len(got_last)
    for a, b in zip(ref_last, got_last):
        try:
            fa, fb = float(a), float(b)
        except ValueError:
            assert a == b
            continue
        assert abs(fa - fb) <= 2e-4 * max(abs(fa), 1e-3) + 1e-6
    assert abs(fit.loglik - ref["loglik"]) <= 1e-9 * max(1.0, abs(ref["loglik"]))
    assert abs(fit.pi0 - ref["pi0"][0]) <= 1e-8
    np.testing.assert_allclose(fit.grid_wts, ref["grid"], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(fit.config_prior, ref["config"], rtol=1e-8, atol=1e-12)
    if sc.get("getbf"):
        post = hm.posteriors(fit)
        np.testing.assert_allclose(post["gene_post"], [g[2] for g in ref["genes"]], rtol=1e-8, atol=1e-13)
        np.testing.assert_allclose(post["gene_bf"], [g[3] for g in ref["genes"]], rtol=0, atol=1e-9)
        np.testing.assert_allclose(post["snp_bf"], [s[1] for s in ref["snps"]], rtol=0, atol=1e-9)
        np.testing.assert_allclose(post["snp_post"], [s[2] for s in ref["snps"]], rtol=1e-8, atol=1e-13)
        np.testing.assert_allclose(post["cfg_bf"], [s[3] for s in ref["snps"]], rtol=0, atol=1e-9)
        np.testing.assert_allclose(post["gene_cfg_post"], ref["postcfg"], rtol=1e-8, atol=1e-13)
    if sc.get("getci"):
        hm.profile_ci(fit)
        # intervals move in ticks of 0.001 from the estimate (eqtlbma_hm.cpp:1348-1573): same number of ticks
        assert abs(fit.pi0_ci[0] - ref["pi0"][1]) <= 1e-7 and abs(fit.pi0_ci[1] - ref["pi0"][2]) <= 1e-7
        np.testing.assert_allclose(fit.grid_ci[:, 0], ref["grid_left"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(fit.grid_ci[:, 1], ref["grid_right"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(fit.config_ci[:, 0], ref["config_left"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(fit.config_ci[:, 1], ref["config_right"], rtol=0, atol=1e-7)
    assert hm.launch_count > 0
    hm.close()


@pytest.mark.parametrize("shape", [dict(n_genes=3000, snps_lo=1, snps_hi=40, n_subgroups=3, grid=10),
                                   dict(n_genes=400, snps_lo=50, snps_hi=3000, n_subgroups=2, grid=5),   # many units per gene
                                   dict(n_genes=500, snps_lo=1, snps_hi=30, n_subgroups=5, grid=25),    # dim 31
                                   dict(n_genes=300, snps_lo=1, snps_hi=20, n_subgroups=6, grid=7),     # dim 63, generic grid
                                   dict(n_genes=200, snps_lo=1, snps_hi=5, n_subgroups=3, grid=1, singletons_only=True)])
def test_hm_likelihood_and_esums_match_oracle(cuda_lib, shape):
    """eqb_hm_loglik / eqb_hm_esums (the three device operations) on ragged shapes against the numpy restatement."""
    from eqtlbma_b200.hm_synth import make_hm_dataset
    ds = make_hm_dataset(seed=77, round_text=False, **shape)
    o = HmOracle(ds.B, ds.gene_off)
    hm = _engine(ds.B, ds.gene_off)
    rs = np.random.RandomState(5)
    for trial in range(3):
        pi0 = float(rs.uniform(0.05, 0.95))
        gw = rs.dirichlet(np.ones(ds.grid))
        cp = rs.dirichlet(np.ones(ds.dim) * 0.5)
        lik = hm.loglik(pi0, gw, cp, keep=True)
        ref = o.loglik(pi0, gw, cp)
        assert abs(lik - ref) <= 1e-10 * abs(ref)
        sums = hm.esums(pi0, gw, cp)
        n_pi0, n_gw, n_cp = o.fixedpoint(pi0, gw, cp, dict(pi0=False, grid=False, configs=False))
        assert abs(sums[0] / ds.n_genes - n_pi0) <= 1e-10
        if ds.dim > 1:
            t = sums[1:1 + ds.dim] + np.log10(cp)
            got = 10.0 ** (t - np.log10(np.sum(10.0 ** (t - t.max()))) - t.max())
            np.testing.assert_allclose(got, n_cp, rtol=1e-9, atol=1e-15)
        t = sums[1 + ds.dim:] + np.log10(gw)
        got = 10.0 ** (t - np.log10(np.sum(10.0 ** (t - t.max()))) - t.max())
        np.testing.assert_allclose(got, n_gw, rtol=1e-9, atol=1e-15)
    hm.close()


def test_hm_extreme_values_take_the_clamped_path(cuda_lib):
    """Values outside +-1e6 (finite): the data set is not `ranged` and every exponential is clamped; same results as
    the restatement (10^(x - max) underflows to zero there)."""
    from eqtlbma_b200.hm_synth import make_hm_dataset
    ds = make_hm_dataset(seed=78, n_genes=300, snps_lo=1, snps_hi=30, round_text=False)
    B = ds.B.copy()
    rs = np.random.RandomState(9)
    idx = rs.choice(B.size, 200, replace=False)
    B.ravel()[idx[:100]] = -5e6 * rs.uniform(1, 1e3, 100)
    B.ravel()[idx[100:150]] = -1e300
    B.ravel()[idx[150:]] = 3e7 * rs.uniform(1, 5, 50)
    o = HmOracle(B, ds.gene_off)
    hm = _engine(B, ds.gene_off)
    gw = rs.dirichlet(np.ones(ds.grid))
    cp = rs.dirichlet(np.ones(ds.dim))
    lik, ref = hm.loglik(0.3, gw, cp, keep=True), o.loglik(0.3, gw, cp)
    assert abs(lik - ref) <= 1e-10 * abs(ref)
    sums = hm.esums(0.3, gw, cp)
    n_pi0, n_gw, n_cp = o.fixedpoint(0.3, gw, cp, dict(pi0=False, grid=False, configs=False))
    assert abs(sums[0] / ds.n_genes - n_pi0) <= 1e-10
    t = sums[1 + ds.dim:] + np.log10(gw)
    np.testing.assert_allclose(10.0 ** (t - np.log10(np.sum(10.0 ** (t - t.max()))) - t.max()), n_gw, rtol=1e-9, atol=1e-15)
    hm.close()


def test_hm_em_monotone_and_append_in_pieces(cuda_lib):
    """Size-independent properties on a larger input: the likelihood never decreases along the EM (the reference aborts
    otherwise, eqtlbma_hm.cpp:1091-1095), weights stay on the simplex, and loading the genes file by file
    (eqb_hm_append called three times) gives bit-identical results to loading them at once."""
    from eqtlbma_b200.hm import HmEngine, HmFit
    from eqtlbma_b200.hm_synth import make_hm_dataset
    ds = make_hm_dataset(seed=3, n_genes=20000, snps_lo=1, snps_hi=60, round_text=False)
    a = _engine(ds.B, ds.gene_off)
    b = HmEngine(ds.dim, ds.grid)
    cuts = [0, 7000, 7001, ds.n_genes]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        p0, p1 = int(ds.gene_off[lo]), int(ds.gene_off[hi])
        b.append(ds.B[p0:p1], ds.gene_off[lo:hi + 1] - p0)
    b.finalize()
    fits = []
    for e in (a, b):
        fit = e.em(HmFit(0.5, np.full(ds.grid, 1.0 / ds.grid), np.full(ds.dim, 1.0 / ds.dim)), thresh=0.01, stepmax=1.0)
        liks = [float(ln.split()[3]) for ln in fit.log_lines if ln.startswith("iter ")]
        assert all(y >= x - 1e-6 for x, y in zip(liks[:-1], liks[1:])) and len(liks) > 5
        assert abs(fit.grid_wts.sum() - 1) < 1e-12 and abs(fit.config_prior.sum() - 1) < 1e-12 and 0 < fit.pi0 < 1
        fits.append(fit)
    assert fits[0].loglik == fits[1].loglik and fits[0].pi0 == fits[1].pi0
    assert np.array_equal(fits[0].grid_wts, fits[1].grid_wts) and np.array_equal(fits[0].config_prior, fits[1].config_prior)
    o = HmOracle(ds.B, ds.gene_off)
    ref = o.loglik(fits[0].pi0, fits[0].grid_wts, fits[0].config_prior)
    assert abs(fits[0].loglik - ref) <= 1e-10 * abs(ref)
    a.close()
    b.close()


def test_hm_rejects_nonfinite_and_unfinalized(cuda_lib):
    from eqtlbma_b200.hm import HmEngine
    from eqtlbma_b200.hm_synth import make_hm_dataset
    ds = make_hm_dataset(seed=4, n_genes=20)
    B = ds.B.copy()
    B[5, 2, 3] = np.nan
    hm = HmEngine(ds.dim, ds.grid)
    hm.append(B, ds.gene_off)
    with pytest.raises(RuntimeError, match="NaN"):
        hm.finalize()
    hm.close()
    hm = HmEngine(ds.dim, ds.grid)
    hm.append(ds.B, ds.gene_off)
    with pytest.raises(RuntimeError, match="finalize"):
        hm.loglik(0.5, np.full(ds.grid, 1.0 / ds.grid), np.full(ds.dim, 1.0 / ds.dim))
    with pytest.raises(RuntimeError):
        hm.append(ds.B, np.array([0, 3, 3, ds.n_pairs], dtype=np.int64))  # a gene without pairs
    hm.close()


def test_hm_from_device_resident_raw_abfs(cuda_lib):
    """eqtlbma_bf -> eqtlbma_hm without the text round trip: the raw ABFs of eqb_run stay on the device
    (eqb_raw_abfs_device) and are appended device-to-device; the fit is bit-identical to the one on the host copies."""
    import eqtlbma_b200
    from eqtlbma_b200.hm import HmEngine, HmFit
    from eqtlbma_b200.synth import make_dataset
    ds = make_dataset(seed=21, n_subgroups=3, n_inds=120, n_genes=60, snps_per_gene=6)
    eng = eqtlbma_b200.Engine(ds, analysis="join", bfs="all")
    res = eng.run()
    ptr, n_pairs, ids, off = eng.raw_abfs_device()
    C, K = res.abf_cfg.shape[1], res.abf_cfg.shape[2]
    assert n_pairs == res.abf_cfg.shape[0] and off[-1] == n_pairs
    fits = []
    for mode in ("device", "host"):
        hm = HmEngine(C, K)
        if mode == "device":
            hm.append_device(ptr, n_pairs, off)
        else:
            hm.append(res.abf_cfg, off)
        hm.finalize()
        fits.append(hm.em(HmFit(0.5, np.full(K, 1.0 / K), np.full(C, 1.0 / C)), thresh=0.05))
        hm.close()
    assert fits[0].loglik == fits[1].loglik and fits[0].pi0 == fits[1].pi0 and fits[0].iters == fits[1].iters
    assert np.array_equal(fits[0].config_prior, fits[1].config_prior)
    o = HmOracle(res.abf_cfg, off)
    ref = o.loglik(fits[0].pi0, fits[0].grid_wts, fits[0].config_prior)
    assert abs(fits[0].loglik - ref) <= 1e-10 * max(1.0, abs(ref))
    eng.close()
