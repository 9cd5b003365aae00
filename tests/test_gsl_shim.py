"""Pins the GSL shim (oracle/gsl_shim) against INDEPENDENT implementations.

Every golden fixture of this repository is the unmodified reference linked against oracle/gsl_shim instead of GNU GSL
(absent from the image), and the CPU oracle links the same shim: an error in the shim would be common-mode.  These tests
check the primitives the hot path reaches -- gsl_cdf_tdist_P/Q (gene_snp_pair.cpp:274, utils_math.cpp:202),
gsl_cdf_ugaussian_Pinv (gene_snp_pair.cpp:274, utils_math.cpp:92), gsl_cdf_fdist_Q / gsl_cdf_chisq_Qinv (MVLR.cpp:404-405),
gsl_multifit_linear + _rank (utils_math.cpp:189-190), gsl_rng_mt19937 / gsl_rng_uniform_int / gsl_ran_shuffle
(gene.cpp:394,523,639), gsl_sort_index (utils_math.cpp:87), gsl_combination_next (gene_snp_pair.cpp:476-548),
gsl_sf_choose (gene_snp_pair.cpp:591), gsl_linalg_SV_decomp / gsl_linalg_LU_* (the hybrid error model: utils_math.cpp:249-301,
gene_snp_pair.cpp:793-812,1196-1224) -- against mpmath (50 digits), scipy.stats, numpy.linalg and numpy's own MT19937.
Residual risk that remains: real GSL's large-nu Cornish-Fisher shortcut inside gsl_cdf_tdist_P is deliberately not
emulated, so the shim is, if anything, closer to the exact distribution than GSL."""
import ctypes as C
import itertools

import numpy as np
import pytest


@pytest.fixture(scope="module")
def shim(oracle_lib):
    lib = oracle_lib
    for name, nargs in (("gsl_cdf_tdist_P", 2), ("gsl_cdf_tdist_Q", 2), ("gsl_cdf_ugaussian_Pinv", 1), ("gsl_cdf_ugaussian_P", 1),
                        ("gsl_cdf_fdist_Q", 3), ("gsl_cdf_chisq_Qinv", 2), ("gsl_cdf_chisq_Q", 2), ("gsl_stats_tss", None)):
        f = getattr(lib, name)
        f.restype = C.c_double
        if nargs:
            f.argtypes = [C.c_double] * nargs
    lib.gsl_sf_choose.restype = C.c_double
    lib.gsl_sf_choose.argtypes = [C.c_uint, C.c_uint]
    return lib


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def test_student_t_cdf_against_mpmath(shim):
    import mpmath as mp
    mp.mp.dps = 50
    worst = 0.0
    for nu in (1, 2, 3, 7, 28, 147, 286, 397, 1000, 5000):
        for x in (0.0, 1e-9, 0.3, 1.0, 2.5, 6.0, 12.0, 25.0, 40.0, 80.0):
            # P(T <= -x) = I_{nu/(nu+x^2)}(nu/2, 1/2) / 2  (exact, regularised incomplete beta)
            tail = mp.betainc(mp.mpf(nu) / 2, mp.mpf(1) / 2, 0, mp.mpf(nu) / (nu + mp.mpf(x) ** 2), regularized=True) / 2
            for sign in (-1.0, 1.0):
                P = shim.gsl_cdf_tdist_P(sign * x, float(nu))
                Q = shim.gsl_cdf_tdist_Q(sign * x, float(nu))
                eP = tail if sign < 0 else 1 - tail
                eQ = 1 - tail if sign < 0 else tail
                if tail < mp.mpf("1e-300"):  # below the double range: the shim must underflow to (nearly) zero as well
                    assert min(P, Q) < 1e-290
                    continue
                worst = max(worst, float(abs(P - eP) / eP), float(abs(Q - eQ) / eQ))
    assert worst < 5e-13, worst


def test_student_t_cdf_against_scipy(shim):
    from scipy import stats
    rng = np.random.default_rng(0)
    for nu in (5.0, 50.0, 287.0):
        x = rng.normal(0, 4, 200)
        got = np.array([shim.gsl_cdf_tdist_P(float(v), nu) for v in x])
        assert np.allclose(got, stats.t.cdf(x, nu), rtol=1e-11, atol=0)
        gotq = np.array([shim.gsl_cdf_tdist_Q(float(v), nu) for v in x])
        assert np.allclose(gotq, stats.t.sf(x, nu), rtol=1e-11, atol=0)


def _exact_normal_quantile(mp, P, x0):
    """root of log Phi(x) = log P (no cancellation in either tail: mpmath's ncdf goes through erfc)"""
    P = mp.mpf(P)
    if P > mp.mpf("0.5"):
        return -mp.findroot(lambda x: mp.log(mp.ncdf(x)) - mp.log(1 - P), -mp.mpf(x0))
    return mp.findroot(lambda x: mp.log(mp.ncdf(x)) - mp.log(P), mp.mpf(x0))


def test_normal_quantile_against_mpmath(shim):
    import mpmath as mp
    mp.mp.dps = 50
    worst = 0.0
    for P in (1e-300, 1e-200, 1e-100, 1e-50, 1e-20, 1e-10, 1e-5, 0.01, 0.02425, 0.1, 0.3, 0.7, 0.97575, 0.99, 1 - 1e-10):
        got = shim.gsl_cdf_ugaussian_Pinv(P)
        exact = _exact_normal_quantile(mp, P, got)
        worst = max(worst, float(abs(got - exact) / abs(exact)))
    assert worst < 1e-13, worst
    assert shim.gsl_cdf_ugaussian_Pinv(0.5) == 0.0
    # the composition used by the standardisation (gene_snp_pair.cpp:274): z = Pinv(T_nu(-|t|)), far into the tail
    for t, nu in ((5.0, 147.0), (20.0, 286.0), (37.0, 397.0)):
        p = mp.betainc(mp.mpf(nu) / 2, mp.mpf(1) / 2, 0, mp.mpf(nu) / (nu + mp.mpf(t) ** 2), regularized=True) / 2
        got = shim.gsl_cdf_ugaussian_Pinv(shim.gsl_cdf_tdist_P(-t, nu))
        exact = _exact_normal_quantile(mp, p, got)
        assert abs(got - exact) / abs(exact) < 1e-13


def test_f_and_chisq_against_scipy(shim):
    from scipy import stats
    for x, n1, n2 in ((0.5, 1, 190), (3.0, 1, 190), (25.0, 1, 196), (120.0, 1, 150), (1e-3, 1, 99), (7.7, 3, 40)):
        assert _rel(shim.gsl_cdf_fdist_Q(x, float(n1), float(n2)), stats.f.sf(x, n1, n2)) < 1e-10
    for Q, nu in ((0.5, 1), (0.05, 1), (1e-6, 1), (1e-30, 1), (0.999, 1), (0.2, 4)):
        assert _rel(shim.gsl_cdf_chisq_Qinv(Q, float(nu)), stats.chi2.isf(Q, nu)) < 1e-10
        assert _rel(shim.gsl_cdf_chisq_Q(stats.chi2.isf(Q, nu), float(nu)), Q) < 1e-10
    assert shim.gsl_sf_choose(9, 4) == 126.0 and shim.gsl_sf_choose(20, 10) == 184756.0 and shim.gsl_sf_choose(5, 0) == 1.0


class GslVector(C.Structure):
    _fields_ = [("size", C.c_size_t), ("stride", C.c_size_t), ("data", C.POINTER(C.c_double)), ("block", C.c_void_p),
                ("owner", C.c_int)]


class GslMatrix(C.Structure):
    _fields_ = [("size1", C.c_size_t), ("size2", C.c_size_t), ("tda", C.c_size_t), ("data", C.POINTER(C.c_double)),
                ("block", C.c_void_p), ("owner", C.c_int)]


def _multifit(lib, X, y):
    n, p = X.shape
    lib.gsl_matrix_alloc.restype = C.POINTER(GslMatrix)
    lib.gsl_vector_alloc.restype = C.POINTER(GslVector)
    lib.gsl_multifit_linear_alloc.restype = C.c_void_p
    lib.gsl_multifit_linear_rank.restype = C.c_size_t
    mX, cov = lib.gsl_matrix_alloc(C.c_size_t(n), C.c_size_t(p)), lib.gsl_matrix_alloc(C.c_size_t(p), C.c_size_t(p))
    vy, vc = lib.gsl_vector_alloc(C.c_size_t(n)), lib.gsl_vector_alloc(C.c_size_t(p))
    for i in range(n):
        vy.contents.data[i] = y[i]
        for j in range(p):
            mX.contents.data[i * mX.contents.tda + j] = X[i, j]
    w = C.c_void_p(lib.gsl_multifit_linear_alloc(C.c_size_t(n), C.c_size_t(p)))
    chisq = C.c_double(0)
    rc = lib.gsl_multifit_linear(mX, vy, vc, cov, C.byref(chisq), w)
    assert rc == 0
    rank = lib.gsl_multifit_linear_rank(C.c_double(2.2204460492503131e-16), w)
    c = np.array([vc.contents.data[j] for j in range(p)])
    cv = np.array([[cov.contents.data[i * cov.contents.tda + j] for j in range(p)] for i in range(p)])
    lib.gsl_multifit_linear_free(w)
    for m in (mX, cov):
        lib.gsl_matrix_free(m)
    for v in (vy, vc):
        lib.gsl_vector_free(v)
    return c, cv, chisq.value, int(rank)


def _balanced_min_norm(X, y):
    """numpy restatement of gsl_multifit_linear for a rank-deficient design: columns scaled by the power of two that
    brings their 1-norm into (0.5, 1] (gsl_linalg_balance_columns), minimum-norm least squares on the balanced matrix with
    singular values <= DBL_EPSILON * s_0 dropped, coefficients unscaled."""
    D = np.ones(X.shape[1])
    for j in range(X.shape[1]):
        s = np.abs(X[:, j]).sum()
        if s == 0 or not np.isfinite(s):
            continue
        f = 1.0
        while s > 1.0:
            s /= 2.0
            f *= 2.0
        while s < 0.5:
            s *= 2.0
            f /= 2.0
        D[j] = f
    U, sv, Vt = np.linalg.svd(X / D, full_matrices=False)
    keep = sv > 2.2204460492503131e-16 * sv[0]
    c = (Vt.T[:, keep] / sv[keep]) @ (U[:, keep].T @ y)
    return c / D, int(keep.sum())


def test_multifit_linear_against_numpy(shim):
    rng = np.random.default_rng(3)
    for n, q in ((200, 0), (300, 11), (57, 3), (15, 10)):
        g = rng.choice(3, n, p=[0.49, 0.42, 0.09]).astype(float)
        X = np.column_stack([np.ones(n), g] + [rng.normal(0, 1, n) for _ in range(q)])
        y = 4 + 0.3 * g + rng.normal(0, 1, n)
        c, cov, chisq, rank = _multifit(shim, X, y)
        ref, res, rk, _ = np.linalg.lstsq(X, y, rcond=None)
        assert rank == X.shape[1] == rk
        assert np.allclose(c, ref, rtol=1e-10, atol=1e-12)
        r = y - X @ ref
        assert abs(chisq - r @ r) <= 1e-10 * (r @ r)
        s2 = (r @ r) / (n - rank)
        assert np.allclose(cov, s2 * np.linalg.inv(X.T @ X), rtol=1e-8, atol=1e-14)
    # rank-deficient designs: constant genotype, and a covariate equal to the genotype
    n = 40
    for X in (np.column_stack([np.ones(n), np.full(n, 2.0)]),
              np.column_stack([np.ones(n), np.arange(n) % 3, (np.arange(n) % 3).astype(float), rng.normal(0, 1, n)])):
        y = rng.normal(3, 1, n)
        c, cov, chisq, rank = _multifit(shim, X.astype(float), y)
        ref, rk = _balanced_min_norm(X.astype(float), y)
        assert rank == rk == X.shape[1] - 1
        assert np.allclose(c, ref, rtol=1e-9, atol=1e-12)
        r = y - X @ c
        r0 = y - X @ np.linalg.lstsq(X, y, rcond=None)[0]
        assert abs(r @ r - r0 @ r0) <= 1e-10 * (r0 @ r0)   # it IS a least-squares solution
    t = rng.normal(0, 3, 50)
    shim.gsl_stats_tss.argtypes = [C.POINTER(C.c_double), C.c_size_t, C.c_size_t]
    assert abs(shim.gsl_stats_tss((C.c_double * 50)(*t), 1, 50) - ((t - t.mean()) ** 2).sum()) < 1e-10


class GslRng(C.Structure):
    _fields_ = [("type", C.c_void_p), ("state", C.c_void_p)]


def _rng(lib, seed):
    lib.gsl_rng_alloc.restype = C.POINTER(GslRng)
    lib.gsl_rng_alloc.argtypes = [C.c_void_p]
    lib.gsl_rng_get.restype = C.c_ulong
    lib.gsl_rng_uniform_int.restype = C.c_ulong
    lib.gsl_rng_uniform_int.argtypes = [C.POINTER(GslRng), C.c_ulong]
    lib.gsl_rng_set.argtypes = [C.POINTER(GslRng), C.c_ulong]
    lib.gsl_ran_flat.restype = C.c_double
    lib.gsl_ran_flat.argtypes = [C.POINTER(GslRng), C.c_double, C.c_double]
    T = C.c_void_p.in_dll(lib, "gsl_rng_mt19937")
    r = lib.gsl_rng_alloc(T)
    lib.gsl_rng_set(r, seed)
    return r


def _numpy_stream(seed, n):
    bg = np.random.MT19937()
    bg._legacy_seeding(np.uint32(seed if seed != 0 else 4357))  # gsl_rng_set(r, 0) uses 4357 (mt19937 2002 seeding)
    return [int(v) for v in bg.random_raw(n)]


def test_mt19937_stream_uniform_int_and_shuffle(shim):
    # known answers: first output of MT19937 for init_genrand(5489), and GSL's documented default (seed 0 -> 4357)
    r = _rng(shim, 5489)
    assert shim.gsl_rng_get(r) == 3499211612
    r = _rng(shim, 0)
    assert shim.gsl_rng_get(r) == 4293858116
    for seed in (1859, 1, 0, 123456789):
        r = _rng(shim, seed)
        assert [shim.gsl_rng_get(r) for _ in range(2000)] == _numpy_stream(seed, 2000)
    # gsl_rng_uniform_int(n): scale = 0xffffffff / n; k = get() / scale until k < n.  gsl_ran_shuffle: for i = n-1..1:
    # j = uniform_int(i + 1); swap(i, j) (GSL manual, "Shuffling and Sampling"), replayed on numpy's own generator
    seed, N, reps = 1859, 450, 7
    raw = iter(_numpy_stream(seed, 20 * N * reps))

    def uniform_int(n):
        scale = 0xFFFFFFFF // n
        while True:
            k = next(raw) // scale
            if k < n:
                return k

    expect = list(range(N))
    r = _rng(shim, seed)
    buf = (C.c_size_t * N)(*range(N))
    shim.gsl_ran_shuffle.argtypes = [C.POINTER(GslRng), C.c_void_p, C.c_size_t, C.c_size_t]
    for _ in range(reps):  # cumulative shuffles, as Gene::MakePermutationsJoin applies them (gene.cpp:639)
        for i in range(N - 1, 0, -1):
            j = uniform_int(i + 1)
            expect[i], expect[j] = expect[j], expect[i]
        shim.gsl_ran_shuffle(r, buf, N, C.sizeof(C.c_size_t))
        assert list(buf) == expect
    # gsl_ran_flat(a, b) = a (1 - u) + b u with u = get() / 2^32 (Gene::CalcPermutationPvalue, gene.cpp:358)
    u = next(raw) / 4294967296.0
    assert shim.gsl_ran_flat(r, 0.25, 0.75) == 0.25 * (1 - u) + 0.75 * u


def test_sort_index_and_combinations(shim):
    rng = np.random.default_rng(9)
    x = np.round(rng.normal(0, 1, 300), 1)  # many ties: the tie order of the index heapsort is part of --qnorm
    p = (C.c_size_t * 300)()
    shim.gsl_sort_index.argtypes = [C.POINTER(C.c_size_t), C.POINTER(C.c_double), C.c_size_t, C.c_size_t]
    shim.gsl_sort_index(p, (C.c_double * 300)(*x), 1, 300)
    idx = np.array(list(p))
    assert sorted(idx.tolist()) == list(range(300)) and np.all(np.diff(x[idx]) >= 0)

    class Comb(C.Structure):
        _fields_ = [("n", C.c_size_t), ("k", C.c_size_t), ("data", C.POINTER(C.c_size_t))]

    shim.gsl_combination_calloc.restype = C.POINTER(Comb)
    for n, k in ((5, 2), (9, 4), (3, 3)):
        c = shim.gsl_combination_calloc(C.c_size_t(n), C.c_size_t(k))
        got = [tuple(c.contents.data[i] for i in range(k))]
        while shim.gsl_combination_next(c) == 0:
            got.append(tuple(c.contents.data[i] for i in range(k)))
        assert got == list(itertools.combinations(range(n), k))  # lexicographic order (configuration names / order)
        shim.gsl_combination_free(c)


class GslPermutation(C.Structure):
    _fields_ = [("size", C.c_size_t), ("data", C.POINTER(C.c_size_t))]


def _to_gsl(lib, A):
    lib.gsl_matrix_alloc.restype = C.POINTER(GslMatrix)
    m = lib.gsl_matrix_alloc(C.c_size_t(A.shape[0]), C.c_size_t(A.shape[1]))
    for i in range(A.shape[0]):
        for j in range(A.shape[1]):
            m.contents.data[i * m.contents.tda + j] = A[i, j]
    return m


def _from_gsl(m):
    c = m.contents
    return np.array([[c.data[i * c.tda + j] for j in range(c.size2)] for i in range(c.size1)])


def test_svd_and_lu_against_numpy(shim):
    """gsl_linalg_SV_decomp and gsl_linalg_LU_decomp / _invert / _lndet: what the hybrid error model reaches through
    mygsl_linalg_pseudoinverse, mygsl_linalg_invert and CalcLog10AbfMvlr (utils_math.cpp:249-301, gene_snp_pair.cpp:793-812,
    1196-1224).  Singular values, the pseudo-inverse V D^-1 U' and the LU inverse / log-determinant against numpy.linalg."""
    lib = shim
    lib.gsl_vector_alloc.restype = C.POINTER(GslVector)
    lib.gsl_permutation_alloc.restype = C.POINTER(GslPermutation)
    lib.gsl_linalg_LU_lndet.restype = C.c_double
    rng = np.random.default_rng(11)
    for n, p in ((70, 2), (120, 4), (6, 6), (3, 3)):
        A = rng.normal(size=(n, p))
        A[:, 0] = 1.0  # an intercept column, like every design matrix of the path
        U, V = _to_gsl(lib, A), _to_gsl(lib, np.zeros((p, p)))
        S, work = lib.gsl_vector_alloc(C.c_size_t(p)), lib.gsl_vector_alloc(C.c_size_t(p))
        assert lib.gsl_linalg_SV_decomp(U, V, S, work) == 0
        sv = np.array([S.contents.data[j] for j in range(p)])
        Um, Vm = _from_gsl(U), _from_gsl(V)
        assert np.allclose(np.sort(sv)[::-1], np.linalg.svd(A, compute_uv=False), rtol=1e-12, atol=0)
        assert np.allclose((Um * sv) @ Vm.T, A, rtol=0, atol=1e-12 * np.abs(A).max() * n)
        assert np.allclose((Vm / sv) @ Um.T, np.linalg.pinv(A), rtol=1e-10, atol=1e-13)
        for m in (U, V):
            lib.gsl_matrix_free(m)
        for v in (S, work):
            lib.gsl_vector_free(v)
    for n in (2, 3, 5, 9):
        B = rng.normal(size=(n, n)) + n * np.eye(n)
        LU, inv = _to_gsl(lib, B), _to_gsl(lib, np.zeros((n, n)))
        perm = lib.gsl_permutation_alloc(C.c_size_t(n))
        sign = C.c_int(0)
        assert lib.gsl_linalg_LU_decomp(LU, perm, C.byref(sign)) == 0
        assert abs(lib.gsl_linalg_LU_lndet(LU) - np.linalg.slogdet(B)[1]) < 1e-12
        assert lib.gsl_linalg_LU_invert(LU, perm, inv) == 0
        assert np.allclose(_from_gsl(inv), np.linalg.inv(B), rtol=1e-11, atol=1e-14)
        lib.gsl_permutation_free(perm)
        for m in (LU, inv):
            lib.gsl_matrix_free(m)
