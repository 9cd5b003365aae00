/*
 * eqtlbma_oracle.cpp -- CPU ORACLE of the eqtlbma_bf hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain, loop-by-loop restatement of the reference's algorithm (timflutre/eqtlbma v1.3.3) over
 * the flat layouts of include/eqtlbma_b200.h.  Each function cites the reference file:line it
 * follows.  The numerical primitives the reference takes from GNU GSL (absent from this image)
 * come from oracle/gsl_shim (MT19937 / shuffle / combination bit-exact, floating routines to
 * ~1e-14).  Parity pin: tests/test_oracle_vs_reference.py compares this file, on identical
 * inputs and seeds, with the UNMODIFIED reference compiled into oracle/_ref/ (full-precision
 * dumps, committed under tests/golden/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product path never does.
 */
#include "eqtlbma_oracle.h"
#include "gsl_shim/gsl_shim.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const double kNaN = std::numeric_limits<double>::quiet_NaN();
const double kInf = std::numeric_limits<double>::infinity();

struct Sub {
  int geno_id = 0, n_exp_cols = 0, Q = 0, n_cov_cols = 0;
  std::vector<int32_t> all2geno, all2exp, all2cov;
  std::vector<uint8_t> snp_has, gene_has;
  std::vector<double> Y, C;
};

/* everything GeneSnpPair keeps for one pair (gene_snp_pair.hpp:57-75) */
struct Pair {
  std::vector<int> n;          /* subgroup2samplesize_, 0 = no entry */
  std::vector<int> ncov;       /* subgroup2nbcovariates_ */
  std::vector<double> pve, sigmahat, betahat, sebetahat, pval;
  std::vector<std::vector<double> > raw_gen; /* 3 x L */
  std::vector<std::vector<double> > raw_cfg; /* C x K */
  double w_gen[3];
  double w_gensin, w_all;
  std::vector<double> w_cfg;
  explicit Pair(int S)
      : n(S, 0), ncov(S, 0), pve(S, kNaN), sigmahat(S, kNaN), betahat(S, kNaN), sebetahat(S, kNaN),
        pval(S, kNaN), w_gensin(kNaN), w_all(kNaN)
  {
    w_gen[0] = w_gen[1] = w_gen[2] = kNaN;
  }
  bool has(int s) const { return n[s] > 0; } /* HasResults, gene_snp_pair.cpp:61-68 */
};

} // namespace

struct eqo_ctx {
  eqb_config cfg;
  std::vector<std::vector<double> > genos;
  std::vector<int> geno_cols;
  std::vector<Sub> subs;
  std::vector<double> phi2L, oma2L, phi2S, oma2S;
  std::vector<int64_t> cb, ce;
  std::vector<std::vector<int> > configs_all; /* gamma vectors, k = 1..S lexicographic */
  std::string err;
  int threads = 1;
  bool fatal = false;
};

namespace {

bool is_nan(double x) { return !(x == x); } /* utils_math.cpp:42 */

/* utils::log10_weighted_sum, uniform weights (utils_math.cpp:100-131) */
double log10_weighted_sum(const double *vec, size_t size)
{
  double max = vec[0];
  for (size_t i = 0; i < size; ++i)
    if (vec[i] > max) max = vec[i];
  const double w = (double)(1 / ((double)size));
  double sum = 0.0;
  for (size_t i = 0; i < size; ++i) {
    if (is_nan(vec[i])) continue;
    sum += w * pow(10, vec[i] - max);
  }
  double res = max + log10(sum);
  if (std::abs(res) <= DBL_EPSILON) res = 0.0;
  return res;
}

/* utils::log10_weighted_sum, given weights (utils_math.cpp:135-159) */
double log10_weighted_sum(const double *vec, const double *weights, size_t size)
{
  double max = vec[0];
  for (size_t i = 0; i < size; ++i)
    if (vec[i] > max) max = vec[i];
  double sum = 0.0;
  for (size_t i = 0; i < size; ++i) {
    if (is_nan(vec[i])) continue;
    sum += weights[i] * pow(10, vec[i] - max);
  }
  double res = max + log10(sum);
  if (std::abs(res) <= DBL_EPSILON) res = 0.0;
  return res;
}

/* utils::qqnorm (utils_math.cpp:80-96) */
void qqnorm(double *data, size_t n)
{
  std::vector<size_t> order(n);
  gsl_sort_index(order.data(), data, 1, n);
  const double a = (n <= 10 ? 0.375 : 0.5);
  for (size_t i = 0; i < n; ++i) {
    const double q = (i + 1 - a) / (n + 1 - 2 * a);
    data[order[i]] = gsl_cdf_ugaussian_Pinv(q);
  }
}

/* utils::median (utils_math.hpp:61-80) */
double median(std::vector<double> v)
{
  if (v.empty()) return kNaN;
  const size_t size = v.size(), mid = size / 2;
  std::nth_element(v.begin(), v.begin() + mid, v.end());
  if (size % 2 != 0) return v[mid];
  const double a = v[mid];
  std::nth_element(v.begin(), v.begin() + mid - 1, v.end());
  return (a + v[mid - 1]) / 2.0;
}

/* GeneSnpPair::FillStlContainers for ONE subgroup (gene_snp_pair.cpp:94-169):
 * expression index from the permuted sample, genotype and covariate index from the unpermuted one */
bool gather(eqo_ctx *c, int64_t g, int64_t m, int s, const size_t *perm, std::vector<double> &y,
            std::vector<double> &x, std::vector<std::vector<double> > &cov)
{
  const Sub &sb = c->subs[s];
  const int N_all = c->cfg.n_samples_all;
  const double *Yg = &sb.Y[(size_t)g * sb.n_exp_cols];
  const double *Gm = &c->genos[sb.geno_id][(size_t)m * c->geno_cols[sb.geno_id]];
  y.clear();
  x.clear();
  std::vector<int> kept;
  for (int i = 0; i < N_all; ++i) {
    const int idx_all = perm ? (int)perm[i] : i;
    const int e = sb.all2exp[idx_all];
    const int gi = sb.all2geno[i];
    if (e >= 0 && gi >= 0 && !is_nan(Yg[e])) {
      y.push_back(Yg[e]);
      x.push_back(Gm[gi]);
      kept.push_back(i);
    }
  }
  if (y.empty()) return false;
  if (c->cfg.qnorm) qqnorm(y.data(), y.size());
  cov.assign(sb.Q, std::vector<double>());
  for (int q = 0; q < sb.Q; ++q) {
    for (size_t r = 0; r < kept.size(); ++r) {
      const int ci = sb.all2cov.empty() ? -1 : sb.all2cov[kept[r]];
      if (ci < 0) { /* gene_snp_pair.cpp:138-144: fatal in the reference */
        c->fatal = true;
        c->err = "missing covariate for a sample kept in the regression";
        cov[q].push_back(kNaN);
      } else
        cov[q].push_back(sb.C[(size_t)q * sb.n_cov_cols + ci]);
    }
  }
  return true;
}

/* utils::FitSingleGeneWithSingleSnp (utils_math.cpp:166-209) */
void fit_ols(const std::vector<double> &yv, const std::vector<double> &xv,
             const std::vector<std::vector<double> > &cov, double &pve, double &sigmahat,
             double &betahat, double &sebetahat, double &pval)
{
  const size_t N = yv.size(), P = 2 + cov.size();
  if (N < P + 1) {
    pve = sigmahat = betahat = sebetahat = pval = kNaN;
    return;
  }
  gsl_matrix *X = gsl_matrix_alloc(N, P);
  gsl_vector *y = gsl_vector_alloc(N);
  for (size_t i = 0; i < N; ++i) {
    gsl_vector_set(y, i, yv[i]);
    gsl_matrix_set(X, i, 0, 1.0);
    gsl_matrix_set(X, i, 1, xv[i]);
    for (size_t j = 0; j < cov.size(); ++j) gsl_matrix_set(X, i, j + 2, cov[j][i]);
  }
  gsl_vector *B = gsl_vector_alloc(P);
  gsl_matrix *covB = gsl_matrix_alloc(P, P);
  gsl_multifit_linear_workspace *work = gsl_multifit_linear_alloc(N, P);
  double rss;
  gsl_multifit_linear(X, y, B, covB, &rss, work);
  const size_t rank = gsl_multifit_linear_rank(GSL_DBL_EPSILON, work);
  pve = 1 - rss / gsl_stats_tss(y->data, y->stride, y->size);
  sigmahat = sqrt(rss / (double)(N - rank));
  betahat = gsl_vector_get(B, 1);
  sebetahat = sqrt(gsl_matrix_get(covB, 1, 1));
  pval = 2 * gsl_cdf_tdist_Q(fabs(betahat / sebetahat), N - rank);
  gsl_vector_free(B);
  gsl_matrix_free(covB);
  gsl_multifit_linear_free(work);
  gsl_matrix_free(X);
  gsl_vector_free(y);
}

/* GeneSnpPair::CalcSstatsOneSbgrp, normal likelihood (gene_snp_pair.cpp:175-208) */
void calc_sstats_one(eqo_ctx *c, int64_t g, int64_t m, int s, const size_t *perm, Pair &pr)
{
  std::vector<double> y, x;
  std::vector<std::vector<double> > cov;
  if (!gather(c, g, m, s, perm, y, x, cov)) return;
  pr.n[s] = (int)y.size();
  pr.ncov[s] = (int)cov.size();
  fit_ols(y, x, cov, pr.pve[s], pr.sigmahat[s], pr.betahat[s], pr.sebetahat[s], pr.pval[s]);
}

/* GeneSnpPair::StandardizeSstatsAndCorrectSmallSampleSize (gene_snp_pair.cpp:256-290) */
void standardize(const Pair &pr, int S, std::vector<std::vector<double> > &std_)
{
  std_.assign(S, std::vector<double>(3, kNaN));
  for (int s = 0; s < S; ++s) {
    if (!pr.has(s)) continue;
    const double N = pr.n[s];
    double bhat = pr.betahat[s] / pr.sigmahat[s], sebhat = pr.sebetahat[s] / pr.sigmahat[s],
           t = bhat / sebhat;
    if (is_nan(t)) continue;
    const double nu = N - 2 - pr.ncov[s];
    t = gsl_cdf_gaussian_Pinv(gsl_cdf_tdist_P(-fabs(bhat / sebhat), nu), 1.0);
    if (fabs(t) > 1e-8) {
      const double sigmahat = fabs(pr.betahat[s]) / (fabs(t) * sebhat);
      bhat = pr.betahat[s] / sigmahat;
      sebhat = fabs(bhat / t);
    } else {
      bhat = 0;
      sebhat = kInf;
    }
    std_[s][0] = bhat;
    std_[s][1] = sebhat;
    std_[s][2] = t;
  }
}

/* CalcLog10AbfUvlr (gene_snp_pair.cpp:297-356) */
double abf_uvlr(const std::vector<int> &gamma, const std::vector<std::vector<double> > &st,
                double phi2, double oma2)
{
  double l10AbfAll = 0.0, num = 0.0, denom = 0.0, varbbarhat = 0.0;
  std::vector<double> singles;
  for (size_t s = 0; s < gamma.size(); ++s) {
    if (gamma[s] == 0) continue;
    const double bhat = st[s][0], varbhat = pow(st[s][1], 2), t = st[s][2];
    double l;
    if (fabs(t) < 1e-8)
      l = 0;
    else {
      num += bhat / (varbhat + phi2);
      denom += 1 / (varbhat + phi2);
      varbbarhat += 1 / (varbhat + phi2);
      l = 0.5 * log10(varbhat) - 0.5 * log10(varbhat + phi2) +
          (0.5 * pow(t, 2) * phi2 / (varbhat + phi2)) / log(10);
    }
    singles.push_back(l);
  }
  const double bbarhat = (denom != 0.0) ? num / denom : 0.0;
  varbbarhat = (varbbarhat != 0.0) ? 1 / varbbarhat : kInf;
  if (bbarhat != 0.0 && varbbarhat < kInf) {
    const double T2 = pow(bbarhat, 2.0) / varbbarhat;
    const double lbar = (T2 != 0) ? 0.5 * log10(varbbarhat) - 0.5 * log10(varbbarhat + oma2) +
                                        (0.5 * T2 * oma2 / (varbbarhat + oma2)) / log(10)
                                  : 0;
    l10AbfAll = lbar;
    for (size_t i = 0; i < singles.size(); ++i) l10AbfAll += singles[i];
  } else
    l10AbfAll = 0.0;
  return l10AbfAll;
}

/* gsl_combination enumeration used everywhere configurations are listed
 * (gene_snp_pair.cpp:469-485,504-550; eqtlbma_bf.cpp:1190-1222) */
void enumerate_configs(int S, bool singletons_only, std::vector<std::vector<int> > &out)
{
  out.clear();
  for (int k = 1; k <= S; ++k) {
    gsl_combination *comb = gsl_combination_calloc(S, k);
    while (true) {
      std::vector<int> gamma(S, 0);
      for (int i = 0; i < k; ++i) gamma[gsl_combination_get(comb, i)] = 1;
      out.push_back(gamma);
      if (gsl_combination_next(comb) != GSL_SUCCESS) break;
    }
    gsl_combination_free(comb);
    if (singletons_only) break;
  }
}

/* CalcAbfsUvlrForConsistentConfiguration (gene_snp_pair.cpp:364-416) */
void abfs_consistent(eqo_ctx *c, Pair &pr, const std::vector<std::vector<double> > &std_)
{
  const int S = c->cfg.n_subgroups;
  const size_t L = c->phi2L.size();
  std::vector<int> gamma(S, 0);
  std::vector<std::vector<double> > st;
  for (int s = 0; s < S; ++s) {
    if (pr.has(s)) {
      gamma[s] = 1;
      st.push_back(std_[s]);
    } else
      st.push_back(std::vector<double>(3, 0));
  }
  pr.raw_gen.assign(3, std::vector<double>(L, kNaN));
  for (size_t k = 0; k < L; ++k) {
    pr.raw_gen[0][k] = abf_uvlr(gamma, st, c->phi2L[k], c->oma2L[k]);
    pr.raw_gen[1][k] = abf_uvlr(gamma, st, 0.0, c->phi2L[k] + c->oma2L[k]);
    pr.raw_gen[2][k] = abf_uvlr(gamma, st, c->phi2L[k] + c->oma2L[k], 0.0);
  }
  for (int j = 0; j < 3; ++j)
    pr.w_gen[j] = L ? log10_weighted_sum(pr.raw_gen[j].data(), L) : kNaN;
}

/* CalcAbfsUvlrForSingletons (gene_snp_pair.cpp:422-464) / ForEachConfiguration (:504-550) */
void abfs_configs(eqo_ctx *c, Pair &pr, const std::vector<std::vector<double> > &std_,
                  bool singletons_only)
{
  const int S = c->cfg.n_subgroups;
  const size_t K = c->phi2S.size();
  std::vector<std::vector<int> > configs;
  enumerate_configs(S, singletons_only, configs);
  pr.raw_cfg.assign(configs.size(), std::vector<double>(K, 0.0));
  pr.w_cfg.assign(configs.size(), kNaN);
  for (size_t ci = 0; ci < configs.size(); ++ci) {
    std::vector<int> gamma = configs[ci];
    std::vector<std::vector<double> > st;
    bool any = false;
    for (int s = 0; s < S; ++s) {
      if (gamma[s] == 1 && pr.has(s)) {
        st.push_back(std_[s]);
        any = true;
      } else {
        gamma[s] = 0;
        st.push_back(std::vector<double>(3, 0.0));
      }
    }
    /* singleton of a subgroup without results: vector of zeros (gene_snp_pair.cpp:436-457);
     * a general configuration whose active subgroups all lack results evaluates to zeros too */
    if (!singletons_only || any)
      for (size_t k = 0; k < K; ++k) pr.raw_cfg[ci][k] = abf_uvlr(gamma, st, c->phi2S[k], c->oma2S[k]);
    pr.w_cfg[ci] = log10_weighted_sum(pr.raw_cfg[ci].data(), K);
  }
}

/* CalcBMAlite (gene_snp_pair.cpp:552-570): singletons are the first S configurations */
void bma_lite(Pair &pr, int S)
{
  std::vector<double> v, w;
  for (int s = 0; s < S; ++s) {
    v.push_back(pr.w_cfg[s]);
    w.push_back((1.0 / 2.0) * (1.0 / S));
  }
  v.push_back(pr.w_gen[0]);
  w.push_back(1.0 / 2.0);
  pr.w_gensin = log10_weighted_sum(v.data(), w.data(), v.size());
}

/* CalcBMA (gene_snp_pair.cpp:572-602) */
void bma(Pair &pr, int S, const std::vector<std::vector<int> > &configs)
{
  std::vector<double> v, w;
  for (size_t ci = 0; ci < configs.size(); ++ci) {
    int k = 0;
    for (int s = 0; s < S; ++s) k += configs[ci][s];
    v.push_back(pr.w_cfg[ci]);
    w.push_back((1.0 / (double)S) * (1.0 / gsl_sf_choose(S, k)));
  }
  pr.w_all = log10_weighted_sum(v.data(), w.data(), v.size());
}

/* GeneSnpPair::CalcAbfsUvlr (gene_snp_pair.cpp:604-622); which: EQB_PBF_* / EQB_BFS_*+1 semantics:
 * 1 = "gen", 2 = contains "sin", 3 = "all" */
void calc_abfs_uvlr(eqo_ctx *c, Pair &pr, int which)
{
  const int S = c->cfg.n_subgroups;
  std::vector<std::vector<double> > std_;
  standardize(pr, S, std_);
  abfs_consistent(c, pr, std_);
  if (which == 2) {
    abfs_configs(c, pr, std_, true);
    bma_lite(pr, S);
  } else if (which == 3) {
    abfs_configs(c, pr, std_, false);
    bma_lite(pr, S);
    bma(pr, S, c->configs_all);
  }
}

/* ------------------------------------------------------------------ MVLR (MVLR.cpp) */

struct Mat {
  int r, c;
  std::vector<double> a;
  Mat() : r(0), c(0) {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  double &operator()(int i, int j) { return a[(size_t)i * c + j]; }
  double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
};

Mat mul(const Mat &A, bool ta, const Mat &B, bool tb)
{
  const int M = ta ? A.c : A.r, K = ta ? A.r : A.c, N = tb ? B.r : B.c;
  Mat C(M, N);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      double acc = 0.0;
      for (int k = 0; k < K; ++k) acc += (ta ? A(k, i) : A(i, k)) * (tb ? B(j, k) : B(k, j));
      C(i, j) = acc;
    }
  return C;
}

/* LU decomposition helpers through the shim (gsl_linalg_LU_decomp / _invert / _lndet) */
Mat lu_inverse(const Mat &A, double *lndet)
{
  const int n = A.r;
  gsl_matrix *t = gsl_matrix_alloc(n, n), *inv = gsl_matrix_calloc(n, n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) gsl_matrix_set(t, i, j, A(i, j));
  gsl_permutation *pp = gsl_permutation_alloc(n);
  int ss;
  gsl_linalg_LU_decomp(t, pp, &ss);
  if (lndet) *lndet = gsl_linalg_LU_lndet(t);
  gsl_linalg_LU_invert(t, pp, inv);
  Mat R(n, n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) R(i, j) = gsl_matrix_get(inv, i, j);
  gsl_permutation_free(pp);
  gsl_matrix_free(t);
  gsl_matrix_free(inv);
  return R;
}

struct Mvlr {
  int s, n, q, m;
  double alpha; /* sigma_option = --fiterr */
  Mat Y, g, Xc, T, Sigma0, Sigma0_inv;

  /* MVLR::init + compute_common + compute_Sigma_null (MVLR.cpp:26-77,102-144,176-195) */
  void init(const std::vector<std::vector<double> > &Yin, const std::vector<double> &xg,
            const std::vector<std::vector<double> > &cov, double fiterr)
  {
    s = (int)Yin.size();
    n = (int)Yin[0].size();
    q = (int)cov.size() + 1;
    alpha = fiterr;
    Y = Mat(n, s);
    g = Mat(n, 1);
    Xc = Mat(n, q);
    for (int i = 0; i < s; ++i)
      for (int j = 0; j < n; ++j) Y(j, i) = Yin[i][j];
    for (int j = 0; j < n; ++j) {
      g(j, 0) = xg[j];
      Xc(j, 0) = 1.0;
      for (int i = 1; i < q; ++i) Xc(j, i) = cov[i - 1][j];
    }
    m = q + s + 1;
    Mat XtX = mul(Xc, true, Xc, false), XtXi(q, q);
    if (q == 1)
      XtXi(0, 0) = 1.0 / XtX(0, 0);
    else
      XtXi = lu_inverse(XtX, NULL);
    Mat t1 = mul(XtXi, false, Xc, true);
    Mat t2 = mul(Xc, false, t1, false);
    T = Mat(n, n);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) T(i, j) = -t2(i, j) + (i == j ? 1.0 : 0.0);
    Mat TY = mul(T, false, Y, false);
    Sigma0 = mul(Y, true, TY, false);
    const double sc = 1.0 / (n + m - q - s - 1);
    for (size_t i = 0; i < Sigma0.a.size(); ++i) Sigma0.a[i] *= sc;
    Sigma0_inv = lu_inverse(Sigma0, NULL);
  }

  /* MVLR::compute_residual (MVLR.cpp:323-437) for one column of Y */
  std::vector<double> residual(int col, bool with_g, double &factor)
  {
    const int size = q + (with_g ? 1 : 0);
    Mat X(n, size), y(n, 1);
    for (int j = 0; j < n; ++j) {
      for (int k = 0; k < q; ++k) X(j, k) = Xc(j, k);
      if (with_g) X(j, q) = g(j, 0);
      y(j, 0) = Y(j, col);
    }
    Mat XtX = mul(X, true, X, false);
    gsl_matrix *A = gsl_matrix_alloc(size, size), *V = gsl_matrix_calloc(size, size);
    gsl_vector *Sv = gsl_vector_calloc(size), *work = gsl_vector_calloc(size);
    for (int i = 0; i < size; ++i)
      for (int j = 0; j < size; ++j) gsl_matrix_set(A, i, j, XtX(i, j));
    gsl_linalg_SV_decomp(A, V, Sv, work);
    Mat Vm(size, size), D(size, size);
    for (int i = 0; i < size; ++i) {
      for (int j = 0; j < size; ++j) Vm(i, j) = gsl_matrix_get(V, i, j);
      const double v = gsl_vector_get(Sv, i);
      if (v > 1e-8) D(i, i) = 1 / v;
    }
    gsl_matrix_free(A);
    gsl_matrix_free(V);
    gsl_vector_free(Sv);
    gsl_vector_free(work);
    Mat XtXi = mul(mul(Vm, false, D, false), false, Vm, true);
    Mat hB = mul(mul(XtXi, false, X, true), false, y, false);
    Mat fy = mul(X, false, hB, false);
    std::vector<double> res(n);
    for (int j = 0; j < n; ++j) res[j] = y(j, 0) - fy(j, 0);
    factor = 1;
    if (size > q) {
      double rr = 0.0;
      for (int j = 0; j < n; ++j) rr += res[j] * res[j];
      const double sigma1 = rr / (n - size);
      Mat Tg = mul(g, true, T, false); /* 1 x n */
      Mat gTg = mul(Tg, false, g, false);
      const double b = hB(q, 0);
      const double T2 = (b * gTg(0, 0) * b) / pow(sigma1, 2); /* sigma1 squared: MVLR.cpp:397 */
      const double v1 = size - q, v2 = n - size;
      const double F = (v2 - v1 + 1) * T2 / (v1 * v2);
      const double qv = gsl_cdf_fdist_Q(F, v1, v2 - v1 + 1);
      const double newF = gsl_cdf_chisq_Qinv(qv, v1) / v1;
      factor = (F < 1e-8) ? 1 : F / newF;
    }
    return res;
  }

  /* MVLR::compute_Sigma / compute_Sigma_mle (MVLR.cpp:148-298) */
  void sigma(const std::vector<int> &gamma, Mat &Sigma, Mat &Sigma_inv)
  {
    if (alpha < 1e-6) {
      Sigma = Sigma0;
      Sigma_inv = Sigma0_inv;
      return;
    }
    Mat E(n, s);
    std::vector<double> fac(s);
    for (int i = 0; i < s; ++i) {
      double factor = 1;
      std::vector<double> r = residual(i, gamma[i] == 1, factor);
      fac[i] = sqrt(factor);
      for (int j = 0; j < n; ++j) E(j, i) = r[j];
    }
    Mat Se = mul(E, true, E, false);
    Sigma = Mat(s, s);
    for (int i = 0; i < s; ++i)
      for (int j = 0; j < s; ++j)
        Sigma(i, j) = (i == j ? 1e-4 * double(m) / double(m + n) : 0.0) + Se(i, j) * (double(1.0) / double(m + n));
    /* Sigma <- Sigma * diag(fac): only the right multiplication survives (MVLR.cpp:268-273) */
    for (int i = 0; i < s; ++i)
      for (int j = 0; j < s; ++j) Sigma(i, j) *= fac[j];
    for (int i = 0; i < s; ++i)
      for (int j = 0; j < s; ++j) Sigma(i, j) = alpha * Sigma(i, j) + (1.0 - alpha) * Sigma0(i, j);
    Sigma_inv = lu_inverse(Sigma, NULL);
  }

  /* MVLR::compute_log10_ABF_vec (MVLR.cpp:715-764) with compute_stats (:517-554),
   * construct_meta_Gamma (:473-494), set_Wg (:497-512), compute_log10_ABF(Wg) (:559-605) */
  std::vector<double> abf_vec(const std::vector<int> &gamma, const std::vector<double> &phi2,
                              const std::vector<double> &oma2)
  {
    Mat Sigma, Sigma_inv;
    sigma(gamma, Sigma, Sigma_inv);
    Mat Gm = mul(T, false, g, false);        /* n x 1 */
    Mat K = mul(Gm, true, Gm, false);        /* 1 x 1 : (Tg)'(Tg) */
    Mat Vinv(s, s);
    for (int i = 0; i < s; ++i)
      for (int j = 0; j < s; ++j) Vinv(i, j) = K(0, 0) * Sigma_inv(i, j);
    Mat t1 = mul(Sigma_inv, false, Y, true); /* s x n */
    Mat b = mul(t1, false, Gm, false);       /* s x 1 */
    Mat Gamma(s, s);
    for (int i = 0; i < s; ++i)
      for (int j = 0; j < s; ++j)
        Gamma(i, j) = (gamma[i] * sqrt(Sigma(i, i))) * (gamma[j] * sqrt(Sigma(j, j)));
    std::vector<double> out;
    for (size_t k = 0; k < oma2.size(); ++k) {
      Mat W(s, s);
      for (int i = 0; i < s; ++i)
        for (int j = 0; j < s; ++j) W(i, j) = Gamma(i, j) * oma2[k];
      for (int j = 0; j < s; ++j) W(j, j) = (oma2[k] + phi2[k]) * Gamma(j, j);
      Mat A = mul(Vinv, false, W, false);
      for (int i = 0; i < s; ++i) A(i, i) += 1;
      double lndet = 0.0;
      Mat Ai = lu_inverse(A, &lndet);
      Mat t3 = mul(W, false, Ai, false);
      Mat t4 = mul(b, true, t3, false);
      Mat t5 = mul(t4, false, b, false);
      double rst = .5 * t5(0, 0);
      rst += -0.5 * lndet;
      out.push_back(rst / log(10.0));
    }
    return out;
  }
};

/* GeneSnpPair::CalcAbfsMvlr (gene_snp_pair.cpp:624-758): a fresh MVLR object per configuration */
void calc_abfs_mvlr(eqo_ctx *c, int64_t g, int64_t m, const size_t *perm, int which, Pair &pr)
{
  const int S = c->cfg.n_subgroups;
  std::vector<std::vector<double> > Y;
  std::vector<double> xg;
  std::vector<std::vector<double> > cov0;
  for (int s = 0; s < S; ++s) { /* FillStlContainers(..., same_individuals=true, ...) */
    std::vector<double> y, x;
    std::vector<std::vector<double> > cov;
    if (!gather(c, g, m, s, perm, y, x, cov)) continue;
    if (Y.empty()) {
      xg = x;
      cov0 = cov;
    }
    Y.push_back(y);
    pr.n[s] = (int)y.size();
    pr.ncov[s] = (int)cov.size();
  }
  const size_t L = c->phi2L.size(), K = c->phi2S.size();
  std::vector<double> fixp(L), fixo(L), maxp(L), maxo(L);
  for (size_t k = 0; k < L; ++k) { /* grid.cpp:46-55 */
    fixp[k] = 0.0;
    fixo[k] = c->phi2L[k] + c->oma2L[k];
    maxp[k] = c->phi2L[k] + c->oma2L[k];
    maxo[k] = 0.0;
  }
  const int s_ = (int)Y.size();
  {
    Mvlr mv;
    mv.init(Y, xg, cov0, c->cfg.fiterr);
    std::vector<int> ones(s_, 1);
    pr.raw_gen.assign(3, std::vector<double>());
    pr.raw_gen[0] = mv.abf_vec(ones, c->phi2L, c->oma2L);
    pr.raw_gen[1] = mv.abf_vec(ones, fixp, fixo);
    pr.raw_gen[2] = mv.abf_vec(ones, maxp, maxo);
    for (int j = 0; j < 3; ++j) pr.w_gen[j] = log10_weighted_sum(pr.raw_gen[j].data(), L);
  }
  if (which == 1) return;
  std::vector<std::vector<int> > configs;
  enumerate_configs(s_, which == 2, configs);
  pr.raw_cfg.assign(configs.size(), std::vector<double>());
  pr.w_cfg.assign(configs.size(), kNaN);
  for (size_t ci = 0; ci < configs.size(); ++ci) {
    Mvlr mv;
    mv.init(Y, xg, cov0, c->cfg.fiterr);
    pr.raw_cfg[ci] = mv.abf_vec(configs[ci], c->phi2S, c->oma2S);
    pr.w_cfg[ci] = log10_weighted_sum(pr.raw_cfg[ci].data(), K);
  }
  bma_lite(pr, S);
  if (which == 3) bma(pr, S, c->configs_all);
}

/* ------------------------------------------------------------------ hybrid (gene_snp_pair.cpp:760-1423) */

/* utils::mygsl_linalg_pseudoinverse (utils_math.cpp:249-275) through the shim's gsl_linalg_SV_decomp;
 * optionally returns V D^-2 V' and the singular values (CalcBetahatsAndDiagsPerSubgroup needs them) */
Mat pinv_svd(const Mat &A, Mat *VD2Vt = NULL, std::vector<double> *sv = NULL)
{
  const int M = A.r, N = A.c;
  gsl_matrix *U = gsl_matrix_alloc(M, N), *V = gsl_matrix_alloc(N, N);
  gsl_vector *D = gsl_vector_alloc(N), *work = gsl_vector_alloc(N);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) gsl_matrix_set(U, i, j, A(i, j));
  gsl_linalg_SV_decomp(U, V, D, work);
  if (sv) {
    sv->resize(N);
    for (int j = 0; j < N; ++j) (*sv)[j] = gsl_vector_get(D, j);
  }
  Mat Vm(N, N), VDinv(N, N), Um(M, N);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      Vm(i, j) = gsl_matrix_get(V, i, j);
      VDinv(i, j) = Vm(i, j) * pow(gsl_vector_get(D, j), -1.0);
    }
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) Um(i, j) = gsl_matrix_get(U, i, j);
  if (VD2Vt) {
    Mat VD2(N, N);
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) VD2(i, j) = Vm(i, j) * pow(pow(gsl_vector_get(D, j), -1.0), 2.0);
    *VD2Vt = mul(VD2, false, Vm, true);
  }
  gsl_matrix_free(U);
  gsl_matrix_free(V);
  gsl_vector_free(D);
  gsl_vector_free(work);
  return mul(VDinv, false, Um, true);
}

/* utils::CalcMleErrorCovariance (utils_math.cpp:306-348): Sigma = Y' (I - X (X'X)^+ X') Y / N */
Mat mle_error_covariance(const Mat &Y, const Mat &X)
{
  const int N = X.r;
  const Mat XtXinv = pinv_svd(mul(X, true, X, false));
  const Mat H = mul(mul(X, false, XtXinv, false), false, X, true);
  Mat T(N, N);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) T(i, j) = -H(i, j) + (i == j ? 1.0 : 0.0);
  Mat Sg = mul(Y, true, mul(T, false, Y, false), false);
  for (size_t e = 0; e < Sg.a.size(); ++e) Sg.a[e] *= 1 / (double)N;
  return Sg;
}

/* CalcLog10AbfMvlr of gene_snp_pair.cpp:1165-1255 (the hybrid model's own, not X. Wen's class) */
double abf_hybrid(const std::vector<int> &gamma, const std::vector<double> &bhat, const Mat &Sigma, const Mat &Vg,
                  double phi2, double oma2)
{
  const int S = (int)gamma.size();
  Mat Wg(S, S);
  for (int i = 0; i < S; ++i)
    for (int j = 0; j < S; ++j) Wg(i, j) = ((i == j) ? phi2 + oma2 : oma2) * (double)(gamma[i] * gamma[j]);
  const Mat Vinv = lu_inverse(Vg, NULL);
  Mat b(S, 1);
  for (int i = 0; i < S; ++i) b(i, 0) = bhat[i];
  const Mat bVg = mul(b, true, Vinv, false); /* 1 x S */
  Mat sd(S, S);
  for (int i = 0; i < S; ++i) sd(i, i) = pow(Sigma(i, i), 0.5);
  Wg = mul(mul(sd, false, Wg, false), false, sd, false);
  Mat ivw = mul(Vinv, false, Wg, false);
  for (int i = 0; i < S; ++i) ivw(i, i) += 1.0;
  double lndet = 0.0;
  const Mat ivw_inv = lu_inverse(ivw, &lndet);
  double l10 = -0.5 * lndet;
  const Mat t6 = mul(mul(mul(bVg, false, Wg, false), false, ivw_inv, false), false, bVg, true);
  l10 += 0.5 * t6(0, 0);
  return l10 / log(10);
}

/* GeneSnpPair::CalcAbfsHybrid (gene_snp_pair.cpp:1384-1418) with CalcSstatsHybrid (:1134-1163),
 * CalcBetahatsAndDiagsPerSubgroup (:760-879), CalcOffDiagCovarsFromPairsOfSubgroups (:1059-1132) */
void calc_abfs_hybrid(eqo_ctx *c, int64_t g, int64_t m, const size_t *perm, int which, Pair &pr)
{
  const int S = c->cfg.n_subgroups, N_all = c->cfg.n_samples_all;
  const double f = c->cfg.fiterr;
  std::vector<double> bhat(S, kNaN);
  Mat Sigma(S, S), Vg(S, S);
  /* diagonals: each subgroup on its own individuals, permuted / quantile-normalised like the uvlr statistics */
  for (int s = 0; s < S; ++s) {
    std::vector<double> y, x;
    std::vector<std::vector<double> > cov;
    if (!gather(c, g, m, s, perm, y, x, cov)) continue;
    const int N = (int)y.size(), Q = (int)cov.size(), Q1 = Q + 1, Q2 = Q + 2;
    pr.n[s] = N;
    pr.ncov[s] = Q;
    Mat X(N, Q2), Xc(N, Q1), yv(N, 1);
    for (int i = 0; i < N; ++i) {
      yv(i, 0) = y[i];
      X(i, 0) = Xc(i, 0) = 1.0;
      X(i, 1) = x[i];
      for (int j = 0; j < Q; ++j) X(i, j + 2) = Xc(i, j + 1) = cov[j][i];
    }
    Mat VD2Vt;
    std::vector<double> sv;
    const Mat Xps = pinv_svd(X, &VD2Vt, &sv);
    size_t rank = 0;
    for (size_t j = 0; j < sv.size(); ++j)
      if (sv[j] > GSL_DBL_EPSILON) rank += 1;
    const Mat B = mul(Xps, false, yv, false);
    const Mat XB = mul(X, false, B, false);
    double rss_full = 0.0;
    for (int i = 0; i < N; ++i) rss_full += (yv(i, 0) - XB(i, 0)) * (yv(i, 0) - XB(i, 0));
    const double sigma2_full = rss_full / (double)N;
    pr.pve[s] = 1 - rss_full / gsl_stats_tss(y.data(), 1, y.size());
    pr.sigmahat[s] = sqrt(rss_full / (double)(N - rank));
    pr.betahat[s] = B(1, 0);
    pr.sebetahat[s] = pr.sigmahat[s] * sqrt(VD2Vt(1, 1));
    pr.pval[s] = 2 * gsl_cdf_tdist_Q(fabs(pr.betahat[s] / pr.sebetahat[s]), N - rank);
    const Mat Bn = mul(pinv_svd(Xc), false, yv, false);
    const Mat XBn = mul(Xc, false, Bn, false);
    double sigma2_null = 0.0;
    for (int i = 0; i < N; ++i) sigma2_null += (yv(i, 0) - XBn(i, 0)) * (yv(i, 0) - XBn(i, 0));
    sigma2_null /= (double)N;
    bhat[s] = B(1, 0);
    Sigma(s, s) = f * sigma2_full + (1 - f) * sigma2_null;
    Vg(s, s) = Sigma(s, s) * VD2Vt(1, 1);
  }
  /* off-diagonals: pairs of subgroups on their common individuals -- never permuted, never quantile-normalised,
   * genotype and covariates of the FIRST subgroup of the pair, covariates indexed by the all-sample index itself
   * (gene_snp_pair.cpp:881-983; samples.cpp:148-192) */
  for (int s1 = 0; s1 < S - 1; ++s1)
    for (int s2 = s1 + 1; s2 < S; ++s2) {
      const Sub &a = c->subs[s1], &b = c->subs[s2];
      const double *Ya = &a.Y[(size_t)g * a.n_exp_cols], *Yb = &b.Y[(size_t)g * b.n_exp_cols];
      const double *Gm = &c->genos[a.geno_id][(size_t)m * c->geno_cols[a.geno_id]];
      std::vector<int> both, only1, only2;
      for (int i = 0; i < N_all; ++i) {
        const bool p1 = a.all2geno[i] >= 0 && a.all2exp[i] >= 0 && !is_nan(Ya[a.all2exp[i]]);
        const bool p2 = b.all2geno[i] >= 0 && b.all2exp[i] >= 0 && !is_nan(Yb[b.all2exp[i]]);
        if (p1 && p2) both.push_back(i);
        else if (p1) only1.push_back(i);
        else if (p2) only2.push_back(i);
      }
      if (both.empty()) {
        c->fatal = true;
        c->err = "ERROR: two subgroups have no individuals in common";
        return;
      }
      const int Q = a.Q, Q2 = Q + 2;
      struct Fill {
        static Mat design(eqo_ctx *c, const Sub &a, const double *Gm, const std::vector<int> &inds, int Q)
        {
          Mat X((int)inds.size(), Q + 2);
          for (size_t r = 0; r < inds.size(); ++r) {
            X((int)r, 0) = 1.0;
            if (a.all2geno[inds[r]] < 0) {
              c->fatal = true;
              c->err = "--error hybrid: an individual unique to the second subgroup has no genotype in the first";
              X((int)r, 1) = kNaN;
            } else
              X((int)r, 1) = Gm[a.all2geno[inds[r]]];
            for (int q = 0; q < Q; ++q) {
              if (inds[r] >= a.n_cov_cols) {
                c->fatal = true;
                c->err = "--error hybrid: covariate index out of range";
                X((int)r, 2 + q) = kNaN;
              } else
                X((int)r, 2 + q) = a.C[(size_t)q * a.n_cov_cols + inds[r]];
            }
          }
          return X;
        }
      };
      const Mat X12 = Fill::design(c, a, Gm, both, Q);
      Mat Y12((int)both.size(), 2);
      for (size_t r = 0; r < both.size(); ++r) {
        Y12((int)r, 0) = Ya[a.all2exp[both[r]]];
        Y12((int)r, 1) = Yb[b.all2exp[both[r]]];
      }
      const Mat tXX = mul(X12, true, X12, false);
      Mat A[2];
      for (int u = 0; u < 2; ++u) { /* GetMatrixA (:985-1014) */
        const std::vector<int> &only = u == 0 ? only1 : only2;
        Mat G = tXX;
        if (!only.empty()) {
          const Mat Xu = Fill::design(c, a, Gm, only, Q);
          const Mat tXuXu = mul(Xu, true, Xu, false);
          for (size_t e = 0; e < G.a.size(); ++e) G.a[e] += tXuXu.a[e];
        }
        A[u] = mul(pinv_svd(G), false, X12, true);
      }
      if (c->fatal) return;
      const Mat Sfull = mle_error_covariance(Y12, X12);
      Mat Xc12(X12.r, Q2 - 1);
      for (int i = 0; i < X12.r; ++i) {
        Xc12(i, 0) = 1.0;
        for (int j = 2; j < Q2; ++j) Xc12(i, j - 1) = X12(i, j);
      }
      const Mat Snull = mle_error_covariance(Y12, Xc12);
      Sigma(s1, s2) = Sigma(s2, s1) = f * Sfull(0, 1) + (1 - f) * Snull(0, 1);
      const Mat cov12 = mul(A[0], false, A[1], true);
      Vg(s1, s2) = Vg(s2, s1) = Sigma(s1, s2) * cov12(1, 1);
    }
  const size_t L = c->phi2L.size(), K = c->phi2S.size();
  const std::vector<int> ones(S, 1);
  pr.raw_gen.assign(3, std::vector<double>(L));
  for (size_t k = 0; k < L; ++k) { /* CalcAbfsHybridForConsistentConfiguration (:1257-1308) */
    const double ph = c->phi2L[k], om = c->oma2L[k];
    pr.raw_gen[0][k] = abf_hybrid(ones, bhat, Sigma, Vg, ph, om);
    pr.raw_gen[1][k] = abf_hybrid(ones, bhat, Sigma, Vg, 0.0, ph + om);
    pr.raw_gen[2][k] = abf_hybrid(ones, bhat, Sigma, Vg, ph + om, 0.0);
  }
  for (int j = 0; j < 3; ++j) pr.w_gen[j] = log10_weighted_sum(pr.raw_gen[j].data(), L);
  if (which == 1) return;
  std::vector<std::vector<int> > configs;
  enumerate_configs(S, which == 2, configs);
  pr.raw_cfg.assign(configs.size(), std::vector<double>(K));
  pr.w_cfg.assign(configs.size(), kNaN);
  for (size_t ci = 0; ci < configs.size(); ++ci) {
    for (size_t k = 0; k < K; ++k) pr.raw_cfg[ci][k] = abf_hybrid(configs[ci], bhat, Sigma, Vg, c->phi2S[k], c->oma2S[k]);
    pr.w_cfg[ci] = log10_weighted_sum(pr.raw_cfg[ci].data(), K);
  }
  bma_lite(pr, S);
  if (which == 3) bma(pr, S, c->configs_all);
}

/* ------------------------------------------------------------------ gene level */

bool gene_has_all(const eqo_ctx *c, int64_t g)
{
  for (int s = 0; s < c->cfg.n_subgroups; ++s)
    if (!c->subs[s].gene_has[g]) return false;
  return true;
}

bool snp_has_all(const eqo_ctx *c, int64_t m)
{
  for (int s = 0; s < c->cfg.n_subgroups; ++s)
    if (!c->subs[s].snp_has[m]) return false;
  return true;
}

/* Gene::HasAtLeastOneCisSnpInAtLeastOneSubgroup (gene.cpp:202-216) + the mvlr skip
 * (eqtlbma_bf.cpp:747-762) */
bool gene_analyzed(const eqo_ctx *c, int64_t g)
{
  bool any = false;
  for (int64_t m = c->cb[g]; m < c->ce[g] && !any; ++m)
    for (int s = 0; s < c->cfg.n_subgroups; ++s)
      if (c->subs[s].gene_has[g] && c->subs[s].snp_has[m]) {
        any = true;
        break;
      }
  if (!any) return false;
  if (c->cfg.analysis == EQB_ANALYSIS_JOIN && c->cfg.error_model != EQB_ERROR_UVLR && !gene_has_all(c, g))
    return false;
  return true;
}

/* one pair of Gene::TestForAssociations (gene.cpp:285-331) or of a permutation loop */
void test_pair(eqo_ctx *c, int64_t g, int64_t m, const size_t *perm, int which, Pair &pr)
{
  const int S = c->cfg.n_subgroups;
  const bool join = c->cfg.analysis == EQB_ANALYSIS_JOIN;
  if (!join || c->cfg.error_model == EQB_ERROR_UVLR) {
    for (int s = 0; s < S; ++s)
      if (c->subs[s].gene_has[g] && c->subs[s].snp_has[m]) calc_sstats_one(c, g, m, s, perm, pr);
    if (join) calc_abfs_uvlr(c, pr, which);
  } else {
    if (!snp_has_all(c, m)) return; /* gene.cpp:315-321 */
    if (c->cfg.error_model == EQB_ERROR_HYBRID)
      calc_abfs_hybrid(c, g, m, perm, which, pr);
    else
      calc_abfs_mvlr(c, g, m, perm, which, pr);
  }
}

double weighted_for_pbf(const Pair &pr, int pbf)
{
  switch (pbf) {
    case EQB_PBF_GEN: return pr.w_gen[0];
    case EQB_PBF_GEN_SIN: return pr.w_gensin;
    case EQB_PBF_ALL: return pr.w_all;
  }
  return kNaN;
}

/* Gene::CalcPermutationPvalue (gene.cpp:348-364) */
double perm_pvalue(size_t total, size_t sofar, double more_extreme, size_t cutoff, const gsl_rng *rng)
{
  if (sofar == total) return more_extreme / (total + 1);
  return gsl_ran_flat(rng, ((1 + cutoff) / ((double)(sofar + 2))), ((1 + cutoff) / ((double)(sofar + 1))));
}

struct PermState {
  size_t nb_permutations; /* the caller's variable, decremented on NaN statistics (App. B #5) */
  gsl_rng *rngPerm, *rngTrick;
};

/* Gene::MakePermutationsJoin (gene.cpp:598-717) */
void perm_join(eqo_ctx *c, int64_t g, const eqb_perm_config *pc, PermState &ps, int64_t out_idx,
               eqb_perm_results *res)
{
  const int N_all = c->cfg.n_samples_all;
  const int64_t nsnp = c->ce[g] - c->cb[g];
  const int S = c->cfg.n_subgroups;
  const int which_true = c->cfg.bfs + 1, which_perm = pc->pbf;
  gsl_permutation *perm = gsl_permutation_calloc(N_all);
  size_t nbperms = 0;
  double count = 1;
  std::vector<double> stats_perms;

  /* FindMaxTrueL10Abf (:575-582) / AvgTrueL10Abfs (:587-596) */
  std::vector<double> truew(nsnp);
  for (int64_t j = 0; j < nsnp; ++j) {
    Pair pr(S);
    test_pair(c, g, c->cb[g] + j, NULL, std::max(which_true, which_perm), pr);
    truew[j] = weighted_for_pbf(pr, pc->pbf);
  }
  double true_stat;
  if (pc->maxbf) {
    true_stat = -kInf;
    for (int64_t j = 0; j < nsnp; ++j)
      if (truew[j] > true_stat) true_stat = truew[j];
  } else {
    std::vector<double> v;
    for (int64_t j = 0; j < nsnp; ++j)
      if (!is_nan(truew[j])) v.push_back(truew[j]);
    true_stat = log10_weighted_sum(v.data(), v.size());
  }

  bool shuffle_only = false;
  const size_t nb_all = ps.nb_permutations;
  std::vector<double> snps(nsnp);
  for (size_t perm_id = 0; perm_id < nb_all; ++perm_id) {
    gsl_ran_shuffle(ps.rngPerm, perm->data, perm->size, sizeof(size_t));
    if (shuffle_only) continue;
    std::fill(snps.begin(), snps.end(), 0.0);
#pragma omp parallel for num_threads(c->threads) schedule(static)
    for (int64_t j = 0; j < nsnp; ++j) {
      const int64_t m = c->cb[g] + j;
      if (c->cfg.error_model != EQB_ERROR_UVLR && !snp_has_all(c, m)) continue;
      Pair pr(S);
      test_pair(c, g, m, perm->data, which_perm, pr);
      snps[j] = weighted_for_pbf(pr, pc->pbf);
    }
    double stat;
    if (pc->maxbf)
      stat = *std::max_element(snps.begin(), snps.end());
    else
      stat = log10_weighted_sum(snps.data(), snps.size());
    if (res->perm_stats) res->perm_stats[out_idx * pc->nperm + perm_id] = stat;
    if (is_nan(stat)) {
      ps.nb_permutations--;
      continue;
    }
    ++nbperms;
    if (stat >= true_stat) ++count;
    stats_perms.push_back(stat);
    if (pc->trick != 0 && count == 1 + pc->tricut) {
      if (pc->trick == 1)
        break;
      else if (pc->trick == 2)
        shuffle_only = true;
    }
  }
  if (res->count) res->count[out_idx] = (int64_t)count;
  if (res->nperm_done) res->nperm_done[out_idx] = (int64_t)nbperms;
  if (res->true_stat) res->true_stat[out_idx] = true_stat;
  if (res->pval) res->pval[out_idx] = perm_pvalue(ps.nb_permutations, nbperms, count, pc->tricut, ps.rngTrick);
  /* the reference takes the median over one element MORE than it stored (gene.cpp:713-714), i.e.
   * it reads the vector's spare capacity; with glibc that element is 0.0 (verified against
   * oracle/_ref on this image), so the emulation is "stored statistics plus one 0.0" */
  if (res->median_perm) {
    stats_perms.push_back(0.0);
    res->median_perm[out_idx] = median(stats_perms);
  }
  gsl_permutation_free(perm);
}

/* Gene::MakePermutationsSepAllSubgroups (gene.cpp:493-570) and ...SepPerSubgroup (:380-450);
 * only_s < 0 means all subgroups */
void perm_sep(eqo_ctx *c, int64_t g, int only_s, const eqb_perm_config *pc, PermState &ps,
              int64_t out_idx, eqb_perm_results *res)
{
  const int N_all = c->cfg.n_samples_all;
  const int64_t nsnp = c->ce[g] - c->cb[g];
  const int S = c->cfg.n_subgroups;
  gsl_permutation *perm = gsl_permutation_calloc(N_all);
  size_t nbperms = 0;
  double count = 1;

  /* FindMinTruePvalue{PerSubgroup,AllSubgroups} (gene.cpp:369-378,480-491) */
  double true_min = 1.0;
  for (int64_t j = 0; j < nsnp; ++j) {
    Pair pr(S);
    test_pair(c, g, c->cb[g] + j, NULL, 0, pr);
    for (int s = 0; s < S; ++s) {
      if (only_s >= 0 && s != only_s) continue;
      if (!c->subs[s].gene_has[g]) continue;
      if (pr.has(s) && pr.pval[s] < true_min) true_min = pr.pval[s];
    }
  }

  bool shuffle_only = false;
  const size_t nb_all = ps.nb_permutations;
  std::vector<double> pv(nsnp);
  for (size_t perm_id = 0; perm_id < nb_all; ++perm_id) {
    gsl_ran_shuffle(ps.rngPerm, perm->data, perm->size, sizeof(size_t));
    if (shuffle_only) continue;
    std::fill(pv.begin(), pv.end(), 1.0);
#pragma omp parallel for num_threads(c->threads) schedule(static)
    for (int64_t j = 0; j < nsnp; ++j) {
      const int64_t m = c->cb[g] + j;
      Pair pr(S);
      double pmin = 1;
      for (int s = 0; s < S; ++s) {
        if (only_s >= 0 && s != only_s) continue;
        if (c->subs[s].gene_has[g] && c->subs[s].snp_has[m]) {
          calc_sstats_one(c, g, m, s, perm->data, pr);
          /* GetBetapvalGeno of a subgroup without entry is undefined in the reference; a kept
           * subgroup always has one here because the gather cannot be empty for both passes */
          const double p = pr.pval[s];
          if (only_s >= 0)
            pmin = p; /* gene.cpp:425: plain assignment */
          else if (p < pmin)
            pmin = p; /* gene.cpp:542-543 */
        }
      }
      pv[j] = pmin;
    }
    const double stat = *std::min_element(pv.begin(), pv.end());
    const int64_t stride = (only_s >= 0) ? S : 1;
    if (res->perm_stats)
      res->perm_stats[(out_idx * stride + (only_s >= 0 ? only_s : 0)) * pc->nperm + perm_id] = stat;
    if (is_nan(stat)) {
      ps.nb_permutations--;
      continue;
    }
    ++nbperms;
    if (stat <= true_min) ++count;
    if (pc->trick != 0 && count == 1 + pc->tricut) {
      if (pc->trick == 1)
        break;
      else if (pc->trick == 2)
        shuffle_only = true;
    }
  }
  const int64_t o = (only_s >= 0) ? out_idx * S + only_s : out_idx;
  if (res->count) res->count[o] = (int64_t)count;
  if (res->nperm_done) res->nperm_done[o] = (int64_t)nbperms;
  if (res->true_stat) res->true_stat[o] = true_min;
  if (res->pval) res->pval[o] = perm_pvalue(ps.nb_permutations, nbperms, count, pc->tricut, ps.rngTrick);
  gsl_permutation_free(perm);
}

} // namespace

/* ------------------------------------------------------------------ C interface */

extern "C" {

int eqo_create(eqo_ctx **ctx, const eqb_config *cfg)
{
  if (!ctx || !cfg || cfg->abi_version != EQB_ABI_VERSION) return 1;
  eqo_ctx *c = new eqo_ctx();
  c->cfg = *cfg;
  c->subs.resize(cfg->n_subgroups);
  if (cfg->bfs == EQB_BFS_ALL) enumerate_configs(cfg->n_subgroups, false, c->configs_all);
  *ctx = c;
  return 0;
}

void eqo_destroy(eqo_ctx *ctx) { delete ctx; }

const char *eqo_last_error(const eqo_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

void eqo_set_threads(eqo_ctx *ctx, int32_t n) { ctx->threads = n > 0 ? n : 1; }

int eqo_set_genotypes(eqo_ctx *c, int32_t geno_id, const double *G, int64_t n_snps, int32_t n_cols)
{
  if (geno_id < 0 || n_snps != c->cfg.n_snps) {
    c->err = "bad genotype matrix";
    return 1;
  }
  if ((size_t)geno_id >= c->genos.size()) {
    c->genos.resize(geno_id + 1);
    c->geno_cols.resize(geno_id + 1, 0);
  }
  c->genos[geno_id].assign(G, G + (size_t)n_snps * n_cols);
  c->geno_cols[geno_id] = n_cols;
  return 0;
}

int eqo_set_subgroup(eqo_ctx *c, int32_t s, const eqb_subgroup *sg)
{
  if (s < 0 || s >= c->cfg.n_subgroups) {
    c->err = "bad subgroup index";
    return 1;
  }
  Sub &sb = c->subs[s];
  const int N_all = c->cfg.n_samples_all;
  const int64_t M = c->cfg.n_snps, G = c->cfg.n_genes;
  sb.geno_id = sg->geno_id;
  sb.n_exp_cols = sg->n_exp_cols;
  sb.Q = sg->n_covariates;
  sb.n_cov_cols = sg->n_cov_cols;
  sb.all2geno.assign(sg->all2geno, sg->all2geno + N_all);
  sb.all2exp.assign(sg->all2exp, sg->all2exp + N_all);
  if (sg->all2cov) sb.all2cov.assign(sg->all2cov, sg->all2cov + N_all);
  else sb.all2cov.clear();
  if (sg->snp_has_geno) sb.snp_has.assign(sg->snp_has_geno, sg->snp_has_geno + M);
  else sb.snp_has.assign(M, 1);
  if (sg->gene_has_exp) sb.gene_has.assign(sg->gene_has_exp, sg->gene_has_exp + G);
  else sb.gene_has.assign(G, 1);
  sb.Y.assign(sg->Y, sg->Y + (size_t)G * sg->n_exp_cols);
  if (sb.Q > 0) sb.C.assign(sg->C, sg->C + (size_t)sb.Q * sb.n_cov_cols);
  else sb.C.clear();
  return 0;
}

int eqo_set_grids(eqo_ctx *c, const double *phi2L, const double *oma2L, int32_t L,
                  const double *phi2S, const double *oma2S, int32_t K)
{
  c->phi2L.assign(phi2L, phi2L + L);
  c->oma2L.assign(oma2L, oma2L + L);
  c->phi2S.assign(phi2S, phi2S + K);
  c->oma2S.assign(oma2S, oma2S + K);
  return 0;
}

/* Snp::IsInCis (snp.cpp:274-297) */
static int is_in_cis(uint64_t pos, uint64_t start, uint64_t end, int anchor, uint64_t radius)
{
  int res = -1;
  const uint64_t hi = (anchor == EQB_ANCHOR_TSS_TES ? end : start) + radius;
  if (((start >= radius && pos >= start - radius) || (start < radius)) && pos <= hi)
    res = 0;
  else if (pos > hi)
    res = 1;
  return res;
}

/* Gene::SetCisSnps (gene.cpp:140-157): linear scan over the chromosome's position-sorted SNPs */
int eqo_build_cis_windows(eqo_ctx *c, const int32_t *gene_chr, const int64_t *gene_start,
                          const int64_t *gene_end, const int32_t *snp_chr, const int64_t *snp_pos,
                          int32_t anchor, int64_t radius, int64_t *begin_out, int64_t *end_out)
{
  const int64_t G = c->cfg.n_genes, M = c->cfg.n_snps;
  c->cb.assign(G, 0);
  c->ce.assign(G, 0);
  for (int64_t g = 0; g < G; ++g) {
    int64_t b = -1, e = -1;
    for (int64_t m = 0; m < M; ++m) {
      if (snp_chr[m] != gene_chr[g]) continue;
      const int r = is_in_cis((uint64_t)snp_pos[m], (uint64_t)gene_start[g], (uint64_t)gene_end[g],
                              anchor, (uint64_t)radius);
      if (r == 1) break;
      if (r == -1) continue;
      if (b < 0) b = m;
      e = m + 1;
    }
    if (b >= 0) {
      c->cb[g] = b;
      c->ce[g] = e;
    }
    if (begin_out) begin_out[g] = c->cb[g];
    if (end_out) end_out[g] = c->ce[g];
  }
  return 0;
}

int eqo_set_cis_windows(eqo_ctx *c, const int64_t *begin, const int64_t *end)
{
  c->cb.assign(begin, begin + c->cfg.n_genes);
  c->ce.assign(end, end + c->cfg.n_genes);
  return 0;
}

int eqo_finalize(eqo_ctx *c)
{
  if ((int64_t)c->cb.size() != c->cfg.n_genes) {
    c->err = "cis windows not set";
    return 1;
  }
  return 0;
}

int64_t eqo_n_configs(const eqo_ctx *c)
{
  const int S = c->cfg.n_subgroups;
  if (c->cfg.analysis != EQB_ANALYSIS_JOIN || c->cfg.bfs == EQB_BFS_GEN) return 0;
  if (c->cfg.bfs == EQB_BFS_SIN) return S;
  return ((int64_t)1 << S) - 1;
}

int eqo_pair_offsets(eqo_ctx *c, int64_t gene_lo, int64_t gene_hi, int64_t *offsets)
{
  int64_t acc = 0;
  for (int64_t g = gene_lo; g < gene_hi; ++g) {
    offsets[g - gene_lo] = acc;
    if (gene_analyzed(c, g)) acc += c->ce[g] - c->cb[g];
  }
  offsets[gene_hi - gene_lo] = acc;
  return 0;
}

int eqo_run(eqo_ctx *c, int64_t gene_lo, int64_t gene_hi, eqb_results *res)
{
  const int S = c->cfg.n_subgroups;
  const int64_t C = eqo_n_configs(c);
  const size_t L = c->phi2L.size(), K = c->phi2S.size();
  const int which = c->cfg.bfs + 1;
  std::vector<int64_t> off(gene_hi - gene_lo + 1);
  eqo_pair_offsets(c, gene_lo, gene_hi, off.data());
  c->fatal = false;
  for (int64_t g = gene_lo; g < gene_hi; ++g) {
    const bool an = gene_analyzed(c, g);
    if (res->gene_analyzed) res->gene_analyzed[g - gene_lo] = an ? 1 : 0;
    if (!an) continue;
    const int64_t nsnp = c->ce[g] - c->cb[g];
#pragma omp parallel for num_threads(c->threads) schedule(dynamic, 4)
    for (int64_t j = 0; j < nsnp; ++j) {
      const int64_t p = off[g - gene_lo] + j;
      Pair pr(S);
      test_pair(c, g, c->cb[g] + j, NULL, which, pr);
      for (int s = 0; s < S; ++s) {
        if (res->n) res->n[p * S + s] = pr.n[s];
        if (res->sstats) {
          double *o = &res->sstats[(p * S + s) * 5];
          o[0] = pr.pve[s];
          o[1] = pr.sigmahat[s];
          o[2] = pr.betahat[s];
          o[3] = pr.sebetahat[s];
          o[4] = pr.pval[s];
        }
      }
      if (c->cfg.analysis != EQB_ANALYSIS_JOIN) continue;
      const bool empty = pr.raw_gen.empty(); /* mvlr pair skipped: no ABF at all */
      if (res->abf_gen)
        for (int j3 = 0; j3 < 3; ++j3)
          for (size_t k = 0; k < L; ++k)
            res->abf_gen[(p * 3 + j3) * L + k] = empty ? kNaN : pr.raw_gen[j3][k];
      if (res->abf_cfg)
        for (int64_t ci = 0; ci < C; ++ci)
          for (size_t k = 0; k < K; ++k)
            res->abf_cfg[(p * C + ci) * K + k] = (empty || pr.raw_cfg.empty()) ? kNaN : pr.raw_cfg[ci][k];
      if (res->abf_w) {
        double *o = &res->abf_w[p * (5 + C)];
        o[0] = pr.w_gen[0];
        o[1] = pr.w_gen[1];
        o[2] = pr.w_gen[2];
        o[3] = pr.w_gensin;
        o[4] = pr.w_all;
        for (int64_t ci = 0; ci < C; ++ci) o[5 + ci] = pr.w_cfg.empty() ? kNaN : pr.w_cfg[ci];
      }
    }
  }
  return c->fatal ? 2 : 0;
}

/* makePermutations / makePermutationsSep / makePermutationsJoin (eqtlbma_bf.cpp:773-917) */
int eqo_run_permutations(eqo_ctx *c, int64_t gene_lo, int64_t gene_hi, const eqb_perm_config *pc,
                         eqb_perm_results *res)
{
  if (pc->wrtsize <= 0 || gene_lo % pc->wrtsize != 0) {
    c->err = "gene_lo must be a multiple of wrtsize";
    return 1;
  }
  const int S = c->cfg.n_subgroups;
  const bool join = c->cfg.analysis == EQB_ANALYSIS_JOIN;
  const int64_t n = gene_hi - gene_lo;
  const int64_t per_gene = (!join && pc->permsep == 2) ? S : 1;
  for (int64_t i = 0; i < n * per_gene; ++i) {
    if (res->pval) res->pval[i] = kNaN;
    if (res->nperm_done) res->nperm_done[i] = 0;
    if (res->count) res->count[i] = 0;
    if (res->true_stat) res->true_stat[i] = kNaN;
    if (res->median_perm) res->median_perm[i] = kNaN;
  }
  if (res->perm_stats)
    for (int64_t i = 0; i < n * per_gene * pc->nperm; ++i) res->perm_stats[i] = kNaN;

  PermState ps;
  ps.nb_permutations = (size_t)pc->nperm;
  gsl_rng_env_setup();
  ps.rngPerm = gsl_rng_alloc(gsl_rng_default);
  ps.rngTrick = gsl_rng_alloc(gsl_rng_default);
  c->fatal = false;
  for (int64_t g0 = gene_lo; g0 < gene_hi; g0 += pc->wrtsize) {
    const int64_t g1 = std::min<int64_t>(g0 + pc->wrtsize, gene_hi);
    if (join) {
      if (pc->pbf == EQB_PBF_NONE) continue;
      gsl_rng_set(ps.rngPerm, pc->seed);
      if (pc->trick != 0) gsl_rng_set(ps.rngTrick, pc->seed);
      for (int64_t g = g0; g < g1; ++g) {
        if (!gene_analyzed(c, g)) continue; /* includes the mvlr all-subgroups condition */
        perm_join(c, g, pc, ps, g - gene_lo, res);
      }
    } else if (pc->permsep == 1) {
      gsl_rng_set(ps.rngPerm, pc->seed);
      if (pc->trick != 0) gsl_rng_set(ps.rngTrick, pc->seed);
      for (int64_t g = g0; g < g1; ++g) {
        if (!gene_analyzed(c, g)) continue;
        perm_sep(c, g, -1, pc, ps, g - gene_lo, res);
      }
    } else if (pc->permsep == 2) {
      for (int s = 0; s < S; ++s) {
        gsl_rng_set(ps.rngPerm, pc->seed);
        if (pc->trick != 0) gsl_rng_set(ps.rngTrick, pc->seed);
        for (int64_t g = g0; g < g1; ++g) {
          if (!gene_analyzed(c, g)) continue;
          perm_sep(c, g, s, pc, ps, g - gene_lo, res);
        }
      }
    }
  }
  gsl_rng_free(ps.rngPerm);
  gsl_rng_free(ps.rngTrick);
  return c->fatal ? 2 : 0;
}

void eqo_shuffle_table(uint64_t seed, int64_t n_skip_shuffles, int64_t n_perm, int32_t n, int32_t *perms)
{
  gsl_rng_env_setup();
  gsl_rng *r = gsl_rng_alloc(gsl_rng_default);
  gsl_rng_set(r, seed);
  gsl_permutation *p = gsl_permutation_calloc(n);
  for (int64_t k = 0; k < n_skip_shuffles; ++k) gsl_ran_shuffle(r, p->data, p->size, sizeof(size_t));
  for (int i = 0; i < n; ++i) p->data[i] = i;
  for (int64_t k = 0; k < n_perm; ++k) {
    gsl_ran_shuffle(r, p->data, p->size, sizeof(size_t));
    for (int i = 0; i < n; ++i) perms[k * n + i] = (int32_t)p->data[i];
  }
  gsl_permutation_free(p);
  gsl_rng_free(r);
}

} /* extern "C" */
