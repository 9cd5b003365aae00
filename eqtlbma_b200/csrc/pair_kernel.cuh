// pair_kernel.cuh -- the fused gene x SNP x (permutation) kernel of the eqtlbma_bf hot path.
//
// One CTA = one (gene, permutation) work item; the identity permutation is the "true" pass.
//   phase 1 (K1, covariate projection): each warp takes subgroups s = warp, warp+8, ...: gathers the
//            (permuted) expression row, builds the keep-mask (ragged individuals, NaN expression,
//            absent genes), optional --qnorm, and an orthonormal basis of [1, covariates] on the
//            kept rows (CGS2), then residualises the phenotype.         gene_snp_pair.cpp:79-170
//   phase 2 (K2, contraction): each warp takes cis SNPs m = begin+warp, ...: the genotype row is
//            masked, residualised against the basis and contracted with the residual phenotype
//            -> betahat, sebetahat, sigmahat, pve, p-value per subgroup.
//                                          gene_snp_pair.cpp:175-208, utils_math.cpp:166-209
//   phase 3 (K3, Bayes factors): the same warp standardises the S summary statistics
//            (gene_snp_pair.cpp:256-290) and evaluates the closed-form log10 ABFs over grid x
//            configurations with log-sum-exp averaging (gene_snp_pair.cpp:297-622).
//   phase 4 (K4, permutation statistic): per-CTA reduction over the gene's SNPs of the chosen
//            weighted ABF (max or log10-mean) or of the minimum p-value (gene.cpp:380-717).
// Data layout: everything lives in the sorted all-sample index space, rows padded to ldn doubles.
#pragma once

#include <cfloat>
#include <cstdint>

#include "device_math.cuh"

namespace eqb {

constexpr int MAXS = 64;  // subgroups (configuration masks are 64-bit)
constexpr int MAXQ = 31;  // covariates per subgroup
constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
constexpr double LN10 = 2.302585092994045684;

struct SubDev {
  const double *X;         // [M][ldn] genotypes in all-sample space (0 where absent)
  const double *Yall;      // [G][ldn] expression in all-sample space (NaN where absent/missing)
  const double *Call;      // [Q][ldn] covariates in all-sample space
  const uint8_t *gmask;    // [ldn] sample has a genotype in this subgroup
  const uint8_t *cmask;    // [ldn] sample has covariates in this subgroup
  const uint8_t *snp_has;  // [M]
  const uint8_t *gene_has; // [G]
  int Q;
  int pad;
};

struct DevParams {
  int S, N, ldn, analysis, bfs, qnorm, error_model, L, K, Qmax;
  double fiterr;
  long long M, G, C;
  const double *phi2L, *oma2L, *phi2S, *oma2S;
  const unsigned long long *cfg_mask; // [C] subgroup bitmask of each configuration, reference order
  const double *cfg_weight;           // [C] (1/S)(1/choose(S,|config|)), gene_snp_pair.cpp:590-592
  double size_weight[MAXS + 1];       // the same weight as a function of the configuration size
  const long long *cis_begin, *cis_end;
  SubDev sub[MAXS];
};

enum { STAT_NONE = 0, STAT_JOIN_MAX = 1, STAT_JOIN_AVG = 2, STAT_SEP_ALL = 3, STAT_SEP_PER = 4 };

struct LaunchArgs {
  const int *genes;  // work list of analysed genes
  int n_genes;
  int perms_per_gene; // 0 = identity permutation only (true pass)
  long long p0, P_total;
  const unsigned short *perm_tab; // [slot][P_total][N]
  const int *gene_slot;           // [n_genes] slot of the gene inside its write-group
  int which;                      // 1 = gen, 2 = sin / gen-sin, 3 = all
  int stat_kind;
  int want_outputs;
  int true_rules; // statistic of the TRUE data (NaN entries dropped first: gene.cpp:575-596)
  const long long *pair_off; // [n_genes]
  int *out_n;
  double *out_ss, *out_gen, *out_cfg, *out_w;
  double *out_stat; // [n_genes*per][P_total] (perms) or [n_genes*per] (identity)
  double *basis_ws; // global workspace (nullptr -> dynamic shared memory)
  double *table_ws;
  int *err_flag;
  double *hy_off; // --error hybrid: Vg_12 cache of the launch [gene][hy_stride][S (S - 1) / 2] (hybrid_offdiag_kernel)
  int hy_stride;  // SNPs per gene slot of hy_off (largest cis window of the launch)
};

// sizes (in doubles) of the two workspaces
__host__ __device__ inline int basis_rows(int Qmax, int qnorm) { return Qmax + 2 + (qnorm ? 2 : 0); }
__host__ __device__ inline size_t basis_doubles(int S, int Qmax, int ldn, int qnorm)
{
  return (size_t)S * basis_rows(Qmax, qnorm) * ldn;
}
__host__ __device__ inline size_t table_doubles(int S, int K) { return (size_t)3 * S + (size_t)3 * K * S; }

// utils::log10_weighted_sum semantics (utils_math.cpp:100-159) as an online accumulator:
// the running maximum starts at element 0 (a NaN there poisons the result), NaN elements are
// skipped, |result| <= DBL_EPSILON snaps to 0.
struct Lse {
  double m, acc;
  bool first_nan, any;
  __device__ void init()
  {
    m = -INFINITY;
    acc = 0.0;
    first_nan = false;
    any = false;
  }
  __device__ __noinline__ void add(double v, double w, bool is_first)
  {
    if (isnan(v)) {
      if (is_first) first_nan = true;
      return;
    }
    any = true;
    if (v > m) {
      acc = acc * exp10(m - v) + w; // exp10(-inf) = 0 on the first element
      m = v;
    } else
      acc += w * exp10(v - m);
  }
  __device__ void merge(const Lse &o)
  {
    first_nan = first_nan || o.first_nan;
    if (!o.any) return;
    if (!any) {
      m = o.m;
      acc = o.acc;
      any = true;
      return;
    }
    if (o.m > m) {
      acc = acc * exp10(m - o.m) + o.acc;
      m = o.m;
    } else
      acc += o.acc * exp10(o.m - m);
  }
  __device__ double result() const
  {
    if (first_nan) return nan("");
    double res = m + log10(acc);
    if (fabs(res) <= DBL_EPSILON) res = 0.0;
    return res;
  }
};

__device__ inline Lse warp_merge(Lse a)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Lse b;
    b.m = __shfl_xor_sync(0xffffffffu, a.m, o);
    b.acc = __shfl_xor_sync(0xffffffffu, a.acc, o);
    b.first_nan = __shfl_xor_sync(0xffffffffu, (int)a.first_nan, o);
    b.any = __shfl_xor_sync(0xffffffffu, (int)a.any, o);
    a.merge(b);
  }
  return a;
}

// CalcLog10AbfUvlr (gene_snp_pair.cpp:297-356) from the per-(grid point, subgroup) table
// tab[(k*S+s)*3 + {0,1,2}] = { 1/(v+phi2), bhat/(v+phi2), single-subgroup log10 ABF }
static __device__ __noinline__ double abf_from_table(const double *tab_k, unsigned long long mask, double oma2)
{
  double num = 0.0, den = 0.0, sing = 0.0;
  while (mask) {
    const int s = __ffsll((long long)mask) - 1;
    mask &= mask - 1;
    const double *e = tab_k + 3 * s;
    den += e[0];
    num += e[1];
    sing += e[2];
  }
  const double bbar = (den != 0.0) ? num / den : 0.0;
  const double V = (den != 0.0) ? 1.0 / den : INFINITY;
  if (bbar != 0.0 && V < INFINITY) {
    const double T2 = bbar * bbar / V;
    const double lbar =
        (T2 != 0.0) ? 0.5 * log10(V) - 0.5 * log10(V + oma2) + (0.5 * T2 * oma2 / (V + oma2)) / LN10 : 0.0;
    return lbar + sing;
  }
  return 0.0;
}

// one table entry; subgroups with |t| < 1e-8 contribute nothing (gene_snp_pair.cpp:314-316)
__device__ __forceinline__ void table_entry(double b, double v, double t, double phi2, double *e)
{
  if (fabs(t) < 1e-8) {
    e[0] = 0.0;
    e[1] = 0.0;
    e[2] = 0.0;
  } else {
    e[0] = 1.0 / (v + phi2);
    e[1] = b / (v + phi2);
    e[2] = 0.5 * log10(v) - 0.5 * log10(v + phi2) + (0.5 * t * t * phi2 / (v + phi2)) / LN10;
  }
}

// direct evaluation for the consistent configuration (one use per grid point: no table)
static __device__ __noinline__ double abf_direct(const double *st, int S, unsigned long long mask, double phi2, double oma2)
{
  double num = 0.0, den = 0.0, sing = 0.0;
  while (mask) {
    const int s = __ffsll((long long)mask) - 1;
    mask &= mask - 1;
    double e[3];
    table_entry(st[s], st[S + s], st[2 * S + s], phi2, e);
    den += e[0];
    num += e[1];
    sing += e[2];
  }
  const double bbar = (den != 0.0) ? num / den : 0.0;
  const double V = (den != 0.0) ? 1.0 / den : INFINITY;
  if (bbar != 0.0 && V < INFINITY) {
    const double T2 = bbar * bbar / V;
    const double lbar =
        (T2 != 0.0) ? 0.5 * log10(V) - 0.5 * log10(V + oma2) + (0.5 * T2 * oma2 / (V + oma2)) / LN10 : 0.0;
    return lbar + sing;
  }
  return 0.0;
}

// gsl_sort_index (index heapsort, the tie order of utils::qqnorm, utils_math.cpp:87), sequential
__device__ inline void heapsort_index(int *p, const double *data, int n)
{
  if (n == 0) return;
  for (int i = 0; i < n; ++i) p[i] = i;
  int N = n - 1;
  int k = N / 2;
  k++;
  do {
    k--;
    { // downheap
      int kk = k;
      const int pki = p[kk];
      while (kk <= N / 2) {
        int j = 2 * kk;
        if (j < N && data[p[j]] < data[p[j + 1]]) j++;
        if (!(data[pki] < data[p[j]])) break;
        p[kk] = p[j];
        kk = j;
      }
      p[kk] = pki;
    }
  } while (k > 0);
  while (N > 0) {
    const int tmp = p[0];
    p[0] = p[N];
    p[N] = tmp;
    N--;
    int kk = 0;
    const int pki = p[kk];
    while (kk <= N / 2) {
      int j = 2 * kk;
      if (j < N && data[p[j]] < data[p[j + 1]]) j++;
      if (!(data[pki] < data[p[j]])) break;
      p[kk] = p[j];
      kk = j;
    }
    p[kk] = pki;
  }
}

template <int NPL>
__global__ void __launch_bounds__(THREADS) pair_kernel(const DevParams *__restrict__ prm_, const LaunchArgs la)
{
  const DevParams &prm = *prm_;
  extern __shared__ double dyn_smem[];
  __shared__ int s_n[MAXS];
  __shared__ int s_rankz[MAXS];
  __shared__ unsigned int s_colvalid[MAXS];
  __shared__ double s_yy[MAXS], s_tss[MAXS], s_ybar[MAXS];
  __shared__ double w_part[WARPS][2];   // join statistic partials (m, acc) or max
  __shared__ int w_flag[WARPS][3];      // first_nan, any, count_nonnan
  __shared__ double w_sep[WARPS][MAXS]; // sep per-subgroup partial minima
  __shared__ int w_sep_nan[MAXS];

  const int S = prm.S, N = prm.N, ldn = prm.ldn, Qmax = prm.Qmax;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ppg = la.perms_per_gene > 0 ? la.perms_per_gene : 1;
  const int gi = blockIdx.x / ppg;
  const long long p = la.perms_per_gene > 0 ? la.p0 + (blockIdx.x % ppg) : -1;
  const int g = la.genes[gi];
  const long long mbeg = prm.cis_begin[g], mend = prm.cis_end[g];

  // workspaces
  const int brows = basis_rows(Qmax, prm.qnorm);
  const size_t nb = basis_doubles(S, Qmax, ldn, prm.qnorm), nt = table_doubles(S, prm.K);
  double *dyn = dyn_smem;
  double *basis = la.basis_ws ? la.basis_ws + (size_t)blockIdx.x * nb : dyn;
  if (!la.basis_ws) dyn += nb;
  double *tables = la.table_ws ? la.table_ws + (size_t)blockIdx.x * nt * WARPS : dyn;
  double *mytab = tables + (size_t)warp * nt;
  const unsigned short *perm =
      (p >= 0) ? la.perm_tab + ((size_t)la.gene_slot[gi] * la.P_total + p) * N : nullptr;

  if (threadIdx.x < MAXS) w_sep_nan[threadIdx.x] = 0;

  // ------------------------------------------------------------------ phase 1: per-subgroup setup
  for (int s = warp; s < S; s += WARPS) {
    const SubDev &sb = prm.sub[s];
    double *q = basis + (size_t)s * brows * ldn; // rows 0..Q: basis, row Qmax+1: residual phenotype
    double *yt = q + (size_t)(Qmax + 1) * ldn;
    int n = 0;
    if (sb.gene_has[g]) {
      const double *Yg = sb.Yall + (size_t)g * ldn;
      for (int i = lane; i < ldn; i += 32) {
        double yv = 0.0;
        bool keep = false;
        if (i < N) {
          const int j = perm ? (int)perm[i] : i;
          yv = Yg[j];
          keep = sb.gmask[i] && !isnan(yv);
        }
        yt[i] = keep ? yv : 0.0;
        q[i] = keep ? 1.0 : 0.0;
        n += keep ? 1 : 0;
      }
      n = warp_sum_int(n);
    }
    __syncwarp();
    if (n == 0) {
      if (lane == 0) {
        s_n[s] = 0;
        s_rankz[s] = 0;
        s_colvalid[s] = 0;
        s_yy[s] = 0.0;
        s_tss[s] = 0.0;
        s_ybar[s] = 0.0;
      }
      continue;
    }
    // --qnorm (utils_math.cpp:80-96): ranks of the kept values -> normal scores
    if (prm.qnorm) {
      double *vals = q + (size_t)(Qmax + 2) * ldn;     // scratch row: the kept values, compacted
      int *ord = (int *)(q + (size_t)(Qmax + 3) * ldn); // scratch row: sort permutation
      if (lane == 0) {
        int c = 0;
        for (int i = 0; i < N; ++i)
          if (q[i] != 0.0) vals[c++] = yt[i];
        heapsort_index(ord, vals, n);
        const double a = (n <= 10 ? 0.375 : 0.5);
        for (int r = 0; r < n; ++r) vals[ord[r]] = ugaussian_Pinv((r + 1 - a) / (n + 1 - 2 * a));
        c = 0;
        for (int i = 0; i < N; ++i)
          if (q[i] != 0.0) yt[i] = vals[c++];
      }
      __syncwarp();
    }
    // basis column 0: the intercept on the kept rows
    const double inv_sqrt_n = 1.0 / sqrt((double)n);
    for (int i = lane; i < ldn; i += 32) q[i] = (q[i] != 0.0) ? inv_sqrt_n : 0.0;
    __syncwarp();
    unsigned int colvalid = 1u;
    int rankz = 1;
    const int Q = sb.Q;
    for (int k = 1; k <= Q; ++k) {
      double *qk = q + (size_t)k * ldn;
      const double *Ck = sb.Call + (size_t)(k - 1) * ldn;
      double nrm0 = 0.0;
      int missing = 0;
      for (int i = lane; i < ldn; i += 32) {
        const bool keep = q[i] != 0.0;
        const double v = keep ? Ck[i] : 0.0;
        if (keep && !sb.cmask[i]) missing = 1;
        qk[i] = v;
        nrm0 += v * v;
      }
      nrm0 = warp_sum(nrm0);
      if (__any_sync(0xffffffffu, missing) && lane == 0) atomicExch(la.err_flag, 1); // gene_snp_pair.cpp:138-144
      __syncwarp();
      for (int pass = 0; pass < 2; ++pass) {
        for (int j = 0; j < k; ++j) {
          if (!((colvalid >> j) & 1u)) continue;
          const double *qj = q + (size_t)j * ldn;
          double h = 0.0;
          for (int i = lane; i < ldn; i += 32) h += qj[i] * qk[i];
          h = warp_sum(h);
          for (int i = lane; i < ldn; i += 32) qk[i] -= h * qj[i];
          __syncwarp();
        }
      }
      double nrm1 = 0.0;
      for (int i = lane; i < ldn; i += 32) nrm1 += qk[i] * qk[i];
      nrm1 = warp_sum(nrm1);
      if (nrm1 > 1e-20 * nrm0 && nrm1 > 0.0) {
        const double inv = 1.0 / sqrt(nrm1);
        for (int i = lane; i < ldn; i += 32) qk[i] *= inv;
        colvalid |= (1u << k);
        rankz++;
      } else {
        for (int i = lane; i < ldn; i += 32) qk[i] = 0.0;
      }
      __syncwarp();
    }
    // total sum of squares of the kept phenotype (gsl_stats_tss, utils_math.cpp:198)
    double ysum = 0.0;
    for (int i = lane; i < ldn; i += 32) ysum += yt[i];
    ysum = warp_sum(ysum);
    const double ybar = ysum / n;
    double tss = 0.0;
    for (int i = lane; i < ldn; i += 32)
      if (q[i] != 0.0) {
        const double d = yt[i] - ybar;
        tss += d * d;
      }
    tss = warp_sum(tss);
    // residual phenotype
    for (int pass = 0; pass < 2; ++pass) {
      for (int j = 0; j <= Q; ++j) {
        if (!((colvalid >> j) & 1u)) continue;
        const double *qj = q + (size_t)j * ldn;
        double h = 0.0;
        for (int i = lane; i < ldn; i += 32) h += qj[i] * yt[i];
        h = warp_sum(h);
        for (int i = lane; i < ldn; i += 32) yt[i] -= h * qj[i];
        __syncwarp();
      }
    }
    double yy = 0.0;
    for (int i = lane; i < ldn; i += 32) yy += yt[i] * yt[i];
    yy = warp_sum(yy);
    if (lane == 0) {
      s_n[s] = n;
      s_rankz[s] = rankz;
      s_colvalid[s] = colvalid;
      s_yy[s] = yy;
      s_tss[s] = tss;
      s_ybar[s] = ybar;
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ phases 2-3: per-SNP work
  const bool join = prm.analysis == 1;
  const int L = prm.L, K = prm.K;
  const long long C = (la.which == 1) ? 0 : ((la.which == 2) ? S : prm.C);
  double *st = mytab;          // [3][S] standardised bhat, var, t
  double *tab = mytab + 3 * S; // [K][S][3]

  Lse acc_stat; // join avg
  acc_stat.init();
  double max_stat = -INFINITY; // join max
  bool first_nan = false;
  int cnt_nonnan = 0;
  double sep_all_min = 1.0; // lane 0 meaningful

  if (la.stat_kind == STAT_SEP_PER)
    for (int s = lane; s < S; s += 32) w_sep[warp][s] = INFINITY;
  __syncwarp();

  for (long long m = mbeg + warp; m < mend; m += WARPS) {
    const bool is_first = (m == mbeg);
    unsigned long long has_mask = 0ull;
    double snp_pmin = 1.0;
    const long long pair = la.want_outputs ? la.pair_off[gi] + (m - mbeg) : 0;

    for (int s = 0; s < S; ++s) {
      const SubDev &sb = prm.sub[s];
      const int n = s_n[s];
      const bool have = (n > 0) && sb.snp_has[m];
      double pve = nan(""), sigmahat = nan(""), betahat = nan(""), se = nan(""), pval = nan("");
      if (have) {
        const double *q = basis + (size_t)s * brows * ldn;
        const double *yt = q + (size_t)(Qmax + 1) * ldn;
        const double *Xm = sb.X + (size_t)m * ldn;
        const int Q = sb.Q;
        const unsigned int colvalid = s_colvalid[s];
        const int rankz = s_rankz[s];
        if (n >= (2 + Q) + 1) { // utils_math.cpp:175: at least one residual degree of freedom
          double xr[NPL];
          double xraw2 = 0.0, xsum = 0.0;
#pragma unroll
          for (int j = 0; j < NPL; ++j) {
            const int i = lane + 32 * j;
            double v = 0.0;
            if (i < ldn) v = (q[i] != 0.0) ? Xm[i] : 0.0;
            xr[j] = v;
            xraw2 += v * v;
            xsum += v;
          }
          xraw2 = warp_sum(xraw2);
          xsum = warp_sum(xsum);
          for (int pass = 0; pass < 2; ++pass) {
            for (int k = 0; k <= Q; ++k) {
              if (!((colvalid >> k) & 1u)) continue;
              const double *qk = q + (size_t)k * ldn;
              double h = 0.0;
#pragma unroll
              for (int j = 0; j < NPL; ++j) {
                const int i = lane + 32 * j;
                if (i < ldn) h += qk[i] * xr[j];
              }
              h = warp_sum(h);
#pragma unroll
              for (int j = 0; j < NPL; ++j) {
                const int i = lane + 32 * j;
                if (i < ldn) xr[j] -= h * qk[i];
              }
            }
          }
          double xx = 0.0, xy = 0.0;
#pragma unroll
          for (int j = 0; j < NPL; ++j) {
            const int i = lane + 32 * j;
            if (i < ldn) {
              xx += xr[j] * xr[j];
              xy += xr[j] * yt[i];
            }
          }
          xx = warp_sum(xx);
          xy = warp_sum(xy);
          const double tssv = s_tss[s];
          if (xx > 1e-24 * xraw2 && xraw2 > 0.0) {
            // full-rank design: FWL form of the least-squares fit
            betahat = xy / xx;
            double rss = 0.0;
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
              const int i = lane + 32 * j;
              if (i < ldn) {
                const double r = yt[i] - betahat * xr[j];
                rss += r * r;
              }
            }
            rss = warp_sum(rss);
            const int rank = rankz + 1;
            pve = 1.0 - rss / tssv;
            sigmahat = sqrt(rss / (double)(n - rank));
            se = sigmahat / sqrt(xx);
            if (lane == 0) pval = 2.0 * tdist_Q(fabs(betahat / se), (double)(n - rank));
            pval = __shfl_sync(0xffffffffu, pval, 0);
          } else {
            // genotype inside span([1, covariates]) on the kept rows: rank-deficient design.
            // gsl_multifit_linear keeps the minimum-norm solution on the column-balanced matrix
            // (SURVEY.md App. B #9).
            const double rss = s_yy[s];
            const int rank = rankz;
            pve = 1.0 - rss / tssv;
            sigmahat = sqrt(rss / (double)(n - rank));
            if (xraw2 == 0.0) {
              betahat = 0.0;
              se = 0.0;
              pval = nan("");
            } else if (Q == 0) {
              // X = [1, c*1]: closed form of the balanced rank-1 pseudo-inverse
              const double cst = xsum / n;
              double f0 = 1.0, f1 = 1.0, s0 = (double)n, s1 = fabs(cst) * n;
              while (s0 > 1.0) { s0 /= 2.0; f0 *= 2.0; }
              while (s0 < 0.5) { s0 *= 2.0; f0 /= 2.0; }
              while (s1 > 1.0) { s1 /= 2.0; f1 *= 2.0; }
              while (s1 < 0.5) { s1 *= 2.0; f1 /= 2.0; }
              const double a = 1.0 / f0, b = cst / f1, ab2 = a * a + b * b;
              const double ybar = s_ybar[s];
              betahat = b * ybar / (ab2 * f1);
              const double s2 = rss / (double)(n - rank);
              se = sqrt(s2 * b * b / (ab2 * ab2 * n) / (f1 * f1));
              if (lane == 0) pval = 2.0 * tdist_Q(fabs(betahat / se), (double)(n - rank));
              pval = __shfl_sync(0xffffffffu, pval, 0);
            } else {
              betahat = nan("");
              se = nan("");
              pval = nan("");
              if (lane == 0) atomicOr(la.err_flag + 1, 1); // documented unsupported degenerate design
            }
          }
        }
        has_mask |= (1ull << s);
        if (lane == 0) {
          // inputs of the standardisation, stored in the per-warp table rows
          st[s] = betahat;
          st[S + s] = se;
          st[2 * S + s] = sigmahat;
        }
        if (pval < snp_pmin) snp_pmin = pval; // gene.cpp:542-543
      }
      if (la.want_outputs && lane == 0) {
        if (la.out_n) la.out_n[pair * S + s] = have ? n : 0;
        if (la.out_ss) {
          double *o = la.out_ss + (pair * S + s) * 5;
          o[0] = pve;
          o[1] = sigmahat;
          o[2] = betahat;
          o[3] = se;
          o[4] = pval;
        }
      }
      if (la.stat_kind == STAT_SEP_PER && lane == 0) {
        // gene.cpp:414-429: 1.0 without data, else the p-value (possibly NaN)
        const double v = (sb.gene_has[g] && sb.snp_has[m]) ? (have ? pval : nan("")) : 1.0;
        if (isnan(v)) {
          if (is_first) atomicExch(&w_sep_nan[s], 1);
        } else if (v < w_sep[warp][s])
          w_sep[warp][s] = v;
      }
    }
    if (la.stat_kind == STAT_SEP_ALL && snp_pmin < sep_all_min) sep_all_min = snp_pmin;
    if (!join) continue;

    // -------- standardisation (gene_snp_pair.cpp:256-290), one lane per subgroup
    __syncwarp();
    for (int s = lane; s < S; s += 32) {
      if (!((has_mask >> s) & 1ull)) continue;
      const double beta = st[s], se_ = st[S + s], sg = st[2 * S + s];
      double bhat = beta / sg, sebhat = se_ / sg, t = bhat / sebhat;
      double ob = nan(""), ov = nan(""), ot = nan("");
      if (!isnan(t)) {
        const double nu = (double)s_n[s] - 2.0 - prm.sub[s].Q;
        t = ugaussian_Pinv(tdist_P(-fabs(bhat / sebhat), nu));
        if (fabs(t) > 1e-8) {
          const double sg2 = fabs(beta) / (fabs(t) * sebhat);
          bhat = beta / sg2;
          sebhat = fabs(bhat / t);
        } else {
          bhat = 0.0;
          sebhat = INFINITY;
        }
        ob = bhat;
        ov = sebhat * sebhat;
        ot = t;
      }
      st[s] = ob;
      st[S + s] = ov;
      st[2 * S + s] = ot;
    }
    __syncwarp();

    // -------- consistent configuration on gridL: gen, gen-fix, gen-maxh (gene_snp_pair.cpp:364-416)
    double w_gen[3] = {nan(""), nan(""), nan("")};
    const int nvar = (p >= 0) ? 1 : 3; // permutations only ever read "gen" of these three
    for (int var = 0; var < nvar; ++var) {
      Lse a;
      a.init();
      for (int k = lane; k < L; k += 32) {
        const double ph = prm.phi2L[k], om = prm.oma2L[k];
        const double phi2 = (var == 0) ? ph : ((var == 1) ? 0.0 : ph + om);
        const double oma2 = (var == 0) ? om : ((var == 1) ? ph + om : 0.0);
        const double v = abf_direct(st, S, has_mask, phi2, oma2);
        if (la.want_outputs && la.out_gen) la.out_gen[(pair * 3 + var) * L + k] = v;
        a.add(v, 1.0 / (double)L, k == 0);
      }
      a = warp_merge(a);
      w_gen[var] = (L > 0) ? a.result() : nan("");
    }
    double w_gensin = nan(""), w_all = nan("");
    if (la.which >= 2) {
      // -------- per-(grid point, subgroup) table on gridS
      for (int e = lane; e < K * S; e += 32) {
        const int k = e / S, s = e % S;
        double *te = tab + (size_t)e * 3;
        if ((has_mask >> s) & 1ull)
          table_entry(st[s], st[S + s], st[2 * S + s], prm.phi2S[k], te);
        else {
          te[0] = 0.0;
          te[1] = 0.0;
          te[2] = 0.0;
        }
      }
      __syncwarp();
      // -------- configurations (gene_snp_pair.cpp:422-550): one lane per configuration
      Lse lite, bma;
      lite.init();
      bma.init();
      for (long long c = lane; c < C; c += 32) {
        const unsigned long long cm = (la.which == 2) ? (1ull << c) : prm.cfg_mask[c];
        const unsigned long long mask = cm & has_mask;
        Lse a;
        a.init();
        for (int k = 0; k < K; ++k) {
          const double v = abf_from_table(tab + (size_t)k * S * 3, mask, prm.oma2S[k]);
          if (la.want_outputs && la.out_cfg) la.out_cfg[(pair * C + c) * K + k] = v;
          a.add(v, 1.0 / (double)K, k == 0);
        }
        const double wc = a.result();
        if (la.want_outputs && la.out_w) la.out_w[pair * (5 + C) + 5 + c] = wc;
        if (c < S) lite.add(wc, 0.5 / (double)S, c == 0);         // CalcBMAlite, gene_snp_pair.cpp:552-570
        if (la.which == 3) bma.add(wc, prm.cfg_weight[c], c == 0); // CalcBMA, gene_snp_pair.cpp:572-602
      }
      lite = warp_merge(lite);
      lite.add(w_gen[0], 0.5, false);
      w_gensin = lite.result();
      if (la.which == 3) {
        bma = warp_merge(bma);
        w_all = bma.result();
      }
    }
    if (la.want_outputs && la.out_w && lane == 0) {
      double *o = la.out_w + pair * (5 + C);
      o[0] = w_gen[0];
      o[1] = w_gen[1];
      o[2] = w_gen[2];
      o[3] = w_gensin;
      o[4] = w_all;
    }
    // -------- running permutation statistic over the SNPs of the gene
    if (la.stat_kind == STAT_JOIN_MAX || la.stat_kind == STAT_JOIN_AVG) {
      const double v = (la.which == 1) ? w_gen[0] : ((la.which == 2) ? w_gensin : w_all);
      if (isnan(v)) {
        if (is_first) first_nan = true;
      } else {
        cnt_nonnan++;
        if (v > max_stat) max_stat = v;
        acc_stat.add(v, 1.0, false);
      }
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------ phase 4: CTA reduction
  if (la.stat_kind == STAT_NONE) return;
  const int per = (la.stat_kind == STAT_SEP_PER) ? S : 1;
  double *out = (p >= 0) ? la.out_stat + ((size_t)gi * per) * la.P_total + p : la.out_stat + (size_t)gi * per;
  const size_t ostride = (p >= 0) ? (size_t)la.P_total : 1;
  if (lane == 0) {
    if (la.stat_kind == STAT_JOIN_MAX) {
      w_part[warp][0] = max_stat;
    } else if (la.stat_kind == STAT_JOIN_AVG) {
      w_part[warp][0] = acc_stat.m;
      w_part[warp][1] = acc_stat.acc;
    } else if (la.stat_kind == STAT_SEP_ALL) {
      w_part[warp][0] = sep_all_min;
    }
    w_flag[warp][0] = first_nan ? 1 : 0;
    w_flag[warp][1] = acc_stat.any ? 1 : 0;
    w_flag[warp][2] = cnt_nonnan;
  }
  __syncthreads();
  if (la.stat_kind == STAT_SEP_PER) {
    for (int s = threadIdx.x; s < S; s += THREADS) {
      double v = INFINITY;
      for (int w = 0; w < WARPS; ++w) v = fmin(v, w_sep[w][s]);
      // min_element semantics (gene.cpp:429): NaN only if the first SNP's value is NaN
      if (la.true_rules) { // FindMinTruePvaluePerSubgroup (gene.cpp:369-378): starts at 1.0, NaN never wins
        if (!(v < 1.0)) v = 1.0;
      } else if (w_sep_nan[s])
        v = nan("");
      else if (mend == mbeg || isinf(v))
        v = 1.0;
      out[(size_t)s * ostride] = v;
    }
    return;
  }
  if (threadIdx.x == 0) {
    const long long Mg = mend - mbeg;
    bool fn = false;
    int nn = 0;
    for (int w = 0; w < WARPS; ++w) {
      fn = fn || w_flag[w][0];
      nn += w_flag[w][2];
    }
    double res;
    if (la.stat_kind == STAT_JOIN_MAX) {
      double v = -INFINITY;
      for (int w = 0; w < WARPS; ++w) v = fmax(v, w_part[w][0]);
      res = (fn && !la.true_rules) ? nan("") : v; // max_element (gene.cpp:678) vs FindMaxTrueL10Abf (:575)
    } else if (la.stat_kind == STAT_JOIN_AVG) {
      Lse t;
      t.init();
      for (int w = 0; w < WARPS; ++w) {
        Lse o;
        o.m = w_part[w][0];
        o.acc = w_part[w][1];
        o.any = w_flag[w][1] != 0;
        o.first_nan = false;
        t.merge(o);
      }
      // log10_weighted_sum over all SNPs (gene.cpp:690) or over the non-NaN ones (gene.cpp:587-596)
      const double size = la.true_rules ? (double)nn : (double)Mg;
      if ((fn && !la.true_rules) || nn == 0)
        res = nan("");
      else {
        res = t.m + log10(t.acc * (1.0 / size));
        if (fabs(res) <= DBL_EPSILON) res = 0.0;
      }
    } else {
      double v = 1.0;
      for (int w = 0; w < WARPS; ++w) v = fmin(v, w_part[w][0]);
      res = v;
    }
    out[0] = res;
  }
}

} // namespace eqb
