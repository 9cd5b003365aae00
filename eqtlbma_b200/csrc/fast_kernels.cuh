// fast_kernels.cuh -- the split K1 / K2+K3 path for the common case in which the set of
// individuals entering a subgroup's regression does not depend on the gene (no NaN expression
// inside the subgroup, gene expressed there) and the data are not permuted.
//
//   prep_basis_kernel  (K1a) one CTA per subgroup: orthonormal basis of [1, covariates] on the
//                      subgroup's individuals (CGS2), its rank.              gene_snp_pair.cpp:130-150
//   prep_y_kernel      (K1b) one warp per (gene, subgroup): residual phenotype, yy, tss, mean; flags
//                      the (gene, subgroup) cells that need the general path (NaN, absent gene).
//   prep_x_kernel      (K1c) one warp per SNP, all subgroups in one pass over the genotype row:
//                      residual genotype sum of squares xx = |x - QQ'x|^2 (+ raw moments).
//   fast_pair_kernel   (K2+K3) one CTA per gene, tiles of T cis SNPs, three compact phases:
//                      A  warp per SNP: x . ytil_s for every subgroup (the cis-banded contraction)
//                      B  thread per (pair, subgroup): betahat, se, sigmahat, pve, p-value and the
//                         standardisation (one Student tail + one normal quantile)
//                                           utils_math.cpp:166-209, gene_snp_pair.cpp:256-290
//                      C  thread per (pair, row): ABFs over the grid with online log-sum-exp
//                                           gene_snp_pair.cpp:297-622
// Genes with any non-generic (gene, subgroup) cell are routed to pair_kernel (general path).
#pragma once

#include "pair_kernel.cuh"

namespace eqb {

struct FastSub {
  const double *Bs;    // [Q+1][ldn] orthonormal basis (generic mask)
  const double *Ytil;  // [G][ldn] residual phenotype (valid where ystat flag = 1)
  const double *ystat; // [G][4] yy, tss, ybar, generic-flag
  const double *xstat; // [M][3] xx, xraw2, xsum
  int n, rankz;        // kept individuals, rank of [1, covariates]
  unsigned int colvalid;
  int pad;
};

struct FastParams {
  FastSub sub[MAXS];
};

struct FastArgs {
  const int *genes;            // fast genes of this launch
  int n_genes;
  int T;                       // pairs per tile
  int which;                   // 1 gen, 2 sin, 3 all
  long long n_pairs;           // pairs of the fast genes
  const long long *fast_base;  // [n_genes] first compact pair index of each fast gene
  const long long *pair_off;   // [n_genes] first OUTPUT pair index of each fast gene
  int *out_n;
  double *out_ss, *out_gen, *out_cfg, *out_w;
};

// ---------------------------------------------------------------- K1a
__global__ void __launch_bounds__(32) prep_basis_kernel(const DevParams *__restrict__ prm_, double *const *Bs_all,
                                                        const uint8_t *const *emask_all, int *__restrict__ n_out,
                                                        int *__restrict__ rankz_out, unsigned int *__restrict__ colvalid_out,
                                                        int *__restrict__ err_flag)
{
  const DevParams &prm = *prm_;
  const int s = blockIdx.x, lane = threadIdx.x, ldn = prm.ldn, N = prm.N;
  const SubDev &sb = prm.sub[s];
  double *q = Bs_all[s];
  const uint8_t *emask = emask_all[s];
  int n = 0;
  for (int i = lane; i < ldn; i += 32) {
    const bool keep = (i < N) && sb.gmask[i] && emask[i];
    q[i] = keep ? 1.0 : 0.0;
    n += keep ? 1 : 0;
  }
  n = warp_sum_int(n);
  __syncwarp();
  unsigned int colvalid = 0;
  int rankz = 0;
  if (n > 0) {
    const double inv_sqrt_n = 1.0 / sqrt((double)n);
    for (int i = lane; i < ldn; i += 32) q[i] = (q[i] != 0.0) ? inv_sqrt_n : 0.0;
    __syncwarp();
    colvalid = 1u;
    rankz = 1;
    for (int k = 1; k <= sb.Q; ++k) {
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
      if (__any_sync(0xffffffffu, missing) && lane == 0) atomicExch(err_flag + 2, 1); // only fatal if a gene uses it
      __syncwarp();
      for (int pass = 0; pass < 2; ++pass)
        for (int j = 0; j < k; ++j) {
          if (!((colvalid >> j) & 1u)) continue;
          const double *qj = q + (size_t)j * ldn;
          double h = 0.0;
          for (int i = lane; i < ldn; i += 32) h += qj[i] * qk[i];
          h = warp_sum(h);
          for (int i = lane; i < ldn; i += 32) qk[i] -= h * qj[i];
          __syncwarp();
        }
      double nrm1 = 0.0;
      for (int i = lane; i < ldn; i += 32) nrm1 += qk[i] * qk[i];
      nrm1 = warp_sum(nrm1);
      if (nrm1 > 1e-20 * nrm0 && nrm1 > 0.0) {
        const double inv = 1.0 / sqrt(nrm1);
        for (int i = lane; i < ldn; i += 32) qk[i] *= inv;
        colvalid |= (1u << k);
        rankz++;
      } else
        for (int i = lane; i < ldn; i += 32) qk[i] = 0.0;
      __syncwarp();
    }
  }
  if (lane == 0) {
    n_out[s] = n;
    rankz_out[s] = rankz;
    colvalid_out[s] = colvalid;
  }
}

// ---------------------------------------------------------------- K1b
__global__ void __launch_bounds__(THREADS) prep_y_kernel(const DevParams *__restrict__ prm_, const FastParams *__restrict__ fp_,
                                                         double *const *Ytil_all, double *const *ystat_all)
{
  const DevParams &prm = *prm_;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * WARPS + warp;
  const int S = prm.S, ldn = prm.ldn;
  if (item >= prm.G * S) return;
  const long long g = item / S;
  const int s = (int)(item % S);
  const SubDev &sb = prm.sub[s];
  const FastSub &fs = fp_->sub[s];
  double *yt = Ytil_all[s] + (size_t)g * ldn;
  double *ys = ystat_all[s] + (size_t)g * 4;
  const double *q = fs.Bs;
  const int n = fs.n;
  bool generic = sb.gene_has[g] && n > 0 && !prm.qnorm; // --qnorm goes through the general path
  double ysum = 0.0;
  if (generic) {
    const double *Yg = sb.Yall + (size_t)g * ldn;
    int bad = 0;
    for (int i = lane; i < ldn; i += 32) {
      const bool keep = q[i] != 0.0;
      const double v = keep ? Yg[i] : 0.0;
      if (keep && isnan(v)) bad = 1;
      yt[i] = v;
      ysum += v;
    }
    if (__any_sync(0xffffffffu, bad)) generic = false;
  }
  if (!generic) {
    if (lane == 0) {
      ys[0] = 0.0;
      ys[1] = 0.0;
      ys[2] = 0.0;
      // 2 = nothing to compute in this cell (gene not expressed here / no individual), 0 = general path
      ys[3] = (!sb.gene_has[g] || n == 0) ? 2.0 : 0.0;
    }
    for (int i = lane; i < ldn; i += 32) yt[i] = 0.0;
    return;
  }
  __syncwarp();
  ysum = warp_sum(ysum);
  const double ybar = ysum / n;
  double tss = 0.0;
  for (int i = lane; i < ldn; i += 32)
    if (q[i] != 0.0) {
      const double d = yt[i] - ybar;
      tss += d * d;
    }
  tss = warp_sum(tss);
  const int Q = sb.Q;
  for (int pass = 0; pass < 2; ++pass)
    for (int j = 0; j <= Q; ++j) {
      if (!((fs.colvalid >> j) & 1u)) continue;
      const double *qj = q + (size_t)j * ldn;
      double h = 0.0;
      for (int i = lane; i < ldn; i += 32) h += qj[i] * yt[i];
      h = warp_sum(h);
      for (int i = lane; i < ldn; i += 32) yt[i] -= h * qj[i];
      __syncwarp();
    }
  double yy = 0.0;
  for (int i = lane; i < ldn; i += 32) yy += yt[i] * yt[i];
  yy = warp_sum(yy);
  if (lane == 0) {
    ys[0] = yy;
    ys[1] = tss;
    ys[2] = ybar;
    ys[3] = 1.0;
  }
}

// ---------------------------------------------------------------- K1c
// One warp per SNP.  The projection on the basis is classical Gram-Schmidt applied twice (CGS2):
// inside a pass the Q+1 dot products are independent, so their partial sums and the warp
// reductions overlap instead of forming a dependent chain.  Subgroups that share the genotype
// matrix, the mask and the covariates (dup_of[s] >= 0) reuse the result of the earlier subgroup.
template <int NPL>
__global__ void __launch_bounds__(THREADS) prep_x_kernel(const DevParams *__restrict__ prm_, const FastParams *__restrict__ fp_,
                                                         double *const *xstat_all, const int *__restrict__ dup_of)
{
  const DevParams &prm = *prm_;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m = (long long)blockIdx.x * WARPS + warp;
  if (m >= prm.M) return;
  const int S = prm.S, ldn = prm.ldn;
  for (int s = 0; s < S; ++s) {
    const SubDev &sb = prm.sub[s];
    const FastSub &fs = fp_->sub[s];
    double *xs = xstat_all[s] + (size_t)m * 3;
    if (!sb.snp_has[m] || fs.n == 0) {
      if (lane == 0) {
        xs[0] = 0.0;
        xs[1] = 0.0;
        xs[2] = 0.0;
      }
      continue;
    }
    if (dup_of[s] >= 0 && prm.sub[dup_of[s]].snp_has[m]) {
      // same genotype row, same individuals, same covariates: identical residual
      if (lane == 0) {
        const double *src = xstat_all[dup_of[s]] + (size_t)m * 3;
        xs[0] = src[0];
        xs[1] = src[1];
        xs[2] = src[2];
      }
      continue;
    }
    const double *q = fs.Bs;
    const double *Xm = sb.X + (size_t)m * ldn;
    double xr[NPL];
    double xraw2 = 0.0, xsum = 0.0;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int i = lane + 32 * j;
      const double v = (i < ldn && q[i] != 0.0) ? Xm[i] : 0.0;
      xr[j] = v;
      xraw2 += v * v;
      xsum += v;
    }
    xraw2 = warp_sum(xraw2);
    xsum = warp_sum(xsum);
    const int Q = sb.Q;
    for (int pass = 0; pass < 2; ++pass) {
      for (int k0 = 0; k0 <= Q; k0 += 4) {
        double h[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
          const int i = lane + 32 * j;
          if (i < ldn) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
              if (k0 + a <= Q) h[a] += q[(size_t)(k0 + a) * ldn + i] * xr[j];
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int a = 0; a < 4; ++a) h[a] += __shfl_xor_sync(0xffffffffu, h[a], o);
        }
        // NB: the four projections of a block are removed together (CGS inside the block)
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
          const int i = lane + 32 * j;
          if (i < ldn) {
            double v = xr[j];
#pragma unroll
            for (int a = 0; a < 4; ++a)
              if (k0 + a <= Q && ((fs.colvalid >> (k0 + a)) & 1u)) v -= h[a] * q[(size_t)(k0 + a) * ldn + i];
            xr[j] = v;
          }
        }
      }
    }
    double xx = 0.0;
#pragma unroll
    for (int j = 0; j < NPL; ++j) xx += xr[j] * xr[j];
    xx = warp_sum(xx);
    if (lane == 0) {
      xs[0] = xx;
      xs[1] = xraw2;
      xs[2] = xsum;
    }
  }
}

// ---------------------------------------------------------------- K2 + K3
// per-thread standardisation + summary statistics of one (pair, subgroup)
struct PairStat {
  double pve, sigmahat, betahat, se, pval; // outputs of utils::FitSingleGeneWithSingleSnp
  double b, v, t;                          // standardised (gene_snp_pair.cpp:256-290)
};

__device__ __noinline__ void stats_from_dots(double xy, double xx, double xraw2, double xsum, double yy, double tss,
                                             double ybar, int n, int Q, int rankz, PairStat &o)
{
  const double qn = nan("");
  o.pve = o.sigmahat = o.betahat = o.se = o.pval = qn;
  o.b = o.v = o.t = qn;
  if (n < (2 + Q) + 1) return; // utils_math.cpp:175
  double tail = qn; // two-sided Student tail of |betahat/se| with n - rank degrees of freedom
  int rank;
  if (xx > 1e-24 * xraw2 && xraw2 > 0.0) {
    rank = rankz + 1;
    o.betahat = xy / xx;
    double rss = yy - xy * o.betahat;
    if (rss < 0.0) rss = 0.0;
    o.pve = 1.0 - rss / tss;
    o.sigmahat = sqrt(rss / (double)(n - rank));
    o.se = o.sigmahat / sqrt(xx);
  } else {
    rank = rankz;
    const double rss = yy;
    o.pve = 1.0 - rss / tss;
    o.sigmahat = sqrt(rss / (double)(n - rank));
    if (xraw2 == 0.0) {
      o.betahat = 0.0;
      o.se = 0.0;
    } else if (Q == 0) {
      const double cst = xsum / n;
      double f0 = 1.0, f1 = 1.0, s0 = (double)n, s1 = fabs(cst) * n;
      while (s0 > 1.0) { s0 /= 2.0; f0 *= 2.0; }
      while (s0 < 0.5) { s0 *= 2.0; f0 /= 2.0; }
      while (s1 > 1.0) { s1 /= 2.0; f1 *= 2.0; }
      while (s1 < 0.5) { s1 *= 2.0; f1 /= 2.0; }
      const double a = 1.0 / f0, b = cst / f1, ab2 = a * a + b * b;
      o.betahat = b * ybar / (ab2 * f1);
      const double s2 = rss / (double)(n - rank);
      o.se = sqrt(s2 * b * b / (ab2 * ab2 * n) / (f1 * f1));
    } else
      return; // documented unsupported degenerate design (NaN)
  }
  const double tt = o.betahat / o.se;
  double central;
  if (!isnan(tt)) {
    tdist_tails(fabs(tt), (double)(n - rank), tail, central);
    o.pval = (fabs(tt) > 0.0) ? tail : 1.0; // 2 * gsl_cdf_tdist_Q(|t|, n - rank)
  }
  // standardisation
  double bhat = o.betahat / o.sigmahat, sebhat = o.se / o.sigmahat, t = bhat / sebhat;
  if (isnan(t)) return;
  const double nu = (double)n - 2.0 - Q;
  double P;
  if (nu == (double)(n - rank) && !isnan(tail))
    P = (t == 0.0) ? 0.5 : 0.5 * tail; // gsl_cdf_tdist_P(-|t|, nu) shares the tail
  else
    P = tdist_P(-fabs(t), nu);
  t = ugaussian_Pinv(P);
  if (fabs(t) > 1e-8) {
    const double sg2 = fabs(o.betahat) / (fabs(t) * sebhat);
    bhat = o.betahat / sg2;
    sebhat = fabs(bhat / t);
  } else {
    bhat = 0.0;
    sebhat = INFINITY;
  }
  o.b = bhat;
  o.v = sebhat * sebhat;
  o.t = t;
}

// unique phi2 values of the consistent-configuration rows (gen / gen-fix / gen-maxh on gridL):
// uphi[UL]; row r, grid point k uses uphi[idxL[r*L+k]] and omaL[r*L+k]
struct GridTab {
  const double *uphi;
  const int *idxL;
  const double *omaL;
  int UL;
  int pad;
};

// ES-model log10 ABF from the sums over the active subgroups (CalcLog10AbfUvlr, gene_snp_pair.cpp:332-355)
// lbar = 0.5 log10(V) - 0.5 log10(V+oma2) + 0.5 T2 oma2/(V+oma2)/ln10 with V = 1/den, T2 = num^2/den,
// rewritten as -0.5 log10(1 + oma2 den) + 0.5 num^2 oma2 / (1 + oma2 den) / ln10
__device__ __forceinline__ double abf_from_sums(double den, double num, double sing, double oma2)
{
  const double bbar = (den != 0.0) ? num / den : 0.0;
  const double V = (den != 0.0) ? 1.0 / den : INFINITY;
  if (bbar != 0.0 && V < INFINITY) {
    const double T2 = bbar * bbar / V;
    if (T2 == 0.0) return sing;
    const double z = 1.0 + oma2 * den;
    return sing + (-0.5 * log1p(oma2 * den) + 0.5 * num * num * oma2 / z) / LN10;
  }
  return 0.0;
}

// one (subgroup, phi2) term: { 1/(v+phi2), b/(v+phi2), single-subgroup log10 ABF }, the latter as
// -0.5 log10(1 + phi2/v) + 0.5 t^2 phi2/(v+phi2)/ln10  (gene_snp_pair.cpp:314-324)
__device__ __forceinline__ void term_entry(double b, double v, double t, double phi2, double &d, double &bd, double &sg)
{
  if (fabs(t) < 1e-8) {
    d = 0.0;
    bd = 0.0;
    sg = 0.0;
  } else {
    const double inv = 1.0 / (v + phi2);
    d = inv;
    bd = b * inv;
    sg = (-0.5 * log1p(phi2 / v) + 0.5 * t * t * phi2 * inv) / LN10;
  }
}

__global__ void __launch_bounds__(THREADS) fast_pair_kernel(const DevParams *__restrict__ prm_,
                                                            const FastParams *__restrict__ fp_, const FastArgs fa,
                                                            const GridTab gt)
{
  const DevParams &prm = *prm_;
  extern __shared__ double fsm[];
  const int S = prm.S, ldn = prm.ldn, L = prm.L, K = prm.K, T = fa.T, UL = gt.UL;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long C = (fa.which == 1) ? 0 : ((fa.which == 2) ? S : prm.C);
  const bool join = prm.analysis == 1;
  const int nrow_small = 3 + ((fa.which == 2) ? S : 0); // rows whose values are staged in shared memory
  const int vals_per_pair = 3 * L + ((fa.which == 2) ? S * K : 0);
  // shared memory carve-up
  double *xy = fsm;                                          // [T][S]
  double *st = xy + (size_t)T * S;                           // [T][3][S]
  double *agg = st + (size_t)T * 3 * S;                      // [T][UL][3]
  double *vals = agg + (size_t)T * UL * 3;                   // [T][vals_per_pair]
  double *wrow = vals + (size_t)T * vals_per_pair;           // [T][3+S] weighted small rows
  double *tab = wrow + (size_t)T * (3 + S);                  // [T][K][S][3] (which == 3)
  unsigned long long *hasm = (unsigned long long *)(tab + ((fa.which == 3) ? (size_t)T * K * S * 3 : 0)); // [T]
  long long *s_pair = (long long *)(hasm + T);               // [T] output pair index
  long long *s_m = s_pair + T;                               // [T] SNP index
  int *s_gene = (int *)(s_m + T);                            // [T] gene id

  const long long q0 = (long long)blockIdx.x * T;
  const int tn = (int)min((long long)T, fa.n_pairs - q0);
  // map the tile's pairs to (gene, SNP, output index)
  for (int j = threadIdx.x; j < tn; j += THREADS) {
    const long long q = q0 + j;
    int lo = 0, hi = fa.n_genes; // last gene with fast_base <= q
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (fa.fast_base[mid] <= q) lo = mid;
      else hi = mid;
    }
    const int g = fa.genes[lo];
    const long long off = q - fa.fast_base[lo];
    s_gene[j] = g;
    s_m[j] = prm.cis_begin[g] + off;
    s_pair[j] = fa.pair_off[lo] + off;
    hasm[j] = 0ull;
  }
  __syncthreads();
  // ---------------- phase A: contraction x . ytil_s (one warp per pair)
  for (int j = warp; j < tn; j += WARPS) {
    const long long m = s_m[j];
    const size_t grow = (size_t)s_gene[j] * ldn;
    for (int s0 = 0; s0 < S; s0 += 8) {
      double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const int sn = min(8, S - s0);
      bool same = true;
      for (int a = 1; a < sn; ++a) same = same && (prm.sub[s0 + a].X == prm.sub[s0].X);
      if (same) {
        const double *Xm = prm.sub[s0].X + (size_t)m * ldn;
        for (int i = lane; i < ldn; i += 32) {
          const double x = Xm[i];
#pragma unroll
          for (int a = 0; a < 8; ++a)
            if (a < sn) acc[a] += x * fp_->sub[s0 + a].Ytil[grow + i];
        }
      } else {
        for (int a = 0; a < sn; ++a) {
          const double *Xa = prm.sub[s0 + a].X + (size_t)m * ldn;
          const double *Ya = fp_->sub[s0 + a].Ytil + grow;
          for (int i = lane; i < ldn; i += 32) acc[a] += Xa[i] * Ya[i];
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int a = 0; a < 8; ++a) acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], o);
      }
      if (lane < sn) {
        double r = acc[0];
#pragma unroll
        for (int a = 1; a < 8; ++a)
          if (lane == a) r = acc[a];
        xy[(size_t)j * S + s0 + lane] = r;
      }
    }
  }
  __syncthreads();
  // ---------------- phase B: thread per (pair, subgroup): summary statistics + standardisation
  for (int it = threadIdx.x; it < tn * S; it += THREADS) {
    const int j = it / S, s = it % S;
    const long long m = s_m[j];
    const int g = s_gene[j];
    const SubDev &sb = prm.sub[s];
    const FastSub &fs = fp_->sub[s];
    const double *ys = fs.ystat + (size_t)g * 4;
    const bool have = sb.gene_has[g] && sb.snp_has[m] && fs.n > 0;
    PairStat ps;
    ps.pve = ps.sigmahat = ps.betahat = ps.se = ps.pval = nan("");
    ps.b = ps.v = ps.t = nan("");
    if (have) {
      const double *xs = fs.xstat + (size_t)m * 3;
      stats_from_dots(xy[(size_t)j * S + s], xs[0], xs[1], xs[2], ys[0], ys[1], ys[2], fs.n, sb.Q, fs.rankz, ps);
      atomicOr(&hasm[j], 1ull << s);
    }
    st[((size_t)j * 3 + 0) * S + s] = ps.b;
    st[((size_t)j * 3 + 1) * S + s] = ps.v;
    st[((size_t)j * 3 + 2) * S + s] = ps.t;
    const long long pair = s_pair[j];
    if (fa.out_n) fa.out_n[pair * S + s] = have ? fs.n : 0;
    if (fa.out_ss) {
      double *o = fa.out_ss + (pair * S + s) * 5;
      o[0] = ps.pve;
      o[1] = ps.sigmahat;
      o[2] = ps.betahat;
      o[3] = ps.se;
      o[4] = ps.pval;
    }
  }
  __syncthreads();
  if (!join) return;
  // ---------------- phase C0: sums over the subgroups with results, per unique phi2 (consistent configuration)
  for (int it = threadIdx.x; it < tn * UL; it += THREADS) {
    const int j = it / UL, u = it % UL;
    const double *stj = st + (size_t)j * 3 * S;
    unsigned long long mask = hasm[j];
    const double phi2 = gt.uphi[u];
    double den = 0.0, num = 0.0, sing = 0.0;
    while (mask) {
      const int s = __ffsll((long long)mask) - 1;
      mask &= mask - 1;
      double d, bd, sg;
      term_entry(stj[s], stj[S + s], stj[2 * S + s], phi2, d, bd, sg);
      den += d;
      num += bd;
      sing += sg;
    }
    double *a = agg + ((size_t)j * UL + u) * 3;
    a[0] = den;
    a[1] = num;
    a[2] = sing;
  }
  if (fa.which == 3) {
    for (int it = threadIdx.x; it < tn * K * S; it += THREADS) {
      const int j = it / (K * S), e = it % (K * S), k = e / S, s = e % S;
      const double *stj = st + (size_t)j * 3 * S;
      double *te = tab + ((size_t)j * K * S + e) * 3;
      if ((hasm[j] >> s) & 1ull)
        term_entry(stj[s], stj[S + s], stj[2 * S + s], prm.phi2S[k], te[0], te[1], te[2]);
      else {
        te[0] = 0.0;
        te[1] = 0.0;
        te[2] = 0.0;
      }
    }
  }
  __syncthreads();
  // ---------------- phase C1: thread per (pair, value): the 3L consistent values (+ S*K singleton values)
  for (int it = threadIdx.x; it < tn * vals_per_pair; it += THREADS) {
    const int j = it / vals_per_pair, e = it % vals_per_pair;
    const long long pair = s_pair[j];
    double v;
    if (e < 3 * L) {
      const double *a = agg + ((size_t)j * UL + gt.idxL[e]) * 3;
      v = abf_from_sums(a[0], a[1], a[2], gt.omaL[e]);
      if (fa.out_gen) fa.out_gen[pair * 3 * L + e] = v;
    } else {
      const int e2 = e - 3 * L, c = e2 / K, k = e2 % K;
      const double *stj = st + (size_t)j * 3 * S;
      v = 0.0;
      if ((hasm[j] >> c) & 1ull) {
        double d, bd, sg;
        term_entry(stj[c], stj[S + c], stj[2 * S + c], prm.phi2S[k], d, bd, sg);
        v = abf_from_sums(d, bd, sg, prm.oma2S[k]);
      }
      if (fa.out_cfg) fa.out_cfg[(pair * C + c) * K + k] = v;
    }
    vals[(size_t)j * vals_per_pair + e] = v;
  }
  __syncthreads();
  // ---------------- phase C2: log10_weighted_sum of each staged row (utils_math.cpp:100-131)
  for (int it = threadIdx.x; it < tn * nrow_small; it += THREADS) {
    const int j = it / nrow_small, r = it % nrow_small;
    const int nk = (r < 3) ? L : K;
    const double *v = vals + (size_t)j * vals_per_pair + ((r < 3) ? r * L : 3 * L + (r - 3) * K);
    Lse a;
    a.init();
    for (int k = 0; k < nk; ++k) a.add(v[k], 1.0 / (double)nk, k == 0);
    const double w = (nk > 0) ? a.result() : nan("");
    wrow[(size_t)j * (3 + S) + r] = w;
    double *o = fa.out_w + s_pair[j] * (5 + C);
    if (r < 3) o[r] = w;
    else o[5 + (r - 3)] = w;
  }
  if (fa.which == 3) {
    // all configurations: thread per (pair, configuration), values from the shared table
    for (long long it = threadIdx.x; it < (long long)tn * C; it += THREADS) {
      const int j = (int)(it / C);
      const long long c = it % C;
      const long long pair = s_pair[j];
      const unsigned long long mask = prm.cfg_mask[c] & hasm[j];
      Lse a;
      a.init();
      for (int k = 0; k < K; ++k) {
        const double *tk = tab + ((size_t)j * K + k) * S * 3;
        double den = 0.0, num = 0.0, sing = 0.0;
        unsigned long long mm = mask;
        while (mm) {
          const int s = __ffsll((long long)mm) - 1;
          mm &= mm - 1;
          den += tk[3 * s];
          num += tk[3 * s + 1];
          sing += tk[3 * s + 2];
        }
        const double v = abf_from_sums(den, num, sing, prm.oma2S[k]);
        if (fa.out_cfg) fa.out_cfg[(pair * C + c) * K + k] = v;
        a.add(v, 1.0 / (double)K, k == 0);
      }
      fa.out_w[pair * (5 + C) + 5 + c] = a.result();
    }
  }
  __syncthreads();
  // ---------------- phase C3: BMAlite / BMA (gene_snp_pair.cpp:552-602)
  if (fa.which == 1) {
    for (int j = threadIdx.x; j < tn; j += THREADS) {
      double *o = fa.out_w + s_pair[j] * (5 + C);
      o[3] = nan("");
      o[4] = nan("");
    }
    return;
  }
  for (int j = warp; j < tn; j += WARPS) {
    const long long pair = s_pair[j];
    const double *wcfg = fa.out_w + pair * (5 + C) + 5;
    Lse lite, bma;
    lite.init();
    bma.init();
    for (long long c = lane; c < C; c += 32) {
      const double wc = (fa.which == 2) ? wrow[(size_t)j * (3 + S) + 3 + c] : wcfg[c];
      if (c < S) lite.add(wc, 0.5 / (double)S, c == 0);
      if (fa.which == 3) bma.add(wc, prm.cfg_weight[c], c == 0);
    }
    lite = warp_merge(lite);
    lite.add(wrow[(size_t)j * (3 + S) + 0], 0.5, false);
    const double w_gensin = lite.result();
    double w_all = nan("");
    if (fa.which == 3) {
      bma = warp_merge(bma);
      w_all = bma.result();
    }
    if (lane == 0) {
      double *o = fa.out_w + pair * (5 + C);
      o[3] = w_gensin;
      o[4] = w_all;
    }
  }
}

// shared-memory bytes of fast_pair_kernel for a tile of T pairs
__host__ __device__ inline size_t fast_smem_bytes(int T, int S, int L, int K, int UL, int which)
{
  size_t d = (size_t)T * S + (size_t)T * 3 * S + (size_t)T * UL * 3 +
             (size_t)T * (3 * L + (which == 2 ? S * K : 0)) + (size_t)T * (3 + S);
  if (which == 3) d += (size_t)T * K * S * 3;
  return d * 8 + (size_t)T * (8 + 8 + 8 + 4) + 16;
}

} // namespace eqb
