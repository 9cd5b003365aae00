// mvlr_kernel.cuh -- multivariate (MVLR, --error mvlr) joint Bayes factors on the device.
//
// X. Wen's MVLR class (reference MVLR.cpp) rebuilds an n x n projection T and dense n-sized products
// for every configuration of every pair.  Everything it computes is a function of the sufficient
// statistics  YtY = Y'TY (S x S, per gene[, permutation]),  k = g'Tg  and  b0 = Y'Tg  (per pair)
// (SURVEY.md App. A.6), where T = I - QQ' projects out [1, covariates] on the common individuals:
//   Sigma0 = YtY / n                                                    MVLR.cpp:176-195
//   Sigma(gamma) = fiterr * [ (H m + E'E)/(m+n) ] diag(sqrt(factor)) + (1 - fiterr) Sigma0   :201-298
//       E'E_ij = YtY_ij - (g_i + g_j - g_i g_j) b0_i b0_j / k,  factor_i = F / chisq_Qinv(fdist_Q(F))  :323-437
//   b = Sigma^-1 b0,  V^-1 = k Sigma^-1,  Gamma = u u', u_i = gamma_i sqrt(Sigma_ii)               :473-554
//   log10 ABF = [ b'W (I + V^-1 W)^-1 b / 2 - ln det(I + V^-1 W) / 2 ] / ln 10                      :559-605
// Only the rows/columns of the active subgroups of W are non-zero, so the determinant and the
// quadratic form are evaluated on the |gamma| x |gamma| active block (same value up to rounding).
// One CTA = one (gene, permutation); warps take SNPs; lanes take configurations.
#pragma once

#include "pair_kernel.cuh"

namespace eqb {

constexpr int MV_MAXS = 16; // subgroups supported by the per-thread dense algebra

// in-place LU with partial pivoting (row interchanges) of an n x n row-major matrix, stride MV_MAXS
__device__ inline void mv_lu(double *A, int n, int *piv)
{
  for (int i = 0; i < n; ++i) piv[i] = i;
  for (int j = 0; j + 1 < n; ++j) {
    double mx = fabs(A[j * MV_MAXS + j]);
    int ip = j;
    for (int i = j + 1; i < n; ++i) {
      const double a = fabs(A[i * MV_MAXS + j]);
      if (a > mx) {
        mx = a;
        ip = i;
      }
    }
    if (ip != j) {
      for (int c = 0; c < n; ++c) {
        const double t = A[j * MV_MAXS + c];
        A[j * MV_MAXS + c] = A[ip * MV_MAXS + c];
        A[ip * MV_MAXS + c] = t;
      }
      const int t = piv[j];
      piv[j] = piv[ip];
      piv[ip] = t;
    }
    const double ajj = A[j * MV_MAXS + j];
    if (ajj != 0.0)
      for (int i = j + 1; i < n; ++i) {
        const double f = A[i * MV_MAXS + j] / ajj;
        A[i * MV_MAXS + j] = f;
        for (int c = j + 1; c < n; ++c) A[i * MV_MAXS + c] -= f * A[j * MV_MAXS + c];
      }
  }
}

// x = A^-1 rhs given the LU factors (rhs indexed in the original row order)
__device__ inline void mv_lu_solve(const double *LU, const int *piv, int n, const double *rhs, double *x)
{
  for (int i = 0; i < n; ++i) {
    double acc = rhs[piv[i]];
    for (int c = 0; c < i; ++c) acc -= LU[i * MV_MAXS + c] * x[c];
    x[i] = acc;
  }
  for (int i = n - 1; i >= 0; --i) {
    double acc = x[i];
    for (int c = i + 1; c < n; ++c) acc -= LU[i * MV_MAXS + c] * x[c];
    x[i] = acc / LU[i * MV_MAXS + i];
  }
}

__device__ inline void mv_inverse(double *A /* destroyed */, int n, double *inv)
{
  int piv[MV_MAXS];
  mv_lu(A, n, piv);
  double e[MV_MAXS], x[MV_MAXS];
  for (int c = 0; c < n; ++c) {
    for (int i = 0; i < n; ++i) e[i] = (i == c) ? 1.0 : 0.0;
    mv_lu_solve(A, piv, n, e, x);
    for (int i = 0; i < n; ++i) inv[i * MV_MAXS + c] = x[i];
  }
}

struct MvGene {       // per (gene, permutation), shared memory
  double YtY[MV_MAXS * MV_MAXS];
  double Sig0[MV_MAXS * MV_MAXS];
  double Sig0inv[MV_MAXS * MV_MAXS];
  int n, q;           // common individuals, 1 + covariates
};

// ABFs of one configuration over a grid; writes raw values (optional) and returns the weighted ABF
static __device__ __noinline__ double mvlr_config(const MvGene &G, int S, unsigned long long gamma, double k, const double *b0,
                                           double alpha, const double *phi2, const double *oma2, int nk, int variant,
                                           double *raw_out)
{
  const int n = G.n, q = G.q, m = q + S + 1;
  double Sig[MV_MAXS * MV_MAXS], Sinv[MV_MAXS * MV_MAXS], tmp[MV_MAXS * MV_MAXS];
  if (alpha < 1e-6) {
    for (int i = 0; i < S; ++i)
      for (int j = 0; j < S; ++j) {
        Sig[i * MV_MAXS + j] = G.Sig0[i * MV_MAXS + j];
        Sinv[i * MV_MAXS + j] = G.Sig0inv[i * MV_MAXS + j];
      }
  } else {
    double fac[MV_MAXS];
    const int size = q + 1;
    for (int i = 0; i < S; ++i) {
      fac[i] = 1.0;
      if ((gamma >> i) & 1ull) {
        // residual of y_i on [covariates, g]; generalized inverse drops a null genotype direction
        const bool drop = !(k > 1e-8);
        const double beta = drop ? 0.0 : b0[i] / k;
        const double ee = G.YtY[i * MV_MAXS + i] - (drop ? 0.0 : b0[i] * b0[i] / k);
        const double sigma1 = ee / (double)(n - size);
        const double T2 = beta * k * beta / (sigma1 * sigma1); // sigma1 squared: MVLR.cpp:397
        const double v1 = 1.0, v2 = (double)(n - size);
        const double F = (v2 - v1 + 1.0) * T2 / (v1 * v2);
        double factor = 1.0;
        if (!(F < 1e-8)) {
          const double qv = fdist_Q(F, v1, v2 - v1 + 1.0);
          const double newF = chisq_Qinv_1df(qv) / v1;
          factor = F / newF;
        }
        fac[i] = sqrt(factor);
      }
    }
    const bool drop = !(k > 1e-8);
    for (int i = 0; i < S; ++i)
      for (int j = 0; j < S; ++j) {
        const int gi = (int)((gamma >> i) & 1ull), gj = (int)((gamma >> j) & 1ull);
        const double cross = drop ? 0.0 : (double)(gi + gj - gi * gj) * b0[i] * b0[j] / k;
        double v = (G.YtY[i * MV_MAXS + j] - cross) * (double(1.0) / double(m + n));
        if (i == j) v += 1e-4 * double(m) / double(m + n);
        v *= fac[j]; // Sigma * diag(sqrt(factor)): right multiplication only (MVLR.cpp:268-273)
        Sig[i * MV_MAXS + j] = alpha * v + (1.0 - alpha) * G.Sig0[i * MV_MAXS + j];
        tmp[i * MV_MAXS + j] = Sig[i * MV_MAXS + j];
      }
    mv_inverse(tmp, S, Sinv);
  }
  // b = Sigma^-1 b0; active subgroups
  double b[MV_MAXS], u[MV_MAXS];
  int idx[MV_MAXS], na = 0;
  for (int i = 0; i < S; ++i) {
    double acc = 0.0;
    for (int j = 0; j < S; ++j) acc += Sinv[i * MV_MAXS + j] * b0[j];
    b[i] = acc;
    if ((gamma >> i) & 1ull) {
      idx[na] = i;
      u[na] = sqrt(Sig[i * MV_MAXS + i]);
      ++na;
    }
  }
  Lse acc;
  acc.init();
  for (int g = 0; g < nk; ++g) {
    const double ph = phi2[g], om = oma2[g];
    const double p2 = (variant == 0) ? ph : ((variant == 1) ? 0.0 : ph + om);
    const double o2 = (variant == 0) ? om : ((variant == 1) ? ph + om : 0.0);
    // W_aa and A = I + k Sinv_aa W_aa
    double W[MV_MAXS * MV_MAXS], A[MV_MAXS * MV_MAXS];
    for (int a = 0; a < na; ++a)
      for (int c = 0; c < na; ++c) W[a * MV_MAXS + c] = (a == c) ? (o2 + p2) * u[a] * u[a] : o2 * u[a] * u[c];
    for (int a = 0; a < na; ++a)
      for (int c = 0; c < na; ++c) {
        double s = 0.0;
        for (int e = 0; e < na; ++e) s += k * Sinv[idx[a] * MV_MAXS + idx[e]] * W[e * MV_MAXS + c];
        A[a * MV_MAXS + c] = s + ((a == c) ? 1.0 : 0.0);
      }
    int piv[MV_MAXS];
    mv_lu(A, na, piv);
    double lndet = 0.0;
    for (int a = 0; a < na; ++a) lndet += log(fabs(A[a * MV_MAXS + a]));
    // quad = b_a' W A^-1 b_a : solve A x = b_a, then b_a' W x
    double ba[MV_MAXS], x[MV_MAXS];
    for (int a = 0; a < na; ++a) ba[a] = b[idx[a]];
    mv_lu_solve(A, piv, na, ba, x);
    double quad = 0.0;
    for (int a = 0; a < na; ++a) {
      double s = 0.0;
      for (int c = 0; c < na; ++c) s += W[a * MV_MAXS + c] * x[c];
      quad += ba[a] * s;
    }
    const double v = (0.5 * quad - 0.5 * lndet) / LN10;
    if (raw_out) raw_out[g] = v;
    acc.add(v, 1.0 / (double)nk, g == 0);
  }
  return (nk > 0) ? acc.result() : nan("");
}

template <int NPL>
__global__ void __launch_bounds__(THREADS) mvlr_kernel(const DevParams *__restrict__ prm_, const LaunchArgs la)
{
  const DevParams &prm = *prm_;
  extern __shared__ double dyn_smem[];
  __shared__ MvGene MG;
  __shared__ unsigned int s_colvalid;
  __shared__ int s_bad;
  __shared__ double w_part[WARPS][2];
  __shared__ int w_flag[WARPS][3];
  const int S = prm.S, N = prm.N, ldn = prm.ldn, Qmax = prm.Qmax, L = prm.L, K = prm.K;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ppg = la.perms_per_gene > 0 ? la.perms_per_gene : 1;
  const int gi = blockIdx.x / ppg;
  const long long p = la.perms_per_gene > 0 ? la.p0 + (blockIdx.x % ppg) : -1;
  const int g = la.genes[gi];
  const long long mbeg = prm.cis_begin[g], mend = prm.cis_end[g];
  const unsigned short *perm = (p >= 0) ? la.perm_tab + ((size_t)la.gene_slot[gi] * la.P_total + p) * N : nullptr;
  // dynamic shared memory: basis [Qmax+1][ldn], ytil [S][ldn], per-warp b0 [WARPS][MV_MAXS]
  double *q = dyn_smem;
  double *yt = q + (size_t)(Qmax + 1) * ldn;
  double *wb0 = yt + (size_t)S * ldn;
  const SubDev &sb0 = prm.sub[0];
  const int Q = sb0.Q;

  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  // ---- phase 1: common mask (from subgroup 0), basis, residual phenotypes
  for (int s = warp; s < S; s += WARPS) {
    const SubDev &sb = prm.sub[s];
    const double *Yg = sb.Yall + (size_t)g * ldn;
    const double *Y0 = sb0.Yall + (size_t)g * ldn;
    int bad = 0;
    for (int i = lane; i < ldn; i += 32) {
      double yv = 0.0;
      bool keep = false, keep0 = false;
      if (i < N) {
        const int j = perm ? (int)perm[i] : i;
        yv = Yg[j];
        keep = sb.gmask[i] && !isnan(yv);
        keep0 = sb0.gmask[i] && !isnan(Y0[j]);
      }
      if (keep != keep0) bad = 1; // MVLR needs the same individuals in every subgroup
      yt[(size_t)s * ldn + i] = keep0 ? yv : 0.0;
      if (s == 0) q[i] = keep0 ? 1.0 : 0.0;
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicExch(&s_bad, 1);
  }
  __syncthreads();
  if (warp == 0) {
    int n = 0;
    for (int i = lane; i < ldn; i += 32) n += (q[i] != 0.0) ? 1 : 0;
    n = warp_sum_int(n);
    const double inv_sqrt_n = n > 0 ? 1.0 / sqrt((double)n) : 0.0;
    for (int i = lane; i < ldn; i += 32) q[i] = (q[i] != 0.0) ? inv_sqrt_n : 0.0;
    __syncwarp();
    unsigned int colvalid = 1u;
    for (int k = 1; k <= Q; ++k) {
      double *qk = q + (size_t)k * ldn;
      const double *Ck = sb0.Call + (size_t)(k - 1) * ldn;
      double nrm0 = 0.0;
      int missing = 0;
      for (int i = lane; i < ldn; i += 32) {
        const bool keep = q[i] != 0.0;
        const double v = keep ? Ck[i] : 0.0;
        if (keep && !sb0.cmask[i]) missing = 1;
        qk[i] = v;
        nrm0 += v * v;
      }
      nrm0 = warp_sum(nrm0);
      if (__any_sync(0xffffffffu, missing) && lane == 0) atomicExch(la.err_flag, 1);
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
      } else
        for (int i = lane; i < ldn; i += 32) qk[i] = 0.0;
      __syncwarp();
    }
    if (lane == 0) {
      MG.n = n;
      MG.q = Q + 1;
      s_colvalid = colvalid;
    }
  }
  __syncthreads();
  if (s_bad) {
    if (threadIdx.x == 0) atomicExch(la.err_flag + 3, 1);
    return;
  }
  for (int s = warp; s < S; s += WARPS) { // residual phenotypes T y_s
    double *ys = yt + (size_t)s * ldn;
    for (int pass = 0; pass < 2; ++pass)
      for (int j = 0; j <= Q; ++j) {
        if (!((s_colvalid >> j) & 1u)) continue;
        const double *qj = q + (size_t)j * ldn;
        double h = 0.0;
        for (int i = lane; i < ldn; i += 32) h += qj[i] * ys[i];
        h = warp_sum(h);
        for (int i = lane; i < ldn; i += 32) ys[i] -= h * qj[i];
        __syncwarp();
      }
  }
  __syncthreads();
  for (int e = warp; e < S * S; e += WARPS) { // YtY
    const int i = e / S, j = e % S;
    double acc = 0.0;
    for (int t = lane; t < ldn; t += 32) acc += yt[(size_t)i * ldn + t] * yt[(size_t)j * ldn + t];
    acc = warp_sum(acc);
    if (lane == 0) MG.YtY[i * MV_MAXS + j] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int n = MG.n, qq = MG.q, m = qq + S + 1;
    double tmp[MV_MAXS * MV_MAXS];
    for (int i = 0; i < S; ++i)
      for (int j = 0; j < S; ++j) {
        MG.Sig0[i * MV_MAXS + j] = MG.YtY[i * MV_MAXS + j] * (1.0 / (double)(n + m - qq - S - 1));
        tmp[i * MV_MAXS + j] = MG.Sig0[i * MV_MAXS + j];
      }
    mv_inverse(tmp, S, MG.Sig0inv);
  }
  __syncthreads();

  // ---- phases 2-3: per SNP
  const long long C = (la.which == 1) ? 0 : ((la.which == 2) ? S : prm.C);
  Lse acc_stat;
  acc_stat.init();
  double max_stat = -INFINITY;
  bool first_nan = false;
  int cnt_nonnan = 0;
  double *b0 = wb0 + warp * MV_MAXS;
  for (long long m = mbeg + warp; m < mend; m += WARPS) {
    const bool is_first = (m == mbeg);
    const long long pair = la.want_outputs ? la.pair_off[gi] + (m - mbeg) : 0;
    bool all_geno = true;
    for (int s = 0; s < S; ++s) all_geno = all_geno && prm.sub[s].snp_has[m];
    double w_gen[3] = {nan(""), nan(""), nan("")}, w_gensin = nan(""), w_all = nan("");
    double stat_v = 0.0; // skipped pairs leave 0.0 in the permutation vector (gene.cpp:643,663-664)
    if (all_geno && MG.n > 0) { // gene.cpp:315-321
      // residual genotype, k = g'Tg, b0 = Y'Tg
      const double *Xm = sb0.X + (size_t)m * ldn;
      double xr[NPL];
#pragma unroll
      for (int j = 0; j < NPL; ++j) {
        const int i = lane + 32 * j;
        xr[j] = (i < ldn && q[i] != 0.0) ? Xm[i] : 0.0;
      }
      for (int pass = 0; pass < 2; ++pass)
        for (int k = 0; k <= Q; ++k) {
          if (!((s_colvalid >> k) & 1u)) continue;
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
      double kk = 0.0;
#pragma unroll
      for (int j = 0; j < NPL; ++j) kk += xr[j] * xr[j];
      kk = warp_sum(kk);
      for (int s = 0; s < S; ++s) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
          const int i = lane + 32 * j;
          if (i < ldn) acc += xr[j] * yt[(size_t)s * ldn + i];
        }
        acc = warp_sum(acc);
        if (lane == 0) b0[s] = acc;
      }
      __syncwarp();
      const unsigned long long ones = (S >= 64) ? ~0ull : ((1ull << S) - 1ull);
      // consistent configuration: gen, gen-fix, gen-maxh (gene_snp_pair.cpp:624-656)
      const int nvar = (p >= 0) ? 1 : 3;
      if (lane < nvar) {
        double *raw = (la.want_outputs && la.out_gen) ? la.out_gen + (pair * 3 + lane) * L : nullptr;
        w_gen[0] = mvlr_config(MG, S, ones, kk, b0, prm.fiterr, prm.phi2L, prm.oma2L, L, lane, raw);
      }
      w_gen[1] = __shfl_sync(0xffffffffu, w_gen[0], 1);
      w_gen[2] = __shfl_sync(0xffffffffu, w_gen[0], 2);
      w_gen[0] = __shfl_sync(0xffffffffu, w_gen[0], 0);
      if (p >= 0) w_gen[1] = w_gen[2] = nan("");
      if (la.which >= 2) {
        Lse lite, bma;
        lite.init();
        bma.init();
        for (long long c = lane; c < C; c += 32) {
          const unsigned long long cm = (la.which == 2) ? (1ull << c) : prm.cfg_mask[c];
          double *raw = (la.want_outputs && la.out_cfg) ? la.out_cfg + (pair * C + c) * K : nullptr;
          const double wc = mvlr_config(MG, S, cm, kk, b0, prm.fiterr, prm.phi2S, prm.oma2S, K, 0, raw);
          if (la.want_outputs && la.out_w) la.out_w[pair * (5 + C) + 5 + c] = wc;
          if (c < S) lite.add(wc, 0.5 / (double)S, c == 0);
          if (la.which == 3) bma.add(wc, prm.cfg_weight[c], c == 0);
        }
        lite = warp_merge(lite);
        lite.add(w_gen[0], 0.5, false);
        w_gensin = lite.result();
        if (la.which == 3) {
          bma = warp_merge(bma);
          w_all = bma.result();
        }
      }
      stat_v = (la.which == 1) ? w_gen[0] : ((la.which == 2) ? w_gensin : w_all);
    } else if (la.want_outputs) {
      // pair skipped by the reference: no ABF at all (NaN rows)
      for (int e = lane; e < 3 * L; e += 32)
        if (la.out_gen) la.out_gen[pair * 3 * L + e] = nan("");
      for (long long e = lane; e < C * K; e += 32)
        if (la.out_cfg) la.out_cfg[pair * C * K + e] = nan("");
      for (long long e = lane; e < C; e += 32)
        if (la.out_w) la.out_w[pair * (5 + C) + 5 + e] = nan("");
    }
    if (la.want_outputs && lane == 0) {
      if (la.out_w) {
        double *o = la.out_w + pair * (5 + C);
        o[0] = w_gen[0];
        o[1] = w_gen[1];
        o[2] = w_gen[2];
        o[3] = w_gensin;
        o[4] = w_all;
      }
      for (int s = 0; s < S; ++s) {
        if (la.out_n) la.out_n[pair * S + s] = (all_geno && MG.n > 0) ? MG.n : 0;
        if (la.out_ss)
          for (int e = 0; e < 5; ++e) la.out_ss[(pair * S + s) * 5 + e] = nan("");
      }
    }
    if (la.stat_kind == STAT_JOIN_MAX || la.stat_kind == STAT_JOIN_AVG) {
      const double v = (la.true_rules && !(all_geno && MG.n > 0)) ? nan("") : stat_v;
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
  if (la.stat_kind == STAT_NONE) return;
  double *out = (p >= 0) ? la.out_stat + (size_t)gi * la.P_total + p : la.out_stat + (size_t)gi;
  if (lane == 0) {
    w_part[warp][0] = (la.stat_kind == STAT_JOIN_MAX) ? max_stat : acc_stat.m;
    w_part[warp][1] = acc_stat.acc;
    w_flag[warp][0] = first_nan ? 1 : 0;
    w_flag[warp][1] = acc_stat.any ? 1 : 0;
    w_flag[warp][2] = cnt_nonnan;
  }
  __syncthreads();
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
      res = (fn && !la.true_rules) ? nan("") : v;
    } else {
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
      const double size = la.true_rules ? (double)nn : (double)Mg;
      if ((fn && !la.true_rules) || nn == 0)
        res = nan("");
      else {
        res = t.m + log10(t.acc * (1.0 / size));
        if (fabs(res) <= DBL_EPSILON) res = 0.0;
      }
    }
    out[0] = res;
  }
}

__host__ __device__ inline size_t mvlr_smem_doubles(int S, int Qmax, int ldn)
{
  return (size_t)(Qmax + 1) * ldn + (size_t)S * ldn + (size_t)WARPS * MV_MAXS;
}

} // namespace eqb
