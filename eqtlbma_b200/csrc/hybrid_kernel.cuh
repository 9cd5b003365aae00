// hybrid_kernel.cuh -- --error hybrid on the device (gene_snp_pair.cpp:760-1423).
//
// The hybrid model keeps every subgroup's own individuals and lets the errors of two subgroups be
// correlated through the individuals they share.  Per pair (gene, SNP) it needs
//   diagonals  (CalcBetahatsAndDiagsPerSubgroup :760-879), subgroup s on its own kept rows (permuted /
//              quantile-normalised exactly like the uvlr statistics: FillStlContainers :79-169):
//                betahat_s = x~'y~ / x~'x~,  rss_full = y~'y~ - (x~'y~)^2 / x~'x~,  rss_null = y~'y~
//                Sigma_ss = (f rss_full + (1 - f) rss_null) / n_s,   Vg_ss = Sigma_ss / x~'x~
//              with x~, y~ the residuals on [1, covariates] (Frisch-Waugh form of the reference's SVD solve; the
//              [1][1] entry of V D^-2 V' = (X'X)^-1 is 1 / x~'x~),
//   off-diagonals (CalcOffDiagCovarsFromPairsOfSubgroups :1059-1132), subgroups s1 < s2 on the individuals present in
//              both -- NEVER permuted and never quantile-normalised, genotype and covariates of s1 for all three sets
//              (FillGslStructuresForPairOfSubgroup :881-983; samples.cpp:148-192):
//                Sigma_12 = f [y1'(I - H)y2 / n12] + (1 - f) [y1'(I - Hc)y2 / n12]            (CalcMleErrorCovariance)
//                Vg_12 = Sigma_12 * [ (G12 + Gu1)^-1 G12 (G12 + Gu2)^-1 ][1][1]               (GetMatricesA)
//              all of it a function of the Gram matrices of z = [1, g, covariates] over the three sets (both, only s1,
//              only s2) and of z'y1, z'y2, y1'y2 over the common set,
//   the model's own ABF (CalcLog10AbfMvlr :1165-1255):
//                W = D (gamma gamma' o [phi2 I + oma2 11']) D,  D = diag(sqrt(Sigma_ss))
//                log10 ABF = [ b'Vg^-1 W (I + Vg^-1 W)^-1 Vg^-1 b / 2 - ln det(I + Vg^-1 W) / 2 ] / ln 10.
// Two kernels per launch:
//   hybrid_offdiag_kernel  grid = genes x SNP slices, warp per SNP: the off-diagonal blocks do not depend on the permutation,
//                          so they are computed once per (gene, SNP) -- the warp accumulates the dot products, lane 0 does
//                          the (2 + Q)-sized algebra -- into a cache [gene][largest window][S (S - 1) / 2];
//   hybrid_kernel<NPL>     CTA = (gene, permutation) (x SNP slices in output-only launches), warp per SNP: phase 1 = the
//                          uvlr setup, per SNP the diagonals, Vg assembled and inverted once, then lanes take
//                          (configuration, grid point) items with the ABF evaluated on the active block.
// The reference's pseudo-inverses are plain inverses here: a rank-deficient Gram matrix (monomorphic genotype on a set)
// is reported through the degenerate-design flag and the pair's Bayes factors are NaN (same documented tie as the uvlr
// path).
#pragma once

#include "mvlr_kernel.cuh"

namespace eqb {

constexpr int HY_MAXQ2 = 8; // 2 + covariates of the off-diagonal designs
constexpr int HY_NT = HY_MAXQ2 * (HY_MAXQ2 + 1) / 2;

// per-warp state of one pair, shared memory
struct HyPair {
  double b[MV_MAXS];    // betahat
  double sd[MV_MAXS];   // sqrt(Sigma_ss)
  double bVg[MV_MAXS];  // b' Vg^-1
  double Vg[MV_MAXS * MV_MAXS];
  double Vinv[MV_MAXS * MV_MAXS];
  double scr[64];       // raw values of the (configuration, grid point) items of one pass
};

// inverse of an n x n matrix (row-major, stride HY_MAXQ2) by Gauss-Jordan with partial pivoting
static __device__ __noinline__ bool hy_inverse(const double *A, int n, double *inv)
{
  double T[HY_MAXQ2 * HY_MAXQ2];
  double scale = 0.0;
  for (int i = 0; i < n; ++i) {
    scale = fmax(scale, fabs(A[i * HY_MAXQ2 + i]));
    for (int j = 0; j < n; ++j) {
      T[i * HY_MAXQ2 + j] = A[i * HY_MAXQ2 + j];
      inv[i * HY_MAXQ2 + j] = (i == j) ? 1.0 : 0.0;
    }
  }
  for (int j = 0; j < n; ++j) {
    int ip = j;
    double mx = fabs(T[j * HY_MAXQ2 + j]);
    for (int i = j + 1; i < n; ++i) {
      const double a = fabs(T[i * HY_MAXQ2 + j]);
      if (a > mx) {
        mx = a;
        ip = i;
      }
    }
    if (!(mx > 1e-13 * scale)) return false;
    if (ip != j)
      for (int c = 0; c < n; ++c) {
        double t = T[j * HY_MAXQ2 + c];
        T[j * HY_MAXQ2 + c] = T[ip * HY_MAXQ2 + c];
        T[ip * HY_MAXQ2 + c] = t;
        t = inv[j * HY_MAXQ2 + c];
        inv[j * HY_MAXQ2 + c] = inv[ip * HY_MAXQ2 + c];
        inv[ip * HY_MAXQ2 + c] = t;
      }
    const double d = 1.0 / T[j * HY_MAXQ2 + j];
    for (int c = 0; c < n; ++c) {
      T[j * HY_MAXQ2 + c] *= d;
      inv[j * HY_MAXQ2 + c] *= d;
    }
    for (int i = 0; i < n; ++i) {
      if (i == j) continue;
      const double f = T[i * HY_MAXQ2 + j];
      if (f == 0.0) continue;
      for (int c = 0; c < n; ++c) {
        T[i * HY_MAXQ2 + c] -= f * T[j * HY_MAXQ2 + c];
        inv[i * HY_MAXQ2 + c] -= f * inv[j * HY_MAXQ2 + c];
      }
    }
  }
  return true;
}

// a' M b for n-vectors and an n x n matrix of stride HY_MAXQ2
__device__ inline double hy_quad(const double *a, const double *M, const double *b, int n)
{
  double acc = 0.0;
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) s += M[i * HY_MAXQ2 + j] * b[j];
    acc += a[i] * s;
  }
  return acc;
}

// One (configuration, grid point) of CalcLog10AbfMvlr (gene_snp_pair.cpp:1165-1255).  W is non-zero on the active
// subgroups only, so with B = I + W_aa (Vg^-1)_aa on the active block
//   ln det(I + Vg^-1 W) = ln det B   (Sylvester),   b'Vg^-1 W (I + Vg^-1 W)^-1 Vg^-1 b = c_a' B^-1 W_aa c_a,  c = Vg^-1 b
// (same value up to rounding as the reference's S x S products).
static __device__ __noinline__ double hybrid_value(const HyPair &H, int S, unsigned long long gamma, double p2, double o2)
{
  int idx[MV_MAXS], na = 0;
  for (int i = 0; i < S; ++i)
    if ((gamma >> i) & 1ull) idx[na++] = i;
  if (na == 1) {
    const int i = idx[0];
    const double w = H.sd[i] * (p2 + o2) * H.sd[i], bb = 1.0 + w * H.Vinv[i * MV_MAXS + i], c = H.bVg[i];
    return (-0.5 * log(fabs(bb)) + 0.5 * c * w * c / bb) / LN10;
  }
  double W[MV_MAXS * MV_MAXS], B[MV_MAXS * MV_MAXS], rhs[MV_MAXS], x[MV_MAXS];
  for (int a = 0; a < na; ++a)
    for (int c = 0; c < na; ++c) W[a * MV_MAXS + c] = H.sd[idx[a]] * ((a == c) ? p2 + o2 : o2) * H.sd[idx[c]];
  for (int a = 0; a < na; ++a) {
    double r = 0.0;
    for (int c = 0; c < na; ++c) {
      double sacc = 0.0;
      for (int e = 0; e < na; ++e) sacc += W[a * MV_MAXS + e] * H.Vinv[idx[e] * MV_MAXS + idx[c]];
      B[a * MV_MAXS + c] = sacc + ((a == c) ? 1.0 : 0.0);
      r += W[a * MV_MAXS + c] * H.bVg[idx[c]];
    }
    rhs[a] = r;
  }
  int piv[MV_MAXS];
  mv_lu(B, na, piv);
  double lndet = 0.0;
  for (int a = 0; a < na; ++a) lndet += log(fabs(B[a * MV_MAXS + a]));
  mv_lu_solve(B, piv, na, rhs, x);
  double quad = 0.0;
  for (int a = 0; a < na; ++a) quad += H.bVg[idx[a]] * x[a];
  return (-0.5 * lndet + 0.5 * quad) / LN10;
}

// ABFs of one configuration over a grid (CalcLog10AbfMvlr, gene_snp_pair.cpp:1165-1255); writes the raw values
// (optional) and returns the grid-averaged ABF
static __device__ __noinline__ double hybrid_config(const HyPair &H, int S, unsigned long long gamma, const double *phi2,
                                                    const double *oma2, int nk, int variant, double *raw_out)
{
  Lse acc;
  acc.init();
  for (int g = 0; g < nk; ++g) {
    const double ph = phi2[g], om = oma2[g];
    const double p2 = (variant == 0) ? ph : ((variant == 1) ? 0.0 : ph + om);
    const double o2 = (variant == 0) ? om : ((variant == 1) ? ph + om : 0.0);
    double W[MV_MAXS * MV_MAXS], A[MV_MAXS * MV_MAXS];
    for (int i = 0; i < S; ++i)
      for (int j = 0; j < S; ++j) {
        const bool on = ((gamma >> i) & 1ull) && ((gamma >> j) & 1ull);
        W[i * MV_MAXS + j] = on ? H.sd[i] * ((i == j) ? p2 + o2 : o2) * H.sd[j] : 0.0;
      }
    for (int i = 0; i < S; ++i)
      for (int j = 0; j < S; ++j) {
        double s = 0.0;
        for (int e = 0; e < S; ++e) s += H.Vinv[i * MV_MAXS + e] * W[e * MV_MAXS + j];
        A[i * MV_MAXS + j] = s + ((i == j) ? 1.0 : 0.0);
      }
    int piv[MV_MAXS];
    mv_lu(A, S, piv);
    double lndet = 0.0;
    for (int a = 0; a < S; ++a) lndet += log(fabs(A[a * MV_MAXS + a]));
    double x[MV_MAXS];
    mv_lu_solve(A, piv, S, H.bVg, x); // (I + Vg^-1 W)^-1 (b'Vg^-1)'
    double quad = 0.0;
    for (int j = 0; j < S; ++j) {
      double s = 0.0;
      for (int i = 0; i < S; ++i) s += H.bVg[i] * W[i * MV_MAXS + j];
      quad += s * x[j];
    }
    const double v = (-0.5 * lndet + 0.5 * quad) / LN10;
    if (raw_out) raw_out[g] = v;
    acc.add(v, 1.0 / (double)nk, g == 0);
  }
  return (nk > 0) ? acc.result() : nan("");
}

// Vg_12 of one pair of subgroups for one (gene, SNP) (CalcOffDiagCovarsFromPairsOfSubgroups, gene_snp_pair.cpp:1059-1132):
// the warp accumulates the Gram matrices of z = [1, g, covariates of s1] over the individuals common to / unique to the
// two subgroups and z'y1, z'y2, y1'y2 over the common ones; lane 0 does the (2 + Q)-sized algebra.  NaN = degenerate.
static __device__ __noinline__ double hy_offdiag(const DevParams &prm, int g, long long m, int s1, int s2, double fit, int lane,
                                                 int *err_flag)
{
  const int N = prm.N, ldn = prm.ldn;
  bool degenerate = false;
  double vg12 = nan("");
  const SubDev &sa = prm.sub[s1], &sc = prm.sub[s2];
  const int Q = sa.Q, Q2 = Q + 2, NT = Q2 * (Q2 + 1) / 2;
  const double *Y1 = sa.Yall + (size_t)g * ldn, *Y2 = sc.Yall + (size_t)g * ldn;
  const double *Xm = sa.X + (size_t)m * ldn;
  double acc[3 * HY_NT + 2 * HY_MAXQ2 + 1];
  for (int e = 0; e < 3 * NT + 2 * Q2 + 1; ++e) acc[e] = 0.0;
  int n12 = 0, bad = 0;
  for (int i = lane; i < N; i += 32) {
    const double y1 = Y1[i], y2 = Y2[i];
    const bool p1 = sa.gmask[i] && !isnan(y1);
    const bool p2 = sc.gmask[i] && !isnan(y2);
    if (!p1 && !p2) continue;
    if (!sa.gmask[i] || (Q > 0 && !sa.cmask[i])) { // the reference reads past its vectors here
      bad = 1;
      continue;
    }
    double z[HY_MAXQ2];
    z[0] = 1.0;
    z[1] = Xm[i];
    for (int k = 0; k < Q; ++k) z[2 + k] = sa.Call[(size_t)k * ldn + i];
    double *G = acc + ((p1 && p2) ? 0 : (p1 ? 1 : 2)) * NT;
    int t = 0;
    for (int a = 0; a < Q2; ++a)
      for (int b = 0; b <= a; ++b) G[t++] += z[a] * z[b];
    if (p1 && p2) {
      double *h = acc + 3 * NT;
      for (int a = 0; a < Q2; ++a) {
        h[a] += z[a] * y1;
        h[Q2 + a] += z[a] * y2;
      }
      h[2 * Q2] += y1 * y2;
      ++n12;
    }
  }
  for (int e = 0; e < 3 * NT + 2 * Q2 + 1; ++e) acc[e] = warp_sum(acc[e]);
  n12 = warp_sum_int(n12);
  if (__any_sync(0xffffffffu, bad)) {
    if (lane == 0) atomicExch(err_flag + 4, 1);
    degenerate = true;
  }
  if (n12 == 0) { // "have no individuals in common": fatal in the reference (gene_snp_pair.cpp:897-901)
    if (lane == 0) atomicExch(err_flag + 5, 1);
    degenerate = true;
  }
  bool ok = true;
  if (lane == 0 && !degenerate) {
    double G12[HY_MAXQ2 * HY_MAXQ2], G1[HY_MAXQ2 * HY_MAXQ2], G2[HY_MAXQ2 * HY_MAXQ2];
    int t = 0;
    for (int a = 0; a < Q2; ++a)
      for (int b = 0; b <= a; ++b, ++t) {
        G12[a * HY_MAXQ2 + b] = G12[b * HY_MAXQ2 + a] = acc[t];
        G1[a * HY_MAXQ2 + b] = G1[b * HY_MAXQ2 + a] = acc[t] + acc[NT + t];
        G2[a * HY_MAXQ2 + b] = G2[b * HY_MAXQ2 + a] = acc[t] + acc[2 * NT + t];
      }
    const double *h1 = acc + 3 * NT, *h2 = h1 + Q2;
    const double y12 = acc[3 * NT + 2 * Q2];
    double I12[HY_MAXQ2 * HY_MAXQ2], I1[HY_MAXQ2 * HY_MAXQ2], I2[HY_MAXQ2 * HY_MAXQ2];
    // (three separate calls of the out-of-line routine: inlined into one short-circuit chain the second and third
    // inverses came back as the identity in the sm_100a build of nvcc 12.9)
    const bool ok12 = hy_inverse(G12, Q2, I12);
    const bool ok1 = hy_inverse(G1, Q2, I1);
    const bool ok2 = hy_inverse(G2, Q2, I2);
    ok = ok12 && ok1 && ok2;
    if (ok) {
      const double s_full = (y12 - hy_quad(h1, I12, h2, Q2)) / (double)n12;
      // null model: the same without the genotype column
      double Gc[HY_MAXQ2 * HY_MAXQ2], Ic[HY_MAXQ2 * HY_MAXQ2], c1[HY_MAXQ2], c2[HY_MAXQ2];
      for (int a = 0, ra = 0; a < Q2; ++a) {
        if (a == 1) continue;
        c1[ra] = h1[a];
        c2[ra] = h2[a];
        for (int b = 0, rb = 0; b < Q2; ++b) {
          if (b == 1) continue;
          Gc[ra * HY_MAXQ2 + rb] = G12[a * HY_MAXQ2 + b];
          ++rb;
        }
        ++ra;
      }
      ok = hy_inverse(Gc, Q2 - 1, Ic);
      if (ok) {
        const double s_null = (y12 - hy_quad(c1, Ic, c2, Q2 - 1)) / (double)n12;
        const double sig12 = fit * s_full + (1.0 - fit) * s_null;
        double cov11 = 0.0; // [(G12 + Gu1)^-1 G12 (G12 + Gu2)^-1][1][1]
        for (int a = 0; a < Q2; ++a) {
          double sacc = 0.0;
          for (int b = 0; b < Q2; ++b) sacc += G12[a * HY_MAXQ2 + b] * I2[b * HY_MAXQ2 + 1];
          cov11 += I1[1 * HY_MAXQ2 + a] * sacc;
        }
        vg12 = sig12 * cov11;
      }
    }
  }
  ok = __shfl_sync(0xffffffffu, (int)ok, 0) != 0;
  vg12 = __shfl_sync(0xffffffffu, vg12, 0);
  return (ok && !degenerate) ? vg12 : nan("");
}

// off-diagonal cache of one launch: [gene of the work list][SNP of its window][pair of subgroups]
__global__ void __launch_bounds__(THREADS) hybrid_offdiag_kernel(const DevParams *__restrict__ prm_, const LaunchArgs la)
{
  const DevParams &prm = *prm_;
  const int S = prm.S, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gi = blockIdx.x, g = la.genes[gi];
  const long long mbeg = prm.cis_begin[g], mend = prm.cis_end[g];
  const int npsub = S * (S - 1) / 2;
  for (long long m = mbeg + (long long)blockIdx.y * WARPS + warp; m < mend; m += (long long)WARPS * gridDim.y) {
    bool all_geno = true;
    for (int s = 0; s < S; ++s) all_geno = all_geno && prm.sub[s].snp_has[m];
    if (!all_geno) continue; // pair skipped by the reference (gene.cpp:315-321)
    double *off = la.hy_off + ((size_t)gi * la.hy_stride + (size_t)(m - mbeg)) * npsub;
    int t = 0;
    for (int s1 = 0; s1 + 1 < S; ++s1)
      for (int s2 = s1 + 1; s2 < S; ++s2, ++t) {
        const double v = hy_offdiag(prm, g, m, s1, s2, prm.fiterr, lane, la.err_flag);
        if (lane == 0) off[t] = v;
      }
  }
}

__host__ __device__ inline size_t hybrid_smem_bytes(int S, int Qmax, int ldn, int qnorm, bool basis_in_smem)
{
  return (basis_in_smem ? basis_doubles(S, Qmax, ldn, qnorm) * sizeof(double) : 0) + (size_t)WARPS * sizeof(HyPair);
}

template <int NPL>
__global__ void __launch_bounds__(THREADS) hybrid_kernel(const DevParams *__restrict__ prm_, const LaunchArgs la)
{
  const DevParams &prm = *prm_;
  extern __shared__ double dyn_smem[];
  __shared__ int s_n[MV_MAXS];
  __shared__ int s_rankz[MV_MAXS];
  __shared__ unsigned int s_colvalid[MV_MAXS];
  __shared__ double s_yy[MV_MAXS], s_tss[MV_MAXS];
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
  const double fit = prm.fiterr;

  const int brows = basis_rows(Qmax, prm.qnorm);
  const size_t nb = basis_doubles(S, Qmax, ldn, prm.qnorm);
  HyPair *hy_all = (HyPair *)dyn_smem;
  HyPair &H = hy_all[warp];
  double *basis = la.basis_ws ? la.basis_ws + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * nb : (double *)(hy_all + WARPS);

  // ------------------------------------------------------------------ phase 1: per-subgroup bases and residual
  // phenotypes on the (permuted) kept rows -- the uvlr setup (pair_kernel phase 1)
  for (int s = warp; s < S; s += WARPS) {
    const SubDev &sb = prm.sub[s];
    double *q = basis + (size_t)s * brows * ldn;
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
      }
      continue;
    }
    if (prm.qnorm) { // utils_math.cpp:80-96
      double *vals = q + (size_t)(Qmax + 2) * ldn;
      int *ord = (int *)(q + (size_t)(Qmax + 3) * ldn);
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
    for (int pass = 0; pass < 2; ++pass)
      for (int j = 0; j <= Q; ++j) {
        if (!((colvalid >> j) & 1u)) continue;
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
      s_n[s] = n;
      s_rankz[s] = rankz;
      s_colvalid[s] = colvalid;
      s_yy[s] = yy;
      s_tss[s] = tss;
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ phases 2-3: per SNP
  const long long C = (la.which == 1) ? 0 : ((la.which == 2) ? S : prm.C);
  Lse acc_stat;
  acc_stat.init();
  double max_stat = -INFINITY;
  bool first_nan = false;
  int cnt_nonnan = 0;
  // (gridDim.y > 1 only in output-only launches: the SNPs of a gene are then cut into slices so that a true pass over few
  // genes with long windows fills the device; a launch that reduces a statistic over the gene keeps one CTA per gene)
  for (long long m = mbeg + (long long)blockIdx.y * WARPS + warp; m < mend; m += (long long)WARPS * gridDim.y) {
    const bool is_first = (m == mbeg);
    const long long pair = la.want_outputs ? la.pair_off[gi] + (m - mbeg) : 0;
    bool all_geno = true;
    for (int s = 0; s < S; ++s) all_geno = all_geno && prm.sub[s].snp_has[m];
    double w_gen[3] = {nan(""), nan(""), nan("")}, w_gensin = nan(""), w_all = nan("");
    double stat_v = 0.0; // skipped pairs leave 0.0 in the permutation vector (gene.cpp:643,663-664)
    if (all_geno) { // gene.cpp:315-321
      bool degenerate = false;
      // -------- diagonals
      for (int s = 0; s < S; ++s) {
        const SubDev &sb = prm.sub[s];
        const int n = s_n[s];
        const int Q = sb.Q;
        double pve = nan(""), sigmahat = nan(""), betahat = nan(""), se = nan(""), pval = nan("");
        double sig_ss = nan(""), vg_ss = nan("");
        if (n >= (2 + Q) + 1) {
          const double *q = basis + (size_t)s * brows * ldn;
          const double *yt = q + (size_t)(Qmax + 1) * ldn;
          const double *Xm = sb.X + (size_t)m * ldn;
          const unsigned int colvalid = s_colvalid[s];
          double xr[NPL];
          double xraw2 = 0.0;
#pragma unroll
          for (int j = 0; j < NPL; ++j) {
            const int i = lane + 32 * j;
            double v = 0.0;
            if (i < ldn) v = (q[i] != 0.0) ? Xm[i] : 0.0;
            xr[j] = v;
            xraw2 += v * v;
          }
          xraw2 = warp_sum(xraw2);
          for (int pass = 0; pass < 2; ++pass)
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
          if (xx > 1e-24 * xraw2 && xraw2 > 0.0 && s_rankz[s] == Q + 1) {
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
            const int rank = Q + 2;
            pve = 1.0 - rss / s_tss[s];
            sigmahat = sqrt(rss / (double)(n - rank));
            se = sigmahat * sqrt(1.0 / xx);
            if (la.want_outputs) { // the p-value is only printed (no permutation statistic of the join analysis uses it)
              if (lane == 0) pval = 2.0 * tdist_Q(fabs(betahat / se), (double)(n - rank));
              pval = __shfl_sync(0xffffffffu, pval, 0);
            }
            sig_ss = fit * (rss / (double)n) + (1.0 - fit) * (s_yy[s] / (double)n);
            vg_ss = sig_ss * (1.0 / xx);
          } else
            degenerate = true;
        } else
          degenerate = true;
        if (lane == 0) {
          H.b[s] = betahat;
          H.sd[s] = sqrt(sig_ss);
          H.Vg[s * MV_MAXS + s] = vg_ss;
        }
        if (la.want_outputs && lane == 0) {
          if (la.out_n) la.out_n[pair * S + s] = n;
          if (la.out_ss) {
            double *o = la.out_ss + (pair * S + s) * 5;
            o[0] = pve;
            o[1] = sigmahat;
            o[2] = betahat;
            o[3] = se;
            o[4] = pval;
          }
        }
      }
      __syncwarp();
      // -------- off-diagonals: Vg_12 of every pair of subgroups from the per-launch cache (hybrid_offdiag_kernel; they
      // do not depend on the permutation)
      {
        const int npsub = S * (S - 1) / 2;
        const double *off = la.hy_off + ((size_t)gi * la.hy_stride + (size_t)(m - mbeg)) * npsub;
        int t = 0, bad = 0;
        for (int s1 = 0; s1 + 1 < S; ++s1)
          for (int s2 = s1 + 1; s2 < S; ++s2, ++t) {
            const double v = off[t];
            if (isnan(v)) bad = 1;
            if (lane == 0) H.Vg[s1 * MV_MAXS + s2] = H.Vg[s2 * MV_MAXS + s1] = v;
          }
        if (bad) degenerate = true;
      }
      __syncwarp();
      if (degenerate) {
        if (lane == 0) atomicOr(la.err_flag + 1, 1); // documented unsupported degenerate design
      } else if (lane == 0) {
        double tmp[MV_MAXS * MV_MAXS];
        for (int i = 0; i < S; ++i)
          for (int j = 0; j < S; ++j) tmp[i * MV_MAXS + j] = H.Vg[i * MV_MAXS + j];
        mv_inverse(tmp, S, H.Vinv);
        for (int j = 0; j < S; ++j) {
          double sacc = 0.0;
          for (int i = 0; i < S; ++i) sacc += H.b[i] * H.Vinv[i * MV_MAXS + j];
          H.bVg[j] = sacc;
        }
      }
      __syncwarp();
      const unsigned long long ones = (S >= 64) ? ~0ull : ((1ull << S) - 1ull);
      if (!degenerate) {
        // lanes take (configuration, grid point) items: floor(32 / nk) configurations per pass (one configuration over
        // two half-passes when 32 < nk <= 64); the owner lane of a configuration then averages its nk values in grid
        // order (utils::log10_weighted_sum).  Larger grids: one lane per configuration (hybrid_config).
        auto eval_set = [&](long long nconf, bool gen_set, const double *phi2, const double *oma2, int nk, double *raw,
                            auto &&own) {
          if (nk > 64 || nk < 1) {
            for (long long c = lane; c < nconf; c += 32) {
              const unsigned long long cm = gen_set ? ones : ((la.which == 2) ? (1ull << c) : prm.cfg_mask[c]);
              own(c, hybrid_config(H, S, cm, phi2, oma2, nk, gen_set ? (int)c : 0, raw ? raw + c * nk : nullptr));
            }
            return;
          }
          const int npack = (nk <= 32) ? 32 / nk : 1, span = (nk <= 32) ? nk : 32;
          for (long long c0 = 0; c0 < nconf; c0 += npack) {
            for (int half = 0; half * 32 < nk; ++half) {
              const int slot = lane / span, gp = (nk <= 32) ? lane % span : lane + 32 * half;
              const long long c = c0 + slot;
              if (slot < npack && c < nconf && gp < nk) {
                const unsigned long long cm = gen_set ? ones : ((la.which == 2) ? (1ull << c) : prm.cfg_mask[c]);
                const int variant = gen_set ? (int)c : 0;
                const double ph = phi2[gp], om = oma2[gp];
                const double p2 = (variant == 0) ? ph : ((variant == 1) ? 0.0 : ph + om);
                const double o2 = (variant == 0) ? om : ((variant == 1) ? ph + om : 0.0);
                const double v = hybrid_value(H, S, cm, p2, o2);
                if (raw) raw[c * nk + gp] = v;
                H.scr[slot * span + gp] = v;
              }
            }
            __syncwarp();
            if (lane < npack && c0 + lane < nconf) {
              Lse a;
              a.init();
              const double *vals = H.scr + lane * span;
              for (int g2 = 0; g2 < nk; ++g2) a.add(vals[g2], 1.0 / (double)nk, g2 == 0);
              own(c0 + lane, a.result());
            }
            __syncwarp();
          }
        };
        // consistent configuration: gen, gen-fix, gen-maxh (gene_snp_pair.cpp:1257-1308)
        const int nvar = (p >= 0) ? 1 : 3;
        double w_own[3] = {nan(""), nan(""), nan("")};
        unsigned int has = 0u; // variants whose average this lane owns
        eval_set(nvar, true, prm.phi2L, prm.oma2L, L, (la.want_outputs && la.out_gen) ? la.out_gen + pair * 3 * L : nullptr,
                 [&](long long c, double w) {
                   if (c == 0) w_own[0] = w;
                   else if (c == 1) w_own[1] = w;
                   else w_own[2] = w;
                   has |= 1u << (int)c;
                 });
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const unsigned int who = __ballot_sync(0xffffffffu, (has >> j) & 1u);
          const double v = __shfl_sync(0xffffffffu, w_own[j], who ? __ffs(who) - 1 : 0);
          w_gen[j] = who ? v : nan("");
        }
        if (la.which >= 2) { // singletons / every configuration on gridS (:1310-1382), BMAlite, BMA
          Lse lite, bma;
          lite.init();
          bma.init();
          eval_set(C, false, prm.phi2S, prm.oma2S, K, (la.want_outputs && la.out_cfg) ? la.out_cfg + pair * C * K : nullptr,
                   [&](long long c, double wc) {
                     if (la.want_outputs && la.out_w) la.out_w[pair * (5 + C) + 5 + c] = wc;
                     if (c < S) lite.add(wc, 0.5 / (double)S, c == 0);
                     if (la.which == 3) bma.add(wc, prm.cfg_weight[c], c == 0);
                   });
          lite = warp_merge(lite);
          lite.add(w_gen[0], 0.5, false);
          w_gensin = lite.result();
          if (la.which == 3) {
            bma = warp_merge(bma);
            w_all = bma.result();
          }
        }
        stat_v = (la.which == 1) ? w_gen[0] : ((la.which == 2) ? w_gensin : w_all);
      } else {
        stat_v = nan("");
        if (la.want_outputs) {
          for (int e = lane; e < 3 * L; e += 32)
            if (la.out_gen) la.out_gen[pair * 3 * L + e] = nan("");
          for (long long e = lane; e < C * K; e += 32)
            if (la.out_cfg) la.out_cfg[pair * C * K + e] = nan("");
          for (long long e = lane; e < C; e += 32)
            if (la.out_w) la.out_w[pair * (5 + C) + 5 + e] = nan("");
        }
      }
    } else if (la.want_outputs) {
      // pair skipped by the reference: no statistic, no ABF (NaN rows)
      for (int e = lane; e < 3 * L; e += 32)
        if (la.out_gen) la.out_gen[pair * 3 * L + e] = nan("");
      for (long long e = lane; e < C * K; e += 32)
        if (la.out_cfg) la.out_cfg[pair * C * K + e] = nan("");
      for (long long e = lane; e < C; e += 32)
        if (la.out_w) la.out_w[pair * (5 + C) + 5 + e] = nan("");
      for (int s = lane; s < S; s += 32) {
        if (la.out_n) la.out_n[pair * S + s] = 0;
        if (la.out_ss)
          for (int e = 0; e < 5; ++e) la.out_ss[(pair * S + s) * 5 + e] = nan("");
      }
    }
    if (la.want_outputs && lane == 0 && la.out_w) {
      double *o = la.out_w + pair * (5 + C);
      o[0] = w_gen[0];
      o[1] = w_gen[1];
      o[2] = w_gen[2];
      o[3] = w_gensin;
      o[4] = w_all;
    }
    if (la.stat_kind == STAT_JOIN_MAX || la.stat_kind == STAT_JOIN_AVG) {
      const double v = (la.true_rules && !all_geno) ? nan("") : stat_v;
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

} // namespace eqb
