// perm_kernel.cuh -- K4: the permutation kernel (uvlr), one CTA per (gene, permutation).
//
// Same semantics as pair_kernel (it replaces it for permuted evaluations and for the statistic of
// the true data, computed with the identity permutation by this same code so that `>=` ties are
// exact), but the contraction is done on the FP64 tensor cores in Gram form:
//   phase 1  per subgroup: (permuted) phenotype gather, keep-mask, CGS2 basis of [1, covariates] on
//            the kept rows, residual phenotype -> rows of a shared-memory matrix B (row stride
//            = 1 mod 16 doubles) + one 0/1 mask row per subgroup            gene_snp_pair.cpp:79-170
//   phase 2  per warp, tiles of 8 cis SNPs: H = X_tile * B' and R2 = (X_tile o X_tile) * Mask' with
//            mma.sync m8n8k4 f64; lane (row, k-slot) streams a contiguous quarter of its genotype row
//            from global/L2, B fragments come from shared memory
//   phase 3  thread per (SNP, subgroup): x~'x~ = R2 - sum_k H_k^2, x~'y~ = H_y, summary statistics and
//            standardisation (tabulated t->z map when the subgroup's degrees of freedom match, exact
//            otherwise); entries whose Gram form cancelled (x~'x~ < 1e-2 R2) are redone with explicit CGS2
//   phase 4  warp per SNP: ABFs + log-sum-exp (as pair_kernel), running statistic over the gene's SNPs,
//            CTA reduction                                                      gene.cpp:380-717
#pragma once

#include "fast_kernels.cuh"

namespace eqb {

__host__ __device__ inline int perm_lds(int ldn) { return ldn + 1; }
// rows of the shared B matrix: S*(Qmax+2) data rows, then 8 zero rows of padding (partial tiles read them)
__host__ __device__ inline int perm_brows(int S, int Qmax) { return S * (Qmax + 2) + 8; }
// per-warp scratch (doubles): H[8][W], st[8][3][S], tab[K][S][3], flags[8], agg[8][UL][3], vals[8][L], wc[8][S], wg[8], has[8]
// per-warp scratch (doubles): [H[8][W] aliased with agg[8][UL][3]], st[8][3][S], flags[8], vals[8][L], wc[8][S],
// wg[8], has[8]; for --pbf all also tab[K][S][3] and mitm[K][2^SA][3] (sums over the low SA subgroups; S <= 10, K <= 12)
__host__ __device__ inline int perm_mitm_sa(int S, int K) { return (S <= 10 && K <= 12 && S > 5) ? S - 5 : (S <= 5 && K <= 12 ? 0 : -1); }
__host__ __device__ inline size_t perm_u1_doubles(int S, int Qmax, int UL)
{
  const size_t W = (size_t)(((S * (Qmax + 2) + 7) / 8) * 8 + ((S + 7) / 8) * 8);
  const size_t a = 8 * W, b = (size_t)8 * UL * 3;
  return a > b ? a : b;
}
__host__ __device__ inline size_t perm_warp_doubles(int S, int Qmax, int K, int L, int UL, int which)
{
  const int sa = perm_mitm_sa(S, K);
  size_t d = perm_u1_doubles(S, Qmax, UL) + 8 * 3 * S + 8 + (size_t)8 * L + (size_t)8 * S + 8 + 8;
  if (which == 3) d += (size_t)3 * K * S + (sa >= 0 ? (size_t)3 * ((K + 3) & ~3) * (1u << sa) : 0); // mitm padded to 4 grid points
  return d;
}
__host__ __device__ inline size_t perm_smem_doubles(int S, int Qmax, int ldn, int K, int L, int UL, int which, int pw)
{
  return (size_t)perm_brows(S, Qmax) * perm_lds(ldn) + (size_t)pw * perm_warp_doubles(S, Qmax, K, L, UL, which);
}

// column c of the block is B row min(c * rstride, zero_row): rstride = 1 for the data rows, = Qmax+2 to pick
// the intercept row of every subgroup (the mask operand of the sums of squares, scaled by 1/sqrt(n))
__device__ __forceinline__ void dmma_block(const double *__restrict__ xrow, const double *__restrict__ Bsm, int lds,
                                           int ldn4, int rstride, int zero_row, int ntile, bool square,
                                           double *__restrict__ Hw, int W, int col_out, int g8, int kk)
{
  for (int t0 = 0; t0 < ntile; t0 += 4) {
    const int nt = min(4, ntile - t0);
    double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    const double *brow4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      brow4[u] = Bsm + (size_t)min(((t0 + u) * 8 + g8) * rstride, zero_row) * lds + (size_t)kk * ldn4;
    int t = 0;
    for (; t + 4 <= ldn4; t += 4) {
      double2 a0 = *reinterpret_cast<const double2 *>(xrow + t);
      double2 a1 = *reinterpret_cast<const double2 *>(xrow + t + 2);
      if (square) {
        a0.x *= a0.x;
        a0.y *= a0.y;
        a1.x *= a1.x;
        a1.y *= a1.y;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < nt) {
          const double *b = brow4[u] + t;
          dmma_m8n8k4(c[u][0], c[u][1], a0.x, b[0]);
          dmma_m8n8k4(c[u][0], c[u][1], a0.y, b[1]);
          dmma_m8n8k4(c[u][0], c[u][1], a1.x, b[2]);
          dmma_m8n8k4(c[u][0], c[u][1], a1.y, b[3]);
        }
    }
    for (; t < ldn4; ++t) {
      double a = xrow[t];
      if (square) a *= a;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < nt) dmma_m8n8k4(c[u][0], c[u][1], a, brow4[u][t]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < nt) {
        Hw[g8 * W + col_out + (t0 + u) * 8 + 2 * kk] = c[u][0];
        Hw[g8 * W + col_out + (t0 + u) * 8 + 2 * kk + 1] = c[u][1];
      }
  }
}

// ALLCFG: the instantiation that carries the 2^S-1 configuration code (--pbf all); the gen / gen-sin / separate
// statistics use the other one, whose register allocation is not burdened by the meet-in-the-middle tables
template <int PW, bool ALLCFG>
__global__ void __launch_bounds__(PW * 32) perm_kernel(const DevParams *__restrict__ prm_, const FastParams *__restrict__ fp_,
                                                       const LaunchArgs la, const GridTab gt)
{
  constexpr int PTHREADS = PW * 32;
  const DevParams &prm = *prm_;
  extern __shared__ double dyn_smem[];
  __shared__ int s_n[MAXS];
  __shared__ int s_rankz[MAXS];
  __shared__ unsigned int s_colvalid[MAXS];
  __shared__ double s_yy[MAXS], s_tss[MAXS], s_ybar[MAXS];
  __shared__ double w_part[PW][2];
  __shared__ int w_flag[PW][3];
  __shared__ double w_sep[PW][MAXS];
  __shared__ int w_sep_nan[MAXS];

  const int S = prm.S, N = prm.N, ldn = prm.ldn, Qmax = prm.Qmax, L = prm.L, K = prm.K;
  const int lds = perm_lds(ldn), ldn4 = ldn >> 2, R = Qmax + 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g8 = lane >> 2, kk = lane & 3;
  const int ppg = la.perms_per_gene > 0 ? la.perms_per_gene : 1;
  const int gi = blockIdx.x / ppg;
  const long long p = la.perms_per_gene > 0 ? la.p0 + (blockIdx.x % ppg) : -1;
  const int g = la.genes[gi];
  const long long mbeg = prm.cis_begin[g], mend = prm.cis_end[g];
  const unsigned short *perm = (p >= 0) ? la.perm_tab + ((size_t)la.gene_slot[gi] * la.P_total + p) * N : nullptr;

  const int nrowB = S * R, ntB = (nrowB + 7) / 8, ntM = (S + 7) / 8, W = (ntB + ntM) * 8;
  double *Bsm = dyn_smem;                                  // [perm_brows][lds]
  double *wbase = Bsm + (size_t)perm_brows(S, Qmax) * lds; // per-warp scratch
  const int UL = gt.UL;
  double *Hw = wbase + (size_t)warp * perm_warp_doubles(S, Qmax, K, L, UL, la.which);
  double *agg = Hw;                                    // [8][UL][3], aliases H (dead after phase 3 in join mode)
  double *stw = Hw + perm_u1_doubles(S, Qmax, UL);     // [8][3][S] standardised statistics of the tile's SNPs
  unsigned long long *flagw = (unsigned long long *)(stw + 8 * 3 * S); // [8] subgroups needing the explicit path
  double *valw = (double *)(flagw + 8);                // [8][L]
  double *wcw = valw + 8 * L;                          // [8][S]
  double *wgw = wcw + 8 * S;                           // [8]
  unsigned long long *hasw = (unsigned long long *)(wgw + 8); // [8]
  double *tab = (double *)(hasw + 8);                  // [K][S][3]   (--pbf all only)
  double *mitm = tab + 3 * K * S;                      // [K][2^SA][3] (--pbf all only)

  if (threadIdx.x < MAXS) w_sep_nan[threadIdx.x] = 0;
  // zero the padding rows
  for (int i = threadIdx.x; i < 8 * lds; i += PTHREADS) Bsm[(size_t)nrowB * lds + i] = 0.0;

  // ------------------------------------------------------------------ phase 1
  for (int s = warp; s < S; s += PW) {
    const SubDev &sb = prm.sub[s];
    double *q = Bsm + (size_t)s * R * lds;
    double *yt = q + (size_t)(Qmax + 1) * lds;
    for (int k = 1; k <= Qmax; ++k) // unused basis rows stay zero
      for (int i = lane; i < lds; i += 32) q[(size_t)k * lds + i] = 0.0;
    int n = 0;
    if (sb.gene_has[g]) {
      const double *Yg = sb.Yall + (size_t)g * ldn;
      for (int i = lane; i < lds; i += 32) {
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
    } else {
      for (int i = lane; i < lds; i += 32) {
        yt[i] = 0.0;
        q[i] = 0.0;
      }
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
    if (prm.qnorm) {
      // scratch: the padding rows are not big enough for a sort; --qnorm permutations use rows of the
      // NEXT subgroup's block only when it is not yet built -> keep it simple: sequential on the y row
      // using the (still zero) basis rows 1.. as scratch when Qmax >= 2, else the general kernel is used
      // (the host routes --qnorm with Qmax < 2 to pair_kernel)
      double *vals = q + (size_t)1 * lds;
      int *ord = (int *)(q + (size_t)2 * lds);
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
        for (int i = 0; i < lds; ++i) {
          q[(size_t)1 * lds + i] = 0.0;
          q[(size_t)2 * lds + i] = 0.0;
        }
      }
      __syncwarp();
    }
    const double inv_sqrt_n = 1.0 / sqrt((double)n);
    for (int i = lane; i < lds; i += 32) q[i] = (q[i] != 0.0) ? inv_sqrt_n : 0.0;
    __syncwarp();
    unsigned int colvalid = 1u;
    int rankz = 1;
    const int Q = sb.Q;
    for (int k = 1; k <= Q; ++k) {
      double *qk = q + (size_t)k * lds;
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
      if (__any_sync(0xffffffffu, missing) && lane == 0) atomicExch(la.err_flag, 1);
      __syncwarp();
      for (int pass = 0; pass < 2; ++pass)
        for (int j = 0; j < k; ++j) {
          if (!((colvalid >> j) & 1u)) continue;
          const double *qj = q + (size_t)j * lds;
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
        const double *qj = q + (size_t)j * lds;
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
      s_ybar[s] = ybar;
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ phases 2-4
  const bool join = prm.analysis == 1;
  const long long C = (la.which == 1) ? 0 : ((la.which == 2) ? S : prm.C);
  bool same_x = true;
  for (int s = 1; s < S; ++s) same_x = same_x && (prm.sub[s].X == prm.sub[0].X);
  Lse acc_stat;
  acc_stat.init();
  double max_stat = -INFINITY;
  bool first_nan = false;
  int cnt_nonnan = 0;
  double sep_all_min = 1.0;
  if (la.stat_kind == STAT_SEP_PER)
    for (int s = lane; s < S; s += 32) w_sep[warp][s] = INFINITY;
  __syncwarp();

  for (long long m0 = mbeg + (long long)warp * 8; m0 < mend; m0 += (long long)PW * 8) {
    const int tn = (int)min((long long)8, mend - m0);
    const long long mrow = min(m0 + g8, mend - 1);
    // ---- phase 2: DMMA contraction of the tile against every B row and mask row
    if (same_x) {
      const double *xrow = prm.sub[0].X + (size_t)mrow * ldn + (size_t)kk * ldn4;
      dmma_block(xrow, Bsm, lds, ldn4, 1, nrowB, ntB, false, Hw, W, 0, g8, kk);
      dmma_block(xrow, Bsm, lds, ldn4, R, nrowB, ntM, true, Hw, W, ntB * 8, g8, kk);
    } else {
      for (int s = 0; s < S; ++s) { // subgroups with different genotype matrices: one block each
        const double *xrow = prm.sub[s].X + (size_t)mrow * ldn + (size_t)kk * ldn4;
        // tiles starting at the subgroup's first row; the extra columns of a partial tile are ignored
        // compute the R rows of subgroup s (ceil(R/8) tiles) into a temporary window at the end of Hw is not
        // available; instead evaluate tile by tile and copy the needed columns
        const int nts = (R + 7) / 8;
        for (int t0 = 0; t0 < nts; ++t0) {
          double c0 = 0.0, c1 = 0.0;
          const double *brow = Bsm + (size_t)(s * R + t0 * 8 + g8) * lds + (size_t)kk * ldn4;
          for (int t = 0; t < ldn4; ++t) dmma_m8n8k4(c0, c1, xrow[t], brow[t]);
          const int cbase = s * R + t0 * 8 + 2 * kk;
          if (t0 * 8 + 2 * kk < R) Hw[g8 * W + cbase] = c0;
          if (t0 * 8 + 2 * kk + 1 < R) Hw[g8 * W + cbase + 1] = c1;
        }
        {
          double c0 = 0.0, c1 = 0.0;
          const double *mrowp = Bsm + (size_t)(g8 == 0 ? s * R : nrowB) * lds + (size_t)kk * ldn4; // column 0 = intercept row
          for (int t = 0; t < ldn4; ++t) {
            const double a = xrow[t];
            dmma_m8n8k4(c0, c1, a * a, mrowp[t]);
          }
          if (kk == 0) Hw[g8 * W + ntB * 8 + s] = c0; // column 0 of this block = mask row s
        }
        __syncwarp();
      }
    }
    __syncwarp();
    // ---- phase 3: thread per (SNP of the tile, subgroup): statistics + standardisation
    if (lane < 8) flagw[lane] = 0ull;
    __syncwarp();
    for (int it = lane; it < 8 * S; it += 32) {
      const int j = it / S, s = it % S;
      double *stj = stw + j * 3 * S;
      stj[s] = nan("");
      stj[S + s] = nan("");
      stj[2 * S + s] = nan("");
      if (j >= tn) continue;
      const SubDev &sb = prm.sub[s];
      const int n = s_n[s];
      if (!(n > 0 && sb.snp_has[m0 + j])) continue;
      const double *h = Hw + j * W + s * R;
      const double r2 = sqrt((double)n) * Hw[j * W + ntB * 8 + s]; // sum of squares over the kept rows
      double hh = 0.0;
      for (int k = 0; k <= sb.Q; ++k) hh += h[k] * h[k];
      const double xx = r2 - hh;
      if (r2 > 0.0 && xx < 1e-2 * r2) {
        atomicOr(&flagw[j], 1ull << s); // Gram form cancelled: explicit CGS2 below
        continue;
      }
      const FastSub &fs = fp_->sub[s];
      const double nu = (double)n - 2.0 - sb.Q;
      const bool use_tab = (fs.tz != nullptr) && (fs.tz_nu == nu);
      PairStat ps;
      stats_from_dots(h[Qmax + 1], xx, r2, sqrt((double)n) * h[0], s_yy[s], s_tss[s], s_ybar[s], n, sb.Q, s_rankz[s],
                      use_tab ? fs.tz : nullptr, fs.tz_nu, fs.tz_wmax, ps);
      stj[s] = ps.b;
      stj[S + s] = ps.v;
      stj[2 * S + s] = ps.t;
      // the p-value rides along in the (unused) H slot for the separate-analysis statistics
      Hw[j * W + s * R] = ps.pval;
    }
    __syncwarp();
    for (int j = 0; j < tn; ++j) { // explicit path for flagged entries (rare), whole warp per entry
      unsigned long long fl = flagw[j];
      while (fl) {
        const int s = __ffsll((long long)fl) - 1;
        fl &= fl - 1;
        const SubDev &sb = prm.sub[s];
        const double *q = Bsm + (size_t)s * R * lds;
        const double *yt = q + (size_t)(Qmax + 1) * lds;
        const double *Xm = sb.X + (size_t)(m0 + j) * ldn;
        double xraw2 = 0.0, xsum = 0.0, xx = 0.0, xy = 0.0;
        // x~ is rebuilt element-wise in registers-free form: two sweeps per basis column
        // (cost is irrelevant here); scratch = row j of Hw is too small, so recompute projections
        // h_k sequentially and accumulate x~ on the fly into the tab scratch (3*K*S doubles may be
        // smaller than ldn): use lane-strided private accumulation instead
        double xr[16]; // ldn <= 512 supported on this path; larger sample sizes use pair_kernel
        for (int jj = 0; jj < 16; ++jj) {
          const int i = lane + 32 * jj;
          const double v = (i < ldn && q[i] != 0.0) ? Xm[i] : 0.0;
          xr[jj] = v;
          xraw2 += v * v;
          xsum += v;
        }
        xraw2 = warp_sum(xraw2);
        xsum = warp_sum(xsum);
        for (int pass = 0; pass < 2; ++pass)
          for (int k = 0; k <= sb.Q; ++k) {
            if (!((s_colvalid[s] >> k) & 1u)) continue;
            const double *qk = q + (size_t)k * lds;
            double hk = 0.0;
            for (int jj = 0; jj < 16; ++jj) {
              const int i = lane + 32 * jj;
              if (i < ldn) hk += qk[i] * xr[jj];
            }
            hk = warp_sum(hk);
            for (int jj = 0; jj < 16; ++jj) {
              const int i = lane + 32 * jj;
              if (i < ldn) xr[jj] -= hk * qk[i];
            }
          }
        for (int jj = 0; jj < 16; ++jj) {
          const int i = lane + 32 * jj;
          if (i < ldn) {
            xx += xr[jj] * xr[jj];
            xy += xr[jj] * yt[i];
          }
        }
        xx = warp_sum(xx);
        xy = warp_sum(xy);
        if (lane == 0) {
          const int n = s_n[s];
          const FastSub &fs = fp_->sub[s];
          const double nu = (double)n - 2.0 - sb.Q;
          const bool use_tab = (fs.tz != nullptr) && (fs.tz_nu == nu);
          PairStat ps;
          stats_from_dots(xy, xx, xraw2, xsum, s_yy[s], s_tss[s], s_ybar[s], n, sb.Q, s_rankz[s],
                          use_tab ? fs.tz : nullptr, fs.tz_nu, fs.tz_wmax, ps);
          double *stj = stw + j * 3 * S;
          stj[s] = ps.b;
          stj[S + s] = ps.v;
          stj[2 * S + s] = ps.t;
          Hw[j * W + s * R] = ps.pval;
        }
        __syncwarp();
      }
    }
    __syncwarp();
    // ---- phase 4: permutation statistic contributions of the tile
    // 4a: has-masks, separate-analysis minima (lanes 0..7, one SNP each)
    if (lane < 8) {
      unsigned long long hm = 0ull;
      if (lane < tn)
        for (int s = 0; s < S; ++s)
          if ((s_n[s] > 0) && prm.sub[s].snp_has[m0 + lane]) hm |= (1ull << s);
      hasw[lane] = hm;
    }
    __syncwarp();
    if (!join) {
      for (int j = 0; j < tn; ++j) {
        const long long m = m0 + j;
        const bool is_first = (m == mbeg);
        double snp_pmin = 1.0;
        for (int s = 0; s < S; ++s) {
          const bool have = (hasw[j] >> s) & 1ull;
          const double pval = have ? Hw[j * W + s * R] : nan("");
          if (pval < snp_pmin) snp_pmin = pval;
          if (la.stat_kind == STAT_SEP_PER && lane == 0) {
            const double v = (prm.sub[s].gene_has[g] && prm.sub[s].snp_has[m]) ? (have ? pval : nan("")) : 1.0;
            if (isnan(v)) {
              if (is_first) atomicExch(&w_sep_nan[s], 1);
            } else if (v < w_sep[warp][s])
              w_sep[warp][s] = v;
          }
        }
        if (la.stat_kind == STAT_SEP_ALL && snp_pmin < sep_all_min) sep_all_min = snp_pmin;
      }
      __syncwarp();
      continue;
    }
    // 4b: per (SNP, unique phi2 of the "gen" row): sums over the subgroups with results (consistent
    // configuration), one logarithm per item (consistent_sums)
    const int UG = gt.pad; // the unique phi2 values of the gen row come first in gt.uphi
    for (int it = lane; it < 8 * UG; it += 32) {
      const int j = it / UG, u = it % UG;
      if (j >= tn) continue;
      double den, num, sing;
      consistent_sums(stw + j * 3 * S, S, hasw[j], gt.uphi[u], den, num, sing);
      double *a = agg + (j * UL + u) * 3;
      a[0] = den;
      a[1] = num;
      a[2] = sing;
    }
    __syncwarp();
    // 4c: "gen" values on gridL, then their log10_weighted_sum by 4 lanes per SNP (lane = 4 j + q)
    const int qj = lane >> 2, qq = lane & 3;
    for (int it = lane; it < 8 * L; it += 32) {
      const int j = it / L, k = it % L;
      const double *a = agg + (j * UL + gt.idxL[k]) * 3;
      valw[j * L + k] = (j < tn) ? abf_from_sums(a[0], a[1], a[2], gt.omaL[k]) : 0.0;
    }
    __syncwarp();
    {
      const double *vj = valw + qj * L;
      const double wL = 1.0 / (double)L;
      const double w = lws_quad(L, qq, [&](int k) { return vj[k]; }, [&](int) { return wL; });
      if (qq == 0) wgw[qj] = w;
    }
    __syncwarp();
    if (la.which == 2) {
      // 4d: singletons on gridS: one lane per (SNP, subgroup); the K values stay in registers for the two-pass
      // log10_weighted_sum (K <= 16; larger grids use the online form)
      for (int it = lane; it < 8 * S; it += 32) {
        const int j = it / S, c = it % S;
        if (j >= tn) continue;
        const double *stj = stw + j * 3 * S;
        const bool has = (hasw[j] >> c) & 1ull;
        const double b = stj[c], vv = stj[S + c], tt = stj[2 * S + c];
        double res;
        if (K <= 16) {
          double v[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = (k < K && has) ? singleton_value(b, vv, tt, prm.phi2S[k], prm.oma2S[k]) : 0.0;
          double mx = v[0];
#pragma unroll
          for (int k = 1; k < 16; ++k)
            if (k < K) mx = (v[k] > mx) ? v[k] : mx;
          double sum = 0.0;
          const double wK = 1.0 / (double)K;
#pragma unroll
          for (int k = 0; k < 16; ++k)
            if (k < K) {
              const double e = exp10_fast(v[k] - mx);
              sum += isnan(v[k]) ? 0.0 : wK * e;
            }
          res = mx + log10(sum);
          if (fabs(res) <= DBL_EPSILON) res = 0.0;
          if (K == 0) res = nan("");
        } else {
          Lse acc;
          acc.init();
          for (int k = 0; k < K; ++k)
            acc.add(has ? singleton_value(b, vv, tt, prm.phi2S[k], prm.oma2S[k]) : 0.0, 1.0 / (double)K, k == 0);
          res = acc.result();
        }
        wcw[j * S + c] = res;
      }
      __syncwarp();
      { // CalcBMAlite (gene_snp_pair.cpp:552-570): S singleton terms (0.5/S each) then the consistent one (0.5)
        const double *wj = wcw + qj * S;
        const double wg = wgw[qj], wS = 0.5 / (double)S;
        __syncwarp();
        const double w = lws_quad(S + 1, qq, [&](int k) { return (k < S) ? wj[k] : wg; },
                                  [&](int k) { return (k < S) ? wS : 0.5; });
        if (qq == 0) wgw[qj] = w;
      }
      __syncwarp();
    } else if (ALLCFG && la.which == 3) {
      // 4e: all 2^S-1 configurations, warp per SNP.  Fast form (S <= 10): meet in the middle -- lane b owns
      // the subset b of the high min(5,S) subgroups (its per-grid-point sums live in registers), the 2^SA
      // subsets of the low subgroups come from a shared table -- and the BMA sum is accumulated in the
      // linear domain against the bound M = sum_s t_s^2/2 - 350 >= every exponent - 359:
      //   10^abf = exp(A) (1 + oma2 den)^-1/2 exp(num^2 oma2 / (2 (1 + oma2 den)))   (no log, no division)
      const int SA = perm_mitm_sa(S, K);
      for (int j = 0; j < tn; ++j) {
        const double *st = stw + j * 3 * S;
        const unsigned long long has_mask = hasw[j];
        bool fast_ok = SA >= 0;
        double Mref = 0.0;
        for (int s = 0; s < S; ++s)
          if ((has_mask >> s) & 1ull) {
            const double t = st[2 * S + s];
            if (isnan(t) || isnan(st[s]) || isnan(st[S + s]) || isinf(st[s])) fast_ok = false;
            else if (fabs(t) >= 1e-8) Mref += 0.5 * t * t;
          }
        Mref -= 350.0;
        // per-(grid point, subgroup) terms {1/(v+phi2), b/(v+phi2), ln of the single-subgroup ABF}
        for (int e = lane; e < K * S; e += 32) {
          const int k = e / S, s = e % S;
          double *te = tab + (size_t)e * 3;
          te[0] = 0.0;
          te[1] = 0.0;
          te[2] = 0.0;
          if ((has_mask >> s) & 1ull) {
            term_entry(st[s], st[S + s], st[2 * S + s], prm.phi2S[k], te[0], te[1], te[2]);
            if (fast_ok) te[2] *= LN10;
          }
        }
        __syncwarp();
        double res = nan("");
        bool done = false;
        if (fast_ok) {
          const int SB = S - SA;
          const int nA = 1 << SA;
          const int Kp = (K + 3) & ~3; // grid points are evaluated four at a time; the padding entries are neutral
          for (int e = lane; e < Kp * nA; e += 32) { // sums over the subsets of the low subgroups
            const int k = e / nA, a = e % nA;
            double d = 0.0, n_ = 0.0, A = 0.0;
            if (k < K)
              for (int s = 0; s < SA; ++s)
                if ((a >> s) & 1) {
                  const double *te = tab + ((size_t)k * S + s) * 3;
                  d += te[0];
                  n_ += te[1];
                  A += te[2];
                }
            double *m3 = mitm + (size_t)e * 3;
            m3[0] = d;
            m3[1] = n_;
            m3[2] = A;
          }
          __syncwarp();
          double acc = 0.0;
          if (lane < (1 << SB)) {
            double hd[12], hn[12], hA[12], hom[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) {
              hd[k] = hn[k] = 0.0;
              hA[k] = (k < K) ? 0.0 : -1e300; // padding: exp(-inf) = 0
              hom[k] = (k < K) ? prm.oma2S[k] : 0.0;
              if (k < K)
                for (int s = 0; s < SB; ++s)
                  if ((lane >> s) & 1) {
                    const double *te = tab + ((size_t)k * S + SA + s) * 3;
                    hd[k] += te[0];
                    hn[k] += te[1];
                    hA[k] += te[2];
                  }
            }
            const int pb = __popc(lane);
            for (int a = 0; a < nA; ++a) {
              if (a == 0 && lane == 0) continue; // the empty configuration is not part of the model
              const double w = prm.size_weight[pb + __popc(a)];
              double part = 0.0;
#pragma unroll
              for (int k0 = 0; k0 < 12; k0 += 4)
                if (k0 < K) { // warp-uniform; the four evaluations below are independent and branch-free
                  double ev[4];
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const int k = k0 + u;
                    const double *m3 = mitm + ((size_t)k * nA + a) * 3;
                    const double den = hd[k] + m3[0], num = hn[k] + m3[1], A = hA[k] + m3[2];
                    const double r = rsqrt_fast_nb(fma(hom[k], den, 1.0));
                    ev[u] = r * exp_fast_nb(A + 0.5 * num * num * hom[k] * r * r - Mref);
                  }
                  part += (ev[0] + ev[1]) + (ev[2] + ev[3]);
                }
              acc += w * part;
            }
          }
          acc = warp_sum(acc);
          if (acc > 0.0 && acc < INFINITY) {
            res = (Mref + log(acc / (double)K)) / LN10;
            if (fabs(res) <= DBL_EPSILON) res = 0.0;
            done = true;
          } else {
            // out of the representable window (or NaN): redo this SNP in the log domain
            for (int e = lane; e < K * S; e += 32) tab[(size_t)e * 3 + 2] /= LN10;
            __syncwarp();
          }
        }
        if (!done) {
          Lse bma;
          bma.init();
          for (long long c = lane; c < C; c += 32) {
            const unsigned long long mask = prm.cfg_mask[c] & has_mask;
            Lse b;
            b.init();
            for (int k = 0; k < K; ++k) {
              const double *tk = tab + (size_t)k * S * 3;
              double den = 0.0, num = 0.0, sing = 0.0;
              unsigned long long mm = mask;
              while (mm) {
                const int s = __ffsll((long long)mm) - 1;
                mm &= mm - 1;
                den += tk[3 * s];
                num += tk[3 * s + 1];
                sing += tk[3 * s + 2];
              }
              b.add(abf_from_sums(den, num, sing, prm.oma2S[k]), 1.0 / (double)K, k == 0);
            }
            bma.add(b.result(), prm.cfg_weight[c], c == 0); // CalcBMA (gene_snp_pair.cpp:572-602)
          }
          bma = warp_merge(bma);
          res = bma.result();
        }
        if (lane == 0) wgw[j] = res;
        __syncwarp();
      }
    }
    // 4f: running statistic over the SNPs of the gene, in SNP order
    for (int j = 0; j < tn; ++j) {
      const bool is_first = (m0 + j == mbeg);
      const double v = wgw[j];
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

  // ------------------------------------------------------------------ CTA reduction (as pair_kernel)
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
    for (int s = threadIdx.x; s < S; s += PTHREADS) {
      double v = INFINITY;
      for (int w = 0; w < PW; ++w) v = fmin(v, w_sep[w][s]);
      if (la.true_rules) {
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
    for (int w = 0; w < PW; ++w) {
      fn = fn || w_flag[w][0];
      nn += w_flag[w][2];
    }
    double res;
    if (la.stat_kind == STAT_JOIN_MAX) {
      double v = -INFINITY;
      for (int w = 0; w < PW; ++w) v = fmax(v, w_part[w][0]);
      res = (fn && !la.true_rules) ? nan("") : v;
    } else if (la.stat_kind == STAT_JOIN_AVG) {
      Lse t;
      t.init();
      for (int w = 0; w < PW; ++w) {
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
    } else {
      double v = 1.0;
      for (int w = 0; w < PW; ++w) v = fmin(v, w_part[w][0]);
      res = v;
    }
    out[0] = res;
  }
}

} // namespace eqb
