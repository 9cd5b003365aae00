// fast_all_kernel.cuh -- K2+K3 for `--bfs all` (the c3 / GTEx headline path).
//
// Reference: GeneSnpPair::CalcAbfsUvlrForEachConfiguration + CalcBMAlite + CalcBMA (gene_snp_pair.cpp:504-602): for every
// pair, the closed-form ABF of each of the 2^S - 1 configurations on every gridS point (raw values), their grid averages,
// and the two model averages.  One CTA of 8 warps owns a tile of 32 pairs:
//   A  mma.sync.m8n8k4 (DMMA) tile product xy[pairs][S] = X_tile . Ytil_gene^T, one 8-row block per warp (4 warps)
//   B  thread per (pair, subgroup): summary statistics + standardisation
//   C  pair after pair, the whole CTA on one pair; LANE = CONFIGURATION in the reference's order (gsl_combination order),
//      every warp takes every 8th group of 32 consecutive configurations.  A configuration's sums over its subgroups
//      cost O(1): the subgroups are split into a low part (<= 5) and a high part, the sums of every subset of either
//      part are tabulated per grid point in shared memory ONCE PER PAIR for the whole CTA (component-major: the 16 high
//      entries sit in distinct banks, and 32 consecutive configurations share a handful of low entries).  The K raw
//      values of a configuration go to the warp's staging tile and leave as contiguous runs (the output rows of 32
//      consecutive configurations are contiguous); their grid average is accumulated in the linear domain against the
//      likelihood-ratio bound (online log-sum-exp as the fallback: NaN rules of utils::log10_weighted_sum).
//      (First version: one warp per tile with private tables -- 27 KB of shared memory per warp, 7 warps per SM,
//      2.4 ms per tile: latency-bound at 0.37 instructions per clock per SM and no faster than the kernel it replaced.)
// want_raw = 0 (fa.out_cfg == nullptr) skips the emission: compute and emission can be timed separately.
#pragma once

#include "fast_kernels.cuh"

namespace eqb {

constexpr int FA_SL = 5;      // subgroups of the low part
constexpr int FA_MAXS = 10;   // S <= 10 (high part <= 5 subgroups), K <= 16
constexpr int FA_MAXK = 16;
constexpr int FA_MAXG = 4;    // configuration groups per warp: ceil((2^10 - 1) / 32 / 8)

__host__ __device__ inline int fa_low(int S) { return S < FA_SL ? S : FA_SL; }
// doubles of the CTA's phase-C scratch: A[K][3][2^SL] | B[K][3][2^SH] | MA[2^SL] | MB[2^SH] | te[K][S][3] | usum[UL][3] |
// gv[3L] | w0[32] | part[WARPS][3] | stg[WARPS][32][K+1]
__host__ __device__ inline size_t fa_scratch_doubles(int S, int K, int L, int UL)
{
  const int SL = fa_low(S), SH = S - SL;
  return (size_t)K * 3 * (1 << SL) + (size_t)K * 3 * (1 << SH) + (1 << SL) + (1 << SH) + (size_t)K * S * 3 + (size_t)UL * 3 +
         (size_t)3 * L + 32 + (size_t)WARPS * 3 + (size_t)WARPS * 32 * (K + 1);
}
__host__ __device__ inline size_t fast_all_smem_bytes(int S, int K, int L, int UL)
{
  return fast_warp_smem_bytes(S) + fa_scratch_doubles(S, K, L, UL) * 8;
}

// utils::log10_weighted_sum (utils_math.cpp:100-131) of n <= 32 values held one per lane (lanes >= n pass anything),
// by the whole warp: maximum seeded with element 0 (a NaN there poisons the result), NaN elements skipped, weighted sum
// of 10^(x - max), |result| <= DBL_EPSILON snapped to 0.  A serial loop over the values is a chain of ~40 dependent
// instructions per element on ONE thread while the rest of the CTA waits at the next barrier.
__device__ __forceinline__ double warp_lws(double x, double w, int n, int lane, const TabRef T)
{
  const bool in = lane < n;
  const double x0 = __shfl_sync(0xffffffffu, x, 0);
  double mx = (in && x == x) ? x : -INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  double e = (in && x == x) ? w * exp10_tab16<true>(x - mx, T) : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
  if (x0 != x0) return nan("");
  double r = fma(log_tab16(e, T), EQB_INV_LN10, mx);
  if (fabs(r) <= DBL_EPSILON) r = 0.0;
  return r;
}

template <bool DM>
__global__ void __launch_bounds__(THREADS, 2) fast_pair_all_kernel(const DevParams *__restrict__ prm_, const FastParams *__restrict__ fp_,
                                                                   const FastArgs fa, const GridTab gt,
                                                                   const __grid_constant__ GridConst gc)
{
  const DevParams &prm = *prm_;
  extern __shared__ double fsm[];
  const int S = prm.S, ldn = prm.ldn, L = prm.L, K = prm.K, UL = gt.UL;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ BfTabs Tsm;
  bf_tabs_init(Tsm);
  TabRef T;
  T.base = smem_u32(&Tsm);
  const long long tile = blockIdx.x;
  const long long q0 = fa.tile_q0[tile];
  const int tn = (int)(fa.tile_q0[tile + 1] - q0); // 1 .. 32
  const long long C = prm.C;
  const int sst = (3 * S) | 1;
  char *wbase = reinterpret_cast<char *>(fsm);
  double *xy = reinterpret_cast<double *>(wbase);            // [32][S]
  double *st = xy + (size_t)32 * S;                          // [32][sst]  b, v, t per subgroup
  unsigned long long *hasm = (unsigned long long *)(st + (size_t)32 * sst); // [32]
  long long *s_pair = (long long *)(hasm + 32);
  long long *s_m = s_pair + 32;
  int *s_gene = (int *)(s_m + 32);
  double *scr = reinterpret_cast<double *>(wbase + fast_warp_smem_bytes(S));
  const int SL = fa_low(S), SH = S - SL, NA = 1 << SL, NB = 1 << SH;
  double *tA = scr;                                // [K][3][NA]
  double *tB = tA + (size_t)K * 3 * NA;            // [K][3][NB]
  double *mA = tB + (size_t)K * 3 * NB;            // [NA] bound of the low subset: sum t^2 / 2
  double *mB = mA + NA;                            // [NB]
  double *te = mB + NB;                            // [K][S][3] per-(grid point, subgroup) terms (natural-log units)
  double *usum = te + (size_t)K * S * 3;           // [UL][3]
  double *gv = usum + (size_t)UL * 3;              // [3L]
  double *w0 = gv + (size_t)3 * L;                 // [32] grid averages of the first 32 configurations (singletons first)
  double *part = w0 + 32;                          // [WARPS][3] partial model averages (m, acc, poisoned)
  double *stg = part + (size_t)WARPS * 3 + (size_t)warp * 32 * (K + 1); // this warp's [32][K+1] staging tile

  if (threadIdx.x < tn) {
    const long long q = q0 + threadIdx.x;
    int lo = fa.tile_gene[tile];
    while (lo + 1 < fa.n_genes && fa.fast_base[lo + 1] <= q) ++lo;
    const int g = fa.genes[lo];
    const long long off = q - fa.fast_base[lo];
    s_gene[threadIdx.x] = g;
    s_m[threadIdx.x] = prm.cis_begin[g] + off;
    s_pair[threadIdx.x] = fa.pair_off[lo] + off;
    hasm[threadIdx.x] = 0ull;
  }
  // the configurations of this lane (the same for every pair): masks and BMA weights stay in registers
  unsigned long long cmask[FA_MAXG];
  double cwt[FA_MAXG];
#pragma unroll
  for (int jj = 0; jj < FA_MAXG; ++jj) {
    const long long c = ((long long)(warp + jj * WARPS)) * 32 + lane;
    cmask[jj] = (c < C) ? prm.cfg_mask[c] : 0ull;
    cwt[jj] = (c < C) ? prm.cfg_weight[c] : 0.0;
  }
  __syncthreads();
  // ---------------- phase A: contraction, one block of 8 pairs per warp
  if (warp < 4 && warp * 8 < tn) {
    const int r0 = warp * 8, tw = min(8, tn - r0);
    for (int s0 = 0; s0 < S; s0 += 8) {
      const int sn = min(8, S - s0);
      const FastSub *fsub = fp_->sub + s0;
      if (DM)
        contract_tile_dmma(prm.sub[s0].X, fsub, sn, s_m + r0, s_gene + r0, tw, S, ldn, lane, xy + (size_t)r0 * S + s0);
      else
        for (int a = 0; a < sn; ++a)
          contract_tile<1>(prm.sub[s0 + a].X, fsub + a, s_m + r0, s_gene + r0, tw, S, ldn, lane, xy + (size_t)r0 * S + s0 + a);
    }
  }
  __syncthreads();
  // ---------------- phase B: thread per (pair, subgroup)
  for (int it = threadIdx.x; it < tn * S; it += THREADS) {
    const int j = it / S, s = it - j * S;
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
      stats_from_dots(xy[(size_t)j * S + s], xs[0], xs[1], xs[2], ys[0], ys[1], ys[2], fs.n, sb.Q, fs.rankz, fs.tz, fs.tz_nu,
                      fs.tz_wmax, ps);
      atomicOr(&hasm[j], 1ull << s);
    }
    st[(size_t)j * sst + s] = ps.b;
    st[(size_t)j * sst + S + s] = ps.v;
    st[(size_t)j * sst + 2 * S + s] = ps.t;
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
  // ---------------- phase C: pair after pair, lane = configuration
  const int rpi = 32 / K;                         // staged rows copied out per step (their K values are contiguous)
  const int cr = lane / K, ck = lane - cr * K;    // this lane's (row, grid point) in a copy-out step
  const double invK = 1.0 / (double)K, wL = 1.0 / (double)L;
  const int ngroups = (int)((C + 31) / 32);
  for (int j = 0; j < tn; ++j) {
    const double *stj = st + (size_t)j * sst;
    const unsigned long long has = hasm[j];
    const long long pair = s_pair[j];
    double *ow = fa.out_w + pair * (5 + C);
    // ---- step 1: sums of the consistent configuration per unique phi2; per-(grid point, subgroup) terms in
    // natural-log units { 1/(v+phi2), b/(v+phi2), ln single-subgroup ABF }
    for (int u = threadIdx.x; u < UL; u += THREADS) {
      double den, num, sing;
      consistent_sums_t(stj, S, has, gt.uphi[u], den, num, sing, T);
      usum[u * 3] = den;
      usum[u * 3 + 1] = num;
      usum[u * 3 + 2] = sing;
    }
    for (int e = threadIdx.x; e < K * S; e += THREADS) {
      const int k = e / S, s = e - k * S;
      double d = 0.0, bd = 0.0, A = 0.0;
      const double b = stj[s], v = stj[S + s], tt = stj[2 * S + s];
      if (((has >> s) & 1ull) && !(fabs(tt) < 1e-8)) {
        const double phi2 = gc.phiS[k];
        const double inv = rcp_n(v + phi2);
        d = inv;
        bd = b * inv;
        A = (phi2 == 0.0) ? 0.0 : fma(0.5, log_tab16(v * inv, T), 0.5 * tt * tt * phi2 * inv);
      }
      te[e * 3] = d;
      te[e * 3 + 1] = bd;
      te[e * 3 + 2] = A;
    }
    __syncthreads();
    // ---- step 2: the 3L consistent values (gene_snp_pair.cpp:364-416); subset sums of the low / high part for every
    // grid point (component-major) and their likelihood-ratio bounds
    for (int e = threadIdx.x; e < 3 * L; e += THREADS) {
      const double *a = usum + 3 * gt.idxL[e];
      const double v = abf_from_sums_t(a[0], a[1], a[2], gt.omaL[e], T);
      gv[e] = v;
      if (fa.out_gen) fa.out_gen[pair * 3 * L + e] = v;
    }
    for (int it = threadIdx.x; it < K * (NA + NB); it += THREADS) {
      const bool low = it < K * NA;
      const int i2 = low ? it : it - K * NA, sh = low ? SL : SH, nn = low ? NA : NB, s0 = low ? 0 : SL;
      const int k = i2 >> sh, a = i2 & (nn - 1);
      double d = 0.0, n_ = 0.0, A = 0.0;
      for (int s = 0; s < sh; ++s)
        if ((a >> s) & 1) {
          const double *t3 = te + ((size_t)k * S + s0 + s) * 3;
          d += t3[0];
          n_ += t3[1];
          A += t3[2];
        }
      double *tab = (low ? tA : tB) + (size_t)k * 3 * nn + a;
      tab[0] = d;
      tab[nn] = n_;
      tab[2 * nn] = A;
    }
    for (int a = threadIdx.x; a < NA + NB; a += THREADS) {
      const bool low = a < NA;
      const int bits = low ? a : a - NA, s0 = low ? 0 : SL, ns = low ? SL : SH;
      double m = 0.0;
      for (int s = 0; s < ns; ++s)
        if ((bits >> s) & 1) {
          const double tt = stj[2 * S + s0 + s];
          if (((has >> (s0 + s)) & 1ull) && !(fabs(tt) < 1e-8)) m = fma(0.5 * tt, tt, m);
        }
      (low ? mA : mB)[bits] = m;
    }
    __syncthreads();
    // ---- step 3: every configuration on gridS, warp w takes the groups w, w + 8, ...
    if (warp >= WARPS - 3) { // grid averages of gen / gen-fix / gen-maxh: one row per warp, the lanes share the row
      const int r = warp - (WARPS - 3);
      double wr;
      if (L <= 32)
        wr = warp_lws(lane < L ? gv[r * L + lane] : 0.0, wL, L, lane, T);
      else {
        LseTab q;
        q.init();
        for (int k = 0; k < L; ++k) q.add(gv[r * L + k], wL, k == 0, T);
        wr = q.result(T);
      }
      if (lane == 0) {
        ow[r] = wr;
        if (r == 0) part[WARPS * 3 - 1] = wr; // (read back by warp 0 below; slot 3 of the last warp is unused)
      }
    }
    LseTab bma;
    bma.init();
#pragma unroll
    for (int jj = 0; jj < FA_MAXG; ++jj) {
      const int grp = warp + jj * WARPS;
      if (grp >= ngroups) break; // warp-uniform
      const long long c0 = (long long)grp * 32, c = c0 + lane;
      const bool valid = c < C;
      const unsigned long long mask = cmask[jj] & has;
      const int ia = (int)(mask & (unsigned long long)(NA - 1)), ib = (int)(mask >> SL);
      const double Mc = mA[ia] + mB[ib];
      double acc = 0.0;
      double *srow = stg + (size_t)lane * (K + 1);
      const double *pa = tA + ia, *pb = tB + ib;
#pragma unroll 2
      for (int k = 0; k < K; ++k, pa += 3 * NA, pb += 3 * NB) {
        const double den = pa[0] + pb[0], num = pa[NA] + pb[NB], sing = pa[2 * NA] + pb[2 * NB];
        // natural-log ABF; CalcLog10AbfUvlr's guards (see abf_from_sums) as a select: z >= 1 is a normal number whenever
        // den is one, a NaN den (NaN statistics) gives 0 like the reference's "V < +Inf" test
        const double oma2 = gc.omaS[k];
        const double z = fma(oma2, den, 1.0);
        const bool ok = num != 0.0 && den != 0.0 && den == den && z < 1e300;
        const double zz = ok ? z : 1.0;
        double x = sing + fma(-0.5, log_tab16_pos(zz, T), 0.5 * num * num * oma2 * rcp_n(zz));
        x = ok ? x : 0.0;
        srow[k] = x * EQB_INV_LN10;
        acc += exp_tab16<true>(x - Mc, T);
      }
      double w;
      if (acc > 1e-280 && acc < INFINITY) {
        w = (Mc + log_tab16(acc * invK, T)) * EQB_INV_LN10;
        if (fabs(w) <= DBL_EPSILON) w = 0.0;
      } else { // NaN values or a sum outside the representable window: online form on the staged values
        LseTab r;
        r.init();
        for (int k = 0; k < K; ++k) r.add(srow[k], invK, k == 0, T);
        w = r.result(T);
      }
      if (valid) {
        ow[5 + c] = w;
        bma.add(w, cwt[jj], c == 0, T); // CalcBMA (gene_snp_pair.cpp:572-602)
      }
      if (grp == 0) w0[lane] = w;
      __syncwarp();
      if (fa.out_cfg) {
        // rows of 32 consecutive configurations are contiguous in the output: copy rpi rows (rpi * K lanes) per step
        double *dst = fa.out_cfg + (pair * C + c0) * K;
        const int nrow = (int)min((long long)32, C - c0);
        if (cr < rpi)
          for (int r0 = 0; r0 < nrow; r0 += rpi) {
            const int r = r0 + cr;
            if (r < nrow) dst[(size_t)r * K + ck] = stg[(size_t)r * (K + 1) + ck];
          }
      }
      __syncwarp();
    }
    // merge the lanes' partial model averages (online log-sum-exp states), then the warps' through shared memory
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double m2 = __shfl_xor_sync(0xffffffffu, bma.m, o), a2 = __shfl_xor_sync(0xffffffffu, bma.acc, o);
      const int p2 = __shfl_xor_sync(0xffffffffu, (int)bma.poisoned, o);
      const double mx = fmax(bma.m, m2);
      double a = 0.0;
      if (bma.m > -INFINITY) a = fma(bma.acc, exp10_tab16<true>(bma.m - mx, T), a);
      if (m2 > -INFINITY) a = fma(a2, exp10_tab16<true>(m2 - mx, T), a);
      bma.m = mx;
      bma.acc = a;
      bma.poisoned = bma.poisoned || p2;
    }
    if (lane == 0) {
      part[warp * 3] = bma.m;
      part[warp * 3 + 1] = bma.acc;
      if (warp < WARPS - 1) part[warp * 3 + 2] = bma.poisoned ? 1.0 : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      // the warps' partial model averages (lane w holds warp w's state), merged by the lanes together
      const double pm = (lane < WARPS) ? part[lane * 3] : -INFINITY, pa2 = (lane < WARPS) ? part[lane * 3 + 1] : 0.0;
      double mx = pm;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      double a = (pm > -INFINITY) ? pa2 * exp10_tab16<true>(pm - mx, T) : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      LseTab all;
      all.m = mx;
      all.acc = a;
      all.poisoned = part[2] != 0.0; // (configuration 0 belongs to warp 0)
      // CalcBMAlite (gene_snp_pair.cpp:552-570): the S singleton averages (0.5 / S each), then the consistent one (0.5)
      const double term = (lane < S) ? w0[lane] : part[WARPS * 3 - 1];
      const double lite = warp_lws(term, (lane < S) ? 0.5 / (double)S : 0.5, S + 1, lane, T);
      if (lane == 0) {
        ow[3] = lite;
        ow[4] = all.result(T);
      }
    }
    // (the next pair's step 1 writes usum / te only after the barrier that ends its own step 1 ... the tables it replaces
    // are no longer read: every warp passed the barrier above after its last table lookup; part / w0 are rewritten only
    // in the next step 3, two barriers away)
  }
}

} // namespace eqb
