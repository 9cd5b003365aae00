// fast_all_kernel.cuh -- K2+K3 for `--bfs all` (the c3 / GTEx headline path).
//
// Reference: GeneSnpPair::CalcAbfsUvlrForEachConfiguration + CalcBMAlite + CalcBMA (gene_snp_pair.cpp:504-602): for every
// pair, the closed-form ABF of each of the 2^S - 1 configurations on every gridS point (raw values), their grid averages,
// and the two model averages.  Two passes:
//   1  fast_pair_warp_kernel (fast_kernels.cuh, which = 3): warp per tile of 32 pairs -- DMMA contraction, summary statistics
//      and standardisation, the 3L values of the consistent configuration and their grid averages; leaves b, v, t of every
//      (pair, subgroup) in fa.st_all (216 bytes per pair at S = 9).  HBM-bound, a few per cent of the step.
//   2  fast_pair_all_kernel (this file): persistent CTAs of 4 warps, four of them per SM; a CTA takes tiles of
//      PPW = min(4, 32 / K) consecutive pairs:
//   T  LANE = (pair, grid point k): the lane's per-subgroup terms { 1/(v+phi2_k), b/(v+phi2_k), ln ABF_s } depend on nothing
//      else, so the sums over a configuration's subgroups cost O(1): the subgroups are split into three parts and the sums
//      of every subset of a part are tabulated ONCE per (pair, k) in lane-private columns of shared memory (<= 8 + 8 + 8
//      entries for S = 9; one warp per part).  The four warps of the CTA share the tables.
//   C  the configurations in the reference's order (gsl_combination order, masks staged in shared memory: warp-uniform),
//      chunks of 32 dealt round-robin to the warps:
//      C1  lane = (pair, k): 9 conflict-free LDS + 6 DADD per value, then the ES-model ABF; the value goes straight to its
//          raw output slot (the K lanes of a pair write one contiguous run) and into the warp's staging tile [32][33]
//      C2  lane = configuration: for each pair the reference's two-pass log10_weighted_sum over the K staged values
//          (utils_math.cpp:100-131), coalesced store of the grid averages, online accumulation of the BMA (CalcBMA); the
//          first chunk holds the S singletons: BMAlite right there
//   M  the warps' partial BMA states merged through shared memory.
// History (c3 slice, 9 ragged tissues, 511 configurations x 10 grid points): (1) CTA per pair, lane = configuration, tables
// shared by the CTA, four barriers per pair: 174 thread instructions per value, a third of the issue slots at barriers,
// 19 M pairs/s.  (2) this mapping with one CTA of 8 warps per SM, private tables per warp (27 KB each): 17 M pairs/s --
// 8 warps per SM cannot hide the FP64 latency even with four evaluations in flight per warp (issue slots 35 % used, 19 % of
// the time spent by 5 of 8 warps waiting for the contraction).  (3) the tables shared by the 4 warps of a small CTA, 16 warps
// per SM, contraction still inside: 18.7 M pairs/s, 28 % of the time in the 3-row contraction (58 dependent round trips to
// HBM per tile) and at the barriers around it.  (4) contraction and statistics as a first pass over 32-pair tiles.
// fa.out_cfg == nullptr / fa.out_gen == nullptr skip the raw-value emission: compute and emission are timed separately.
#pragma once

#include "fast_all.h"
#include "fast_kernels.cuh"

namespace eqb {

// utils::log10_weighted_sum (utils_math.cpp:100-131) of n <= 32 values held one per lane (lanes >= n pass anything),
// by the whole warp: maximum seeded with element 0 (a NaN there poisons the result), NaN elements skipped, weighted sum
// of 10^(x - max), |result| <= DBL_EPSILON snapped to 0.
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
  const double d1 = e - 1.0; // (weights summing to 1 on equal values: ln(1 + d) = d, not the table's 1e-13)
  double r = fma((fabs(d1) < 1e-8) ? d1 : log_tab16(e, T), EQB_INV_LN10, mx);
  if (fabs(r) <= DBL_EPSILON) r = 0.0;
  return r;
}

__device__ __forceinline__ void lse_merge(LseTab &a, double m2, double acc2, const TabRef T)
{
  const double mx = fmax(a.m, m2);
  double r = 0.0;
  if (a.m > -INFINITY) r = fma(a.acc, exp10_tab16<true>(a.m - mx, T), r);
  if (m2 > -INFINITY) r = fma(acc2, exp10_tab16<true>(m2 - mx, T), r);
  a.m = mx;
  a.acc = r;
}

// PPW: pairs per tile = min(4, 32 / K) (compile-time: the per-pair chains of phase C2 are unrolled side by side).
// fa.use_dmma: every group of 8 subgroups shares one genotype matrix and phase A runs on the tensor cores (the common case).
#ifndef FA_U
#define FA_U 2 // configurations per C1 step (independent chains in flight per warp)
#endif
#ifndef FA_MINB
#define FA_MINB 4
#endif
#ifndef FA_CLAMP
#define FA_CLAMP true
#endif
// LIN (no raw values requested): C1 works in the LINEAR domain against the likelihood-ratio bound Mc = sum_{s in c} t_s^2 / 2 (a
//      Bayes factor cannot exceed the maximised likelihood ratio): the per-subgroup term is tabulated as ln ABF_s - t_s^2 / 2, so
//      the table sum is x - Mc + ln(z) / 2 and the staged value e = exp(.) / sqrt(z) lies in (0, 1] -- no logarithm per value in
//      C1, no maximum pass and no exponential in C2 (K loads and adds, one logarithm per configuration).  A sum that underflows,
//      a NaN and the reference's "b-bar = 0" corner fall back to the log-domain evaluation of that (configuration, pair).
template <int PPW, bool LIN>
__global__ void __launch_bounds__(FA_THREADS, FA_MINB) fast_pair_all_kernel(const DevParams *__restrict__ prm_, const FastParams *__restrict__ fp_,
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
  const long long C = prm.C;
  const int sst = (3 * S) | 1;
  double *st = fsm;                                          // [4][sst]  b, v, t per subgroup
  unsigned long long *hasm = (unsigned long long *)(st + (size_t)4 * sst); // [4]
  long long *s_pair = (long long *)(hasm + 4);               // [4] output pair index
  double *genavg = (double *)(s_pair + 4);                   // [4] grid average of the consistent configuration
  double *part = genavg + 4;                                 // [FA_WARPS][4][3] partial BMA states (m, acc, poisoned)
  double *mcT = part + FA_WARPS * 12;                        // [4][32] LIN: sum of t^2 / 2 over every subset of every part
  char *after = reinterpret_cast<char *>(fsm) + fa_tile_doubles(S) * 8;
  unsigned short *s_mask = reinterpret_cast<unsigned short *>(after); // [C] + zero padding
  double *tab = reinterpret_cast<double *>(after + fa_mask_bytes(C)); // [ne][3][32]
  const FaParts P = fa_parts(S);
  double *stg = tab + (size_t)P.ne * 3 * 32 + (size_t)warp * FA_CH * FA_SROW; // this warp's [FA_CH][FA_SROW] staging tile

  for (long long c = threadIdx.x; c < (long long)(fa_mask_bytes(C) / 2); c += FA_THREADS)
    s_mask[c] = (c < C) ? (unsigned short)prm.cfg_mask[c] : (unsigned short)0; // (the padding is read by the last step)

  // lane = (pair jl, grid point k) in the table and C1 phases
  const int jl = lane / K, k = lane - jl * K;
  const double oma2 = gc.omaS[k], phi2 = gc.phiS[k], hom2 = 0.5 * oma2;
  const double invK = 1.0 / (double)K, wL = 1.0 / (double)L;
  const int sh1 = P.np[0], sh2 = P.np[0] + P.np[1];
  const int mk0 = (1 << P.np[0]) - 1, mk1 = (1 << P.np[1]) - 1;
  const double *tb0 = tab + lane, *tb1 = tab + (size_t)P.off[1] * 96 + lane, *tb2 = tab + (size_t)P.off[2] * 96 + lane;
  const int nchunk = (int)((C + FA_CH - 1) / FA_CH);
  const long long n_tiles = (fa.n_pairs - fa.q_begin + PPW - 1) / PPW;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long q0 = fa.q_begin + tile * PPW;
    const int tn = (int)min((long long)PPW, fa.n_pairs - q0); // 1 .. PPW
    __syncthreads(); // the previous tile is finished (first pass: masks and math tables are in place)
    if (threadIdx.x < tn) {
      const long long q = q0 + threadIdx.x;
      int lo = 0, hi = fa.n_genes - 1; // the last gene whose first compact pair index is <= q (empty genes share a base)
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (fa.fast_base[mid] <= q) lo = mid;
        else hi = mid - 1;
      }
      const long long pair = fa.pair_off[lo] + (q - fa.fast_base[lo]);
      s_pair[threadIdx.x] = pair;
      hasm[threadIdx.x] = fa.has_all[q];
      genavg[threadIdx.x] = fa.out_w[pair * (5 + C)]; // (first pass)
    }
    for (int i = threadIdx.x; i < tn * 3 * S; i += FA_THREADS) {
      const int j = i / (3 * S);
      st[(size_t)j * sst + (i - j * 3 * S)] = fa.st_all[q0 * 3 * S + i];
    }
    __syncthreads();
    // ---------------- T: subset-sum tables, one warp per part, lane-private columns
    const bool active = jl < tn;
    const int jme = active ? jl : 0;
    for (int p = warp; p < 3; p += FA_WARPS) {
      const double *stj = st + (size_t)jme * sst;
      const unsigned long long has = active ? hasm[jme] : 0ull;
      double *tp = tab + (size_t)P.off[p] * 96 + lane;
      tp[0] = 0.0;
      tp[32] = 0.0;
      tp[64] = 0.0;
      for (int i = 0; i < P.np[p]; ++i) {
        const int s = P.start[p] + i;
        double d = 0.0, bd = 0.0, A = 0.0;
        const double b = stj[s], v = stj[S + s], tt = stj[2 * S + s];
        if (((has >> s) & 1ull) && !(fabs(tt) < 1e-8)) { // (gene_snp_pair.cpp:314: |t| < 1e-8 contributes nothing)
          const double inv = rcp_n(v + phi2);
          d = inv;
          bd = b * inv;
          if (LIN) // ln ABF_s - t^2 / 2 = ln(v / (v + phi2)) / 2 - t^2 v / (2 (v + phi2)), formed without cancellation
            A = (phi2 == 0.0) ? -0.5 * tt * tt : fma(0.5, log_tab16(v * inv, T), -0.5 * tt * tt * v * inv);
          else
            A = (phi2 == 0.0) ? 0.0 : fma(0.5, log_tab16(v * inv, T), 0.5 * tt * tt * phi2 * inv);
        }
        double *te = tp + (size_t)(1 << i) * 96;
        te[0] = d;
        te[32] = bd;
        te[64] = A;
      }
      for (int a = 3; a < (1 << P.np[p]); ++a) {
        const int lo = a & (a - 1);
        if (lo == 0) continue; // single subgroup: written above
        const double *e0 = tp + (size_t)lo * 96, *e1 = tp + (size_t)(a & -a) * 96;
        double *ea = tp + (size_t)a * 96;
        ea[0] = e0[0] + e1[0];
        ea[32] = e0[32] + e1[32];
        ea[64] = e0[64] + e1[64];
      }
    }
    if (LIN && warp == FA_WARPS - 1 && lane < 3 * PPW) {
      // bounds of the subsets: lane = (pair, part), 2^np entries each (warp 3 has no table part to build)
      const int jj = lane / 3, p = lane - jj * 3;
      double *mc = mcT + jj * 32 + P.off[p];
      const double *stj = st + (size_t)min(jj, tn - 1) * sst;
      const unsigned long long has = (jj < tn) ? hasm[jj] : 0ull;
      mc[0] = 0.0;
      for (int a = 1; a < (1 << P.np[p]); ++a) {
        const int i = __ffs(a) - 1, s = P.start[p] + i;
        const double tt = stj[2 * S + s];
        const bool on = ((has >> s) & 1ull) && !(fabs(tt) < 1e-8);
        mc[a] = mc[a & (a - 1)] + (on ? 0.5 * tt * tt : 0.0);
      }
    }
    __syncthreads();
    // ---------------- C: chunks of 32 configurations, round-robin over the warps
    double *oc = nullptr; // this lane's raw output column: out_cfg[pair][c][k]
    if (fa.out_cfg && active) oc = fa.out_cfg + s_pair[jme] * C * K + k;
    LseTab bma[PPW];
#pragma unroll
    for (int jj = 0; jj < PPW; ++jj) bma[jj].init();
    for (int ch = warp; ch < nchunk; ch += FA_WARPS) {
      const long long c0 = (long long)ch * FA_CH;
      const int nc = (int)min((long long)FA_CH, C - c0);
      // ---- C1 (FA_U configurations per step, independent straight-line chains)
      const unsigned short *mk = s_mask + c0;
      for (int cl0 = 0; cl0 < nc; cl0 += FA_U) {
        double xv[FA_U];
#pragma unroll
        for (int u = 0; u < FA_U; ++u) {
          const int m = mk[cl0 + u]; // (zero-padded past C)
          const double *p0 = tb0 + (m & mk0) * 96, *p1 = tb1 + ((m >> sh1) & mk1) * 96, *p2 = tb2 + (m >> sh2) * 96;
          const double den = p0[0] + p1[0] + p2[0], num = p0[32] + p1[32] + p2[32], sing = p0[64] + p1[64] + p2[64];
          // natural-log ABF; CalcLog10AbfUvlr's guards (see abf_from_sums) as ONE select on the result: num != 0 implies
          // den != 0 (sums of the same terms), a NaN or infinite den fails z < 1e300 like the reference's "V < +Inf"; the
          // logarithm and the reciprocal of a rejected z are bit manipulations on garbage, never used
          const double z = fma(oma2, den, 1.0);
          if (LIN) {
            // e = exp(x - Mc): a configuration without any active subgroup gives exp(0) / sqrt(1) = 1 by itself; the corners the
            // reference maps to x = 0 although subgroups are active (b-bar = 0, V not finite) and NaN statistics are marked NaN
            const double r = rsqrt_newton1(z);
            const double q = num * r;
            const double e = exp_tab16<true>(fma(q * q, hom2, sing), T) * r;
            const bool plain = (num != 0.0 || den == 0.0) && z < 1e300 && sing == sing;
            xv[u] = plain ? e : nan("");
          } else {
            const bool ok = num != 0.0 && z < 1e300;
            const double x = sing + fma(-0.5, log_tab16_pos(z, T), num * num * (hom2 * rcp_n(z)));
            xv[u] = ok ? x : 0.0;
          }
        }
#pragma unroll
        for (int u = 0; u < FA_U; ++u) {
          const int cl = cl0 + u;
          stg[cl * FA_SROW + lane] = xv[u]; // natural-log units (a row >= nc of the last chunk belongs to a zero-padded mask: unused)
          if (!LIN && oc && cl < nc) oc[(c0 + cl) * K] = xv[u] * EQB_INV_LN10;
        }
      }
      __syncwarp();
      // ---- C2: lane = configuration c0 + lane; the pairs side by side (PPW independent max / sum chains)
      const long long c = c0 + lane;
      const bool valid = lane < nc;
      const double cwt = valid ? prm.cfg_weight[c] : 0.0;
      const double *row = stg + lane * FA_SROW;
      double x0[PPW], mx[PPW], sum[PPW];
      if (LIN) {
        // the bound of this lane's configuration for every pair, then K loads and adds per pair
        const int m = s_mask[valid ? c : 0];
        const int i0 = m & mk0, i1 = P.off[1] + ((m >> sh1) & mk1), i2 = P.off[2] + (m >> sh2);
#pragma unroll
        for (int jj = 0; jj < PPW; ++jj) {
          const double *mc = mcT + jj * 32;
          mx[jj] = mc[i0] + mc[i1] + mc[i2];
          x0[jj] = 0.0;
          sum[jj] = 0.0;
        }
        for (int kk = 0; kk < K; ++kk) {
#pragma unroll
          for (int jj = 0; jj < PPW; ++jj) sum[jj] += row[jj * K + kk];
        }
#pragma unroll
        for (int jj = 0; jj < PPW; ++jj) {
          if (!(sum[jj] > 1e-280)) { // NaN marker or underflow (rare): the log-domain evaluation of the configuration, as below
            double xm = 0.0, x00 = 0.0;
            for (int pass = 0; pass < 2; ++pass) {
              double acc = 0.0;
              for (int kk = 0; kk < K; ++kk) {
                const int col = jj * K + kk;
                const double *p0 = tab + (size_t)i0 * 96 + col, *p1 = tab + (size_t)i1 * 96 + col, *p2 = tab + (size_t)i2 * 96 + col;
                const double den = p0[0] + p1[0] + p2[0], num = p0[32] + p1[32] + p2[32];
                const double sing = p0[64] + p1[64] + p2[64] + mx[jj]; // (the tabulated terms carry - t^2 / 2)
                const double oma2k = gc.omaS[kk];
                const double z = fma(oma2k, den, 1.0);
                const bool ok = num != 0.0 && z < 1e300;
                double x = sing + fma(-0.5, log_tab16(ok ? z : 1.0, T), 0.5 * num * num * oma2k * rcp_n(ok ? z : 1.0));
                x = ok ? x : 0.0;
                if (pass == 0) {
                  if (kk == 0) x00 = xm = x;
                  else xm = fmax(xm, x);
                } else
                  acc += (x == x) ? exp_tab16<true>(x - xm, T) : 0.0;
              }
              if (pass == 1) sum[jj] = acc;
            }
            // (from here on as in the log-domain kernel: maximum xm instead of the bound, poisoned by a NaN first value)
            mx[jj] = xm;
            x0[jj] = x00;
          }
        }
      } else {
#pragma unroll
      for (int jj = 0; jj < PPW; ++jj) {
        x0[jj] = mx[jj] = row[jj * K];
        sum[jj] = 0.0;
      }
      for (int kk = 1; kk < K; ++kk) {
#pragma unroll
        for (int jj = 0; jj < PPW; ++jj) mx[jj] = fmax(mx[jj], row[jj * K + kk]); // (fmax drops a NaN operand; mx is NaN only if x0 is)
      }
      for (int kk = 0; kk < K; ++kk) {
#pragma unroll
        for (int jj = 0; jj < PPW; ++jj) {
          const double v = row[jj * K + kk];
          const double e = exp_tab16<FA_CLAMP>(v - mx[jj], T);
          sum[jj] += (v == v) ? e : 0.0;
        }
      }
      }
#pragma unroll
      for (int jj = 0; jj < PPW; ++jj) {
        // 0 < mean <= 1 (log domain: the maximum contributes 1, so 1/K <= mean); K equal values give 1 and must give exactly the
        // common value (the table logarithm is good to 1e-13 absolute, the reference prints 0 after its DBL_EPSILON snap)
        const double mean = sum[jj] * invK, d1 = mean - 1.0; // (ln(1 + d) = d next to 1)
        double wj = (mx[jj] + ((fabs(d1) < 1e-8) ? d1 : log_tab16_pos(mean, T))) * EQB_INV_LN10;
        if (fabs(wj) <= DBL_EPSILON) wj = 0.0;
        if (x0[jj] != x0[jj]) wj = nan("");
        if (valid) bma[jj].add(wj, cwt, c == 0, T); // CalcBMA (gene_snp_pair.cpp:572-602)
        double *ow = fa.out_w + s_pair[jj] * (5 + C);
        if (valid && jj < tn) ow[5 + c] = wj;
        if (ch == 0) { // warp-uniform (warp 0 only)
          // CalcBMAlite (gene_snp_pair.cpp:552-570): the S singleton averages (0.5 / S each), then the consistent one (0.5)
          const double lite = warp_lws(lane < S ? wj : genavg[jj], lane < S ? 0.5 / (double)S : 0.5, S + 1, lane, T);
          if (lane == 0 && jj < tn) ow[3] = lite;
        }
      }
      __syncwarp();
    }
    // ---------------- M: the lanes' partial model averages (online log-sum-exp states) merged across the warp, then
    // across the warps through shared memory
#pragma unroll
    for (int jj = 0; jj < PPW; ++jj) {
      LseTab b = bma[jj];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double m2 = __shfl_xor_sync(0xffffffffu, b.m, o), a2 = __shfl_xor_sync(0xffffffffu, b.acc, o);
        const int p2 = __shfl_xor_sync(0xffffffffu, (int)b.poisoned, o);
        lse_merge(b, m2, a2, T);
        b.poisoned = b.poisoned || p2;
      }
      if (lane == 0) {
        part[(warp * 4 + jj) * 3] = b.m;
        part[(warp * 4 + jj) * 3 + 1] = b.acc;
        part[(warp * 4 + jj) * 3 + 2] = b.poisoned ? 1.0 : 0.0;
      }
    }
    __syncthreads();
    if (threadIdx.x < tn) {
      const int jj = threadIdx.x;
      LseTab b;
      b.init();
      for (int w = 0; w < FA_WARPS; ++w) {
        lse_merge(b, part[(w * 4 + jj) * 3], part[(w * 4 + jj) * 3 + 1], T);
        b.poisoned = b.poisoned || part[(w * 4 + jj) * 3 + 2] != 0.0;
      }
      fa.out_w[s_pair[jj] * (5 + C) + 4] = b.result(T);
    }
  }
}

} // namespace eqb
