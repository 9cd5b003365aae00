// hm_kernels.cuh -- kernels of the EM of the hierarchical model (eqtlbma_hm --model configs; include/eqtlbma_hm_b200.h).
//
// What the reference computes per EM iteration (src/eqtlbma_hm.cpp:659-868, src/hm_methods.cpp:391-672), with
// B[p][k][l] the raw log10 BF of pair p, configuration k, grid point l, eta the configuration prior, lambda the grid
// weights and uniform SNP priors 1/m_g:
//     a[p][k]   = log10 sum_l lambda_l 10^B[p][k][l]                       snp_eQTL::em_update_config
//     A[g][k]   = log10 (1/m_g) sum_p 10^a[p][k]                           gene_eQTL::em_update_config
//     Gd[g][l]  = log10 (1/m_g) sum_p sum_k eta_k 10^B[p][k][l]            gene_eQTL::em_update_grid
//     BF[g]     = log10 (1/m_g) sum_p sum_k eta_k sum_l lambda_l 10^B      gene_eQTL::compute_log10_BF
//               = log10 sum_l lambda_l 10^Gd[g][l]
// i.e. three groupings of ONE pass over the rows (p, k) of the data.  The reference makes that pass 2 + dim + grid times
// per iteration through nested log10_weighted_sum calls (utils_math.cpp:135-159); hm_estep_kernel makes it once:
//
//   hm_estep_kernel   persistent WARPS (as many CTAs of 4 warps as are resident; the warps of a CTA share only the tables), each
//                     claiming work units from a counter (a unit = a run of whole pairs of one gene, ~1024 rows; the first
//                     rounds of the next unit are in flight while the current one is merged).  A warp takes rounds of 32
//                     consecutive rows (contiguous 32 x grid doubles), staged in shared memory by 1-D TMA bulk copies
//                     (cp.async.bulk + mbarrier, `stages` deep per warp, issued by lane 0); lane = row: row maximum, one
//                     table exponential per element, the row average a (optionally stored), an online log-sum-exp state per
//                     configuration in shared memory, and lane-private linear-domain column sums against a running reference
//                     (the largest row maximum seen), merged over the lanes at the end -> U[unit][dim + grid] (log10 domain).
//                     HBM-bound by design (8 bytes per element read once) with (grid + 2) exponentials per row on the FP64
//                     pipe: the two limits are within 2x of each other at grid = 10.
//   hm_gene_kernel    per gene: log-sum-exp of its units, - log10 m_g, BF[g]           -> PA[j][g], BF[g]
//   hm_lik_kernel     per-gene log10(pi0 + (1 - pi0) BF), fixed-order sum, optional keep
//   hm_sums_kernel    one CTA per output: log-sum-exp over genes of PA[j][g] - lik_g (kept), and sum_g pi0 / 10^lik_g
//   hm_snp_kernel     warp per pair: log10 sum_k eta_k 10^a[p][k] (posterior pass)
//   hm_check_kernel   counts non-finite values at load time
// All reductions have a fixed order (no atomics on data): results do not depend on scheduling.
#pragma once

#include "table_math.cuh"

namespace eqb {

constexpr int HM_WARPS = 4;
constexpr int HM_THREADS = HM_WARPS * 32;
#ifndef HM_MIN_CTAS
#define HM_MIN_CTAS 4
#endif
constexpr int HM_MAXGRID = 32;
constexpr int HM_MAXDIM = 4096;

struct HmArgs {
  const char *B;              // first byte of the data; 16-byte aligned, 16 readable bytes before and after the array
  const long long *unit_row0; // [n_units] first row (pair * dim) of the unit
  const int *unit_rows;       // [n_units] rows of the unit (whole pairs)
  double *U;                  // [n_units][dim + grid] partial log-sum-exps
  const double *cfg;          // [dim] configuration prior
  double *rowA;               // [rows] a[p][k] (posterior pass) or nullptr
  int dim, grid;
  int n_units;   // work units, claimed one at a time by the persistent warps
  int *counter;  // next unclaimed unit (zeroed before every launch)
  int rpr;         // rows per round: 32 (dim >= 32) or (32 / dim) * dim
  int nslot;       // configuration-state slots per warp: dim (dim >= 32) or rpr
  int stages;      // TMA stages per warp
  int stage_bytes; // bytes of one stage (rpr rows + 16, multiple of 16)
  double gw[HM_MAXGRID]; // grid weights: constant-bank operands
};

// 10^y with the base folded into the constants (no y * ln 10 first): 2^(k/16) from the table times 10^g - 1 to g^4,
// |g| <= log10(2)/32, relative error < 4e-11 + 7e-14 |y|/300 like exp10_tab16.  FPCLAMP: any y (-inf, NaN -> ~1e-300);
// otherwise y must be finite with |y| < 1e7 (integer clamp of the binary exponent: results below 2^-1000 come out as ~1e-301)
static __constant__ double HMK[8] = {
    53.150849518197795,    // 0  16 log2(10)
    -0.018814374728998825, // 1  -log10(2) / 16
    2.302585092994046,     // 2  ln 10
    2.650949055239199,     // 3  ln^2 10 / 2
    2.034678592293476,     // 4  ln^3 10 / 6
    1.171255148912267,     // 5  ln^4 10 / 24
    3.321928094887362,     // 6  log2(10)
    0.30102999566398120};  // 7  log10(2)
template <bool FPCLAMP>
__device__ __forceinline__ double hm_exp10(double y, const TabRef T)
{
  if (FPCLAMP) {
    const unsigned int hy = (unsigned int)__double2hiint(y);
    if (hy > 0xC072C000u) y = -300.0; // max(y, -300) on the bit pattern (NaN and -inf are larger)
  }
  const double tm = fma(y, HMK[0], PGK[1]);
  const int k = __double2loint(tm);
  const double kd = tm - PGK[1];
  const double g = fma(kd, HMK[1], y);
  double s = fma(g, HMK[5], HMK[4]);
  s = fma(g, s, HMK[3]);
  s = fma(g, s, HMK[2]);
  const double tj = T.exp16(k & 15);
  const double v = fma(tj * g, s, tj);
  const int e2 = FPCLAMP ? (k >> 4) : max(k >> 4, -1000);
  // the binary exponent is added on the 64-bit pattern (one integer add on the high word, no register shuffling)
  return __longlong_as_double(__double_as_longlong(v) + ((long long)e2 << 52));
}

// N independent exponentials evaluated stage by stage (all N first FMAs, then all N subtractions, ...): the dependency
// chain of one evaluation is eight FP64 operations long, and the instruction scheduler interleaves only about two of them
// when they are written one after the other.  y finite, |y| < 1e7 (the unclamped form of hm_exp10)
template <int N>
__device__ __forceinline__ void hm_exp10_batch(const double (&y)[N], double (&e)[N], const TabRef T)
{
  double tm[N], g[N], s[N], tj[N];
#pragma unroll
  for (int i = 0; i < N; ++i) tm[i] = fma(y[i], HMK[0], PGK[1]);
#pragma unroll
  for (int i = 0; i < N; ++i) g[i] = fma(tm[i] - PGK[1], HMK[1], y[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) tj[i] = T.exp16(__double2loint(tm[i]) & 15);
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = fma(g[i], HMK[5], HMK[4]);
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = fma(g[i], s[i], HMK[3]);
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = fma(g[i], s[i], HMK[2]);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double v = fma(tj[i] * g[i], s[i], tj[i]);
    const int e2 = max(__double2loint(tm[i]) >> 4, -1000);
    e[i] = __longlong_as_double(__double_as_longlong(v) + ((long long)e2 << 52));
  }
}

// 2^d for an integer d <= 0 (exact; 0 below the normal range)
__device__ __forceinline__ double hm_pow2i(int d) { return (d < -1022) ? 0.0 : __hiloint2double((1023 + d) << 20, 0); }
// the same for an integer-valued double d <= 0 (or -inf)
__device__ __forceinline__ double hm_pow2d(double d) { return (d < -1022.0) ? 0.0 : __hiloint2double((1023 + (int)d) << 20, 0); }

__device__ __forceinline__ void hm_mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void hm_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hm_mbar_wait(uint32_t bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "HM_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra HM_DONE;\n"
      "bra HM_WAIT;\n"
      "HM_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// 1-D TMA: `bytes` (multiple of 16) from global `src` (16-byte aligned) to shared `dst`, completion on `bar`
__device__ __forceinline__ void hm_bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ double hm_warp_max(double v)
{
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double hm_warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__host__ __device__ inline size_t hm_align16(size_t x) { return (x + 15) & ~(size_t)15; }
// shared memory of hm_estep_kernel: tables | barriers | configuration states | stages
__host__ __device__ inline size_t hm_smem_bytes(int nslot, int stages, int stage_bytes)
{
  return hm_align16(sizeof(BfTabs)) + hm_align16((size_t)HM_WARPS * stages * 8) + (size_t)HM_WARPS * nslot * 16 +
         (size_t)HM_WARPS * stages * stage_bytes;
}

// G: compile-time number of grid points (EXACT) or their upper bound (the loops are predicated on l < grid)
// RANGED: every value of the data set is within +-1e6 (checked at load time), so differences of two values can take the
// exponential without the floating-point clamp
template <int G, bool EXACT, bool RANGED>
__global__ void __launch_bounds__(HM_THREADS, (G <= 16) ? HM_MIN_CTAS : 3) hm_estep_kernel(const __grid_constant__ HmArgs a)
{
  extern __shared__ __align__(16) unsigned char hm_smem[];
  const int grid = EXACT ? G : a.grid;
  const int dim = a.dim, rpr = a.rpr, nslot = a.nslot, stages = a.stages, stage_bytes = a.stage_bytes;
  BfTabs *tabs = reinterpret_cast<BfTabs *>(hm_smem);
  unsigned char *p_bar = hm_smem + hm_align16(sizeof(BfTabs));
  double2 *kst_all = reinterpret_cast<double2 *>(p_bar + hm_align16((size_t)HM_WARPS * stages * 8));
  unsigned char *stage_all = reinterpret_cast<unsigned char *>(kst_all + (size_t)HM_WARPS * nslot);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bf_tabs_init(*tabs);
  const TabRef T{smem_u32(tabs)};
  const uint32_t bars = smem_u32(p_bar) + (uint32_t)(warp * stages * 8);
  double2 *kst = kst_all + (size_t)warp * nslot;
  unsigned char *stg = stage_all + (size_t)warp * stages * stage_bytes;
  const uint32_t stg_u32 = smem_u32(stg);
  for (int i = lane; i < nslot; i += 32) kst[i] = make_double2(-INFINITY, 0.0);
  if (lane == 0) {
    for (int s = 0; s < stages; ++s) hm_mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads(); // the only CTA-wide barrier: the warps share nothing but the tables from here on

  // lane-private column sums against the running reference mref
  double cs[G];
  const bool small_dim = dim < 32;
  const double cfg_fixed = small_dim ? ((lane < rpr) ? __ldg(a.cfg + lane % dim) : 0.0) : 0.0;
  const long long row_bytes = (long long)grid * 8;
  const int q = small_dim ? rpr / dim : 1;

  // round i of the unit starting at row0 into stage st (lane 0)
  auto issue = [&](long long row0, int nrows, int i, uint32_t st) {
    const int r_first = i * rpr;
    const int nr = min(rpr, nrows - r_first);
    const long long off = (row0 + r_first) * row_bytes;
    const int shift = (int)(off & 15);
    const uint32_t bytes = (uint32_t)((shift + nr * (int)row_bytes + 15) & ~15);
    hm_mbar_expect_tx(bars + 8 * st, bytes);
    hm_bulk_load(stg_u32 + st * (uint32_t)stage_bytes, a.B + (off - shift), bytes, bars + 8 * st);
  };
  auto next_unit = [&]() { // dynamic assignment: one counter per launch, claimed by lane 0
    int u = 0;
    if (lane == 0) u = atomicAdd(a.counter, 1);
    return __shfl_sync(0xffffffffu, u, 0);
  };

  // persistent warps: the pipeline slots of a warp are used consecutively across its units, so the stage index st_cur cycles
  // through 0 .. stages - 1 and the mbarrier parity ph_cur flips at every wrap, throughout; the first rounds of the next unit
  // are issued before the current one is merged
  uint32_t st_cur = 0, ph_cur = 0;
  auto issue_first = [&](long long row0, int nrows, int n_rounds) { // the first rounds of a unit, from the current stage on
    uint32_t st = st_cur;
    for (int i = 0; i < stages && i < n_rounds; ++i) {
      issue(row0, nrows, i, st);
      st = (st + 1 == (uint32_t)stages) ? 0u : st + 1;
    }
  };
  int unit = next_unit();
  long long row0 = 0;
  int nrows = 0, n_rounds = 0;
  if (unit < a.n_units) {
    row0 = a.unit_row0[unit];
    nrows = a.unit_rows[unit];
    n_rounds = (nrows + rpr - 1) / rpr;
    if (lane == 0) issue_first(row0, nrows, n_rounds);
  }
  while (unit < a.n_units) {
    const int nxt = next_unit();
#pragma unroll
    for (int l = 0; l < G; ++l) cs[l] = 0.0;
    double mref = -INFINITY;   // (clamped variant) running reference of the column sums, log10 domain
    int eref = -(1 << 28);     // (RANGED variant) the same as a binary exponent: sums are relative to 2^eref

    for (int i = 0; i < n_rounds; ++i) {
      const uint32_t st = st_cur;
      hm_mbar_wait(bars + 8 * st, ph_cur);
      if (++st_cur == (uint32_t)stages) {
        st_cur = 0;
        ph_cur ^= 1u;
      }
      const int r_first = i * rpr;
      const int nr = min(rpr, nrows - r_first);
      if (lane < nr) {
        const long long off = (row0 + r_first) * row_bytes;
        const double *x = reinterpret_cast<const double *>(stg + (size_t)st * stage_bytes + (int)(off & 15)) + (size_t)lane * grid;
        // even compile-time grids: a row starts on a 16-byte boundary (row_bytes is a multiple of 16) and is read ONCE with
        // 128-bit loads (conflict-free at a lane stride of 80 bytes, where 64-bit loads conflict two ways); otherwise the two
        // passes below read the row from shared memory
        constexpr bool CACHE = EXACT && (G % 2 == 0) && G <= 16;
        double xr[CACHE ? G : 1];
        if constexpr (!CACHE) xr[0] = 0.0; // (never read: HM_X selects x[l])
        if constexpr (CACHE) {
          const double2 *x2 = reinterpret_cast<const double2 *>(x);
#pragma unroll
          for (int l = 0; l < G / 2; ++l) {
            const double2 v = x2[l];
            xr[2 * l] = v.x;
            xr[2 * l + 1] = v.y;
          }
        }
#define HM_X(l) (CACHE ? xr[CACHE ? (l) : 0] : x[l])
        double m = HM_X(0); // (the data are finite, checked at load: a compare-and-select maximum, without fmax()'s NaN handling)
#pragma unroll
        for (int l = 1; l < G; ++l)
          if (EXACT || l < grid) {
            const double v = HM_X(l);
            m = (v > m) ? v : m;
          }
        const int k = small_dim ? 0 : (int)((r_first + lane) % dim);
        const double cfgk = small_dim ? cfg_fixed : __ldg(a.cfg + k);
        const int slot = small_dim ? lane : k;
        double2 s = kst[slot];
        if constexpr (RANGED) {
          // Everything relative to a power of two: E = ceil(m log2 10), so 10^(x - mq) <= 1 with mq = E log10 2, and changes of
          // reference (column sums, per-configuration sums) are exact exponent arithmetic -- no exponential for them and no
          // logarithm per row (the row average a = mq + log10(rs) is only formed when the posterior pass stores it)
          const int E = __double2int_ru(m * HMK[6]);
          const double mq = (double)E * HMK[7];
          const int dE = E - eref;
          double wc;
          if (dE > 0) {
            const double sc = hm_pow2i(-dE);
#pragma unroll
            for (int l = 0; l < G; ++l) cs[l] *= sc;
            eref = E;
            wc = cfgk;
          } else
            wc = cfgk * hm_pow2i(dE);
          double rs = 0.0;
#ifndef HM_EXP_BATCH
#define HM_EXP_BATCH 5
#endif
          if constexpr (CACHE && HM_EXP_BATCH > 1 && (G % HM_EXP_BATCH) == 0) {
#pragma unroll
            for (int l0 = 0; l0 < G; l0 += HM_EXP_BATCH) {
              double yb[HM_EXP_BATCH], eb[HM_EXP_BATCH];
#pragma unroll
              for (int i = 0; i < HM_EXP_BATCH; ++i) yb[i] = xr[l0 + i] - mq;
              hm_exp10_batch<HM_EXP_BATCH>(yb, eb, T);
#pragma unroll
              for (int i = 0; i < HM_EXP_BATCH; ++i) {
                rs = fma(a.gw[l0 + i], eb[i], rs);
                cs[l0 + i] = fma(wc, eb[i], cs[l0 + i]);
              }
            }
          } else {
#pragma unroll
            for (int l = 0; l < G; ++l)
              if (EXACT || l < grid) {
                const double e = hm_exp10<false>(HM_X(l) - mq, T);
                rs = fma(a.gw[l], e, rs);
                cs[l] = fma(wc, e, cs[l]);
              }
          }
          if (a.rowA) a.rowA[row0 + r_first + lane] = fma(log_tab16(rs, T), PGK[11], mq);
          // sum over the rows of this configuration of rs 2^E: state = (exponent as a double, sum relative to it)
          if (!(rs >= 0.0))
            s = make_double2(nan(""), nan("")); // (negative weights of a SQUAREM proposal: the reference's log10 is NaN)
          else if (rs > 0.0 && s.x == s.x) {
            const double Ed = (double)E, dd = Ed - s.x; // +inf for the first row
            if (dd > 0.0) {
              s.y = fma(s.y, hm_pow2d(-dd), rs);
              s.x = Ed;
            } else
              s.y = fma(rs, hm_pow2d(dd), s.y);
          }
        } else {
          // column reference: rescale the sums when this row raises it
          const double d = m - mref; // +inf for the first row
          double wc;
          if (d > 0.0) {
            const double sc = hm_exp10<true>(-d, T);
#pragma unroll
            for (int l = 0; l < G; ++l) cs[l] *= sc;
            mref = m;
            wc = cfgk;
          } else
            wc = cfgk * hm_exp10<true>(d, T);
          double rs = 0.0;
#pragma unroll
          for (int l = 0; l < G; ++l)
            if (EXACT || l < grid) {
              const double e = hm_exp10<true>(HM_X(l) - m, T);
              rs = fma(a.gw[l], e, rs);
              cs[l] = fma(wc, e, cs[l]);
            }
          const double ar = fma(log_tab16(rs, T), PGK[11], m); // a[p][k]; -inf for a zero sum, NaN for a negative one
          if (a.rowA) a.rowA[row0 + r_first + lane] = ar;
          // online log-sum-exp of a over the rows with this configuration
          if (ar != ar)
            s = make_double2(ar, ar); // (negative weights of a SQUAREM proposal: the likelihood is NaN, as in the reference)
          else if (ar > -INFINITY) {
            const double dd = ar - s.x;
            const bool up = dd > 0.0;
            const double e = hm_exp10<true>(up ? -dd : dd, T);
            s.y = up ? fma(s.y, e, 1.0) : s.y + e;
            s.x = up ? ar : s.x;
          }
        }
        kst[slot] = s;
#undef HM_X
      }
      __syncwarp();
      if (lane == 0 && i + stages < n_rounds) issue(row0, nrows, i + stages, st); // the stage just consumed
    }

    // ---- next unit: its first rounds travel while this one is merged (every stage of the warp is free here)
    const int cur = unit;
    unit = nxt;
    if (unit < a.n_units) {
      row0 = a.unit_row0[unit];
      nrows = a.unit_rows[unit];
      n_rounds = (nrows + rpr - 1) / rpr;
      if (lane == 0) issue_first(row0, nrows, n_rounds);
    }

    // ---- merge over the lanes: columns (lane l keeps column l) ...
    double *Uu = a.U + (size_t)cur * (dim + grid);
    if constexpr (RANGED) {
      int EM = eref;
#pragma unroll
      for (int o = 16; o; o >>= 1) EM = max(EM, __shfl_xor_sync(0xffffffffu, EM, o));
      const double fac = hm_pow2i(eref - EM); // (0 for a lane that saw no row)
      double mine = 0.0;
#pragma unroll
      for (int l = 0; l < G; ++l)
        if (EXACT || l < grid) {
          const double v = hm_warp_sum(cs[l] * fac);
          if (lane == l) mine = v;
        }
      if (lane < grid) Uu[dim + lane] = (mine == 0.0) ? -INFINITY : fma((double)EM, HMK[7], log10(mine)); // (NaN sums stay NaN)
    } else {
      const double M = hm_warp_max(mref);
      const double fac = (mref > -INFINITY) ? hm_exp10<true>(mref - M, T) : 0.0;
      double mine = 0.0;
#pragma unroll
      for (int l = 0; l < G; ++l)
        if (EXACT || l < grid) {
          const double v = hm_warp_sum(cs[l] * fac);
          if (lane == l) mine = v;
        }
      if (lane < grid) Uu[dim + lane] = (mine == 0.0) ? -INFINITY : M + log10(mine); // (NaN sums stay NaN)
    }
    // ... and configurations (the q states of a configuration, written by q lanes of this warp)
    for (int k = lane; k < dim; k += 32) {
      double MM = -INFINITY;
      bool bad = false;
      for (int j = 0; j < q; ++j) {
        const double2 s = kst[j * dim + k];
        bad = bad || (s.x != s.x);
        MM = fmax(MM, s.x);
      }
      double sum = 0.0;
      for (int j = 0; j < q; ++j) {
        const double2 s = kst[j * dim + k];
        if (s.x > -INFINITY) sum += RANGED ? s.y * hm_pow2d(s.x - MM) : s.y * exp10(s.x - MM);
      }
      if (RANGED)
        Uu[k] = bad ? nan("") : ((sum == 0.0) ? -INFINITY : fma(MM, HMK[7], log10(sum)));
      else
        Uu[k] = bad ? nan("") : ((sum == 0.0) ? -INFINITY : MM + log10(sum));
    }
    __syncwarp();
    for (int i = lane; i < nslot; i += 32) kst[i] = make_double2(-INFINITY, 0.0);
    __syncwarp();
  }
}

// per gene: merge its units, subtract log10 m_g; PA is output-major [dim + grid][n_genes].  One WARP per gene (lane = output
// j, j + 32, ...), HM_GENE_WARPS genes per CTA: most genes have one or two units and dim + grid <= 32 outputs
constexpr int HM_GENE_WARPS = 8;
__global__ void __launch_bounds__(HM_GENE_WARPS * 32) hm_gene_kernel(const double *__restrict__ U, const long long *__restrict__ gene_unit0,
                                                                    const long long *__restrict__ gene_off, int dim, int grid, long long n_genes,
                                                                    const double *__restrict__ gw, double *__restrict__ PA, double *__restrict__ BF)
{
  const long long g = (long long)blockIdx.x * HM_GENE_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (g >= n_genes) return;
  const long long u0 = gene_unit0[g], u1 = gene_unit0[g + 1];
  const double l10m = log10((double)(gene_off[g + 1] - gene_off[g]));
  const int nout = dim + grid;
  // BF_g = log10 sum_l lambda_l 10^Gd[l]
  double bf_max = -INFINITY;
  bool bf_bad = false;
  for (int j = lane; j < nout; j += 32) {
    double MM = -INFINITY;
    bool bad = false;
    for (long long u = u0; u < u1; ++u) {
      const double v = U[(size_t)u * nout + j];
      bad = bad || (v != v);
      MM = fmax(MM, v);
    }
    double val;
    if (bad)
      val = nan("");
    else if (!(MM > -INFINITY))
      val = -INFINITY;
    else if (u1 - u0 == 1)
      val = MM - l10m;
    else {
      double sum = 0.0;
      for (long long u = u0; u < u1; ++u) sum += exp10(U[(size_t)u * nout + j] - MM);
      val = MM + log10(sum) - l10m;
    }
    PA[(size_t)j * n_genes + g] = val;
    if (j >= dim) {
      bf_bad = bf_bad || (val != val);
      bf_max = fmax(bf_max, val);
    }
  }
  bf_max = hm_warp_max(bf_max);
  bf_bad = __any_sync(0xffffffffu, bf_bad);
  // every lane weighs the grid-point values it produced itself (re-read: its own writes), butterfly sum = fixed order
  double part = 0.0;
  for (int j = lane; j < nout; j += 32)
    if (j >= dim) {
      const double v = PA[(size_t)j * n_genes + g];
      if (v > -INFINITY) part += gw[j - dim] * exp10(v - bf_max);
    }
  part = hm_warp_sum(part);
  if (lane == 0) BF[g] = bf_bad ? nan("") : bf_max + log10(part);
}

// fixed-order block reductions (blockDim.x a power of two <= 1024)
__device__ __forceinline__ double hm_block_sum(double v, double *sh)
{
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int o = blockDim.x >> 1; o; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}
__device__ __forceinline__ double hm_block_max(double v, double *sh)
{
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int o = blockDim.x >> 1; o; o >>= 1) {
    if ((int)threadIdx.x < o) {
      const double b = sh[threadIdx.x + o];
      // NaN-propagating maximum
      sh[threadIdx.x] = (b != b || sh[threadIdx.x] != sh[threadIdx.x]) ? nan("") : fmax(sh[threadIdx.x], b);
    }
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

// gene_eQTL::compute_log10_obs_lik (hm_methods.cpp:476-497) for every gene and their sum: one gene per thread, one partial
// sum per CTA; the CTA that takes the last ticket adds the partials in CTA order (fixed order: no data atomics)
constexpr int HM_LIK_THREADS = 256;
__global__ void __launch_bounds__(HM_LIK_THREADS) hm_lik_kernel(const double *__restrict__ BF, long long n_genes, double pi0, int keep,
                                                                double *__restrict__ kept_lik, double *__restrict__ kept_bf,
                                                                double *__restrict__ partial, unsigned int *__restrict__ ticket,
                                                                double *__restrict__ out)
{
  __shared__ double sh[HM_LIK_THREADS];
  __shared__ bool last;
  const long long g = (long long)blockIdx.x * HM_LIK_THREADS + threadIdx.x;
  double lik = 0.0;
  if (g < n_genes) {
    const double bf = BF[g];
    const double mx = (bf > 0.0) ? bf : 0.0; // max of {0, BF} starting from vec[0] = 0; NaN BF: the sum below is NaN-skipped
    if (bf != bf)
      lik = log10(pi0); // log10_weighted_sum skips a NaN entry that is not the first one (utils_math.cpp:146-150)
    else
      lik = mx + log10(pi0 * exp10(0.0 - mx) + (1.0 - pi0) * exp10(bf - mx));
    if (fabs(lik) <= DBL_EPSILON) lik = 0.0;
    if (keep) {
      kept_lik[g] = lik;
      kept_bf[g] = bf;
    }
  }
  double tot = hm_block_sum(lik, sh);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = tot;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double acc = 0.0;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += HM_LIK_THREADS) acc += __ldcg(partial + i);
  tot = hm_block_sum(acc, sh);
  if (threadIdx.x == 0) {
    out[0] = tot;
    *ticket = 0u; // ready for the next launch (stream order)
  }
}

// CTA j < dim + grid: log10 sum_g 10^(PA[j][g] - lik_g); CTA dim + grid: sum_g 10^(log10 pi0 - lik_g)
constexpr int HM_SUMS_THREADS = 1024;
__global__ void __launch_bounds__(HM_SUMS_THREADS) hm_sums_kernel(const double *__restrict__ PA, const double *__restrict__ kept_lik, long long n_genes,
                                                     int nout, double pi0, double *__restrict__ out)
{
  __shared__ double sh[HM_SUMS_THREADS];
  const int j = blockIdx.x;
  if (j == nout) {
    const double l10pi0 = log10(pi0);
    double acc = 0.0;
    for (long long g = threadIdx.x; g < n_genes; g += blockDim.x) acc += exp10(l10pi0 - kept_lik[g]);
    const double tot = hm_block_sum(acc, sh);
    if (threadIdx.x == 0) out[1] = tot;
    return;
  }
  const double *row = PA + (size_t)j * n_genes;
  double mx = -INFINITY;
  for (long long g = threadIdx.x; g < n_genes; g += blockDim.x) {
    const double v = row[g] - kept_lik[g];
    mx = (v != v || mx != mx) ? nan("") : fmax(mx, v);
  }
  const double MM = hm_block_max(mx, sh);
  double acc = 0.0;
  if (MM > -INFINITY)
    for (long long g = threadIdx.x; g < n_genes; g += blockDim.x) {
      const double v = row[g] - kept_lik[g];
      if (v > -INFINITY) acc += exp10(v - MM);
    }
  const double tot = hm_block_sum(acc, sh);
  if (threadIdx.x == 0) out[2 + j] = (MM != MM) ? nan("") : ((MM > -INFINITY) ? MM + log10(tot) : -INFINITY);
}

// warp per pair: snp_eQTL::compute_log10_BF (hm_methods.cpp:75-128) from the stored row averages
__global__ void __launch_bounds__(128) hm_snp_kernel(const double *__restrict__ rowA, const double *__restrict__ cfg, int dim,
                                                    long long n_pairs, double *__restrict__ snp_bf)
{
  const long long p = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= n_pairs) return;
  const double *ar = rowA + (size_t)p * dim;
  double mx = -INFINITY;
  for (int k = lane; k < dim; k += 32) mx = fmax(mx, ar[k]);
  mx = hm_warp_max(mx);
  double acc = 0.0;
  for (int k = lane; k < dim; k += 32)
    if (ar[k] > -INFINITY) acc += cfg[k] * exp10(ar[k] - mx);
  acc = hm_warp_sum(acc);
  if (lane == 0) {
    double r = mx + log10(acc);
    if (fabs(r) <= DBL_EPSILON) r = 0.0;
    snp_bf[p] = r;
  }
}

// bad[0]: non-finite values; bad[1]: finite values outside +-1e6 (the data set then takes the clamped exponentials)
// ---- multi-GPU exchange over peer memory (NVLink / NVSwitch), one launch per evaluation, no host round trip:
// every rank owns an exchange buffer  flags[2][HM_MAXWORLD] (u64 epochs) | data[2][world][cap] doubles  that its peers map
// through CUDA IPC.  The kernel (one CTA) stores the rank's n partial results into slot [epoch & 1][rank] of EVERY rank's
// buffer (remote stores), fences, publishes the epoch in every rank's flag (release, system scope), waits until all ranks
// have published theirs here (acquire; bounded: ~20 s, then *status = 1), and combines the world x n values in rank order
// into out[] -- every rank computes the same bits from the same values.  Double buffering by epoch parity is enough: a peer
// reaches epoch e + 2 only after this rank has published e + 1, i.e. after it has finished reading e.
constexpr int HM_MAXWORLD = 16;
constexpr int HM_XCHG_DATA_OFF = 2 * HM_MAXWORLD * 8;
struct HmPeers {
  void *base[HM_MAXWORLD];
};
__device__ __forceinline__ double hm_combine_ranks(const double *src, int world, int cap, int j)
{
  if (j < 2) { // plain sums
    double acc = 0.0;
    for (int r = 0; r < world; ++r) acc += __ldcg(src + (size_t)r * cap + j);
    return acc;
  }
  double mx = -INFINITY;
  bool bad = false;
  for (int r = 0; r < world; ++r) {
    const double v = __ldcg(src + (size_t)r * cap + j);
    bad = bad || (v != v);
    mx = fmax(mx, v);
  }
  if (bad) return nan("");
  if (!(mx > -INFINITY) || isinf(mx)) return mx;
  double acc = 0.0;
  for (int r = 0; r < world; ++r) acc += exp10(__ldcg(src + (size_t)r * cap + j) - mx);
  return mx + log10(acc);
}
__global__ void __launch_bounds__(256) hm_xchg_kernel(const HmPeers peers, int world, int rank, int n, int cap, unsigned long long epoch,
                                                      double *__restrict__ out, int *__restrict__ status)
{
  const int par = (int)(epoch & 1ull);
  for (int r = 0; r < world; ++r) {
    double *dst = reinterpret_cast<double *>(static_cast<char *>(peers.base[r]) + HM_XCHG_DATA_OFF) + ((size_t)par * world + rank) * cap;
    for (int j = threadIdx.x; j < n; j += blockDim.x) dst[j] = out[j];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) {
    unsigned long long *f = static_cast<unsigned long long *>(peers.base[threadIdx.x]) + par * HM_MAXWORLD + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
    const unsigned long long *mine = static_cast<const unsigned long long *>(peers.base[rank]) + par * HM_MAXWORLD + threadIdx.x;
    const long long t0 = clock64();
    while (true) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
      if (v >= epoch) break;
      if (clock64() - t0 > 40000000000ll) { // about 20 s: a peer is gone
        *status = 1;
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  const double *src = reinterpret_cast<const double *>(static_cast<const char *>(peers.base[rank]) + HM_XCHG_DATA_OFF) + (size_t)par * world * cap;
  for (int j = threadIdx.x; j < n; j += blockDim.x) out[j] = hm_combine_ranks(src, world, cap, j);
}

__global__ void hm_check_kernel(const double *__restrict__ x, long long n, unsigned long long *__restrict__ bad)
{
  unsigned long long c = 0, w = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = x[i];
    if (!(fabs(v) <= DBL_MAX))
      ++c;
    else if (fabs(v) > 1e6)
      ++w;
  }
  if (c) atomicAdd(bad, c);
  if (w) atomicAdd(bad + 1, w);
}

} // namespace eqb
