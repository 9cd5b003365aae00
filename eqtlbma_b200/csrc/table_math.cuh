// table_math.cuh -- throughput-oriented elementary functions of the Bayes-factor kernels (perm_bf_kernel, phase C of
// fast_pair_warp_kernel): table + short-polynomial exp / log, MUFU-seeded reciprocal and reciprocal square root.
#pragma once

#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

namespace eqb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- throughput-oriented elementary functions
// The BF kernels are bound by the FP64 pipe and the issue slots, every lane runs the same instruction stream, and a
// log10 BF needs 1e-8 ABSOLUTE accuracy: the CUDA library's log / exp10 / division (full range, < 1 ulp) are replaced by
// table + short-polynomial forms good to ~1e-12 (16-entry tables in shared memory: entry j occupies its own bank
// pair, so a warp's 32 independent lookups never conflict).
struct BfTabs {
  double exp16[16];  // 2^(j/16)
  double2 log16[16]; // { 1/m_j, ln m_j }, m_j = 1 + (j + 1/2)/16
};

// polynomial coefficients and scale factors live in the constant bank: a DFMA takes them as a direct operand (a 64-bit
// immediate would cost two uniform-register moves per use -- a third of the instructions of the first version)
static __constant__ double PGK[16] = {
    23.083120654223414,      // 0  16 / ln 2
    6755399441055744.0,      // 1  1.5 * 2^52
    -0.043321698784996581,   // 2  -ln 2 / 16
    1.0 / 24.0,              // 3
    1.0 / 6.0,               // 4
    1.0 / 7.0,               // 5
    -1.0 / 6.0,              // 6
    0.2,                     // 7
    1.0 / 3.0,               // 8
    0.69314718055994530942,  // 9  ln 2
    2.302585092994045684,    // 10 ln 10
    0.43429448190325182765,  // 11 1 / ln 10
    92.332482616893657,      // 12 64 / ln 2
    -0.010830424696249145,   // 13 -ln 2 / 64
    0.0, 0.0};

// the tables are addressed through a 32-bit shared-space address held in a register (one LDS per lookup)
struct TabRef {
  uint32_t base;
  __device__ __forceinline__ double exp16(int j) const
  {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + ((uint32_t)j << 3)));
    return v;
  }
  __device__ __forceinline__ double2 log16(int j) const
  {
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(base + 128u + ((uint32_t)j << 4)));
    return v;
  }
};

// e^x.  FPCLAMP: any x (-inf, NaN -> ~1e-304); otherwise x must be finite with |x| < 1e7 (integer clamp of the binary
// exponent: results below 2^-1000 come out as ~1e-301, i.e. zero for every sum they enter)
template <bool FPCLAMP>
__device__ __forceinline__ double exp_tab16(double x, const TabRef T)
{
  if (FPCLAMP) {
    // max(x, -700) on the bit pattern (negative doubles order like their unsigned high words; NaN and -inf are larger)
    const unsigned int hx = (unsigned int)__double2hiint(x);
    if (hx > 0xC085E000u) x = -700.0;
  }
  const double tm = fma(x, PGK[0], PGK[1]);
  const int k = __double2loint(tm);
  const double kd = tm - PGK[1];
  const double gg = fma(kd, PGK[2], x); // |gg| <= ln2/32
  const double g2 = gg * gg;
  double s = fma(gg, PGK[3], PGK[4]);
  s = fma(gg, s, 0.5);
  const double pp = fma(g2, s, gg); // e^g - 1 to g^4: relative error < 4e-11
  const double tj = T.exp16(k & 15);
  const double v = fma(tj, pp, tj);
  const int e2 = FPCLAMP ? (k >> 4) : max(k >> 4, -1000);
  return __hiloint2double(__double2hiint(v) + (e2 << 20), __double2loint(v));
}
// e^x through a 64-entry table 2^(j/64) (shared-space address `base64`) and a degree-3 polynomial: |g| <= ln2/128, relative
// error g^4/24 < 4e-11 like exp_tab16, one FMA less.  Random lookups into 64 entries conflict on the banks (the 16-entry table
// never does): for the kernel whose FP64 pipe, not its shared-memory pipe, is the limit (perm_bf_kernel, --pbf all).
// x finite with |x| < 1e7; results below 2^-1000 come out as ~1e-301.
__device__ __forceinline__ double exp_tab64(double x, uint32_t base64)
{
  const double tm = fma(x, PGK[12], PGK[1]);
  const int k = __double2loint(tm);
  const double kd = tm - PGK[1];
  const double gg = fma(kd, PGK[13], x);
  const double s = fma(gg, PGK[4], 0.5);
  const double pp = fma(gg * gg, s, gg); // e^g - 1 to g^3
  double tj;
  asm("ld.shared.f64 %0, [%1];" : "=d"(tj) : "r"(base64 + ((uint32_t)(k & 63) << 3)));
  const double v = fma(tj, pp, tj);
  const int e2 = max(k >> 6, -1000);
  return __hiloint2double(__double2hiint(v) + (e2 << 20), __double2loint(v));
}
template <bool FPCLAMP>
__device__ __forceinline__ double exp10_tab16(double x, const TabRef T)
{
  return exp_tab16<FPCLAMP>(x * PGK[10], T);
}

// 1/x for a positive normal x: MUFU seed + one Newton step (relative error < 1e-12)
__device__ __forceinline__ double rcp_n(double x)
{
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return fma(r, fma(-x, r, 1.0), r);
}

// 1/sqrt(x), x normal positive: MUFU seed + one Newton step (relative error ~4e-13)
__device__ __forceinline__ double rsqrt_newton1(double x)
{
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = x * y;
  const double e = fma(-h, y, 1.0);
  return fma(0.5 * y, e, y);
}

// ln x for a NORMAL POSITIVE x (absolute error < 2e-13); the callers guarantee the range (see log_tab16 for the checked form)
__device__ __forceinline__ double log_tab16_pos(double x, const TabRef T)
{
  const int hi = __double2hiint(x);
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x)); // [1, 2)
  const double2 t = T.log16((hi >> 16) & 15);
  const double r = fma(m, t.x, -1.0); // |r| <= 1/33
  double p = fma(r, PGK[5], PGK[6]);
  p = fma(r, p, PGK[7]);
  p = fma(r, p, -0.25);
  p = fma(r, p, PGK[8]);
  p = fma(r, p, -0.5);
  const double lp = fma(r * r, p, r); // log1p(r) to r^7
  return fma((double)((hi >> 20) - 1023), PGK[9], t.y + lp);
}
// any argument: zero, subnormal, negative, Inf, NaN go through the library
__device__ __forceinline__ double log_tab16(double x, const TabRef T)
{
  const int hi = __double2hiint(x);
  if (hi < 0x00100000 || hi >= 0x7ff00000) return log(x);
  return log_tab16_pos(x, T);
}


// log10_weighted_sum accumulated online (utils_math.cpp:100-131) with the table exponential
struct LseTab {
  double m, acc;
  bool poisoned;
  __device__ __forceinline__ void init()
  {
    m = -INFINITY;
    acc = 0.0;
    poisoned = false;
  }
  __device__ __forceinline__ void add(double v, double w, bool is_first, const TabRef T)
  {
    if (v != v) {
      poisoned = poisoned || is_first;
      return;
    }
    const double d = v - m; // +inf on the first element
    const bool up = d > 0.0;
    const double e = exp10_tab16<true>(up ? -d : d, T);
    acc = up ? fma(acc, e, w) : fma(w, e, acc);
    m = up ? v : m;
  }
  __device__ __forceinline__ double result(const TabRef T) const
  {
    if (poisoned) return nan("");
    // m + log10(acc); weights that sum to 1 on equal values give acc = 1 +- 1 ulp and the reference's result is exactly m
    // (0 after its DBL_EPSILON snap): ln(1 + d) = d there, not the table's 1e-13
    const double d1 = acc - 1.0;
    double r = fma((fabs(d1) < 1e-8) ? d1 : log_tab16(acc, T), PGK[11], m);
    if (fabs(r) <= DBL_EPSILON) r = 0.0;
    return r;
  }
};

// fills the tables (first 16 threads of the CTA); the caller synchronises before the first lookup
__device__ __forceinline__ void bf_tabs_init(BfTabs &t)
{
  if (threadIdx.x < 16) {
    const double mj = 1.0 + ((double)threadIdx.x + 0.5) / 16.0;
    t.exp16[threadIdx.x] = exp2((double)threadIdx.x / 16.0);
    t.log16[threadIdx.x] = make_double2(1.0 / mj, log(mj));
  }
}

} // namespace eqb
