// device_math.cuh -- FP64 special functions the hot path needs on the device.
//
// The reference takes these from GNU GSL: gsl_cdf_tdist_P/Q (gene_snp_pair.cpp:274,
// utils_math.cpp:202), gsl_cdf_gaussian_Pinv (gene_snp_pair.cpp:274), gsl_cdf_fdist_Q and
// gsl_cdf_chisq_Qinv (MVLR.cpp:404-405).  They are evaluated here from the defining
// incomplete-beta continued fraction (modified Lentz) and an Acklam start + Halley refinement
// of the normal quantile, accurate to ~1e-14 relative including the far tails (p down to the
// smallest normal double), which is what the 1e-9 / 1e-8 parity budget needs.
#pragma once

#include <cuda_runtime.h>
#include <math.h>


namespace eqb {

__device__ __forceinline__ double lgam_corr(double z)
{
  // lnGamma(z) - [(z-1/2) ln z - z + ln(2 pi)/2] for z >= 10 (Stirling series)
  const double z2 = z * z;
  return (1.0 / 12.0 - (1.0 / 360.0 - (1.0 / 1260.0 - (1.0 / 1680.0 - (1.0 / 1188.0) / z2) / z2) / z2) / z2) / z;
}

// ln[Gamma(a+b) / (Gamma(a) Gamma(b))] without cancelling three large lgamma values
static __device__ __noinline__ double ln_inv_beta(double a, double b)
{
  if (a < b) {
    const double t = a;
    a = b;
    b = t;
  }
  if (b >= 10.0)
    return a * log1p(b / a) + b * log1p(a / b) + 0.5 * (log(a) + log(b) - log(a + b)) -
           0.91893853320467274178 + lgam_corr(a + b) - lgam_corr(a) - lgam_corr(b);
  if (a >= 10.0)
    return (a - 0.5) * log1p(b / a) + b * log(a + b) - b + lgam_corr(a + b) - lgam_corr(a) - lgamma(b);
  return lgamma(a + b) - lgamma(a) - lgamma(b);
}

// continued fraction of the incomplete beta function, modified Lentz
static __device__ __noinline__ double beta_cf(double a, double b, double x)
{
  const double tiny = 1e-300, eps = 1e-16;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < tiny) d = tiny;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 20000; ++m) {
    const double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < eps) break;
  }
  return h;
}

// I_x(a,b) and 1 - I_x(a,b) from x, y = 1-x and their logs (no cancellation in either tail)
__device__ inline void beta_inc_pair(double a, double b, double x, double y, double logx, double logy,
                                     double &I, double &Ic)
{
  if (x <= 0.0) {
    I = 0.0;
    Ic = 1.0;
    return;
  }
  if (y <= 0.0) {
    I = 1.0;
    Ic = 0.0;
    return;
  }
  const double bt = exp(ln_inv_beta(a, b) + a * logx + b * logy);
  if (x < (a + 1.0) / (a + b + 2.0)) {
    I = bt * beta_cf(a, b, x) / a;
    Ic = 1.0 - I;
  } else {
    Ic = bt * beta_cf(b, a, y) / b;
    I = 1.0 - Ic;
  }
}

// two-sided Student tail Pr(|T_nu| > |t|) and its complement
static __device__ __noinline__ void tdist_tails(double t, double nu, double &tail, double &central)
{
  const double t2 = t * t;
  if (t2 == 0.0) {
    tail = 1.0;
    central = 0.0;
    return;
  }
  if (isinf(t2)) {
    tail = 0.0;
    central = 1.0;
    return;
  }
  const double x = nu / (nu + t2), y = t2 / (nu + t2);
  beta_inc_pair(0.5 * nu, 0.5, x, y, -log1p(t2 / nu), -log1p(nu / t2), tail, central);
}

// gsl_cdf_tdist_P
__device__ inline double tdist_P(double x, double nu)
{
  if (isnan(x) || isnan(nu)) return nan("");
  double tail, central;
  tdist_tails(x, nu, tail, central);
  return (x < 0.0) ? 0.5 * tail : 0.5 + 0.5 * central;
}

// gsl_cdf_tdist_Q
__device__ inline double tdist_Q(double x, double nu)
{
  if (isnan(x) || isnan(nu)) return nan("");
  double tail, central;
  tdist_tails(x, nu, tail, central);
  return (x > 0.0) ? 0.5 * tail : 0.5 + 0.5 * central;
}

// gsl_cdf_fdist_Q: upper tail of F(nu1, nu2)
__device__ inline double fdist_Q(double x, double nu1, double nu2)
{
  if (isnan(x)) return nan("");
  if (x <= 0.0) return 1.0;
  const double r = nu1 * x / nu2;
  double I, Ic;
  beta_inc_pair(0.5 * nu2, 0.5 * nu1, 1.0 / (1.0 + r), r / (1.0 + r), -log1p(r), -log1p(1.0 / r), I, Ic);
  return I;
}

// gsl_cdf_ugaussian_Pinv: lower-tail standard normal quantile
static __device__ __noinline__ double ugaussian_Pinv(double P)
{
  if (isnan(P)) return nan("");
  if (P <= 0.0) return (P == 0.0) ? -INFINITY : nan("");
  if (P >= 1.0) return (P == 1.0) ? INFINITY : nan("");
  double sign = 1.0;
  if (P > 0.5) {
    // use symmetry on the complementary probability when it is exactly representable
    const double q = 1.0 - P;
    if (q > 0.0 && (1.0 - q) == P) {
      P = q;
      sign = -1.0;
    }
  }
  const double a0 = -3.969683028665376e+01, a1 = 2.209460984245205e+02, a2 = -2.759285104469687e+02,
               a3 = 1.383577518672690e+02, a4 = -3.066479806614716e+01, a5 = 2.506628277459239e+00;
  const double b0 = -5.447609879822406e+01, b1 = 1.615858368580409e+02, b2 = -1.556989798598866e+02,
               b3 = 6.680131188771972e+01, b4 = -1.328068155288572e+01;
  const double c0 = -7.784894002430293e-03, c1 = -3.223964580411365e-01, c2 = -2.400758277161838e+00,
               c3 = -2.549732539343734e+00, c4 = 4.374664141464968e+00, c5 = 2.938163982698783e+00;
  const double d0 = 7.784695709041462e-03, d1 = 3.224671290700398e-01, d2 = 2.445134137142996e+00,
               d3 = 3.754408661907416e+00;
  const double plow = 0.02425, phigh = 1.0 - plow;
  double x;
  if (P < plow) {
    const double q = sqrt(-2.0 * log(P));
    x = (((((c0 * q + c1) * q + c2) * q + c3) * q + c4) * q + c5) / ((((d0 * q + d1) * q + d2) * q + d3) * q + 1.0);
  } else if (P <= phigh) {
    const double q = P - 0.5, r = q * q;
    x = (((((a0 * r + a1) * r + a2) * r + a3) * r + a4) * r + a5) * q /
        (((((b0 * r + b1) * r + b2) * r + b3) * r + b4) * r + 1.0);
  } else {
    const double q = sqrt(-2.0 * log1p(-P));
    x = -(((((c0 * q + c1) * q + c2) * q + c3) * q + c4) * q + c5) / ((((d0 * q + d1) * q + d2) * q + d3) * q + 1.0);
  }
  for (int it = 0; it < 4; ++it) {
    double u; // (Phi(x) - P) / phi(x)
    if (x < -5.0) {
      // relative residual times the asymptotic Mills ratio: no overflow of exp(x^2/2) in the far tail
      const double Phi = 0.5 * erfc(-x * 0.70710678118654752440);
      const double rel = (Phi - P) / Phi;
      const double x2 = x * x;
      const double mills =
          (-1.0 / x) * (1.0 - 1.0 / x2 + 3.0 / (x2 * x2) - 15.0 / (x2 * x2 * x2) + 105.0 / (x2 * x2 * x2 * x2));
      u = rel * mills;
    } else {
      const double e = 0.5 * erfc(-x * 0.70710678118654752440) - P;
      u = e * 2.50662827463100050242 * exp(0.5 * x * x);
    }
    const double dx = u / (1.0 + 0.5 * x * u);
    x -= dx;
    if (fabs(dx) <= 1e-16 * fabs(x)) break;
  }
  return sign * x;
}

// gsl_cdf_chisq_Qinv for one degree of freedom (the only case MVLR.cpp:405 reaches, p = 1 SNP)
__device__ inline double chisq_Qinv_1df(double Q)
{
  if (isnan(Q)) return nan("");
  if (Q >= 1.0) return 0.0;
  if (Q <= 0.0) return INFINITY;
  const double z = ugaussian_Pinv(0.5 * Q);
  return z * z;
}

// warp reductions ---------------------------------------------------------------------------
// ---------------------------------------------------------------- short-latency elementary functions
// The ABF phases spend most of their instructions in log / exp10 / division.  The forms below (MUFU seed +
// one correction step, Estrin polynomials, library fallback outside the plain positive / finite range) are
// accurate to a few 1e-16 and have shorter dependency chains than the CUDA library versions -- but measured
// on B200 they only pay where several INDEPENDENT evaluations can be interleaved without a branch in between
// (exp_fast_nb / rsqrt_fast_nb in the linear-domain BMA of the permutation kernel: +18%); as drop-in
// replacements inside the branchy per-item ABF code they were a wash (fast_pair_kernel +1%, perm_kernel -8%),
// so rcp_fast / log_fast / exp10_fast forward to the library unless EQB_SHORT_LATENCY_MATH is defined.
// eqb_math_selftest() checks all of them against the library either way.

// 1/x: MUFU seed (upper 20 mantissa bits) + one cubic correction step
__device__ __forceinline__ double rcp_fast_impl(double x)
{
  if (!(fabs(x) >= 1e-290 && fabs(x) <= 1e290)) return 1.0 / x;
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0); // |e| <= ~2^-20
  const double q = fma(e, e, e);    // 1/(1-e) = 1 + e + e^2 + O(e^3), e^3 <= 2^-60
  return fma(r, q, r);
}

// natural logarithm: x = 2^e m, m in [sqrt(1/2), sqrt(2)), log m = 2 atanh(s), s = (m-1)/(m+1), series to s^17
__device__ __forceinline__ double log_fast_impl(double x)
{
  const int hi = __double2hiint(x);
  if (hi < 0x00100000 || hi >= 0x7ff00000) return log(x); // zero, subnormal, negative, Inf, NaN
  int e = (hi >> 20) - 1023;
  int mhi = (hi & 0x000fffff) | 0x3ff00000;
  if (mhi >= 0x3ff6a09f) { // m >= ~sqrt(2): halve
    mhi -= 0x00100000;
    e += 1;
  }
  const double m = __hiloint2double(mhi, __double2loint(x));
  const double f = m - 1.0;
  const double s = f * rcp_fast_impl(2.0 + f);
  const double z = s * s;
  // Estrin evaluation (short dependency chain): p = sum_{k=0..7} c_k z^k, c_k = 2/(2k+3)
  const double z2 = z * z, z4 = z2 * z2;
  const double p01 = fma(2.0 / 5.0, z, 2.0 / 3.0), p23 = fma(2.0 / 9.0, z, 2.0 / 7.0);
  const double p45 = fma(2.0 / 13.0, z, 2.0 / 11.0), p67 = fma(2.0 / 17.0, z, 2.0 / 15.0);
  const double p = fma(z4, fma(z2, p67, p45), fma(z2, p23, p01));
  const double lm = fma(s * z, p, 2.0 * s);
  const double de = (double)e;
  return fma(de, 6.93147180369123816490e-01, fma(de, 1.90821492927058770002e-10, lm)); // ln2 = hi + lo
}

// 10^x = 2^n 2^f, n = rint(x log2 10), f in [-1/2, 1/2]: degree-12 Taylor polynomial of 2^f
__device__ __forceinline__ double exp10_fast_impl(double x)
{
  if (!(x > -300.0 && x < 300.0)) return exp10(x); // also NaN
  const double L2_10_HI = 3.321928094887362182e+00, L2_10_LO = 1.661617516973592e-16;
  const double t = x * L2_10_HI;
  const double magic = 6755399441055744.0; // 1.5 * 2^52: (t + magic) - magic = rint(t), integer in the low word
  const double tm = t + magic;
  const int n = __double2loint(tm);
  const double nd = tm - magic;
  double f = fma(x, L2_10_HI, -nd);
  f = fma(x, L2_10_LO, f);
  // Estrin evaluation of sum_{k=0..12} (ln2^k / k!) f^k
  const double f2 = f * f, f4 = f2 * f2, f8 = f4 * f4;
  const double q01 = fma(6.9314718055994531e-01, f, 1.0);
  const double q23 = fma(5.5504108664821580e-02, f, 2.4022650695910071e-01);
  const double q45 = fma(1.3333558146428443e-03, f, 9.6181291076284772e-03);
  const double q67 = fma(1.5252733804059840e-05, f, 1.5403530393381610e-04);
  const double q89 = fma(1.0178086009239700e-07, f, 1.3215486790144309e-06);
  const double qab = fma(4.4455382718708115e-10, f, 7.0549116208011233e-09);
  const double q03 = fma(f2, q23, q01), q47 = fma(f2, q67, q45), q8b = fma(f2, qab, q89);
  const double qhi = fma(f4, 2.5678435993488205e-11, q8b); // + c12 f^12 = f^8 * (f^4 c12)
  const double p = fma(f8, qhi, fma(f4, q47, q03));
  return p * __hiloint2double((n + 1023) << 20, 0);
}

#if defined(EQB_SHORT_LATENCY_MATH)
__device__ __forceinline__ double rcp_fast(double x) { return rcp_fast_impl(x); }
__device__ __forceinline__ double log_fast(double x) { return log_fast_impl(x); }
__device__ __forceinline__ double exp10_fast(double x) { return exp10_fast_impl(x); }
#else
__device__ __forceinline__ double rcp_fast(double x) { return 1.0 / x; }
__device__ __forceinline__ double log_fast(double x) { return log(x); }
__device__ __forceinline__ double exp10_fast(double x) { return exp10(x); }
#endif

// e^x without a branch (arguments clamped to [-708, 708]; callers exclude NaN): same scheme as exp10_fast.
// Used by the linear-domain BMA of the permutation kernel, where the K grid points of a configuration are
// independent evaluations the compiler can interleave once no branch separates them.
__device__ __forceinline__ double exp_fast_nb(double x)
{
  x = fmin(fmax(x, -708.0), 708.0);
  const double L2E_HI = 1.4426950408889634e+00, L2E_LO = 2.0355273740931033e-17;
  const double magic = 6755399441055744.0;
  const double tm = fma(x, L2E_HI, magic);
  const int n = __double2loint(tm);
  const double nd = tm - magic;
  double f = fma(x, L2E_HI, -nd);
  f = fma(x, L2E_LO, f);
  const double f2 = f * f, f4 = f2 * f2, f8 = f4 * f4;
  const double q01 = fma(6.9314718055994531e-01, f, 1.0);
  const double q23 = fma(5.5504108664821580e-02, f, 2.4022650695910071e-01);
  const double q45 = fma(1.3333558146428443e-03, f, 9.6181291076284772e-03);
  const double q67 = fma(1.5252733804059840e-05, f, 1.5403530393381610e-04);
  const double q89 = fma(1.0178086009239700e-07, f, 1.3215486790144309e-06);
  const double qab = fma(4.4455382718708115e-10, f, 7.0549116208011233e-09);
  const double q03 = fma(f2, q23, q01), q47 = fma(f2, q67, q45), q8b = fma(f2, qab, q89);
  const double qhi = fma(f4, 2.5678435993488205e-11, q8b);
  const double p = fma(f8, qhi, fma(f4, q47, q03));
  return p * __hiloint2double((n + 1023) << 20, 0);
}

// 1/sqrt(x) for a normal positive x, without a branch: MUFU seed + one cubic correction step
__device__ __forceinline__ double rsqrt_fast_nb(double x)
{
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = x * y;
  const double e = fma(-h, y, 1.0);             // 1 - x y^2
  const double q = e * fma(e, 0.375, 0.5);      // (1 - e)^(-1/2) = 1 + e/2 + 3 e^2/8 + O(e^3)
  return fma(y, q, y);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int warp_sum_int(int v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_max_nonan(double v)
{
  // fmax ignores NaN operands
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

} // namespace eqb
