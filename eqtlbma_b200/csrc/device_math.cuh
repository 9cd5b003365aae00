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

// elementary functions of the true-pass kernels: the CUDA math library (full range, < 1 ulp).  The permutation BF
// kernel, which is bound by the FP64 pipe and has no divergence, uses the table-driven forms of perm_gemm.cuh.
__device__ __forceinline__ double rcp_fast(double x) { return 1.0 / x; }
__device__ __forceinline__ double log_fast(double x) { return log(x); }
__device__ __forceinline__ double exp10_fast(double x) { return exp10(x); }

// warp reductions ---------------------------------------------------------------------------
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
