// Scalar float64 device math for the per-walker constants of the render stage.
// Each function cites the reference arithmetic it reproduces (paths relative to
// /root/reference).
#pragma once
#include <math.h>

#include "common.cuh"

namespace psfmc {

#define PSFMC_PI 3.141592653589793238462643383279502884

// Regularised lower incomplete gamma P(a, x) by its power series
//   P(a,x) = x^a e^-x / Gamma(a+1) * sum_k x^k / ((a+1)...(a+k)),
// used only for x <= a + 1 (the median of Gamma(a) is always below a), where
// every term is positive and the sum converges after O(sqrt(a)) terms.
// Group-cooperative: the G lanes of a group call it with identical (a, x) (32/G
// independent groups per warp, every lane of the warp takes part in the shuffles);
// each chunk of G terms costs one division per lane, a product scan and a sum
// reduction instead of G dependent divisions. Every lane of the group returns the
// same value. Loop trip counts are made warp-uniform with __all_sync: the extra
// chunks a faster group runs only add terms below 1e-17 of its sum.
// (G = lanes per group: 8 for large batches, where the float64 pipe is the limit;
//  32 for small ones, where the length of the dependent chain is)

// 1/y for y >= 1 to ~1 ulp: hardware seed (2^-23) + two Newton steps; a full IEEE
// division costs several times as many float64 instructions, and the prepare
// kernel is bound by the float64 pipe.
__device__ __forceinline__ double fast_drcp(double y) {
#ifdef PSFMC_EMU
  return 1.0 / y;
#else
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
  r = fma(r, fma(-y, r, 1.0), r);
  r = fma(r, fma(-y, r, 1.0), r);
  return r;
#endif
}

// Returns the series sum S; P(a, x) = exp(a log x - x - lgamma(a + 1)) * S.
template <int G>
__device__ __forceinline__ double gamma_p_series_group(double a, double x, int glane) {
  double sum = 1.0, tbase = 1.0;
  for (int chunk = 0; chunk < 256; ++chunk) {
    double p = x * fast_drcp(a + (double)(chunk * G + glane + 1));
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
      double q = __shfl_up_sync(0xffffffffu, p, off, G);
      if (glane >= off) p *= q;
    }
    const double term = tbase * p;          // term number chunk*G + glane + 1
    double part = term;
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1)
      part += __shfl_xor_sync(0xffffffffu, part, off, G);
    sum += part;
    tbase = __shfl_sync(0xffffffffu, term, G - 1, G);
    if (__all_sync(0xffffffffu, !(tbase >= sum * 1.0e-17))) break;
  }
  return sum;
}

// kappa = gammaincinv(a, 0.5): the reference's exact Sersic b_n
// (psfMC/ModelComponents/Sersic.py:47-53 with a = 2n). The root of P(a,x) = 1/2
// is unique; Halley iterations on the series above from the Ciotti & Bertin
// (1999) asymptotic seed (large a) or the small-x inversion (small a) converge
// to double precision in 2-3 steps. Warp-cooperative like the series.
// *lgam_a1_out receives lgamma(a + 1) (reused for Gamma(2n) by the caller).
// `active`: false for groups that have no Sersic to work on; they still walk through
// the shuffles (with a harmless a = 1) so that the warp stays convergent.
template <int G>
__device__ __forceinline__ double gammaincinv_half_group(double a, int glane, bool active,
                                                         double *lgam_a1_out) {
  const bool valid = active && (a > 0.0) && isfinite(a);
  if (!valid) a = 1.0;
  const double lgam_a1 = lgamma(a + 1.0);
  *lgam_a1_out = valid ? lgam_a1 : NAN;
  double x;
  if (a >= 0.8) {
    const double ia = 1.0 / a;   // b_n series in 1/n with n = a/2
    x = a - 1.0 / 3.0 + ia * (8.0 / 405.0 + ia * (184.0 / 25515.0 + ia * (1048.0 / 1148175.0 -
        ia * 35115152.0 / 30690717750.0)));
  } else {
    // P(a,x) ~ x^a / Gamma(a+1) for small x
    x = exp((log(0.5) + lgam_a1) / a);
    if (a > 0.3) x = fmax(x, 0.5 * (a - 1.0 / 3.0 + 8.0 / (405.0 * a)));
  }
  if (!(x > 0.0)) x = 1.0e-300;
  double prev_dx = INFINITY;
  bool done = false;
  for (int it = 0; it < 16; ++it) {
    const double series = gamma_p_series_group<G>(a, x, glane);
    if (!done) {
      const double pref = exp(a * log(x) - x - lgam_a1);  // x^a e^-x / Gamma(a+1)
      const double f = pref * series - 0.5;
      const double dens = pref * a / x;                   // dP/dx = x^(a-1) e^-x / Gamma(a)
      if (!(dens > 0.0)) {
        done = true;
      } else {
        double step = f / dens;
        double curv = (a - 1.0) / x - 1.0;                   // P'' / P'
        double denom = 1.0 - 0.5 * step * curv;
        if (denom > 0.25) step /= denom;
        double xn = x - step;
        if (!(xn > 0.0)) xn = 0.5 * x;
        if (xn > a + 1.0) xn = 0.5 * (x + a + 1.0);
        double dx = fabs(xn - x);
        x = xn;
        // Halley's iteration converges cubically: a step below 1e-6 relative leaves
        // an error below 1e-17. (Second test: bouncing between neighbouring doubles.)
        if (dx <= 1.0e-6 * x || (dx <= 1.0e-13 * x && dx >= prev_dx)) done = true;
        prev_dx = dx;
      }
    }
    if (__all_sync(0xffffffffu, done)) break;
  }
  return valid ? x : NAN;
}

// kappa(a) from the piecewise Chebyshev table of log(kappa) over u = log2(a)
// (Clenshaw recurrence); `*ok` is false when a is outside the table (the caller then
// falls back to the iteration).
__device__ __forceinline__ double kappa_from_table(const double *coef, int n_int, double u0,
                                                   double inv_du, double a, bool *ok) {
  *ok = false;
  if (!coef || !(a > 0.0) || !isfinite(a)) return NAN;
  const double pos = (log2(a) - u0) * inv_du;
  if (!(pos >= 0.0) || !(pos < (double)n_int)) return NAN;
  const int k = (int)pos;
  const double t = 2.0 * (pos - (double)k) - 1.0;      // in [-1, 1)
  const double *c = coef + k * PSFMC_KAPPA_DEG;
  double b1 = 0.0, b2 = 0.0;
#pragma unroll
  for (int j = PSFMC_KAPPA_DEG - 1; j >= 1; --j) {
    const double b0 = fma(2.0 * t, b1, c[j] - b2);
    b2 = b1;
    b1 = b0;
  }
  *ok = true;
  return exp(fma(t, b1, c[0] - b2));
}

// psfMC/utils.py:160-164
__device__ __forceinline__ double mag_to_flux(double mag, double mag_zp) {
  return pow(10.0, -0.4 * (mag - mag_zp));
}

// psfMC/ModelComponents/Sersic.py:55-71 (same operation order); gamma_2n is
// Gamma(2n), which the caller derives from the lgamma(2n + 1) the kappa iteration
// already needed: Gamma(a) = exp(lgamma(a + 1)) / a, a few ulp from scipy's gamma.
__device__ __forceinline__ double sersic_sb_eff(double flux, double n, double reff,
                                                double reff_b, double kappa,
                                                double gamma_2n) {
  return flux / (PSFMC_PI * reff * reff_b * 2.0 * n *
                 exp(kappa + log(kappa) * -2.0 * n) * gamma_2n);
}

// psfMC/ModelComponents/PointSource.py:84-97
__device__ __forceinline__ double sinc_ref(double x) {
  return (x != 0.0) ? sin(PSFMC_PI * x) / (PSFMC_PI * x) : 1.0;
}
__device__ __forceinline__ double lanczos3_ref(double x) {
  return (fabs(x) < 3.0) ? sinc_ref(x) * sinc_ref(x / 3.0) : 0.0;
}

// psfMC/ModelComponents/PointSource.py:60-81: clip the position to stay more than
// the kernel radius from the edge, then round half to even (np.round == rint).
__device__ __forceinline__ void stamp_bounds(double pos, double radius, int n,
                                             double *lo, double *hi) {
  double cmin = radius - 0.5, cmax = (double)n - (radius + 0.5);
  double clipped = fmin(fmax(pos, cmin), cmax);   // np.clip = minimum(maximum(.))
  *lo = rint(clipped - radius);
  *hi = rint(clipped + radius);
}

}  // namespace psfmc
