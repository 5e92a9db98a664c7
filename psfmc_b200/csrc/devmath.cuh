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
// Warp-cooperative: ALL 32 lanes call it with identical (a, x); each chunk of 32
// terms costs one division per lane, a product scan and a sum reduction instead
// of 32 dependent divisions. Every lane returns the same value.
__device__ __forceinline__ double gamma_p_series_warp(double a, double x, double lgam_a1,
                                                      int lane) {
  double sum = 1.0, tbase = 1.0;
  for (int chunk = 0; chunk < 64; ++chunk) {
    double p = x / (a + (double)(chunk * 32 + lane + 1));
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      double q = __shfl_up_sync(0xffffffffu, p, off);
      if (lane >= off) p *= q;
    }
    const double term = tbase * p;          // term number chunk*32 + lane + 1
    double part = term;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    sum += part;
    tbase = __shfl_sync(0xffffffffu, term, 31);
    if (tbase < sum * 1.0e-17) break;
  }
  return exp(a * log(x) - x - lgam_a1) * sum;
}

// kappa = gammaincinv(a, 0.5): the reference's exact Sersic b_n
// (psfMC/ModelComponents/Sersic.py:47-53 with a = 2n). The root of P(a,x) = 1/2
// is unique; Halley iterations on the series above from the Ciotti & Bertin
// (1999) asymptotic seed (large a) or the small-x inversion (small a) converge
// to double precision in 2-3 steps. Warp-cooperative like the series.
// *lgam_a1_out receives lgamma(a + 1) (reused for Gamma(2n) by the caller).
__device__ __forceinline__ double gammaincinv_half_warp(double a, int lane,
                                                        double *lgam_a1_out) {
  *lgam_a1_out = NAN;
  if (!(a > 0.0) || !isfinite(a)) return NAN;
  const double lgam_a1 = lgamma(a + 1.0);
  *lgam_a1_out = lgam_a1;
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
  const double lgam_a = lgam_a1 - log(a);
  double prev_dx = INFINITY;
  for (int it = 0; it < 16; ++it) {
    double f = gamma_p_series_warp(a, x, lgam_a1, lane) - 0.5;
    double dens = exp((a - 1.0) * log(x) - x - lgam_a);  // dP/dx
    if (!(dens > 0.0)) break;
    double step = f / dens;
    double curv = (a - 1.0) / x - 1.0;                   // P'' / P'
    double denom = 1.0 - 0.5 * step * curv;
    if (denom > 0.25) step /= denom;
    double xn = x - step;
    if (!(xn > 0.0)) xn = 0.5 * x;
    if (xn > a + 1.0) xn = 0.5 * (x + a + 1.0);
    double dx = fabs(xn - x);
    x = xn;
    // Halley's iteration converges cubically: a step below 1e-6 relative leaves an
    // error below 1e-17. (Second test: bouncing between neighbouring doubles.)
    if (dx <= 1.0e-6 * x || (dx <= 1.0e-13 * x && dx >= prev_dx)) break;
    prev_dx = dx;
  }
  return x;
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
