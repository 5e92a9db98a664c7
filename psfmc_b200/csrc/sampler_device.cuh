// The sampler's half-step on the device (psfmc_ensemble_run with every prior column in one
// of the library's closed-form families, one device; engine.cu: run_ensemble_device).
//
// sampler_host.cuh runs proposals, priors and acceptance on the host around one blocking
// lnL call per half-ensemble: at the reference example's 250 walkers that call costs 57 us
// of which the GPU works 34, at 4096 walkers the host part between two dependent calls
// costs a fifth of the time. Here a half-step is three small kernels around the lnL kernels
// on ONE stream and the host never waits: it only draws the random numbers (numpy's
// MT19937 stream, which does not depend on any result) and the logarithms of the acceptance
// test, a few half-steps ahead, and enqueues.
//
//   propose_kernel   q = c[partner] - zz (c[partner] - s)  (IEEE operations in numpy's order:
//                    __dmul_rn / __dsub_rn, no contraction), the log-prior of q in the
//                    reference's order (priors_host.cuh restated for the device: same
//                    operations, the Weibull column with the device's log / pow), and the
//                    rows the lnL kernels get: q, or -- where the prior is dead -- the
//                    walker's current position (a dead row's lnL is never looked at,
//                    psfMC/models.py:209-211, but its parameters could be anything)
//   (prepare + lnL kernels of the engine on the device pointers)
//   accept_kernel    lnpost = lnL + lnprior (-inf unless both are finite, models.py:238-243);
//                    accept where (D - 1) log zz + lnpost - lnprob > log u; positions,
//                    lnprob and acceptance counts updated in place
//   store_kernel     the ensemble after an iteration into the chain block [k][n_store][D]
//
// What differs from the host loop: log / pow of the Weibull columns are the device's (a few
// ulps from libm, like numpy's SVML ones), and a non-finite float32 lnL is -inf (no float64
// repeat, as for every device-pointer call). Positions are the host loop's unless an
// acceptance is decided by those last bits.
#pragma once
#include "pipeline.cuh"
#include "sampler_host.cuh"

namespace psfmc {

struct DevPriorPlan {
  const psfmc_prior_column *columns;
  const psfmc_prior_term *terms;
  const psfmc_prior_rule *rules;
  int n_columns, n_terms, n_rules, n_components;
};

__device__ __forceinline__ double dev_sub(double a, double b) {
#ifdef PSFMC_EMU
  volatile double r = a - b;
  return r;
#else
  return __dsub_rn(a, b);
#endif
}
__device__ __forceinline__ double dev_add(double a, double b) {
#ifdef PSFMC_EMU
  volatile double r = a + b;
  return r;
#else
  return __dadd_rn(a, b);
#endif
}
__device__ __forceinline__ double dev_mul(double a, double b) {
#ifdef PSFMC_EMU
  volatile double r = a * b;
  return r;
#else
  return __dmul_rn(a, b);
#endif
}
__device__ __forceinline__ double dev_div(double a, double b) {
#ifdef PSFMC_EMU
  volatile double r = a / b;
  return r;
#else
  return __ddiv_rn(a, b);
#endif
}

// rv_continuous.logpdf of one value (priors_host.cuh: prior_columns_host, same operations)
__device__ __forceinline__ double dev_prior_column(const psfmc_prior_column &pc, double x) {
  if (!pc.valid) return NAN;
  const double std_ = dev_div(dev_sub(x, pc.loc), pc.scale);
  if (std_ != std_) return NAN;
  if (pc.family == PSFMC_PRIOR_UNIFORM)
    return (std_ >= 0.0 && std_ <= 1.0) ? dev_sub(0.0, pc.log_scale) : -INFINITY;
  if (pc.family == PSFMC_PRIOR_WEIBULL_MIN) {
    if (!(std_ >= 0.0 && std_ <= INFINITY)) return -INFINITY;
    const double cm1 = dev_sub(pc.shape, 1.0);
    const double xl = (cm1 == 0.0) ? 0.0 : dev_mul(cm1, log(std_));
    const double pw = pow(std_, pc.shape);
    return dev_sub(dev_sub(dev_add(pc.log_shape, xl), pw), pc.log_scale);
  }
  // normal: -x**2 / 2.0 - log(sqrt(2 pi)), then - log(scale)
  const double sq = dev_mul(std_, std_);
  return dev_sub(dev_sub(dev_div(-sq, 2.0), pc.log_norm), pc.log_scale);
}

// One WARP per proposed row: lane j computes coordinate j of the proposal and the
// log-density of prior column j (the division / log / pow chains of the columns run side by
// side: one thread per row took 23 us per 2048 rows, all of it latency), lane 0 adds them up
// in the reference's order. blockDim = 128 (four rows per CTA); rows of more than 32
// columns loop.
#define PSFMC_PROPOSE_MAXD 256
__global__ void propose_kernel(DevPriorPlan pl, const double *__restrict__ pos, long long s0,
                               long long c0, long long ns, int D,
                               const double *__restrict__ zz, const int *__restrict__ partner,
                               double *__restrict__ q, double *__restrict__ q_gpu,
                               double *__restrict__ lnprior) {
  __shared__ double logp_s[4][PSFMC_PROPOSE_MAXD];
  __shared__ int dead_s[4];
  const int lane = threadIdx.x & 31, wrow = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 4 + wrow;
  const bool live = i < ns;
  const long long ii = live ? i : 0;
  const double *cp = pos + (c0 + partner[ii]) * D;
  const double *sp = pos + (s0 + ii) * D;
  double *qp = q + ii * D;
  const double z = zz[ii];
  if (live)
    for (int j = lane; j < D; j += 32) qp[j] = dev_sub(cp[j], dev_mul(z, dev_sub(cp[j], sp[j])));
  __syncwarp();
  if (live)
    for (int c = lane; c < pl.n_columns; c += 32)
      logp_s[wrow][c] = dev_prior_column(pl.columns[c], qp[pl.columns[c].theta_index]);
  __syncwarp();
  if (live && lane == 0) {
    // joint log-prior: per component the sum over its priors (each the sequential sum of
    // its columns), -inf where a rule b > a holds, then the components in model order
    // (priors_host.cuh: prior_sum_host)
    double total = 0.0;
    int t = 0;
    for (int comp = 0; comp < pl.n_components; ++comp) {
      double ctotal = 0.0;
      for (; t < pl.n_terms && pl.terms[t].component == comp; ++t) {
        const int first = pl.terms[t].first_column, n = pl.terms[t].n_columns;
        double s = logp_s[wrow][first];
        for (int k = 1; k < n; ++k) s = dev_add(s, logp_s[wrow][first + k]);
        ctotal = dev_add(ctotal, s);
      }
      for (int r = 0; r < pl.n_rules; ++r) {
        const psfmc_prior_rule &ru = pl.rules[r];
        if (ru.component != comp) continue;
        const double a = ru.a_index >= 0 ? qp[ru.a_index] : ru.a_value;
        const double bb = ru.b_index >= 0 ? qp[ru.b_index] : ru.b_value;
        if (bb > a) ctotal = -INFINITY;
      }
      total = dev_add(total, ctotal);
    }
    lnprior[i] = total;
    dead_s[wrow] = !(total - total == 0.0);
  }
  __syncwarp();
  if (live) {
    const bool dead = dead_s[wrow] != 0;
    double *gp = q_gpu + i * D;
    for (int j = lane; j < D; j += 32) gp[j] = dead ? sp[j] : qp[j];
  }
}

// one thread per proposed row
__global__ void accept_kernel(double *__restrict__ pos, double *__restrict__ lnprob,
                              double *__restrict__ n_accepted, long long s0, long long ns,
                              int D, const double *__restrict__ q,
                              const double *__restrict__ lnl,
                              const double *__restrict__ lnprior,
                              const double *__restrict__ lzz, const double *__restrict__ lu) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  const double lp = lnprior[i], l = lnl[i];
  const bool ok = (lp - lp == 0.0) && (l - l == 0.0);
  const double lnpost = ok ? dev_add(l, lp) : -INFINITY;
  const double lnpdiff = dev_sub(dev_add(lzz[i], lnpost), lnprob[s0 + i]);
  if (lnpdiff > lu[i]) {
    lnprob[s0 + i] = lnpost;
    double *sp = pos + (s0 + i) * D;
    const double *qp = q + i * D;
    for (int j = 0; j < D; ++j) sp[j] = qp[j];
    n_accepted[s0 + i] += 1.0;
  }
}

// chain block [k][n_store][D] / [k][n_store]: one thread per (walker, column)
__global__ void store_kernel(const double *__restrict__ pos, const double *__restrict__ lnprob,
                             long long k, int D, long long n_store, long long slot,
                             double *__restrict__ chain, double *__restrict__ lnprob_chain) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= k * D) return;
  const long long w = e / D;
  const int j = (int)(e - w * D);
  if (chain) chain[(w * n_store + slot) * D + j] = pos[e];
  if (lnprob_chain && j == 0) lnprob_chain[w * n_store + slot] = lnprob[w];
}

}  // namespace psfmc
