// Host-side batched log-priors: the priors stay the
// reference's scipy.stats objects (psfMC/distributions.py) and the Python side decides
// what is evaluated here -- per theta column one of the closed-form families below, with
// every constant (loc, scale, log(scale), the family's normalisation) computed by numpy /
// scipy on the Python side, so that only IEEE-exact operations (+ - * / and comparisons)
// happen here and the result is bit-identical to rv_frozen.logpdf. The Python caller
// verifies that on the first batch and falls back to scipy otherwise
// (psfmc_b200/models.py:_column_logp). No CUDA in this file.
#pragma once
#include <math.h>
#include <stdint.h>

#include <vector>

#include "../../include/psfmc_b200.h"

namespace psfmc {

// rv_continuous.logpdf (scipy/stats/_distn_infrastructure.py): std = (x - loc) / scale;
// inside the support and for valid arguments dist._logpdf(std) - log(scale), outside
// -inf, NaN for NaN input or invalid arguments.
// `families`: bit f set = evaluate columns of family f (default: all native families) --
// the sampler evaluates the closed-form families first (they decide which rows are dead
// before the GPU is started) and the ones with library calls while the GPU computes.
#define PSFMC_PRIOR_FAMILIES_ALL 0xffffffffu
#define PSFMC_PRIOR_FAMILIES_CHEAP ((1u << PSFMC_PRIOR_UNIFORM) | (1u << PSFMC_PRIOR_NORMAL))
inline void prior_columns_host(const psfmc_prior_column *cols, int n_cols, const double *theta,
                               long long n_batch, long long ld, double *logp,
                               long long ld_out, unsigned families = PSFMC_PRIOR_FAMILIES_ALL,
                               unsigned char *dead = nullptr) {
  // dead (optional, [n_batch], zeroed by the caller): set to 1 where an evaluated column is
  // not finite
  // blocks of rows, column by column inside a block: the family switch and the constants
  // stay out of the inner loop, the block's theta rows stay in the L1 cache
  const long long BLOCK = 128;
  for (long long b0 = 0; b0 < n_batch; b0 += BLOCK) {
    const long long b1 = b0 + BLOCK < n_batch ? b0 + BLOCK : n_batch;
    for (int c = 0; c < n_cols; ++c) {
      const psfmc_prior_column &pc = cols[c];
      if (pc.family == PSFMC_PRIOR_OTHER) continue;   // the caller fills this column
      if (!((families >> pc.family) & 1u)) continue;
      const double *x = theta + pc.theta_index;
      double *out = logp + c;
      const double loc = pc.loc, scale = pc.scale;
      if (!pc.valid) {                                 // rv_continuous.badvalue
        for (long long b = b0; b < b1; ++b) out[b * ld_out] = NAN;
        if (dead)
          for (long long b = b0; b < b1; ++b) dead[b] = 1;
        continue;
      }
      if (pc.family == PSFMC_PRIOR_UNIFORM) {
        // uniform_gen._pdf = 1.0 * (x == x); _logpdf = log(_pdf) = 0.0; support [0, 1] of
        // std = (x - loc) / scale. For a finite scale > 0 the rounded quotient lies in
        // [0, 1] exactly when 0 <= d <= scale, d = x - loc (d > scale means d >= scale +
        // ulp(scale), whose quotient exceeds the tie point 1 + 2^-53), and it is NaN exactly
        // when d is: no division.
        const double value = 0.0 - pc.log_scale;
        if (scale > 0.0 && scale <= 1.79769313486231570815e308) {
          for (long long b = b0; b < b1; ++b) {
            const double d = x[b * ld] - loc;
            double v = (d >= 0.0 && d <= scale) ? value : -INFINITY;
            if (d < 0.0 && d > -1.0e-290) {   // (a quotient that underflows to -0.0 is >= 0)
              const double std_ = d / scale;
              v = (std_ >= 0.0 && std_ <= 1.0) ? value : -INFINITY;
            }
            if (d != d) v = NAN;
            out[b * ld_out] = v;
          }
        } else {
          for (long long b = b0; b < b1; ++b) {
            const double std_ = (x[b * ld] - loc) / scale;
            double v = (std_ >= 0.0 && std_ <= 1.0) ? value : -INFINITY;
            if (std_ != std_) v = NAN;
            out[b * ld_out] = v;
          }
        }
      } else if (pc.family == PSFMC_PRIOR_WEIBULL_MIN) {
        // weibull_min_gen._logpdf = np.log(c) + sc.xlogy(c - 1, x) - pow(x, c); support
        // [0, inf). xlogy(a, y) = 0 for a == 0 and y not NaN, else a * log(y). The only
        // family here with library calls (log, pow): see PSFMC_PRIOR_WEIBULL_MIN in the
        // header.
        const double cshape = pc.shape, cm1 = pc.shape - 1.0, log_c = pc.log_shape;
        const double log_scale = pc.log_scale;
        for (long long b = b0; b < b1; ++b) {
          const double std_ = (x[b * ld] - loc) / scale;
          double v = -INFINITY;
          if (std_ >= 0.0 && std_ <= INFINITY) {
            volatile double xl = (cm1 == 0.0) ? 0.0 : cm1 * log(std_);
            volatile double pw = pow(std_, cshape);
            v = ((log_c + xl) - pw) - log_scale;
          }
          if (std_ != std_) v = NAN;
          out[b * ld_out] = v;
        }
      } else {
        // norm_gen._logpdf = -x**2 / 2.0 - log(sqrt(2 pi)); support (-inf, inf)
        const double log_norm = pc.log_norm, log_scale = pc.log_scale;
        for (long long b = b0; b < b1; ++b) {
          const double std_ = (x[b * ld] - loc) / scale;
          volatile double sq = std_ * std_;            // no contraction into an FMA
          double v = (-sq / 2.0 - log_norm) - log_scale;
          if (std_ != std_) v = NAN;
          out[b * ld_out] = v;
        }
      }
      if (dead)
        for (long long b = b0; b < b1; ++b) {
          const double v = out[b * ld_out];
          if (!(v - v == 0.0)) dead[b] = 1;          // -inf, NaN (and +inf): not finite
        }
    }
  }
}

// Joint log-prior per walker from the per-column log-densities, added in the order the
// reference adds them: per component `total += sum(prior.logp(value))` over its priors
// (ComponentBase.py:121-129), a rule "-inf if b > a" per component where one is given
// (Sersic.py:41-45: reff_b > reff), then the components in model order (models.py:187-191).
inline void prior_sum_host(const double *logp, long long n_batch, long long ld_logp,
                           const double *theta, long long ld, const psfmc_prior_term *terms,
                           int n_terms, const psfmc_prior_rule *rules, int n_rules,
                           int n_components, double *lnprior) {
  // per component: its run of terms and its rules (rules sorted by component here)
  std::vector<int> term_end(n_components, 0), rule_begin(n_components + 1, 0), order;
  for (int comp = 0, t = 0; comp < n_components; ++comp) {
    while (t < n_terms && terms[t].component == comp) ++t;
    term_end[comp] = t;
  }
  for (int comp = 0; comp < n_components; ++comp) {
    rule_begin[comp] = (int)order.size();
    for (int r = 0; r < n_rules; ++r)
      if (rules[r].component == comp) order.push_back(r);
  }
  rule_begin[n_components] = (int)order.size();
  for (long long b = 0; b < n_batch; ++b) {
    const double *lp = logp + b * ld_logp;
    const double *row = theta + b * ld;
    double total = 0.0;
    int t = 0;
    for (int comp = 0; comp < n_components; ++comp) {
      double ctotal = 0.0;
      for (; t < term_end[comp]; ++t) {
        // np.sum over the prior's own columns (sequential below 8 elements)
        const int first = terms[t].first_column, n = terms[t].n_columns;
        double s = lp[first];
        for (int k = 1; k < n; ++k) s += lp[first + k];
        ctotal = ctotal + s;
      }
      for (int q = rule_begin[comp]; q < rule_begin[comp + 1]; ++q) {
        const psfmc_prior_rule &ru = rules[order[q]];
        const double a = ru.a_index >= 0 ? row[ru.a_index] : ru.a_value;
        const double bb = ru.b_index >= 0 ? row[ru.b_index] : ru.b_value;
        if (bb > a) ctotal = -INFINITY;
      }
      total = total + ctotal;
    }
    lnprior[b] = total;
  }
}

}  // namespace psfmc
