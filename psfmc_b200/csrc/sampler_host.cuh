// Host side of the sampler's inner loop (include/psfmc_b200.h:
// psfmc_ensemble_run, psfmc_lnpost_batch, psfmc_rng_fill). No CUDA in this file: the lnL of
// a (half-)ensemble is one call into the engine through the callbacks below.
//
// What is restated here, and from where:
//  * emcee 2.2.1 (the reference's pin, /root/reference/environment.yml:25; not vendored in
//    /root/reference, so its published algorithm is restated): EnsembleSampler.sample /
//    _propose_stretch -- per iteration the two halves of the ensemble are updated one after
//    the other, each against the other half:
//        zz      = ((a - 1) * rand(Ns) + 1)**2 / a
//        partner = randint(Nc, size=Ns)
//        q       = c[partner] - zz * (c[partner] - s)
//        lnpdiff = (D - 1) * log(zz) + lnpost(q) - lnpost(s)
//        accept  = lnpdiff > log(rand(Ns))
//    (psfmc_b200/sampler.py is the same loop in numpy; tests compare the two chain for chain).
//  * numpy.random.RandomState (legacy MT19937 stream): random_sample takes two 32-bit
//    draws, (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53; randint(n) draws 32-bit words,
//    masks them with the smallest all-ones mask >= n - 1 and rejects values above n - 1
//    (numpy/random/src/distributions/distributions.c: random_bounded_uint64_fill with
//    use_masked, buffered_bounded_masked_uint32); n == 1 consumes nothing.
//  * psfMC/models.py:205-211, 238-243: lnpost = lnprior + lnL, -inf when either is not finite.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/psfmc_b200.h"
#include "priors_host.cuh"

namespace psfmc {

struct NumpyMT19937 {
  uint32_t *key;   // [624]
  int32_t *pos;
  void reload() {
    const int N = 624, M = 397;
    const uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;
    uint32_t y;
    int i = 0;
    for (; i < N - M; ++i) {
      y = (key[i] & UPPER) | (key[i + 1] & LOWER);
      key[i] = key[i + M] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
    }
    for (; i < N - 1; ++i) {
      y = (key[i] & UPPER) | (key[i + 1] & LOWER);
      key[i] = key[i + (M - N)] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
    }
    y = (key[N - 1] & UPPER) | (key[0] & LOWER);
    key[N - 1] = key[M - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
    *pos = 0;
  }
  uint32_t next32() {
    if (*pos >= 624 || *pos < 0) reload();
    uint32_t y = key[(*pos)++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  double next_double() {
    const int32_t a = (int32_t)(next32() >> 5), b = (int32_t)(next32() >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
  }
  // RandomState.randint(bound), 0 < bound <= 2^32 - 1
  uint32_t next_bounded(uint32_t bound) {
    const uint32_t rng = bound - 1u;
    if (rng == 0u) return 0u;
    uint32_t mask = rng;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    uint32_t val;
    while ((val = (next32() & mask)) > rng) {
    }
    return val;
  }
};

// the engine as the sampler sees it: lnL of n rows in two halves (host work in between)
struct LnlikeCalls {
  void *self;
  int (*begin)(void *self, const double *theta, long long n, long long ld, double *lnl);
  int (*end)(void *self);
};

inline int check_prior_plan(const psfmc_prior_plan *pl, long long ld, const char **why) {
  if (!pl) return 0;   // no priors: lnpost = lnL
  if (pl->n_columns < 0 || pl->n_terms < 0 || pl->n_rules < 0 || pl->n_components < 0) {
    *why = "negative size in the prior plan";
    return 1;
  }
  if ((pl->n_columns && !pl->columns) || (pl->n_terms && !pl->terms) ||
      (pl->n_rules && !pl->rules)) {
    *why = "null table in the prior plan";
    return 1;
  }
  bool other = false;
  for (int c = 0; c < pl->n_columns; ++c) {
    const psfmc_prior_column &pc = pl->columns[c];
    if (pc.family < PSFMC_PRIOR_OTHER || pc.family > PSFMC_PRIOR_WEIBULL_MIN) {
      *why = "unknown prior family";
      return 1;
    }
    if (pc.family == PSFMC_PRIOR_OTHER)
      other = true;
    else if (pc.theta_index < 0 || pc.theta_index >= ld) {
      *why = "prior column outside theta";
      return 1;
    }
  }
  if (other && !pl->other_columns) {
    *why = "the prior plan has PSFMC_PRIOR_OTHER columns but no other_columns callback";
    return 1;
  }
  for (int t = 0; t < pl->n_terms; ++t) {
    const psfmc_prior_term &pt = pl->terms[t];
    if (pt.n_columns < 1 || pt.first_column < 0 || pt.first_column + pt.n_columns > pl->n_columns ||
        pt.component < 0 || pt.component >= pl->n_components ||
        (t > 0 && pt.component < pl->terms[t - 1].component)) {
      *why = "prior terms must lie inside the columns and be grouped by ascending component";
      return 1;
    }
  }
  for (int r = 0; r < pl->n_rules; ++r)
    if (pl->rules[r].a_index >= ld || pl->rules[r].b_index >= ld) {
      *why = "prior rule outside theta";
      return 1;
    }
  return 0;
}

struct LnpostWork {
  std::vector<double> logp, lnprior;
};

// lnpost of n rows: the GPU is started on ALL rows, the priors run on this thread
// meanwhile; rows with a dead prior were evaluated for nothing (models.py:209-211 skips
// them -- same result, a walker's lnL does not depend on its batch).
// Returns 0, an engine error code (> 0, message already set), or -1: the callback failed.
inline int lnpost_rows(const LnlikeCalls &eng, const psfmc_prior_plan *pl, const double *theta,
                       long long n, long long ld, double *lnl, double *lnpost, LnpostWork &wk) {
  if (n <= 0) return 0;
  int rc = eng.begin(eng.self, theta, n, ld, lnl);
  if (rc) return rc;
  int cb = 0;
  if (pl) {
    const long long ldp = pl->n_columns > 0 ? pl->n_columns : 1;
    wk.logp.resize((size_t)n * ldp);
    wk.lnprior.resize((size_t)n);
    prior_columns_host(pl->columns, pl->n_columns, theta, n, ld, wk.logp.data(), ldp);
    bool other = false;
    for (int c = 0; c < pl->n_columns; ++c) other |= pl->columns[c].family == PSFMC_PRIOR_OTHER;
    if (other) cb = pl->other_columns(pl->user, theta, n, ld, wk.logp.data(), ldp);
    if (!cb)
      prior_sum_host(wk.logp.data(), n, ldp, theta, ld, pl->terms, pl->n_terms, pl->rules,
                     pl->n_rules, pl->n_components, wk.lnprior.data());
  }
  rc = eng.end(eng.self);   // (always: the batch in flight must be finished)
  if (rc) return rc;
  if (cb) return -1;
  for (long long b = 0; b < n; ++b) {
    const double lp = pl ? wk.lnprior[b] : 0.0;
    const double sum = lnl[b] + lp;
    lnpost[b] = (std::isfinite(lp) && std::isfinite(lnl[b])) ? sum : -INFINITY;
  }
  return 0;
}

// status of run_ensemble beyond the engine error codes (which are positive)
#define PSFMC_ENS_OK 0
#define PSFMC_ENS_CALLBACK (-1)
#define PSFMC_ENS_POS_INF (-2)
#define PSFMC_ENS_POS_NAN (-3)
#define PSFMC_ENS_LNPROB_NAN (-4)

// q, lnl: [k/2][D] / [k/2] buffers the engine reads / writes (page-locked where the caller
// can: the engine then skips its staging copies and replays one captured graph per call)
inline int run_ensemble(const LnlikeCalls &eng, const psfmc_prior_plan *pl, psfmc_ensemble *e,
                        long long n_iter, double *q, double *lnl) {
  const long long k = e->n_walkers, D = e->n_dim, half = k / 2;
  NumpyMT19937 mt{e->mt_key, e->mt_pos};
  std::vector<double> zz((size_t)half), newlnp((size_t)half);
  std::vector<long long> partner((size_t)half);
  LnpostWork wk;
  const double a = e->a, dm1 = (double)D - 1.0;
  const long long thin = e->thin > 0 ? e->thin : 1;
  for (long long it = 0; it < n_iter; ++it) {
    for (int h = 0; h < 2; ++h) {
      const long long s0 = h == 0 ? 0 : half, c0 = h == 0 ? half : 0;
      const long long ns = h == 0 ? half : k - half, nc = k - ns;
      double *s = e->pos + s0 * D;
      const double *c = e->pos + c0 * D;
      for (long long i = 0; i < ns; ++i) {
        volatile double t = (a - 1.0) * mt.next_double();   // (no contraction into an FMA)
        const double t1 = t + 1.0;
        volatile double sq = t1 * t1;
        zz[i] = sq / a;
      }
      for (long long i = 0; i < ns; ++i) partner[i] = (long long)mt.next_bounded((uint32_t)nc);
      bool has_inf = false, has_nan = false;
      for (long long i = 0; i < ns; ++i) {
        const double *cp = c + partner[i] * D;
        const double *sp = s + i * D;
        double *qp = q + i * D;
        const double z = zz[i];
        for (long long j = 0; j < D; ++j) {
          volatile double m = z * (cp[j] - sp[j]);
          const double v = cp[j] - m;
          qp[j] = v;
          has_inf |= std::isinf(v);
          has_nan |= v != v;
        }
      }
      if (has_inf) return PSFMC_ENS_POS_INF;   // emcee: ValueError
      if (has_nan) return PSFMC_ENS_POS_NAN;
      int rc = lnpost_rows(eng, pl, q, ns, D, lnl, newlnp.data(), wk);
      if (rc) return rc;
      double *lnp = e->lnprob + s0;
      for (long long i = 0; i < ns; ++i) {
        if (newlnp[i] != newlnp[i]) return PSFMC_ENS_LNPROB_NAN;
        volatile double dl = dm1 * log(zz[i]);
        const double lnpdiff = (dl + newlnp[i]) - lnp[i];
        const double lu = log(mt.next_double());
        if (lnpdiff > lu) {
          lnp[i] = newlnp[i];
          memcpy(s + i * D, q + i * D, (size_t)D * sizeof(double));
          if (e->n_accepted) e->n_accepted[s0 + i] += 1.0;
        }
      }
    }
    if (it % thin == 0) {
      const long long ind = e->chain_start + it / thin;
      if (ind < e->chain_len) {
        if (e->chain)
          for (long long w = 0; w < k; ++w)
            memcpy(e->chain + ((size_t)w * e->chain_len + ind) * D, e->pos + w * D,
                   (size_t)D * sizeof(double));
        if (e->lnprob_chain)
          for (long long w = 0; w < k; ++w) e->lnprob_chain[(size_t)w * e->chain_len + ind] = e->lnprob[w];
      }
    }
  }
  return PSFMC_ENS_OK;
}

}  // namespace psfmc
