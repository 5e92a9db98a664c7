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

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
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
  static inline uint32_t temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  // The next n words of the stream at once (the same words n calls of next32 return): the
  // state array is tempered in runs, position and bounds in registers -- the per-word form
  // reloads *pos through the pointer for every word and does not vectorise.
  void fill32(uint32_t *out, long long n) {
    while (n > 0) {
      if (*pos >= 624 || *pos < 0) reload();
      const int p = *pos;
      const long long m = n < (long long)(624 - p) ? n : (long long)(624 - p);
      const uint32_t *src = key + p;
      for (long long i = 0; i < m; ++i) out[i] = temper(src[i]);
      *pos = p + (int)m;
      out += m;
      n -= m;
    }
  }
  // n times random_sample (two words each)
  void fill_double(double *out, long long n) {
    uint32_t words[512];
    while (n > 0) {
      const long long m = n < 256 ? n : 256;
      fill32(words, 2 * m);
      for (long long i = 0; i < m; ++i) {
        const int32_t a = (int32_t)(words[2 * i] >> 5), b = (int32_t)(words[2 * i + 1] >> 6);
        out[i] = (a * 67108864.0 + b) / 9007199254740992.0;
      }
      out += m;
      n -= m;
    }
  }
  // n times randint(bound): masked rejection consumes a data-dependent number of words, so
  // the run over the state array stops at the word that completes the request
  template <typename T>
  void fill_bounded(T *out, long long n, uint32_t bound) {
    const uint32_t rng = bound - 1u;
    if (rng == 0u) {
      for (long long i = 0; i < n; ++i) out[i] = (T)0;
      return;
    }
    uint32_t mask = rng;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    long long done = 0;
    while (done < n) {
      if (*pos >= 624 || *pos < 0) reload();
      int p = *pos;
      while (p < 624 && done < n) {      // (branch-free: a rejected value is overwritten)
        const uint32_t val = temper(key[p++]) & mask;
        out[done] = (T)val;
        done += (long long)(val <= rng);
      }
      *pos = p;
    }
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

// A few persistent host threads for the per-row work of large ensembles (the priors of
// 2048 walkers cost as much host time as their lnL costs GPU time). parallel_rows splits
// [0, n) into contiguous ranges, the calling thread takes the first.
class HostPool {
 public:
  // One pool per process, never destroyed (its threads end with the process: a static
  // destructor would have to join them, and in a forked child -- multiprocessing workers,
  // which inherit the object but not the threads -- that join would never return). A
  // forked child starts with a fresh, empty pool.
  static HostPool &instance() {
    static std::once_flag once;
    std::call_once(once, [] {
      current() = new HostPool();
      pthread_atfork(nullptr, nullptr, [] { current() = new HostPool(); });
    });
    return *current();
  }
  template <typename F>
  void parallel_rows(long long n, long long min_rows_per_thread, const F &fn) {
    int parts = (int)(n / (min_rows_per_thread > 0 ? min_rows_per_thread : 1));
    if (parts > max_threads_) parts = max_threads_;
    if (parts <= 1) {
      fn(0ll, n);
      return;
    }
    std::lock_guard<std::mutex> call_guard(call_m_);   // one parallel region at a time
    ensure_threads(parts - 1);
    const long long step = (n + parts - 1) / parts;
    std::function<void(long long, long long)> job = fn;
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = &job;
      step_ = step;
      n_ = n;
      parts_ = parts;
      pending_.store(parts - 1, std::memory_order_relaxed);
      epoch_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    fn(0ll, step < n ? step : n);
    // the regions are short (tens of microseconds): spin for the others
    for (unsigned spins = 0; pending_.load(std::memory_order_acquire) != 0; ++spins) {
      if (spins > 4096) std::this_thread::yield();
      else cpu_relax();
    }
    job_ = nullptr;
  }
 private:
  static HostPool *&current() {
    static HostPool *pool = nullptr;
    return pool;
  }
  static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
  }
  HostPool() {
    unsigned hw = std::thread::hardware_concurrency();
    // (one process per GPU: the ranks of a node share its cores -- torchrun's
    // LOCAL_WORLD_SIZE; eight spinning pools of eight threads on 32 cores ran the sharded
    // loop at a seventh of the speed of four)
    if (const char *lws = getenv("LOCAL_WORLD_SIZE")) {
      const int ranks = atoi(lws);
      if (ranks > 1) hw = hw / (unsigned)ranks;
    }
    max_threads_ = hw >= 16 ? 8 : (hw >= 8 ? 4 : (hw >= 4 ? 2 : 1));
    if (const char *env = getenv("PSFMC_HOST_THREADS")) {
      const int v = atoi(env);
      if (v >= 1 && v <= 64) max_threads_ = v;
    }
  }
  void ensure_threads(int count) {
    while ((int)threads_.size() < count) {
      const int index = (int)threads_.size() + 1;
      threads_.emplace_back([this, index] { loop(index); });
    }
  }
  // A worker spins for a short while after a region (the next one of the same posterior
  // call follows within microseconds), then sleeps on the condition variable.
  void loop(int index) {
    unsigned long long seen = 0;
    for (;;) {
      bool changed = false;
      for (int spins = 0; spins < 4000; ++spins) {
        if (stop_.load(std::memory_order_relaxed) ||
            epoch_.load(std::memory_order_acquire) != seen) {
          changed = true;
          break;
        }
        cpu_relax();
      }
      if (!changed) {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] {
          return stop_.load() || epoch_.load(std::memory_order_acquire) != seen;
        });
      }
      if (stop_.load()) return;
      std::function<void(long long, long long)> *job = nullptr;
      long long lo = 0, hi = 0;
      {
        // (the region's parameters were published before the epoch moved; taking the
        // mutex orders this read after a region that is still being set up)
        std::lock_guard<std::mutex> lk(m_);
        seen = epoch_.load(std::memory_order_acquire);
        if (index >= parts_) continue;      // not part of this region
        job = job_;
        lo = step_ * index;
        hi = lo + step_ < n_ ? lo + step_ : n_;
      }
      if (job && lo < hi) (*job)(lo, hi);
      pending_.fetch_sub(1, std::memory_order_release);
    }
  }
  std::mutex m_, call_m_;
  std::condition_variable cv_;
  std::vector<std::thread> threads_;
  std::function<void(long long, long long)> *job_ = nullptr;
  long long step_ = 0, n_ = 0;
  int parts_ = 0, max_threads_ = 1;
  std::atomic<int> pending_{0};
  std::atomic<unsigned long long> epoch_{0};
  std::atomic<bool> stop_{false};
};

// PSFMC_ENS_PROFILE=1: where the host time of psfmc_ensemble_run goes (stderr, per run)
struct EnsProfile {
  bool on = false;
  double t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long calls = 0, rows = 0, rows_gpu = 0;
  double t_begin = 0.0;
  EnsProfile() {
    const char *env = getenv("PSFMC_ENS_PROFILE");
    on = env && env[0] == '1';
    if (on) t_begin = now();
  }
  static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch())
        .count();
  }
  void report(const char *what) {
    if (!on || !calls) return;
    fprintf(stderr,
            "[psfmc] %s: %lld posterior calls, %lld rows (%lld on the GPU); per call (us): "
            "draws+proposals %.1f, closed-form priors %.1f, begin %.1f, priors behind the GPU "
            "%.1f, logs + draws + chain storage behind the GPU %.1f, wait %.1f, combine+accept "
            "%.1f, last store %.1f; "
            "whole run %.1f\n",
            what, calls, rows, rows_gpu, 1e6 * t[0] / calls, 1e6 * t[1] / calls,
            1e6 * t[2] / calls, 1e6 * t[3] / calls, 1e6 * t[4] / calls, 1e6 * t[5] / calls,
            1e6 * t[6] / calls, 1e6 * t[7] / calls, 1e6 * (now() - t_begin) / calls);
  }
};

struct LnpostWork {
  EnsProfile prof;
  double last_dead_frac = 1.0;   // share of rows the closed-form priors killed in the last call
  std::vector<double> logp, lnprior;
  std::vector<long long> alive;     // rows sent to the GPU (when some were left out)
  std::vector<unsigned char> dead;
};

// rows up to which a host call is served by replaying a captured graph (engine.cu:
// psfmc_lnlike_batch_begin): the batch size must not change from call to call there
#define PSFMC_ENS_FIXED_BATCH 160

// lnpost of n rows.
//  0. large batches whose previous call lost at most 2 % of its rows to step 1 skip the
//     screening: the GPU is started on all rows at once and step 1 runs behind it with the
//     rest (a converged chain proposes few dead rows; 40 us per 2048 rows not spent in front
//     of the launch);
//  1. the closed-form prior columns (Uniform, Normal) and the component rules decide which
//     rows are dead before the GPU is started: a dead row's lnL is never looked at
//     (psfMC/models.py:209-211 returns before evaluating it). Large batches leave the dead
//     rows out (fewer rows on the GPU); batches of at most PSFMC_ENS_FIXED_BATCH rows keep
//     their size and carry a copy of a live row in the place of a dead one (a dead row's
//     parameters may be anything -- negative radii, a NaN-producing index -- and would be
//     repeated in float64 for nothing). `scratch` [n][ld] (page-locked) takes those rows;
//  2. the GPU is started (begin); while it works: the prior columns with library calls
//     (Weibull), the caller's columns (callback), the sums, and `overlap()` -- the
//     sampler's own logarithms;
//  3. end; lnpost = lnL + lnprior where both are finite, else -inf (models.py:238-243).
// Returns 0, an engine error code (> 0, message already set), or -1: the callback failed.
inline int lnpost_rows(const LnlikeCalls &eng, const psfmc_prior_plan *pl, const double *theta,
                       long long n, long long ld, double *lnl, double *lnpost, LnpostWork &wk,
                       double *scratch, const std::function<void()> *overlap = nullptr) {
  if (n <= 0) return 0;
  HostPool &pool = HostPool::instance();
  EnsProfile &pf = wk.prof;
  double tm = pf.on ? EnsProfile::now() : 0.0;
  auto lap = [&](int slot) {
    if (!pf.on) return;
    const double t = EnsProfile::now();
    pf.t[slot] += t - tm;
    tm = t;
  };
  const long long ldp = (pl && pl->n_columns > 0) ? pl->n_columns : 1;
  bool other = false, costly = false;
  long long n_alive = n;
  const bool screen_first = pl && (n <= PSFMC_ENS_FIXED_BATCH || wk.last_dead_frac > 0.02);
  auto screen = [&]() {
    double *logp = wk.logp.data();
    unsigned char *dead = wk.dead.data();
    pool.parallel_rows(n, 512, [&](long long lo, long long hi) {
      prior_columns_host(pl->columns, pl->n_columns, theta + lo * ld, hi - lo, ld, logp + lo * ldp,
                         ldp, PSFMC_PRIOR_FAMILIES_CHEAP, dead + lo);
      for (int r = 0; r < pl->n_rules; ++r) {
        const psfmc_prior_rule &ru = pl->rules[r];
        for (long long b = lo; b < hi; ++b) {
          const double a = ru.a_index >= 0 ? theta[b * ld + ru.a_index] : ru.a_value;
          const double bb = ru.b_index >= 0 ? theta[b * ld + ru.b_index] : ru.b_value;
          if (bb > a) dead[b] = 1;
        }
      }
    });
    long long alive = 0;
    for (long long b = 0; b < n; ++b) alive += dead[b] ? 0 : 1;
    wk.last_dead_frac = (double)(n - alive) / (double)n;
    return alive;
  };
  if (pl) {
    wk.logp.resize((size_t)n * ldp);
    wk.lnprior.resize((size_t)n);
    wk.dead.assign((size_t)n, 0);
    for (int c = 0; c < pl->n_columns; ++c) {
      other |= pl->columns[c].family == PSFMC_PRIOR_OTHER;
      costly |= pl->columns[c].family == PSFMC_PRIOR_WEIBULL_MIN;
    }
    if (screen_first) n_alive = screen();
  }
  const double *gpu_theta = theta;
  long long gpu_n = n;
  bool compacted = false;
  if (n_alive < n && n_alive > 0 && scratch) {
    if (n <= PSFMC_ENS_FIXED_BATCH) {
      long long first = 0;
      while (wk.dead[first]) ++first;
      for (long long b = 0; b < n; ++b)
        memcpy(scratch + b * ld, theta + (wk.dead[b] ? first : b) * ld, (size_t)ld * sizeof(double));
    } else {
      wk.alive.clear();
      for (long long b = 0; b < n; ++b)
        if (!wk.dead[b]) {
          memcpy(scratch + (long long)wk.alive.size() * ld, theta + b * ld,
                 (size_t)ld * sizeof(double));
          wk.alive.push_back(b);
        }
      gpu_n = n_alive;
      compacted = true;
    }
    gpu_theta = scratch;
  }
  int rc = 0;
  const bool launched = n_alive > 0;
  lap(1);
  if (launched && (rc = eng.begin(eng.self, gpu_theta, gpu_n, ld, lnl))) return rc;
  lap(2);
  ++pf.calls;
  pf.rows += n;
  pf.rows_gpu += launched ? gpu_n : 0;
  int cb = 0;
  if (pl) {
    double *logp = wk.logp.data();
    if (!screen_first) screen();     // (behind the GPU; every row is being evaluated)
    if (costly)
      pool.parallel_rows(n, 256, [&](long long lo, long long hi) {
        prior_columns_host(pl->columns, pl->n_columns, theta + lo * ld, hi - lo, ld,
                           logp + lo * ldp, ldp, 1u << PSFMC_PRIOR_WEIBULL_MIN);
      });
    if (other) cb = pl->other_columns(pl->user, theta, n, ld, logp, ldp);
    if (!cb)
      pool.parallel_rows(n, 1024, [&](long long lo, long long hi) {
        prior_sum_host(logp + lo * ldp, hi - lo, ldp, theta + lo * ld, ld, pl->terms,
                       pl->n_terms, pl->rules, pl->n_rules, pl->n_components,
                       wk.lnprior.data() + lo);
      });
  }
  lap(3);
  if (overlap && *overlap) (*overlap)();
  lap(4);
  if (launched) rc = eng.end(eng.self);   // (always: the batch in flight must be finished)
  lap(5);
  if (rc) return rc;
  if (cb) return -1;
  if (compacted) {
    for (long long b = 0; b < n; ++b) lnpost[b] = -INFINITY;
    for (size_t i = 0; i < wk.alive.size(); ++i) {
      const long long b = wk.alive[i];
      const double lp = wk.lnprior[b], sum = lnl[i] + lp;
      lnpost[b] = (std::isfinite(lp) && std::isfinite(lnl[i])) ? sum : -INFINITY;
    }
    return 0;
  }
  for (long long b = 0; b < n; ++b) {
    const double lp = pl ? wk.lnprior[b] : 0.0;
    const bool gone = !launched || (pl && wk.dead[b]);
    const double sum = gone ? -INFINITY : lnl[b] + lp;
    lnpost[b] = (!gone && std::isfinite(lp) && std::isfinite(lnl[b])) ? sum : -INFINITY;
  }
  return 0;
}

// status of run_ensemble beyond the engine error codes (which are positive)
#define PSFMC_ENS_OK 0
#define PSFMC_ENS_CALLBACK (-1)
#define PSFMC_ENS_POS_INF (-2)
#define PSFMC_ENS_POS_NAN (-3)
#define PSFMC_ENS_LNPROB_NAN (-4)

// q, scratch: [k/2][D], lnl: [k/2] buffers the engine reads / writes (page-locked where the
// caller can: the engine then skips its staging copies and replays one captured graph per
// call)
inline int run_ensemble(const LnlikeCalls &eng, const psfmc_prior_plan *pl, psfmc_ensemble *e,
                        long long n_iter, double *q, double *lnl, double *scratch,
                        LnpostWork &wk) {
  const long long k = e->n_walkers, D = e->n_dim, half = k / 2;
  NumpyMT19937 mt{e->mt_key, e->mt_pos};
  std::vector<double> zz((size_t)half), newlnp((size_t)half), lzz((size_t)half),
      lu((size_t)half), zz_next((size_t)half), lu_next((size_t)half);
  std::vector<long long> partner((size_t)half), partner_next((size_t)half);
  wk.prof = EnsProfile();
  const double a = e->a, dm1 = (double)D - 1.0;
  const long long thin = e->thin > 0 ? e->thin : 1;
  HostPool &pool = HostPool::instance();
  // The draws of a half-step in emcee's order: rand(Ns) for the stretch factors,
  // randint(Nc, size=Ns) for the partners, then -- after the posterior call, which draws
  // nothing -- rand(Ns) for the acceptance. None of them depends on a result, so the draws
  // of the NEXT half-step are made while the GPU works on this one (one serial MT19937
  // stream: ~40 us per 2048 walkers that the GPU would otherwise wait for). The last
  // half-step of the run draws nothing ahead: the state handed back is exactly the one
  // a Python loop would hold.
  auto draw = [&](std::vector<double> &z, std::vector<long long> &p, std::vector<double> &u,
                  long long ns, long long nc) {
    mt.fill_double(z.data(), ns);
    for (long long i = 0; i < ns; ++i) {
      volatile double t = (a - 1.0) * z[i];               // (no contraction into an FMA)
      const double t1 = t + 1.0;
      volatile double sq = t1 * t1;
      z[i] = sq / a;
    }
    mt.fill_bounded(p.data(), ns, (uint32_t)nc);
    mt.fill_double(u.data(), ns);
  };
  // Chain storage. emcee's layout is (walker, iteration, D): one iteration scatters k short
  // rows over the whole array (141 us per 2048 walkers, measured: every row a cache miss).
  // Iterations are collected in a staging block [T][k][D] and written out per walker as
  // runs of T consecutive iterations (by the host threads).
  long long T = 1;
  if (e->chain || e->lnprob_chain) {
    T = (long long)(4u << 20) / (k * D * (long long)sizeof(double));
    T = T < 1 ? 1 : (T > 64 ? 64 : T);
  }
  std::vector<double> stage(e->chain ? (size_t)(T * k * D) : 0);
  std::vector<double> stage_lnp(e->lnprob_chain ? (size_t)(T * k) : 0);
  long long staged = 0, stage_first = 0;
  auto flush = [&]() {
    if (!staged) return;
    pool.parallel_rows(k, 256, [&](long long lo, long long hi) {
      for (long long w = lo; w < hi; ++w)
        for (long long t = 0; t < staged; ++t) {
          if (e->chain)
            memcpy(e->chain + ((size_t)w * e->chain_len + stage_first + t) * D,
                   stage.data() + ((size_t)t * k + w) * D, (size_t)D * sizeof(double));
          if (e->lnprob_chain)
            e->lnprob_chain[(size_t)w * e->chain_len + stage_first + t] =
                stage_lnp[(size_t)t * k + w];
        }
    });
    staged = 0;
  };
  // An iteration is stored while the GPU works on the first half-step of the NEXT one (the
  // positions do not change before that half-step's acceptance); the last one, and
  // whatever is still staged, on the way out -- also of an error return.
  long long pending_it = -1;
  auto store_pending = [&]() {
    if (pending_it < 0) return;
    const long long ind = e->chain_start + pending_it / thin;
    pending_it = -1;
    if (ind >= e->chain_len || !(e->chain || e->lnprob_chain)) return;
    if (!staged) stage_first = ind;
    if (e->chain)
      memcpy(stage.data() + (size_t)staged * k * D, e->pos, (size_t)k * D * sizeof(double));
    if (e->lnprob_chain)
      memcpy(stage_lnp.data() + (size_t)staged * k, e->lnprob, (size_t)k * sizeof(double));
    if (++staged == T) flush();
  };
  struct FlushGuard {
    std::function<void()> fn;
    ~FlushGuard() { fn(); }
  } flush_guard{[&]() {
    store_pending();
    flush();
  }};
  draw(zz, partner, lu, half, k - half);
  for (long long it = 0; it < n_iter; ++it) {
    for (int h = 0; h < 2; ++h) {
      const long long s0 = h == 0 ? 0 : half, c0 = h == 0 ? half : 0;
      const long long ns = h == 0 ? half : k - half, nc = k - ns;
      double *s = e->pos + s0 * D;
      const double *c = e->pos + c0 * D;
      const double t_half = wk.prof.on ? EnsProfile::now() : 0.0;
      std::atomic<int> bad_inf{0}, bad_nan{0};
      pool.parallel_rows(ns, 512, [&](long long lo, long long hi) {
        bool has_inf = false, has_nan = false;
        for (long long i = lo; i < hi; ++i) {
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
        if (has_inf) bad_inf = 1;
        if (has_nan) bad_nan = 1;
      });
      if (bad_inf) return PSFMC_ENS_POS_INF;   // emcee: ValueError
      if (bad_nan) return PSFMC_ENS_POS_NAN;
      const bool more = !(it == n_iter - 1 && h == 1);
      // behind the GPU: the logarithms of this half-step's acceptance test, the next
      // half-step's draws
      const std::function<void()> overlap = [&]() {
        pool.parallel_rows(ns, 512, [&](long long lo, long long hi) {
          for (long long i = lo; i < hi; ++i) {
            lzz[i] = dm1 * log(zz[i]);
            lu[i] = log(lu[i]);
          }
        });
        if (more) draw(zz_next, partner_next, lu_next, nc, ns);
        store_pending();
      };
      if (wk.prof.on) wk.prof.t[0] += EnsProfile::now() - t_half;
      int rc = lnpost_rows(eng, pl, q, ns, D, lnl, newlnp.data(), wk, scratch, &overlap);
      if (rc) return rc;
      const double t_acc = wk.prof.on ? EnsProfile::now() : 0.0;
      double *lnp = e->lnprob + s0;
      std::atomic<int> bad_lnp{0};
      pool.parallel_rows(ns, 512, [&](long long lo, long long hi) {
        for (long long i = lo; i < hi; ++i) {
          if (newlnp[i] != newlnp[i]) {
            bad_lnp = 1;
            continue;
          }
          volatile double part = lzz[i] + newlnp[i];
          const double lnpdiff = part - lnp[i];
          if (lnpdiff > lu[i]) {
            lnp[i] = newlnp[i];
            memcpy(s + i * D, q + i * D, (size_t)D * sizeof(double));
            if (e->n_accepted) e->n_accepted[s0 + i] += 1.0;
          }
        }
      });
      if (bad_lnp) return PSFMC_ENS_LNPROB_NAN;
      if (more) {
        zz.swap(zz_next);
        partner.swap(partner_next);
        lu.swap(lu_next);
      }
      if (wk.prof.on) wk.prof.t[6] += EnsProfile::now() - t_acc;
    }
    if (it % thin == 0) pending_it = it;
  }
  {
    const double t_store = wk.prof.on ? EnsProfile::now() : 0.0;
    store_pending();
    flush();
    if (wk.prof.on) wk.prof.t[7] += EnsProfile::now() - t_store;
  }
  wk.prof.report("ensemble_run");
  return PSFMC_ENS_OK;
}

}  // namespace psfmc
