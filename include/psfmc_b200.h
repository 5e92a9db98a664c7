/*
 * psfmc_b200 -- C ABI of the B200-native batched likelihood engine.
 *
 * Drop-in boundary for the one hot path of mmechtley/psfMC: what emcee calls once
 * per walker per (half-)iteration,
 *     MultiComponentModel.log_posterior            psfMC/models.py:193-243
 * minus the priors (which stay in Python): render the parametric model
 * (psfMC/models.py:245-253), FFT-convolve it with the PSF (psfMC/utils.py:25-32),
 * form the composite inverse-variance map with the PSF-variance term
 * (psfMC/models.py:265-280) and reduce the masked Normal log-likelihood
 * (psfMC/models.py:233-241) -- for a whole batch of parameter vectors per call.
 *
 * Plain C types only: no CUDA, torch or C++ types cross this boundary, so the
 * library can be bound with ctypes (psfmc_b200/_lib.py), cffi, or any FFI.
 * Every entry point returns 0 on success or a PSFMC_ERR_* code; the message is
 * available (per calling thread) from psfmc_last_error(). Numerical failure is
 * NOT an error: a walker whose lnL is NaN/Inf gets -inf, mirroring
 * psfMC/models.py:238-241.
 *
 * Ownership: the caller owns every host buffer; psfmc_engine_create copies what
 * it needs to each device and keeps no caller pointer. Outputs are written into
 * caller-allocated buffers. An engine is not re-entrant: one call at a time.
 */
#ifndef PSFMC_B200_H
#define PSFMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSFMC_ABI_VERSION 2

/* error codes */
#define PSFMC_OK 0
#define PSFMC_ERR_INVALID_ARG 1   /* bad descriptor / shapes / slot indices        */
#define PSFMC_ERR_UNSUPPORTED 2   /* odd width, frame too large, PSF larger than frame */
#define PSFMC_ERR_CUDA 3          /* a CUDA runtime call failed (message has detail) */
#define PSFMC_ERR_NO_DEVICE 4     /* no usable sm_100 device                          */

/* component kinds: psfMC/ModelComponents/{Sky,PointSource,Sersic}.py */
#define PSFMC_SKY 0
#define PSFMC_POINT 1
#define PSFMC_SERSIC 2

/* component flags */
#define PSFMC_FLAG_ANGLE_DEGREES 1 /* Sersic(angle_degrees=True)        Sersic.py:30-39   */
#define PSFMC_FLAG_BILINEAR 2      /* PointSource(shift_method='bilinear'); default is
                                      'lanczos3'                         PointSource.py:18-22 */

/* parameter slots of a component (index into psfmc_component.slot[]) */
#define PSFMC_P_ADU 0    /* Sky.adu                                              */
#define PSFMC_P_X 0      /* xy[0], 0-based pixel coordinates                      */
#define PSFMC_P_Y 1      /* xy[1]                                                */
#define PSFMC_P_MAG 2
#define PSFMC_P_REFF 3   /* Sersic only from here on                              */
#define PSFMC_P_REFF_B 4
#define PSFMC_P_INDEX 5
#define PSFMC_P_ANGLE 6
#define PSFMC_NSLOTS 7

#define PSFMC_MAX_COMPONENTS 32

/* precision modes */
#define PSFMC_PREC_FP64 0 /* everything in float64: parity mode, gated at 1e-10
                             relative against the all-float64 oracle (mode M3)     */
#define PSFMC_PREC_FP32 1 /* float32 render + FFT, float64 chi-square accumulation:
                             the throughput mode, gated at a stated |dlnL|. A walker
                             whose float32 transform is not trustworthy (a convolved
                             model variance below -2^-8 of a pixel's own variance: the
                             rounding noise of a model with a ~1e5 dynamic range) is
                             returned as non-finite and, through psfmc_lnlike_batch,
                             repeated in float64 (PSFMC_DESC_NO_FP64_RESCUE)        */
#define PSFMC_PREC_FP64_RAWF32 2 /* as FP64 but the raw model is rounded to float32
                             after each component is added, which is what the
                             reference does for float32 FITS inputs on numpy 1.x
                             (oracle mode M2; psfMC/models.py:249)                  */

/* descriptor flags */
#define PSFMC_DESC_NO_FP64_RESCUE 1 /* PSFMC_PREC_FP32 only: psfmc_lnlike_batch normally
                             repeats every walker whose float32 result is non-finite in
                             float64 on the GPU (a model whose dynamic range exceeds what
                             a float32 transform resolves gives a negative variance, where
                             the float64 reference is finite); this flag turns that off */

#define PSFMC_DESC_LOW_LATENCY 2 /* the engine will see batches of a few walkers: the
                             row/column kernels are planned with many small CTAs per
                             walker instead of few large ones (the float64 rescue
                             engine is built this way)                               */

/* images psfmc_render_batch can return: the blobs of psfMC/models.py:222-226 */
#define PSFMC_IMG_RAW_MODEL 1u
#define PSFMC_IMG_CONVOLVED_MODEL 2u
#define PSFMC_IMG_RESIDUAL 4u
#define PSFMC_IMG_COMPOSITE_IVM 8u
#define PSFMC_IMG_POINT_SOURCE_SUBTRACTED 16u

/* One parameter of one component: either a constant or an index into theta.
 * (ComponentBase.assign_stochastic: priors vs constants, ComponentBase.py:26-35) */
typedef struct psfmc_slot {
  int32_t theta_index; /* >= 0: value = theta[theta_index]; < 0: value = value */
  int32_t reserved;
  double value;
} psfmc_slot;

typedef struct psfmc_component {
  int32_t kind;  /* PSFMC_SKY / PSFMC_POINT / PSFMC_SERSIC */
  int32_t flags; /* PSFMC_FLAG_* */
  psfmc_slot slot[PSFMC_NSLOTS];
} psfmc_component;

/* Everything psfMC's Configuration + PSFSelector prepare once per model
 * (Configuration.py:38-52, PSFSelector.py:16-43), plus the component program. */
typedef struct psfmc_desc {
  int32_t abi_version; /* PSFMC_ABI_VERSION */
  int32_t height;      /* observation frame, rows                            */
  int32_t width;       /* observation frame, columns: EVEN, like the reference
                          (psfMC/models.py:276). Powers of two 16..1024 are
                          transformed directly; any other size goes through a
                          zero-padded power-of-two transform frame (needs
                          height + psf_height - 1 <= 1024, same for the width)
                          whose linear convolution is folded back modulo
                          (height, width): the reference's circular convolution at
                          the image size (psfMC/utils.py:25-32), exactly           */
  const double *obs_data;  /* [height*width] row-major                          */
  const double *obs_var;   /* [height*width] 1/ivm, +inf at data-bad pixels      */
  const uint8_t *bad_px;   /* [height*width] nonzero = excluded from the sum     */
  int32_t n_psf;           /* K >= 1                                             */
  int32_t psf_height;      /* <= height                                          */
  int32_t psf_width;       /* <= width                                           */
  const double *psf;       /* [K][psf_height*psf_width] normalised PSFs          */
  const double *psf_var;   /* [K][psf_height*psf_width] variance maps            */
  double mag_zeropoint;
  int32_t n_components;
  const psfmc_component *components; /* in model order (models.py:245-253)      */
  psfmc_slot psf_index;    /* PSFSelector.psf_index; constant 0 when K == 1      */
  int32_t precision;       /* PSFMC_PREC_*                                       */
  int32_t n_devices;       /* 0: use the current CUDA device only                */
  const int32_t *devices;  /* CUDA ordinals; batches are split contiguously       */
  int32_t max_batch;       /* hint: largest B per call (0 = grow on demand)      */
  int32_t flags;           /* PSFMC_DESC_*                                        */
} psfmc_desc;

typedef struct psfmc_engine psfmc_engine;

/* Build an engine: uploads the constants to every device, transforms the padded
 * PSFs and variance maps on the device (pad offset pad//2 as utils.py:9-22, the
 * ifftshift of utils.py:32 folded into the spectra) and sizes the scratch. */
int psfmc_engine_create(const psfmc_desc *desc, psfmc_engine **out);
void psfmc_engine_destroy(psfmc_engine *engine);

/* lnL for B parameter vectors held in HOST memory (theta[b*ld + j], ld >= D).
 * Blocking. Rows are split over the engine's devices; the only cross-device
 * traffic is the per-walker lnL copied back. lnl_out[b] = -inf for non-finite
 * results. In PSFMC_PREC_FP32 a non-finite float32 result is first repeated in
 * float64 on the engine's first device (PSFMC_DESC_NO_FP64_RESCUE turns that off).
 * This is what the pool-like map object calls for emcee. */
int psfmc_lnlike_batch(psfmc_engine *engine, const double *theta, int64_t n_batch,
                       int64_t ld, double *lnl_out);

/* The same call in two halves, so that the caller's own per-batch host work (the
 * priors, psfMC/models.py:205-211) runs while the GPU computes: _begin copies / enqueues
 * and returns, _end waits and finishes (including the float64 repeat). theta and
 * lnl_out must stay valid and untouched until _end has returned; one batch in flight
 * per engine. psfmc_lnlike_batch(...) == _begin(...) followed by _end(). */
int psfmc_lnlike_batch_begin(psfmc_engine *engine, const double *theta, int64_t n_batch,
                             int64_t ld, double *lnl_out);
int psfmc_lnlike_batch_end(psfmc_engine *engine);

/* Same computation with DEVICE-resident theta and lnL on device `device_slot`
 * (index into the engine's device list), enqueued on `cuda_stream` (a
 * cudaStream_t passed as void*; NULL = the legacy default stream). Asynchronous:
 * returns after the launches are enqueued. */
int psfmc_lnlike_batch_device(psfmc_engine *engine, int32_t device_slot,
                              const double *theta_dev, int64_t n_batch, int64_t ld,
                              double *lnl_dev, void *cuda_stream);

/* ---- lnL gather over peer memory (one process per GPU, one node) -------------------
 * The pool.map of psfMC/fitting.py:55-58 for a job launched with one rank per GPU:
 * every rank evaluates its contiguous rows of a (half-)ensemble and needs ALL lnL
 * values before it can propose the next one. Each rank owns a mailbox on its GPU
 * ([2][capacity] doubles + one flag per rank); the ranks map one another's mailboxes
 * through CUDA IPC (NVLink peer access), and a call publishes this rank's results with
 * plain stores into EVERY rank's mailbox, signals with one system-scope flag per peer and
 * waits for the peers' flags -- no NCCL, no host round trip. The fused 128 x 128 kernel
 * stores its results in the peers' mailboxes itself (the gather is part of the lnL
 * kernel; one 32-thread kernel behind it exchanges the flags), the other paths publish
 * theirs from one small kernel. Mailboxes alternate between two halves from call to call, so a fast rank
 * can never overwrite values a slow rank is still reading.
 *   psfmc_peer_create   allocates the mailbox (same capacity on every rank) and returns
 *                       its 64-byte IPC handle for the caller to all-gather
 *   psfmc_peer_connect  maps the mailboxes of all `world` ranks (handles[world][64])
 *   psfmc_lnlike_batch_exchange  evaluates this rank's n_rows rows (device-resident
 *                       theta) on `cuda_stream`, publishes them at position row_offset of
 *                       the gathered vector, waits for all ranks and copies the n_total
 *                       gathered values to gathered_dev (device memory; may be null).
 *                       Asynchronous; every rank must make the same sequence of calls. */
#define PSFMC_PEER_HANDLE_BYTES 64
#define PSFMC_PEER_MAX_RANKS 16
int psfmc_peer_create(psfmc_engine *engine, int64_t capacity, void *handle_out);
int psfmc_peer_connect(psfmc_engine *engine, int32_t rank, int32_t world,
                       const void *handles);
int psfmc_lnlike_batch_exchange(psfmc_engine *engine, const double *theta_dev,
                                int64_t n_rows, int64_t ld, int64_t row_offset,
                                int64_t n_total, double *gathered_dev, void *cuda_stream);
/* Where the gathered vector of the LAST exchange sits in this rank's mailbox (device
 * memory; valid until the exchange after the next one reuses that half). Saves the copy
 * to gathered_dev when the consumer can read it in place. */
int psfmc_peer_gathered(psfmc_engine *engine, double **gathered_dev_out);

/* The blob images of psfMC/models.py:213-226 for B parameter vectors: for every
 * image selected in `which` (ascending bit order), out receives
 * [n_selected][B][height*width] doubles (host memory). */
int psfmc_render_batch(psfmc_engine *engine, const double *theta, int64_t n_batch,
                       int64_t ld, uint32_t which, double *out);

/* Posterior-image accumulation on the device (psfMC/models.py:74-97 averages five
 * images per sample; psfMC/analysis/images.py:74-83 re-renders every database row):
 * for every image selected in `which` (ascending bit order) sums_out receives
 * [n_selected][height*width] doubles = the SUM over the B parameter vectors, added
 * up in float64 on the device. The composite IVM is summed as 1/ivm, i.e. in
 * variance space, exactly like the reference's running mean. Only the sums cross
 * the bus (not B images). */
int psfmc_accumulate_batch(psfmc_engine *engine, const double *theta, int64_t n_batch,
                           int64_t ld, uint32_t which, double *sums_out);

/* ---- batched log-priors on the host --------------------
 * The priors stay scipy.stats objects on the Python side (psfMC/distributions.py:
 * 115-128); what crosses this boundary is, per theta column, a closed-form family with
 * constants the caller computed with numpy/scipy, so that only IEEE-exact operations
 * run here and the result is bit-identical to rv_frozen.logpdf (the caller checks that
 * on its first batch). Replaces the per-walker scipy calls of
 * ComponentBase.log_priors (psfMC/ModelComponents/ComponentBase.py:121-129),
 * Sersic.log_priors (Sersic.py:41-45) and MultiComponentModel.log_priors
 * (psfMC/models.py:187-191). No GPU involved. */
#define PSFMC_PRIOR_OTHER 0   /* column evaluated by the caller (any other scipy family) */
#define PSFMC_PRIOR_UNIFORM 1 /* scipy.stats.uniform(loc, scale)                        */
#define PSFMC_PRIOR_NORMAL 2  /* scipy.stats.norm(loc, scale)                           */
#define PSFMC_PRIOR_WEIBULL_MIN 3 /* scipy.stats.weibull_min(c, loc, scale): log(c) +
                                   * xlogy(c - 1, x) - pow(x, c) with the C library's log /
                                   * pow -- what numpy 1.21 (the reference's pin) calls;
                                   * numpy >= 1.22 on AVX-512 hosts uses SVML's pow instead,
                                   * which differs in the last bit for ~5 % of the arguments.
                                   * The Python caller accepts this family only if it agrees
                                   * with rv_frozen.logpdf to 4 ulps on its first batch, and
                                   * keeps it out (PSFMC_PRIOR_OTHER) in strict mode.        */
typedef struct psfmc_prior_column {
  int32_t family;      /* PSFMC_PRIOR_*                                                 */
  int32_t theta_index; /* column of theta                                               */
  int32_t valid;       /* scipy's _argcheck(...) & (scale > 0)                          */
  int32_t reserved;
  double loc, scale;
  double log_scale;    /* numpy.log(scale)                                              */
  double log_norm;     /* NORMAL: scipy's log(sqrt(2 pi)) constant                      */
  double shape;        /* WEIBULL_MIN: c                                                */
  double log_shape;    /* WEIBULL_MIN: numpy.log(c)                                     */
} psfmc_prior_column;
/* logp_out[b*ld_out + c] for every column c whose family is not PSFMC_PRIOR_OTHER. */
int psfmc_prior_columns(const psfmc_prior_column *columns, int32_t n_columns,
                        const double *theta, int64_t n_batch, int64_t ld, double *logp_out,
                        int64_t ld_out);
typedef struct psfmc_prior_term {   /* one prior: n_columns consecutive columns of logp */
  int32_t component, first_column, n_columns, reserved;
} psfmc_prior_term;
typedef struct psfmc_prior_rule {   /* component's joint prior is -inf where b > a      */
  int32_t component, a_index, b_index, reserved;   /* theta columns, or -1: the constant */
  double a_value, b_value;
} psfmc_prior_rule;
/* lnprior_out[b] = sum over components (model order) of the sum over the component's
 * terms (given in evaluation order, grouped by ascending component), in exactly the
 * order the reference adds them. */
int psfmc_prior_sum(const double *logp, int64_t n_batch, int64_t ld_logp, const double *theta,
                    int64_t ld, const psfmc_prior_term *terms, int32_t n_terms,
                    const psfmc_prior_rule *rules, int32_t n_rules, int32_t n_components,
                    double *lnprior_out);

/* ---- the sampler's inner loop on the host side of the library ----
 * psfMC hands `MultiComponentModel.log_posterior` to emcee 2.x's EnsembleSampler
 * (psfMC/fitting.py:55-78); per iteration emcee evaluates two sequentially dependent
 * half-ensembles (stretch move, emcee/ensemble.py: _propose_stretch). At the reference
 * example's ensemble size (250 walkers, examples/run_example.py:9) the GPU needs ~35 us
 * per half-ensemble and a Python loop ~300 us, so the loop itself is offered here:
 * proposals, log-priors, lnL (psfmc_lnlike_batch_begin / _end with the priors evaluated
 * while the GPU computes), acceptance and chain storage for n_iterations without leaving
 * the library. Random numbers come from numpy.random.RandomState's own generator state
 * (MT19937 key + position, handed over and returned), consumed in emcee's call order --
 * rand(Ns), randint(Nc, size=Ns), rand(Ns) per half-step -- with numpy's algorithms
 * (53-bit doubles from two draws, masked rejection for bounded integers), so a seeded
 * run continues bit for bit the stream a Python emcee loop would have drawn. */
typedef struct psfmc_prior_plan {
  const psfmc_prior_column *columns; /* [n_columns], one per theta column             */
  const psfmc_prior_term *terms;     /* [n_terms], see psfmc_prior_sum                 */
  const psfmc_prior_rule *rules;     /* [n_rules]                                      */
  int32_t n_columns, n_terms, n_rules, n_components;
  /* columns of family PSFMC_PRIOR_OTHER are filled by the caller (any scipy family, custom
   * Distribution subclasses, discrete priors): logp[b*ld_logp + c]; non-zero return
   * aborts the run. May be null when no column is OTHER. */
  int (*other_columns)(void *user, const double *theta, int64_t n_batch, int64_t ld,
                       double *logp, int64_t ld_logp);
  void *user;
} psfmc_prior_plan;

/* lnpost_out[b] = lnL + lnprior, or -inf where either is not finite
 * (psfMC/models.py:205-211, 238-243). */
int psfmc_lnpost_batch(psfmc_engine *engine, const psfmc_prior_plan *priors,
                       const double *theta, int64_t n_batch, int64_t ld, double *lnpost_out);

/* The same for a job with one process per GPU (psfmc_peer_create / _connect first): every
 * rank passes the SAME n_batch rows, evaluates its contiguous share of them and receives the
 * values of all rows -- lnL gathered over peer memory, host buffers in and out. priors may
 * be null: lnpost_out is then the lnL itself. Non-finite float32 results are -inf (no
 * float64 repeat). Every rank must make the same sequence of calls. */
int psfmc_lnpost_batch_sharded(psfmc_engine *engine, const psfmc_prior_plan *priors,
                               const double *theta, int64_t n_batch, int64_t ld,
                               double *lnpost_out);

typedef struct psfmc_ensemble {
  int64_t n_walkers;     /* k, even                                                    */
  int64_t n_dim;         /* D = row length of pos                                      */
  double a;              /* stretch scale (emcee default 2.0)                          */
  double *pos;           /* [k][D] in / out                                            */
  double *lnprob;        /* [k]    in / out (must be the lnpost of pos on entry)       */
  uint32_t *mt_key;      /* [624]  numpy RandomState.get_state()[1], in / out          */
  int32_t *mt_pos;       /*        ... get_state()[2], in / out                        */
  double *chain;         /* [k][chain_len][D] (emcee's layout) or null                 */
  double *lnprob_chain;  /* [k][chain_len] or null                                     */
  int64_t chain_len;     /* allocated iterations per walker                            */
  int64_t chain_start;   /* index the first stored iteration goes to                   */
  int64_t thin;          /* store every thin-th iteration (>= 1)                       */
  double *n_accepted;    /* [k], incremented (emcee keeps it as float64)               */
  int64_t flags;         /* PSFMC_ENS_*                                                */
} psfmc_ensemble;
/* One process per GPU (torchrun): every rank makes the same psfmc_ensemble_run call on the
 * same ensemble and random state -- the proposals are then identical on all ranks -- and
 * evaluates only its contiguous share of every half-ensemble; the lnL of all rows is
 * gathered over peer memory (psfmc_peer_create / _connect first; the mailbox capacity must
 * cover n_walkers / 2). Non-finite float32 results are -inf here (no float64 repeat, as
 * for every device-pointer call). */
#define PSFMC_ENS_SHARDED 1
/* Proposals, log-priors and acceptance on the device, around the lnL kernels on one stream;
 * the host only draws the random numbers (and the logarithms of the acceptance test) ahead
 * and enqueues -- no host round trip per half-ensemble. Needs a single-device engine and a
 * prior plan without PSFMC_PRIOR_OTHER columns (else PSFMC_ERR_UNSUPPORTED: the caller
 * falls back to the host loop). Together with PSFMC_ENS_SHARDED every rank runs this loop
 * on its own device -- the proposals are identical on all ranks -- and the lnL kernels
 * store their results into every rank's mailbox themselves. Same chain as the host loop up to the last bits of the
 * Weibull columns' log / pow (the device's) and the missing float64 repeat. */
#define PSFMC_ENS_DEVICE 2
/* Advances the ensemble by n_iterations stretch-move iterations. Errors: a proposal with
 * an infinite / NaN coordinate (emcee raises ValueError), a failing callback. */
int psfmc_ensemble_run(psfmc_engine *engine, const psfmc_prior_plan *priors,
                       psfmc_ensemble *ensemble, int64_t n_iterations);
/* The generator on its own (tests): kind 0 = n doubles of RandomState.random_sample,
 * kind 1 = n values of RandomState.randint(bound) (as doubles). */
int psfmc_rng_fill(uint32_t *mt_key, int32_t *mt_pos, int32_t kind, int64_t n, int64_t bound,
                   double *out);

/* Introspection (roofline bookkeeping for bench.py). */
typedef struct psfmc_info {
  int32_t height, width, n_components, n_sersic, n_point, n_psf, precision;
  int32_t n_devices;
  int32_t path;             /* 0 = staged row/column passes through L2/HBM,
                               1 = fused single-kernel shared-memory path (128 x 128),
                               2 = fused four-CTA-cluster path, frame split over the
                                   shared memory of four SMs (256 x 256),
                               3 = tiled path: 512 x 512 as 4 x 4 interleaved 128 x 128
                                   sub-images through the halves of the fused kernel
                                   and a combine kernel                              */
  int32_t kernels_per_call; /* kernels launched per lnlike call per device        */
  double flops_per_eval;    /* 10 N log2 N + (30 n_sersic + 16) N     */
  double fft_flops_per_eval;   /* 10 N log2 N                                     */
  double hbm_bytes_per_eval;   /* algorithmic bytes through L2/HBM per walker     */
  int64_t launches_total;   /* kernels launched by this engine since creation     */
  int32_t kappa_table;      /* 1: the Sersic kappa comes from the Chebyshev table built
                               and verified at creation; 0: from the iteration       */
  int32_t rescued_total;    /* walkers re-evaluated in float64 by psfmc_lnlike_batch
                               (PSFMC_PREC_FP32, see PSFMC_DESC_NO_FP64_RESCUE)      */
  int32_t graph_replays;    /* psfmc_lnlike_batch calls served by replaying a captured
                               CUDA graph (single-device float32 engines)            */
  int32_t rescued_on_device;   /* of rescued_total: repeated inside such a graph by its
                               conditional node, without a host round trip           */
} psfmc_info;
int psfmc_engine_info(const psfmc_engine *engine, psfmc_info *info);

/* Device-side timing of the dominant kernel of the lnL path (the fused kernel, or
 * the three staged row/column kernels together): when enabled, every launch is
 * bracketed by CUDA events on the stream it is launched on. psfmc_engine_profile_read
 * synchronises, returns the summed duration (ms) and the number of launches since
 * the last read, and resets both. For bench.py's roofline; off by default. */
int psfmc_engine_profile(psfmc_engine *engine, int32_t enable);
int psfmc_engine_profile_read(psfmc_engine *engine, double *kernel_ms_out,
                              int64_t *kernel_launches_out);

/* Measured FP32 FMA throughput of one device (TFLOP/s, FMA = 2 FLOP): the
 * denominator of the FP32 roofline (MEASURED_PEAKS.json has no FP32 entry). */
int psfmc_fp32_peak_probe(int32_t device, double *tflops_out, double *ms_out);

const char *psfmc_last_error(void);
int psfmc_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PSFMC_B200_H */
