// Launch plan + launch sequence of the staged path. Shared by the engine
// (engine.cu, real CUDA launches) and by the CPU emulator harness used in the
// non-GPU test tier (tests/emu/, -DPSFMC_EMU), so that what is tested locally is
// the launch sequence that runs on the B200.
#pragma once
#include <stdlib.h>

#include <utility>

#include "kernels_staged.cuh"

namespace psfmc {

template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                          cudaStream_t stream, Args &&...args) {
#ifdef PSFMC_EMU
  (void)stream;
  emu::launch(grid, block, smem, [&] { kernel(args...); });
#else
  kernel<<<grid, block, smem, stream>>>(std::forward<Args>(args)...);
#endif
}

// Launch with a thread-block cluster of `cluster` CTAs along x.
template <typename... KArgs, typename... Args>
inline void launch_kernel_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                  cudaStream_t stream, unsigned cluster, Args &&...args) {
#ifdef PSFMC_EMU
  (void)stream;
  emu::launch(grid, block, smem, [&] { kernel(args...); }, cluster);
#else
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
#endif
}

inline int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

struct StagedPlan {
  Frame fr;
  int RB, CB;               // rows / columns per CTA
  int n_rowblk, n_colblk;   // CTAs per walker in the row / column kernels
  int threads_rows, threads_cols;
  size_t smem_rows, smem_cols;
  size_t scratch_elems_per_walker;  // complex elements
  long long chunk;          // walkers whose scratch is in flight at once
};

inline bool frame_is_pow2(int H, int W) {
  auto ok = [](int v) { return v >= 16 && v <= 1024 && (v & (v - 1)) == 0; };
  return ok(H) && ok(W);
}

// Transform frame of an observation frame that is not a power of two: the smallest
// powers of two that hold the linear convolution with the PSF stamp.
inline int padded_length(int n, int psf_n) {
  int m = 16;
  while (m < n + psf_n - 1) m <<= 1;
  return m;
}

// Frames the engine covers: powers of two 16..1024 directly; any other height and any
// other EVEN width (the reference needs an even width, psfMC/models.py:276) through a
// padded transform frame of at most 1024 x 1024.
inline bool frame_supported(int H, int W, int psf_h, int psf_w) {
  if (frame_is_pow2(H, W)) return true;
  if (H < 1 || W < 2 || (W & 1)) return false;
  return padded_length(H, psf_h) <= 1024 && padded_length(W, psf_w) <= 1024;
}

// low_latency: plans for batches of a few walkers (the float64 rescue engine): many
// small CTAs per walker instead of few large ones
inline StagedPlan make_staged_plan(int H, int W, int n_components, size_t sizeof_real,
                                   double chunk_mbytes, bool low_latency = false) {
  StagedPlan p;
  p.fr.H = H;
  p.fr.W = W;
  p.fr.Wc = W / 2 + 1;
  p.fr.logH = ilog2(H);
  p.fr.logW = ilog2(W);
  p.fr.Hr = H;
  p.fr.Wr = W;
  p.fr.padded = 0;
  p.fr.fy_hi = p.fr.fx_hi = -1;
  p.fr.fy_lo = p.fr.fx_lo = 1 << 30;
  int rb = (low_latency ? 256 : 2048) / W;
  if (rb < 1) rb = 1;
  if (rb > H) rb = H;
  p.RB = rb;
  p.n_rowblk = H / rb;
  p.threads_rows = rb * (W / 8);
  // columns per CTA: at most 2048/H, chosen to minimise padding of the 2*Wc columns
  int ncol = 2 * p.fr.Wc;
  int cbmax = (low_latency ? 256 : 2048) / H;
  if (cbmax < 1) cbmax = 1;
  int best = cbmax, best_waste = 1 << 30;
  for (int cb = cbmax; cb >= (cbmax + 1) / 2 && cb >= 1; --cb) {
    int waste = ((ncol + cb - 1) / cb) * cb - ncol;
    if (waste < best_waste) {
      best_waste = waste;
      best = cb;
    }
  }
  p.CB = best;
  p.n_colblk = (ncol + best - 1) / best;
  p.threads_cols = best * (H / 8);
  size_t csz = 2 * sizeof_real;
  p.smem_rows = (size_t)(rb * (W + 4) + W) * csz +
                (size_t)(n_components * PSFMC_DERIVED_STRIDE + 64) * sizeof(double) +
                (size_t)(n_components * PSFMC_RC_STRIDE) * sizeof(float);
  p.smem_cols = (size_t)(best * H + H) * csz;
  p.scratch_elems_per_walker = (size_t)ncol * H;
  double per_walker = (double)p.scratch_elems_per_walker * csz;
  long long chunk = (long long)(chunk_mbytes * 1048576.0 / per_walker);
  if (chunk < 1) chunk = 1;
  p.chunk = chunk;
  return p;
}

// Origin of the reference's convolution kernel inside the PSF stamp along one axis:
// conv = ifftshift(irfft2(rfft2(img) * rfft2(psf padded at offset (n - psf_n) // 2)))
// (psfMC/utils.py:9-32) is the circular convolution with the stamp pixel
// n // 2 - (n - psf_n) // 2 at lag 0.
inline int kernel_origin(int n, int psf_n) { return n / 2 - (n - psf_n) / 2; }

// Plan for an observation frame Hr x Wr that is not a power of two (see Frame).
inline StagedPlan make_padded_plan(int Hr, int Wr, int psf_h, int psf_w, int n_components,
                                   size_t sizeof_real, double chunk_mbytes,
                                   bool low_latency = false) {
  StagedPlan p = make_staged_plan(padded_length(Hr, psf_h), padded_length(Wr, psf_w),
                                  n_components, sizeof_real, chunk_mbytes, low_latency);
  p.fr.Hr = Hr;
  p.fr.Wr = Wr;
  p.fr.padded = 1;
  const int oy = kernel_origin(Hr, psf_h), ox = kernel_origin(Wr, psf_w);
  p.fr.fy_hi = psf_h - 2 - oy;
  p.fr.fy_lo = Hr - oy;
  p.fr.fx_hi = psf_w - 2 - ox;
  p.fr.fx_lo = Wr - ox;
  return p;
}

// theta -> per-walker constants. Small batches use 32 lanes per component (the
// kernel is then bound by the length of its dependent float64 chain), large ones 8
// (bound by the float64 pipe).
inline void launch_prepare(const Program &prog, const double *theta, long long n_batch,
                           long long ld, int H, int W, int n_components, double *derived,
                           int *psf_sel, double *wscale, float *rconst, cudaStream_t stream,
                           int *hot = nullptr) {
  const int ncomp = n_components > 0 ? n_components : 1;
  const long long ngroups = n_batch * ncomp;
  const int block = 128;
  const char *env = getenv("PSFMC_PREPARE_GROUP");   // tests: 8 | 32 pins the variant
  const int forced = env ? atoi(env) : 0;
  const bool wide = forced ? forced == 32 : ngroups <= 4096;
  // groups are numbered component-major, each component's run padded to whole CTAs
  const int gpc = block / (wide ? 32 : 8);
  const unsigned grid = (unsigned)(ncomp * ((n_batch + gpc - 1) / gpc));
  // theta rows staged in shared memory: one row per group of the CTA
  size_t smem = (size_t)gpc * (size_t)ld * sizeof(double);
  const int stage = smem <= 40 * 1024 ? 1 : 0;
  if (!stage) smem = 0;
  if (wide)
    launch_kernel(prepare_kernel<32>, dim3(grid), dim3(block), smem, stream, prog, theta,
                  n_batch, ld, H, W, derived, psf_sel, wscale, rconst, stage, hot);
  else
    launch_kernel(prepare_kernel<8>, dim3(grid), dim3(block), smem, stream, prog, theta,
                  n_batch, ld, H, W, derived, psf_sel, wscale, rconst, stage, hot);
}

// Device-resident state the launch sequence needs (one per device per precision).
template <typename T>
struct StagedBuffers {
  const Program *prog;        // device copy (row kernels)
  const Program *prog_host;   // host copy with the device's table pointers (prepare)
  const cplx<T> *tw_w, *tw_h;
  const cplx<T> *spec;        // [K][2*Wc][H]
  const T *obs, *ovar;        // [H*W]
  const unsigned char *bad;   // [H*W]
  double *derived;            // [B][ncomp][STRIDE]
  int *psf_sel;               // [B]
  double *wscale;             // [B] per-walker packing scale
  const double *vscale_inv;   // [K] inverse of the scale folded into the V spectra
  cplx<T> *scratch;           // [chunk][2*Wc][H]
  double *partials;           // [B][n_rowblk]
  float *rconst = nullptr;    // [B][ncomp][PSFMC_RC_STRIDE], float32 engines only
};

// Optional image outputs of one chunk (null = not wanted), [chunk][H*W] each.
template <typename T>
struct ImageOutputs {
  T *raw = nullptr, *conv = nullptr, *resid = nullptr, *ivm = nullptr;
};

template <typename T>
inline int count_launches(const StagedPlan &plan, long long n_batch) {
  long long nchunks = (n_batch + plan.chunk - 1) / plan.chunk;
  return (int)(2 + 3 * nchunks);
}

// theta, lnl: device pointers. n_components: host copy of prog->n_components.
// ps_only renders only the point sources (for the point-source-subtracted image).
template <typename T>
inline void launch_staged_lnlike(const StagedPlan &plan, const StagedBuffers<T> &buf,
                                 int n_components, int precision, const double *theta,
                                 long long n_batch, long long ld, double *lnl,
                                 cudaStream_t stream, bool ps_only = false,
                                 const ImageOutputs<T> *images = nullptr,
                                 long long images_chunk_offset = -1,
                                 cudaEvent_t ev_begin = nullptr,
                                 cudaEvent_t ev_end = nullptr) {
  if (n_batch <= 0) return;
  const Frame fr = plan.fr;
  launch_prepare(*buf.prog_host, theta, n_batch, ld, fr.Hr, fr.Wr, n_components, buf.derived,
                 buf.psf_sel, buf.wscale, buf.rconst, stream);
  if (ev_begin) cudaEventRecord(ev_begin, stream);   // the three row/column kernels
  for (long long start = 0; start < n_batch; start += plan.chunk) {
    long long nb = n_batch - start < plan.chunk ? n_batch - start : plan.chunk;
    const double *der = buf.derived + start * n_components * PSFMC_DERIVED_STRIDE;
    ImageOutputs<T> img;
    if (images && (images_chunk_offset < 0 || images_chunk_offset == start)) img = *images;
    dim3 grid_rows(plan.n_rowblk, (unsigned)nb), grid_cols(plan.n_colblk, (unsigned)nb);
    // float32 lnL path at 256 / 512 columns or rows: instances with the transform
    // length fixed at compile time (the generic passes are bound by index arithmetic)
    const bool fast32 = sizeof(T) == 4 && !ps_only;
    const int lw = fast32 ? fr.logW : 0, lh = fast32 ? fr.logH : 0;
    const double *wsc = (const double *)(buf.wscale + start);
    const int *sel = (const int *)(buf.psf_sel + start);
    double *part = buf.partials + start * plan.n_rowblk;
    const float *rcs = buf.rconst ? buf.rconst + start * n_components * PSFMC_RC_STRIDE
                                  : (const float *)nullptr;
#define PSFMC_ROWS_FWD(KERNEL)                                                               \
  launch_kernel(KERNEL, grid_rows, dim3(plan.threads_rows), plan.smem_rows, stream, fr,      \
                plan.RB, buf.prog, der, wsc, precision, (const double *)nullptr,             \
                (const double *)nullptr, buf.tw_w, buf.scratch, img.raw, rcs)
#define PSFMC_COLS(KERNEL)                                                                   \
  launch_kernel(KERNEL, grid_cols, dim3(plan.threads_cols), plan.smem_cols, stream, fr,      \
                plan.CB, buf.tw_h, buf.spec, sel, buf.scratch, (cplx<T> *)nullptr)
#define PSFMC_ROWS_INV(KERNEL)                                                               \
  launch_kernel(KERNEL, grid_rows, dim3(plan.threads_rows), plan.smem_rows, stream, fr,      \
                plan.RB, buf.tw_w, (const cplx<T> *)buf.scratch, buf.obs, buf.ovar, buf.bad, \
                wsc, sel, buf.vscale_inv, part, img.conv, img.resid, img.ivm)
    if (ps_only)
      PSFMC_ROWS_FWD((rows_fwd_kernel<T, PSFMC_SRC_RENDER_PS>));
    else if (lw == 8)
      PSFMC_ROWS_FWD((rows_fwd_kernel<T, PSFMC_SRC_RENDER, 8>));
    else if (lw == 9)
      PSFMC_ROWS_FWD((rows_fwd_kernel<T, PSFMC_SRC_RENDER, 9>));
    else
      PSFMC_ROWS_FWD((rows_fwd_kernel<T, PSFMC_SRC_RENDER>));
    if (lh == 8)
      PSFMC_COLS((cols_kernel<T, PSFMC_COLS_CONV, 8>));
    else if (lh == 9)
      PSFMC_COLS((cols_kernel<T, PSFMC_COLS_CONV, 9>));
    else
      PSFMC_COLS((cols_kernel<T, PSFMC_COLS_CONV>));
    if (lw == 8)
      PSFMC_ROWS_INV((rows_inv_kernel<T, 8>));
    else if (lw == 9)
      PSFMC_ROWS_INV((rows_inv_kernel<T, 9>));
    else
      PSFMC_ROWS_INV((rows_inv_kernel<T>));
#undef PSFMC_ROWS_FWD
#undef PSFMC_COLS
#undef PSFMC_ROWS_INV
  }
  if (ev_end) cudaEventRecord(ev_end, stream);
  {
    int block = 128;
    unsigned grid = (unsigned)((n_batch + block - 1) / block);
    launch_kernel(finalize_kernel, dim3(grid), dim3(block), 0, stream,
                  (const double *)buf.partials, plan.n_rowblk, (const int *)buf.psf_sel,
                  n_batch, lnl);
  }
}

// Frames of 1024 columns/rows need a little more than the default 48 KB of dynamic
// shared memory: raise the limit of every staged kernel of real type T (and of the
// float64 setup kernels) once per device.
template <typename T>
inline int staged_prepare_device() {
  const int limit = 96 * 1024;
  cudaError_t err = cudaSuccess;
  auto raise = [&](auto kernel) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         limit);
    if (e != cudaSuccess) err = e;
  };
  raise(rows_fwd_kernel<T, PSFMC_SRC_RENDER>);
  raise(rows_fwd_kernel<T, PSFMC_SRC_RENDER_PS>);
  raise(rows_fwd_kernel<double, PSFMC_SRC_PSFPAD>);
  raise(cols_kernel<T, PSFMC_COLS_CONV>);
  raise(cols_kernel<double, PSFMC_COLS_SETUP>);
  raise(rows_inv_kernel<T>);
  return err == cudaSuccess ? 0 : 1;
}

// Setup: spectra of the K padded PSF / variance frames, always float64.
//   pad_psf, pad_var: [K][H][W] doubles on the device; spec_out: [K][2*Wc][H]
inline void launch_staged_setup(const StagedPlan &plan, const cplx<double> *tw_w,
                                const cplx<double> *tw_h, const double *pad_psf,
                                const double *pad_var, int n_psf,
                                cplx<double> *scratch, cplx<double> *spec_out,
                                cudaStream_t stream) {
  const Frame fr = plan.fr;
  dim3 grid_rows(plan.n_rowblk, n_psf), grid_cols(plan.n_colblk, n_psf);
  launch_kernel(rows_fwd_kernel<double, PSFMC_SRC_PSFPAD>, grid_rows,
                dim3(plan.threads_rows), plan.smem_rows, stream, fr, plan.RB,
                (const Program *)nullptr, (const double *)nullptr, (const double *)nullptr, 0,
                pad_psf, pad_var, tw_w, scratch, (double *)nullptr, (const float *)nullptr);
  launch_kernel(cols_kernel<double, PSFMC_COLS_SETUP>, grid_cols, dim3(plan.threads_cols),
                plan.smem_cols, stream, fr, plan.CB, tw_h, (const cplx<double> *)nullptr,
                (const int *)nullptr, scratch, spec_out);
}

// Pass-ordered twiddle table of fft_line_smem (fft.cuh), computed in long double on
// the host; `tw` has room for L entries, of which twiddle_table_entries() are used.
template <typename T>
inline void fill_twiddles(cplx<T> *tw, int L) {
  const long double pi = 3.14159265358979323846264338327950288L;
  auto put = [&](int pos, long long num, long long den) {   // exp(-2 pi i num / den)
    num %= den;
    long double ang = -2.0L * pi * (long double)num / (long double)den;
    long double c = cosl(ang), sn = sinl(ang);
    if ((4 * num) % den == 0) {   // exact values at the quadrant points
      const int q = (int)((4 * num) / den);
      const long double cq[4] = {1, 0, -1, 0}, sq[4] = {0, -1, 0, 1};
      c = cq[q];
      sn = sq[q];
    }
    tw[pos].x = (T)c;
    tw[pos].y = (T)sn;
  };
  for (int k = 0; k < L; ++k) {
    tw[k].x = (T)1;
    tw[k].y = (T)0;
  }
  const int logL = ilog2(L);
  const int n8 = logL / 3, rem = logL - 3 * n8;
  int pos = 0, Ns = 1;
  for (int s = 0; s < n8; ++s) {
    if (s > 0) {
      for (int r = 1; r < 8; ++r)
        for (int k = 0; k < Ns; ++k) put(pos + (r - 1) * Ns + k, (long long)r * k, 8LL * Ns);
      pos += 7 * Ns;
    }
    Ns <<= 3;
  }
  if (rem == 2)
    for (int r = 1; r < 4; ++r)
      for (int k = 0; k < Ns; ++k) put(pos + (r - 1) * Ns + k, (long long)r * k, 4LL * Ns);
  if (rem == 1)
    for (int k = 0; k < Ns; ++k) put(pos + k, k, 2LL * Ns);
}

}  // namespace psfmc
