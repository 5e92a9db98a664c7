// Staged row/column path: works for every supported frame (powers of two 16..1024
// directly, other sizes through a padded transform frame, see Frame) in both precisions; the row spectra of each walker are staged through a
// global scratch buffer that stays L2-resident for the batch chunk in flight.
//
//   rows_fwd : render RB rows -> z = raw + i raw^2 -> FFT_W -> split into the row
//              spectra of the two real images (A: raw, B: raw^2) -> scratch,
//              transposed to column-major so that the column pass is contiguous
//   cols     : FFT_H of each retained column, multiply by the PSF / PSF-variance
//              spectrum, inverse FFT_H, in place in scratch
//   rows_inv : rebuild Y = A' + i B' with Hermitian extension, inverse FFT_W;
//              Re = convolved model, Im = model variance -> residual, composite
//              IVM, masked chi-square terms, float64 partial sums per row block
//   finalize : fixed-order sum of the partials, -0.5 factor, non-finite -> -inf
//
// Scratch layout per walker: S[c][y], c in [0, 2*Wc): c < Wc is column kx = c of
// image A, c >= Wc is column kx = c - Wc of image B; y = 0..H-1 contiguous.
// Spectrum layout per PSF: Spec[c][ky], same column index, including the factor
// (-1)^(kx+ky) / (H*W) (the reference's ifftshift, psfMC/utils.py:32, and the
// inverse-FFT normalisation).
#pragma once
#include "common.cuh"
#include "fft.cuh"
#include "render.cuh"
#include "tma.cuh"

namespace psfmc {

// padding between the rows of a row-kernel tile, in elements (32 bytes)
#define PSFMC_ROW_PAD(T) (32 / (int)sizeof(cplx<T>))

#define PSFMC_SRC_RENDER 0      // rows come from the model renderer
#define PSFMC_SRC_PSFPAD 1      // rows come from padded PSF / variance frames (setup)
#define PSFMC_SRC_RENDER_PS 2   // renderer, point sources only

template <typename T>
__device__ __forceinline__ void load_twiddles(cplx<T> *tw_s, const cplx<T> *tw_g, int L,
                                              int tid, int nthreads) {
  for (int k = tid; k < L; k += nthreads) tw_s[k] = tw_g[k];
}

// ------------------------------------------------------------- rows_fwd --
// grid = (H / RB, B), block = RB * (W / 8)
// dynamic smem: (RB*W + W) cplx<T> + n_components*STRIDE doubles
// LOGW != 0: row length fixed at compile time (specialised instances for the
// throughput path at 256 and 512 columns), 0: taken from `fr`.
// (float32: at most 40 registers, six CTAs of 256 threads per SM -- the kernel is bound by
// occupancy: 66 registers 543 us, 56 registers 537 us, 40 registers 520 us per 255-walker
// chunk at 512 x 512; no spills)
template <typename T, int SRC, int LOGW = 0>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? 6 : 1) rows_fwd_kernel(Frame fr_rt, int RB, const Program *__restrict__ prog,
                                const double *__restrict__ derived,
                                const double *__restrict__ wscale, int precision,
                                const double *__restrict__ pad_a,
                                const double *__restrict__ pad_b,
                                const cplx<T> *__restrict__ tw_w,
                                cplx<T> *__restrict__ scratch,
                                T *__restrict__ raw_out,
                                const float *__restrict__ rconst = nullptr) {
  // rconst (optional, float32 render): the per-walker float32 Sersic constants the
  // prepare kernel wrote ([B][ncomp][PSFMC_RC_STRIDE]); without them every thread
  // derives them from the float64 ones (a float64 log2 per Sersic)
  PSFMC_DYN_SMEM(smem_raw);
  Frame fr = fr_rt;
  if (LOGW) {
    fr.logW = LOGW;
    fr.W = 1 << LOGW;
    fr.Wc = (1 << LOGW) / 2 + 1;
  }
  const int W = fr.W, H = fr.H, Wc = fr.Wc;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const long long b = blockIdx.y;
  const int y0 = blockIdx.x * RB;
  // rows are PITCH elements apart (32 bytes of padding): the transposed accesses
  // below (lanes over the RB rows of one kx) then hit different banks
  const int PITCH = W + PSFMC_ROW_PAD(T);
  const int logRB = 31 - __clz(RB);          // RB is a power of two
  cplx<T> *tile = reinterpret_cast<cplx<T> *>(smem_raw);
  cplx<T> *tw_s = tile + RB * PITCH;
  double *der_s = reinterpret_cast<double *>(tw_s + W);
  // float32 constants behind the float64 ones (16-byte aligned: the stride is even)
  float *rc_s = reinterpret_cast<float *>(
      der_s + (prog ? prog->n_components : 0) * PSFMC_DERIVED_STRIDE);

  load_twiddles<T>(tw_s, tw_w, W, tid, nthreads);
  int ncomp = 0;
  if (SRC != PSFMC_SRC_PSFPAD) {
    ncomp = prog->n_components;
    const double *der_b = derived + b * ncomp * PSFMC_DERIVED_STRIDE;
    for (int k = tid; k < ncomp * PSFMC_DERIVED_STRIDE; k += nthreads) der_s[k] = der_b[k];
    if (rconst)
      for (int k = tid; k < ncomp * PSFMC_RC_STRIDE; k += nthreads)
        rc_s[k] = rconst[b * ncomp * PSFMC_RC_STRIDE + k];
    __syncthreads();
  }

  // ---- fill the tile: z = a + i b  (b = wscale * raw^2 when rendering)
  const int npx = RB * W;
  const T wsc = (SRC != PSFMC_SRC_PSFPAD) ? (T)wscale[b] : (T)1;
  if (SRC == PSFMC_SRC_PSFPAD) {
    for (int e = tid; e < npx; e += nthreads) {
      int r = e >> fr.logW, x = e & (W - 1);
      long long g = (b * H + (y0 + r)) * (long long)W + x;
      tile[r * PITCH + x] = mk<T>((T)pad_a[g], (T)pad_b[g]);
    }
  } else if (sizeof(T) == 8 || SRC == PSFMC_SRC_RENDER_PS) {
    const bool round_f32 = (precision == PSFMC_PREC_FP64_RAWF32);
    for (int e = tid; e < npx; e += nthreads) {
      int r = e >> fr.logW, x = e & (W - 1);
      // (padded frames: nothing is rendered outside the observation frame)
      double val = (x < fr.Wr && y0 + r < fr.Hr)
                       ? raw_pixel_f64(prog, der_s, x, y0 + r, round_f32,
                                       SRC == PSFMC_SRC_RENDER_PS)
                       : 0.0;
      T a = (T)val;
      tile[r * PITCH + x] = mk<T>(a, a * a * wsc);
      if (raw_out) raw_out[(b * H + (y0 + r)) * (long long)W + x] = a;
    }
  } else {
    // float32 throughput path: Sky + Sersic in float, point-source taps in double.
    // RB*W == 8*nthreads, so every thread owns exactly 8 pixels.
    // Pixels are handled in pairs (packed FADD2 / FFMA2 arithmetic): pair j holds the
    // thread's elements 2j and 2j+1.
    cplx<float> acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = mk<float>(0.0f, 0.0f);
    for (int c = 0; c < ncomp; ++c) {
      const double *d = der_s + c * PSFMC_DERIVED_STRIDE;
      const int kind = prog->kind[c];
      if (kind == PSFMC_SKY) {
        const cplx<float> adu = bcast((float)d[D_SKY_ADU]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = acc[j] + adu;
      } else if (kind == PSFMC_POINT) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int e0 = tid + 2 * j * nthreads, e1 = e0 + nthreads;
          acc[j].x += (float)point_pixel(d, e0 & (W - 1), y0 + (e0 >> fr.logW));
          acc[j].y += (float)point_pixel(d, e1 & (W - 1), y0 + (e1 >> fr.logW));
        }
      } else {
        SersicF32 s;
        if (rconst) {
          const float4 *rc4 = reinterpret_cast<const float4 *>(rc_s + c * PSFMC_RC_STRIDE);
          const float4 q0 = rc4[0], q1 = rc4[1], q2 = rc4[2];
          s.xi = q0.x; s.xf = q0.y; s.yi = q0.z; s.yf = q0.w;
          s.a00 = q1.x; s.a01 = q1.y; s.a10 = q1.z; s.a11 = q1.w;
          s.p = q2.x; s.c0 = q2.y; s.c1 = q2.z; s.kq = q2.w;
        } else {
          s = make_sersic_f32(d);
        }
        if (W == 2 * nthreads) {
          // 512-column frames (RB = 4): pair j is row y0 + j at x = tid and tid + W/2 --
          // the x offsets are computed once per component, the row terms are scalars
          // (same arithmetic as the fused kernel's render, see fused_render16)
          const cplx<float> dx =
              (mk<float>((float)tid, (float)(tid + nthreads)) - bcast(s.xi)) - bcast(s.xf);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float dy = ((float)(y0 + j) - s.yi) - s.yf;
            acc[j] = sersic_pair_f32<false>(s, dx, s.a01 * dy, s.a11 * dy, dy * dy, acc[j]);
          }
          continue;
        }
        const cplx<float> xi = bcast(s.xi), xf = bcast(s.xf), yi = bcast(s.yi),
                          yf = bcast(s.yf);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // (coordinate - integer part of the centre) is exact; its fraction goes last
          const int e0 = tid + 2 * j * nthreads, e1 = e0 + nthreads;
          const cplx<float> fx = mk<float>((float)(e0 & (W - 1)), (float)(e1 & (W - 1)));
          const cplx<float> fy = mk<float>((float)(y0 + (e0 >> fr.logW)),
                                           (float)(y0 + (e1 >> fr.logW)));
          const cplx<float> dx = (fx - xi) - xf;
          const cplx<float> dy = (fy - yi) - yf;
          acc[j] = sersic_pair2_f32(s, dx, dy, acc[j]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int e = tid + i * nthreads;
      int r = e >> fr.logW, x = e & (W - 1);
      float a = (i & 1) ? acc[i >> 1].y : acc[i >> 1].x;
      if (x >= fr.Wr || y0 + r >= fr.Hr) a = 0.0f;   // padded frames
      tile[r * PITCH + x] = mk<T>((T)a, (T)(a * a) * wsc);
      if (raw_out) raw_out[(b * H + (y0 + r)) * (long long)W + x] = (T)a;
    }
  }
  __syncthreads();

  // ---- FFT along x of every row of the tile
  {
    const int tpr = W >> 3;  // threads per row
    const int r = tid >> (fr.logW - 3), tl = tid & (tpr - 1);
    fft_line_smem<T, false, LOGW>(tile + r * PITCH, W, fr.logW, tl, tw_s);
  }

  // ---- split z-spectrum into the spectra of the two real rows and store
  // A[kx] = (Z[kx] + conj Z[-kx]) / 2,  B[kx] = (Z[kx] - conj Z[-kx]) / (2i)
  const int nout = RB * 2 * Wc;
  const T half = (T)0.5;
  for (int e = tid; e < nout; e += nthreads) {
    int r = e & (RB - 1), c = e >> logRB;
    int kx = (c < Wc) ? c : c - Wc;
    cplx<T> zk = tile[r * PITCH + kx];
    cplx<T> zm = cconj(tile[r * PITCH + ((W - kx) & (W - 1))]);
    cplx<T> o;
    if (c < Wc) {
      o = mk<T>(half * (zk.x + zm.x), half * (zk.y + zm.y));
    } else {
      cplx<T> dlt = zk - zm;                    // (Zk - Zm) / (2i) = -i/2 * dlt
      o = mk<T>(half * dlt.y, -half * dlt.x);
    }
    scratch[(b * (2 * Wc) + c) * (long long)H + (y0 + r)] = o;
  }
}

// ----------------------------------------------------------------- cols --
#define PSFMC_COLS_CONV 0   // forward, multiply by spectrum, inverse (hot path)
#define PSFMC_COLS_SETUP 1  // forward only, scaled: produces the spectra themselves

// grid = (ceil(2*Wc / CB), B), block = CB * (H / 8); smem (CB*H + H) cplx<T>
template <typename T, int MODE, int LOGH = 0>
__global__ void cols_kernel(Frame fr_rt, int CB, const cplx<T> *__restrict__ tw_h,
                            const cplx<T> *__restrict__ spec,
                            const int *__restrict__ psf_sel,
                            cplx<T> *__restrict__ scratch,
                            cplx<T> *__restrict__ spec_out) {
  PSFMC_DYN_SMEM(smem_raw);
  Frame fr = fr_rt;
  if (LOGH) {
    fr.logH = LOGH;
    fr.H = 1 << LOGH;
  }
  const int H = fr.H, Wc = fr.Wc, ncol = 2 * Wc;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const long long b = blockIdx.y;
  const int c0 = blockIdx.x * CB;
  cplx<T> *tile = reinterpret_cast<cplx<T> *>(smem_raw);
  cplx<T> *tw_s = tile + CB * H;
  load_twiddles<T>(tw_s, tw_h, H, tid, nthreads);

  cplx<T> *base = scratch + (b * ncol + c0) * (long long)H;
  const int nel = CB * H;
  const int nvalid = (ncol - c0 < CB ? ncol - c0 : CB) * H;
#ifndef PSFMC_EMU
  // The CTA's columns are one contiguous run of the column-major scratch: fetch
  // them with a single bulk asynchronous copy (TMA) while the twiddles load.
  __shared__ __align__(8) unsigned long long tile_bar;
  const unsigned tile_bytes = (unsigned)nvalid * (unsigned)sizeof(cplx<T>);
  if (tid == 0) mbar_init(&tile_bar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&tile_bar, tile_bytes);
    bulk_load(tile, base, tile_bytes, &tile_bar);
  }
  for (int e = nvalid + tid; e < nel; e += nthreads) tile[e] = mk<T>((T)0, (T)0);
  mbar_wait(&tile_bar, 0);
#else
  for (int e = tid; e < nel; e += nthreads)
    tile[e] = (e < nvalid) ? base[e] : mk<T>((T)0, (T)0);
#endif
  __syncthreads();

  const int tpc = H >> 3;
  const int col = tid >> (fr.logH - 3), tl = tid & (tpc - 1);
  fft_line_smem<T, false, LOGH>(tile + col * H, H, fr.logH, tl, tw_s);

  if (MODE == PSFMC_COLS_SETUP) {
    // spec_out[b][c][ky] = F * (-1)^(kx+ky) / (H*W)
    const T scale = (T)(1.0 / ((double)fr.H * (double)fr.W));
    cplx<T> *obase = spec_out + (b * ncol + c0) * (long long)H;
    for (int e = tid; e < nvalid; e += nthreads) {
      int cc = c0 + (e >> fr.logH), ky = e & (H - 1);
      int kx = cc < Wc ? cc : cc - Wc;
      // (padded frames: the kernel is laid out with its origin at lag 0, no shift)
      T sgn = (((kx + ky) & 1) && !fr.padded) ? -scale : scale;
      obase[e] = mk<T>(tile[e].x * sgn, tile[e].y * sgn);
    }
    return;
  }

  int sel = psf_sel[b];
  if (sel < 0) sel = 0;  // walker is flagged invalid and overwritten in finalize
  const cplx<T> *sp = spec + ((long long)sel * ncol + c0) * (long long)H;
  for (int e = tid; e < nvalid; e += nthreads) tile[e] = tile[e] * sp[e];
  __syncthreads();

  fft_line_smem<T, true, LOGH>(tile + col * H, H, fr.logH, tl, tw_s);

  if (fr.padded) {
    // fold the linear convolution back modulo Hr (see Frame); reads touch rows >= Hr
    // only, writes rows < Hr only. (fft_line_smem ends with a CTA barrier.)
    const int Hr = fr.Hr, ncols = nvalid / H;
    for (int e = tid; e < ncols * Hr; e += nthreads) {
      const int cc = e / Hr, p = e - cc * Hr;
      cplx<T> v = tile[cc * H + p];
      if (p <= fr.fy_hi) v = v + tile[cc * H + p + Hr];
      if (p >= fr.fy_lo) v = v + tile[cc * H + H + p - Hr];
      tile[cc * H + p] = v;
    }
    __syncthreads();
  }

#ifndef PSFMC_EMU
  // (fft_line_smem ends with a CTA barrier after its last stores) one bulk store
  bulk_store_fence();
  __syncthreads();
  if (tid == 0) bulk_store_and_wait(base, tile, tile_bytes);
#else
  for (int e = tid; e < nvalid; e += nthreads) base[e] = tile[e];
#endif
}

// ------------------------------------------------------------- rows_inv --
template <typename T>
struct Epilogue;

// float64: the reference's expressions (psfMC/models.py:294, :277-279, :235-236)
template <>
struct Epilogue<double> {
  static __device__ __forceinline__ double term(double conv, double mvar, double obs,
                                                double ovar, double *resid,
                                                double *ivm) {
    *resid = obs - conv;
    *ivm = 1.0 / (mvar + ovar);
    return (*resid) * (*resid) * (*ivm) - log(0.5 / PSFMC_PI * (*ivm));
  }
};
// float32 arithmetic, float64 accumulation by the caller
#define PSFMC_VAR_NOISE_TOL 0.00390625f   // 2^-8
template <>
struct Epilogue<float> {
  static __device__ __forceinline__ double term(float conv, float mvar, float obs,
                                                float ovar, float *resid, float *ivm) {
    // -log(ivm / 2pi) = ln2 * log2(tot) + ln(2 pi): the logarithm does not wait
    // for the reciprocal
    *resid = obs - conv;
    const float tot = mvar + ovar;
    *ivm = fast_rcp(tot);
    float lg = fmaf(0.69314718055994530942f, fast_lg2(tot), 1.8378770664093454836f);
    // The convolved model variance is non-negative in exact arithmetic. Where the
    // float32 transform leaves it below -2^-8 of the pixel's own variance, its rounding
    // noise (a model whose brightest pixel outshines the rest by ~1e5) is no longer small
    // against the data: the walker is given up (NaN -> -inf -> repeated in float64 by
    // psfmc_lnlike_batch) instead of returning an lnL tens to hundreds away from the
    // reference's (audit of 65536 prior-drawn walkers).
    if (fmaf(PSFMC_VAR_NOISE_TOL, ovar, mvar) < 0.0f) lg = NAN;
    return (double)fmaf((*resid) * (*resid), *ivm, lg);
  }
};

// grid = (H / RB, B), block = RB * (W / 8); smem (RB*W + W) cplx<T> + 32 doubles
// img_* (optional, may be null): residual / composite IVM / convolved images.
template <typename T, int LOGW = 0>
__global__ void rows_inv_kernel(Frame fr_rt, int RB, const cplx<T> *__restrict__ tw_w,
                                const cplx<T> *__restrict__ scratch,
                                const T *__restrict__ obs, const T *__restrict__ ovar,
                                const unsigned char *__restrict__ bad,
                                const double *__restrict__ wscale,
                                const int *__restrict__ psf_sel,
                                const double *__restrict__ vscale_inv,
                                double *__restrict__ partials,
                                T *__restrict__ img_conv, T *__restrict__ img_resid,
                                T *__restrict__ img_ivm) {
  PSFMC_DYN_SMEM(smem_raw);
  Frame fr = fr_rt;
  if (LOGW) {
    fr.logW = LOGW;
    fr.W = 1 << LOGW;
    fr.Wc = (1 << LOGW) / 2 + 1;
  }
  const int W = fr.W, H = fr.H, Wc = fr.Wc;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const long long b = blockIdx.y;
  const int y0 = blockIdx.x * RB;
  const int PITCH = W + PSFMC_ROW_PAD(T);   // see rows_fwd_kernel
  const int logRB = 31 - __clz(RB);
  cplx<T> *tile = reinterpret_cast<cplx<T> *>(smem_raw);
  cplx<T> *tw_s = tile + RB * PITCH;
  double *red_s = reinterpret_cast<double *>(tw_s + W);
  load_twiddles<T>(tw_s, tw_w, W, tid, nthreads);

  // Y[kx] = A'[kx] + i B'[kx]; Y[W-kx] = conj(A'[kx]) + i conj(B'[kx])
  const cplx<T> *sa = scratch + b * (2 * Wc) * (long long)H;
  const cplx<T> *sb = sa + (long long)Wc * H;
  const int nin = RB * Wc;
  for (int e = tid; e < nin; e += nthreads) {
    int r = e & (RB - 1), kx = e >> logRB;
    cplx<T> a = sa[(long long)kx * H + y0 + r];
    cplx<T> bb = sb[(long long)kx * H + y0 + r];
    if (kx > 0 && kx < W - kx) {
      tile[r * PITCH + kx] = mk<T>(a.x - bb.y, a.y + bb.x);
      tile[r * PITCH + (W - kx)] = mk<T>(a.x + bb.y, bb.x - a.y);
    } else {
      // DC and Nyquist terms of a real row are real: a c2r transform ignores
      // their imaginary parts (numpy.fft.irfft does the same)
      tile[r * PITCH + kx] = mk<T>(a.x, bb.x);
    }
  }
  __syncthreads();

  {
    const int tpr = W >> 3;
    const int r = tid >> (fr.logW - 3), tl = tid & (tpr - 1);
    fft_line_smem<T, true, LOGW>(tile + r * PITCH, W, fr.logW, tl, tw_s);
  }

  if (fr.padded) {
    // fold modulo Wr (see Frame); reads touch columns >= Wr only, writes columns < Wr
    const int Wr = fr.Wr;
    for (int e = tid; e < RB * Wr; e += nthreads) {
      const int r = e / Wr, p = e - r * Wr;
      cplx<T> v = tile[r * PITCH + p];
      if (p <= fr.fx_hi) v = v + tile[r * PITCH + p + Wr];
      if (p >= fr.fx_lo) v = v + tile[r * PITCH + W + p - Wr];
      tile[r * PITCH + p] = v;
    }
    __syncthreads();
  }

  // undo the (exact, power-of-two) channel scalings of the variance image
  int sel = psf_sel[b];
  if (sel < 0) sel = 0;
  const T unscale = (T)(vscale_inv[sel] / wscale[b]);
  double acc = 0.0;
  const int npx = RB * W;
  for (int e = tid; e < npx; e += nthreads) {
    int r = e >> fr.logW, x = e & (W - 1);
    long long g = (long long)(y0 + r) * W + x;
    cplx<T> yv = tile[r * PITCH + x];
    yv.y *= unscale;
    T resid, ivm;
    double t = Epilogue<T>::term(yv.x, yv.y, obs[g], ovar[g], &resid, &ivm);
    if (!bad[g]) acc += t;
    if (img_conv) img_conv[b * (long long)H * W + g] = yv.x;
    if (img_resid) img_resid[b * (long long)H * W + g] = resid;
    if (img_ivm) img_ivm[b * (long long)H * W + g] = ivm;
  }
  // block reduction in float64: warp shuffles, then one value per warp via smem
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
  const int warp = tid >> 5, lane = tid & 31, nwarps = (nthreads + 31) >> 5;
  if (lane == 0) red_s[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int w = 0; w < nwarps; ++w) tot += red_s[w];
    partials[b * gridDim.x + blockIdx.x] = tot;
  }
}

// Posterior-image accumulation: the nb images of a chunk ([nb][npx], walker-major) are
// added to float64 sums; `invert` sums 1/value (the composite IVM is averaged in variance
// space, psfMC/models.py:81-82,96-97). One thread per (pixel, walker slice): slice s
// (blockIdx.y of PSFMC_ACC_SLICES) adds its contiguous share of the walkers, eight loads in
// flight, to its OWN sum acc[s][px] -- no atomics, the order of the additions is fixed --
// and accumulate_reduce_kernel adds the slices up once per call. (Round 2's first version
// had one thread per pixel walk through all 256 walkers of a chunk: 64 CTAs, one dependent
// load after the other, 150 us per image -- five sixths of the posterior-image time of the
// reference's example.)
#define PSFMC_ACC_SLICES 8
template <typename T>
__global__ void accumulate_kernel(const T *__restrict__ img, int nb, long long npx,
                                  int invert, double *__restrict__ acc) {
  const long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= npx) return;
  const int s = blockIdx.y, per = (nb + PSFMC_ACC_SLICES - 1) / PSFMC_ACC_SLICES;
  const int b0 = s * per, b1 = b0 + per < nb ? b0 + per : nb;
  double sum = 0.0;
  int b = b0;
  for (; b + 8 <= b1; b += 8) {
    T v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = img[(long long)(b + k) * npx + px];
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += invert ? 1.0 / (double)v[k] : (double)v[k];
  }
  for (; b < b1; ++b) {
    const double v = (double)img[(long long)b * npx + px];
    sum += invert ? 1.0 / v : v;
  }
  if (b1 > b0) acc[(long long)s * npx + px] += sum;
}
// out[px] = sum over the slices, in slice order
__global__ void accumulate_reduce_kernel(const double *__restrict__ acc, long long npx,
                                         double *__restrict__ out) {
  const long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= npx) return;
  double sum = 0.0;
  for (int s = 0; s < PSFMC_ACC_SLICES; ++s) sum += acc[(long long)s * npx + px];
  out[px] = sum;
}

// one thread per walker: lnL = -0.5 * sum(partials); non-finite => -inf
// (psfMC/models.py:235-241); invalid PSF index => -inf.
__global__ void finalize_kernel(const double *__restrict__ partials, int nblk,
                                const int *__restrict__ psf_sel, long long n_batch,
                                double *__restrict__ lnl) {
  long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_batch) return;
  double tot = 0.0;
  for (int k = 0; k < nblk; ++k) tot += partials[b * nblk + k];
  double v = -0.5 * tot;
  if (!isfinite(v) || psf_sel[b] < 0) v = -INFINITY;
  lnl[b] = v;
}

}  // namespace psfmc
