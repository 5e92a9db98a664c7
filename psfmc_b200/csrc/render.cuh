// Per-walker constants (prepare kernel) and per-pixel rendering of the raw model.
// Reference arithmetic: psfMC/ModelComponents/{Sky,PointSource,Sersic}.py and
// psfMC/models.py:245-253 (paths relative to /root/reference).
#pragma once
#include "common.cuh"
#include "devmath.cuh"

namespace psfmc {

__device__ __forceinline__ double slot_value(const Program *prog, int comp, int slot,
                                             const double *theta_row) {
  int ti = prog->theta_index[comp][slot];
  return ti >= 0 ? theta_row[ti] : prog->value[comp][slot];
}

// ---------------------------------------------------- float32 fast path --
#ifdef PSFMC_EMU
__device__ __forceinline__ float fast_lg2(float x) { return log2f(x); }
__device__ __forceinline__ float fast_ex2(float x) { return exp2f(x); }
__device__ __forceinline__ float fast_rcp(float x) { return 1.0f / x; }
#else
__device__ __forceinline__ float fast_lg2(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#endif

// Per-walker Sersic constants of the float32 path, derived from the float64 ones:
//   value = 2^(c0 - c1 * t) * (1 + (kq * t)^2 / |d|^2),  t = sq^p = 2^(p * log2 sq)
// which is Sersic.py:124-133 with sbeff*exp(kappa) folded into c0 and
// normed_grad * cent_offset = (2 p kappa)^2 / 12 * t^2 / |d|^2 (algebraically
// identical to the reference's g * (sq_delta_r / 12 * g)); kq = 2 p kappa / sqrt(12).
// Range: for a very small index (n ~ 0.006: p = 86, kappa ~ 1e-26) the profile only
// decays where t ~ 1/kappa, so t keeps float32's whole range (an overflow to +inf gives
// 2^-inf = 0, which is what the reference has there); kq * t is clamped at 1e18 so that
// its square stays finite -- the profile is exactly 0 long before (c1 * t = 5 n * kq * t).
// NaN / overflow: nothing on the way from sq to the value may swallow a NaN (fminf
// returns its other operand), and an infinite sq (reff ~ 0) must give NaN like the
// reference's inf * 0 in its gradient term: (sq - sq) is 0 for finite sq, NaN otherwise.
struct SersicF32 {
  float xi, xf, yi, yf;  // centre split into integer + fraction (keeps dx exact)
  float a00, a01, a10, a11;
  float p, c0, c1, kq;
};

__device__ __forceinline__ SersicF32 make_sersic_f32(const double *d) {
  SersicF32 s;
  double xi = rint(d[D_SER_X0]), yi = rint(d[D_SER_Y0]);
  s.xi = (float)xi;
  s.xf = (float)(d[D_SER_X0] - xi);
  s.yi = (float)yi;
  s.yf = (float)(d[D_SER_Y0] - yi);
  s.a00 = (float)d[D_SER_A00];
  s.a01 = (float)d[D_SER_A01];
  s.a10 = (float)d[D_SER_A10];
  s.a11 = (float)d[D_SER_A11];
  double p = d[D_SER_P], kappa = d[D_SER_KAPPA];
  const double log2e = 1.4426950408889634074;
  s.p = (float)p;
  s.c1 = (float)(kappa * log2e);
  s.c0 = (float)(log2(d[D_SER_SBEFF]) + kappa * log2e);
  s.kq = (float)(2.0 * p * kappa * 0.28867513459481288225);   // / sqrt(12)
  return s;
}

__device__ __forceinline__ float sersic_pixel_f32(const SersicF32 &s, float x, float y) {
  float dx = (x - s.xi) - s.xf, dy = (y - s.yi) - s.yf;
  float u = s.a00 * dx + s.a01 * dy;
  float v = s.a10 * dx + s.a11 * dy;
  float sq = u * u + v * v;
  float r2 = dx * dx + dy * dy;
  float t = fast_ex2(s.p * fast_lg2(sq));
  float sb = fast_ex2((s.c0 + (sq - sq)) - s.c1 * t);
  float g = fminf(s.kq * t, 1.0e18f);
  return sb * (1.0f + (g * g) * fast_rcp(r2));
}

__device__ __forceinline__ double sersic_pixel(const double *d, double x, double y);

// The work of ONE G-lane group: component c of walker b, theta row `th` -> derived
// constants (see prepare_kernel, which calls it for a whole batch; the low-latency
// instance of the fused 128 x 128 kernel calls it for its own walker, one warp per
// component). All lanes of a warp must call it together (shuffles); groups that are
// not `live` take part in them and write nothing.
template <int G>
__device__ __forceinline__ void prepare_group(const Program *prog, const double *th,
                                              long long b, int c, int glane, bool live, int H,
                                              int W, double *__restrict__ derived,
                                              int *__restrict__ psf_sel,
                                              double *__restrict__ wscale,
                                              float *__restrict__ rconst,
                                              int *__restrict__ hot) {
  const int ncomp_prog = prog->n_components;
  const int ncomp = ncomp_prog > 0 ? ncomp_prog : 1;
  const bool writer = live && (glane == 0);
  double *out = derived + (b * ncomp + c) * PSFMC_DERIVED_STRIDE;
  const int kind = c < ncomp_prog ? prog->kind[c] : -1;
  const int flags = c < ncomp_prog ? prog->flags[c] : 0;
  if (c == 0 && writer) {
    int sel = 0;
    if (prog->psf_theta_index >= 0) {
      double v = rint(th[prog->psf_theta_index]);
      sel = (v >= 0.0 && v < (double)prog->n_psf) ? (int)v : -1;
    } else {
      double v = rint(prog->psf_value);
      sel = (v >= 0.0 && v < (double)prog->n_psf) ? (int)v : -1;
    }
    psf_sel[b] = sel;
    // Per-walker power-of-two scale for the raw^2 channel of the packed transform
    // (z = raw + i * wscale * raw^2): with wscale ~ 1/flux both channels have
    // comparable magnitude, so rounding errors of the large one do not swamp
    // the small one. Exact (power of two), undone in the epilogue.
    // (only the binary exponent of the total flux matters: float32 exp2 is enough)
    double ftot = 0.0;
    for (int k = 0; k < ncomp_prog; ++k) {
      if (prog->kind[k] == PSFMC_SKY)
        ftot += fabs(slot_value(prog, k, PSFMC_P_ADU, th));
      else
        ftot += (double)exp2f((float)(-0.4 * 3.3219280948873623 *
                                      (slot_value(prog, k, PSFMC_P_MAG, th) - prog->mag_zp)));
    }
    double sc = 1.0;
    if (isfinite(ftot) && ftot > 1e-300) {
      int ex = 0;
      frexp(ftot, &ex);             // ftot = m * 2^ex, m in [0.5, 1)
      if (ex > 500) ex = 500;
      if (ex < -500) ex = -500;
      sc = ldexp(1.0, -ex);
    }
    wscale[b] = sc;
  }
  // kappa = gammaincinv(2n, 1/2): from the engine's Chebyshev table when 2n is inside
  // it, else by the group-cooperative iteration, in which every group of the warp has
  // to take part (shuffles) -- so the iteration runs only if some lane needs it
  const bool is_sersic = live && kind == PSFMC_SERSIC;
  const double n_index = is_sersic ? slot_value(prog, c, PSFMC_P_INDEX, th) : 0.5;
  double lgam_a1 = NAN;
  bool tabulated = false;
  double kappa = kappa_from_table(prog->kappa_coef, prog->kappa_nint, prog->kappa_u0,
                                  prog->kappa_inv_du, 2.0 * n_index, &tabulated);
  if (__any_sync(0xffffffffu, is_sersic && !tabulated)) {
    const double slow = gammaincinv_half_group<G>(2.0 * n_index, glane,
                                                  is_sersic && !tabulated, &lgam_a1);
    if (!tabulated) kappa = slow;
  }
  if (is_sersic && tabulated) lgam_a1 = lgamma(2.0 * n_index + 1.0);
  if (!live) return;
  if (kind == PSFMC_SKY) {
    if (writer) out[D_SKY_ADU] = slot_value(prog, c, PSFMC_P_ADU, th);
  } else if (kind == PSFMC_POINT) {
    double x = slot_value(prog, c, PSFMC_P_X, th);
    double y = slot_value(prog, c, PSFMC_P_Y, th);
    double mag = slot_value(prog, c, PSFMC_P_MAG, th);
    double radius = (flags & PSFMC_FLAG_BILINEAR) ? 0.5 : 3.0;
    double ymin, ymax, xmin, xmax;
    stamp_bounds(y, radius, H, &ymin, &ymax);
    stamp_bounds(x, radius, W, &xmin, &xmax);
    double flux = mag_to_flux(mag, prog->mag_zp);
    // a position that is not finite has no stamp (the reference's weights are NaN for
    // an infinite offset and its slice arithmetic raises for NaN): poison one pixel so
    // that the walker comes out as -inf instead of silently losing the component
    const bool lost = !(isfinite(x) && isfinite(y));
    if (lost) {
      ymin = ymax = xmin = xmax = 0.0;
      flux = NAN;
    }
    if (writer) {
      out[D_PS_X] = x;
      out[D_PS_Y] = y;
      out[D_PS_FLUX] = flux;
      out[D_PS_YMIN] = ymin;
      out[D_PS_YMAX] = ymax;
      out[D_PS_XMIN] = xmin;
      out[D_PS_XMAX] = xmax;
    }
    // separable stamp weights (PointSource.py:40-56: kern = prod over (x, y) of
    // lanczos(diff) or 1 - |diff|), measured from the UNCLIPPED position; lane i
    // of the group computes tap i of the x axis, then tap i of the y axis
    if (glane < 7) {
#pragma unroll
      for (int axis = 0; axis < 2; ++axis) {
        const bool along_y = axis == 1;
        const double pos = (along_y ? ymin : xmin) + (double)glane;
        const double centre = along_y ? y : x;
        double wgt = 0.0;
        if (pos <= (along_y ? ymax : xmax))
          wgt = (flags & PSFMC_FLAG_BILINEAR) ? 1.0 - fabs(pos - centre)
                                              : lanczos3_ref(pos - centre);
        if (lost) wgt = 1.0;
        out[(along_y ? D_PS_WY : D_PS_WX) + glane] = wgt;
      }
    }
  } else if (kind == PSFMC_SERSIC && writer) {  // Sersic.py:73-96 (transform), :47-71 (kappa, sb_eff)
    double x0 = slot_value(prog, c, PSFMC_P_X, th);
    double y0 = slot_value(prog, c, PSFMC_P_Y, th);
    double mag = slot_value(prog, c, PSFMC_P_MAG, th);
    double reff = slot_value(prog, c, PSFMC_P_REFF, th);
    double reff_b = slot_value(prog, c, PSFMC_P_REFF_B, th);
    double n = n_index;
    double angle = slot_value(prog, c, PSFMC_P_ANGLE, th);
    if (flags & PSFMC_FLAG_ANGLE_DEGREES) angle = angle * (PSFMC_PI / 180.0);  // np.deg2rad
    angle += 0.5 * PSFMC_PI;
    double sn = sin(angle), cs = cos(angle);
    // Gamma(2n) = exp(lgamma(2n + 1)) / (2n); overflows to inf for n >= 86 like
    // scipy.special.gamma (-> NaN -> lnL = -inf, as in the reference)
    double gamma_2n = exp(lgam_a1) / (2.0 * n);
    double flux = mag_to_flux(mag, prog->mag_zp);
    out[D_SER_X0] = x0;
    out[D_SER_Y0] = y0;
    out[D_SER_A00] = cs / reff;
    out[D_SER_A01] = sn / reff;
    out[D_SER_A10] = -sn / reff_b;
    out[D_SER_A11] = cs / reff_b;
    out[D_SER_P] = 0.5 / n;
    out[D_SER_KAPPA] = kappa;
    out[D_SER_SBEFF] = sersic_sb_eff(flux, n, reff, reff_b, kappa, gamma_2n);
  }
  // hot[b][c] (optional, fused 128 x 128 path): the pixel (py << 16 | px) the float32
  // kernel takes out of the transform for this component, or -1. A Sersic of index n > 1
  // whose centre lies within 0.05 px of a pixel centre: the reference's centroid
  // correction g^2 |d|^2 / 12 grows like |d|^(2/n - 2) there, so that ONE pixel outshines
  // its neighbours by 10^2 .. 10^5 -- and the rounding noise of a float32 transform,
  // which is relative to the largest element, swamps the rest of the frame (negative
  // convolved variances; round 1 repeated such walkers in float64). The pixel's
  // contribution is convolved exactly instead: (value) x (real-space PSF / PSF-variance
  // stamp) is added in the epilogue (FusedParams::kpv).
  if (hot && writer) {
    int flag = -1;
    if (kind == PSFMC_SERSIC) {
      const double x0 = out[D_SER_X0], y0 = out[D_SER_Y0];
      const double px = rint(x0), py = rint(y0);
      const double r2 = (x0 - px) * (x0 - px) + (y0 - py) * (y0 - py);
      if (n_index > 1.0 && r2 < 0.0025 && px >= 0.0 && px < (double)W && py >= 0.0 &&
          py < (double)H)
        flag = ((int)py << 16) | (int)px;
    }
    hot[b * ncomp + c] = flag;
  }
  if (rconst && writer) {
    float *rc = rconst + (b * ncomp + c) * PSFMC_RC_STRIDE;
    if (kind == PSFMC_SKY) {
      rc[0] = (float)out[D_SKY_ADU];
    } else if (kind == PSFMC_SERSIC) {
      SersicF32 s = make_sersic_f32(out);
      // A very small index (n < 0.01: p = 1/(2n) > 50) overflows the reference's
      // gradient term g * (sdr / 12 * g) in float64 at the far pixels of the frame:
      // sb = 0 there, 0 * inf = NaN, and one NaN in the raw model makes the whole
      // convolution NaN -> lnL = -inf (Sersic.py:129-133, models.py:238-241). The
      // float32 formula clamps kq * t and would come out finite: evaluate the
      // reference's own expression at the four frame corners (where sq, a convex
      // quadratic, is largest) and poison the float32 constants if it is NaN there.
      if (out[D_SER_P] > 20.0) {
        bool poisoned = false;
        for (int k = 0; k < 4; ++k) {
          const double v = sersic_pixel(out, (k & 1) ? (double)(W - 1) : 0.0,
                                        (k & 2) ? (double)(H - 1) : 0.0);
          poisoned = poisoned || (v != v);
        }
        if (poisoned) s.c0 = NAN;
      }
      rc[0] = s.xi; rc[1] = s.xf; rc[2] = s.yi; rc[3] = s.yf;
      rc[4] = s.a00; rc[5] = s.a01; rc[6] = s.a10; rc[7] = s.a11;
      rc[8] = s.p; rc[9] = s.c0; rc[10] = s.c1; rc[11] = s.kq;
    }
  }
}

// One G-lane GROUP per (walker, component): theta -> derived constants, float64.
// The lanes of a group share the scalar work and split the incomplete-gamma series
// of the Sersic kappa (devmath.cuh) and the point-source stamp taps; the four
// groups of a warp walk through the same shuffles (groups without a Sersic idle).
//   grid = ceil(G * B * n_components / blockDim), blockDim a multiple of 32
// wscale[b] receives the packing scale of the walker (see below);
// psf_sel[b] receives the rint-ed PSF index (psfMC/distributions.py:130-138), or
// -1 when it is out of range (the prior is -inf there; the walker gets -inf).
// The program travels as a kernel parameter (constant bank): its per-slot lookups
// are then constant loads instead of dependent global loads.
template <int G>
__global__ void prepare_kernel(const __grid_constant__ Program prog_c,
                               const double *__restrict__ theta, long long n_batch,
                               long long ld, int H, int W, double *__restrict__ derived,
                               int *__restrict__ psf_sel, double *__restrict__ wscale,
                               float *__restrict__ rconst, int stage,
                               int *__restrict__ hot = nullptr) {
  // stage: the theta rows of this CTA's walkers go through shared memory first, read
  // once with coalesced loads. (Tried on the B200: handing the kernel the caller's
  // page-locked HOST rows instead of copying them first -- per C-ABI call 70.6 us against
  // 57.8 us with the H2D copy in front at 100 walkers, 120.6 / 113.4 us at 512, 304 / 303
  // us at 2048: reads over the bus from inside a kernel cost more than the copy engine.)
  PSFMC_DYN_SMEM(smem_raw);
  const Program *prog = &prog_c;
  const int ncomp_prog = prog->n_components;
  const int ncomp = ncomp_prog > 0 ? ncomp_prog : 1;   // an empty model still gets its
                                                       // per-walker PSF index and scale
  const int glane = threadIdx.x & (G - 1);
  // Groups are numbered component-major, every component's run padded to whole CTAs:
  // gid = c * n_pad + b. A warp then holds groups of ONE component kind (the Sky, point
  // source and Sersic branches below no longer run one after the other in every warp)
  // and a CTA the same component of consecutive walkers.
  const int gpc = (int)blockDim.x / G;
  const long long n_pad = (n_batch + gpc - 1) / gpc * gpc;
  const long long gid0 = (long long)blockIdx.x * gpc;
  const long long gid = gid0 + threadIdx.x / G;
  const int c = (int)(gid / n_pad);
  const long long b0 = gid0 - (long long)c * n_pad;   // the CTA's first walker: always live
  long long b = gid - (long long)c * n_pad;
  const bool live = b < n_batch;
  if (!live) b = b0;                        // idle groups shadow the CTA's first group,
                                            // write nothing
  const double *th = theta + b * ld;
  if (stage) {
    double *th_s = reinterpret_cast<double *>(smem_raw);
    long long b1 = b0 + gpc;
    if (b1 > n_batch) b1 = n_batch;
    const long long nel = (b1 - b0) * ld;
    const double *src = theta + b0 * ld;
    for (long long e = threadIdx.x; e < nel; e += blockDim.x) th_s[e] = src[e];
    __syncthreads();
    th = th_s + (b - b0) * ld;
  }
  prepare_group<G>(prog, th, b, c, glane, live, H, W, derived, psf_sel, wscale, rconst, hot);
}

// kappa at n_nodes values of a = 2n by the iteration (one warp per node): used once at
// engine creation to build and to verify the Chebyshev table.
__global__ void kappa_nodes_kernel(const double *__restrict__ a_nodes, int n_nodes,
                                   double *__restrict__ kappa_out) {
  const int node = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  const bool live = node < n_nodes;
  double lgam;
  const double k = gammaincinv_half_group<32>(live ? a_nodes[node] : 1.0, lane, live, &lgam);
  if (live && lane == 0) kappa_out[node] = k;
}

// ------------------------------------------------------------ pixel math --

// float64: the reference's expressions, literally (Sersic.py:124-133,151-153).
__device__ __forceinline__ double sersic_pixel(const double *d, double x, double y) {
  double dx = x - d[D_SER_X0], dy = y - d[D_SER_Y0];
  double u = d[D_SER_A00] * dx + d[D_SER_A01] * dy;
  double v = d[D_SER_A10] * dx + d[D_SER_A11] * dy;
  double sq = u * u + v * v;
  double sdr = sq / (dx * dx + dy * dy);
  double p = d[D_SER_P], kappa = d[D_SER_KAPPA];
  double lg = log(sq);
  double sb = exp(-kappa * expm1(lg * p));
  double g = -kappa * 2.0 * p * exp(lg * (p - 0.5));
  double cent = sdr / 12.0 * g;
  return d[D_SER_SBEFF] * sb * (1.0 + g * cent);
}

// Point-source stamp value at pixel (x, y); 0 outside the stamp
// (PointSource.py:40-56): (w_x * w_y) * flux with the separable weights the
// prepare kernel tabulated. Always float64: at most 49 pixels per source.
__device__ __forceinline__ double point_pixel(const double *d, int x, int y) {
  const int ix = x - (int)d[D_PS_XMIN], iy = y - (int)d[D_PS_YMIN];
  if (ix < 0 || iy < 0 || (double)x > d[D_PS_XMAX] || (double)y > d[D_PS_YMAX]) return 0.0;
  return (d[D_PS_WX + ix] * d[D_PS_WY + iy]) * d[D_PS_FLUX];
}

// Raw model at one pixel in float64 (PSFMC_PREC_FP64), components added in model
// order. round_f32: round the running sum to float32 after every component, i.e.
// the reference's float32 `arr +=` storage (PSFMC_PREC_FP64_RAWF32, oracle M2).
// ps_only: only point sources (models.py:296-306).
__device__ __forceinline__ double raw_pixel_f64(const Program *prog, const double *derived_b,
                                                int x, int y, bool round_f32,
                                                bool ps_only) {
  double acc = 0.0;
  const int ncomp = prog->n_components;
  for (int c = 0; c < ncomp; ++c) {
    const double *d = derived_b + c * PSFMC_DERIVED_STRIDE;
    const int kind = prog->kind[c];
    if (ps_only && kind != PSFMC_POINT) continue;
    if (kind == PSFMC_SKY) {
      acc += d[D_SKY_ADU];
    } else if (kind == PSFMC_POINT) {
      double w = point_pixel(d, x, y);
      // pixels outside the stamp are not touched by the reference's slice `+=`
      if ((double)y >= d[D_PS_YMIN] && (double)y <= d[D_PS_YMAX] &&
          (double)x >= d[D_PS_XMIN] && (double)x <= d[D_PS_XMAX])
        acc += w;
    } else {
      acc += sersic_pixel(d, (double)x, (double)y);
    }
    if (round_f32) acc = (double)(float)acc;
  }
  return acc;
}

// 1/x for a pair on the FMA pipe (the render is bound by the SFU, which already
// carries three transcendentals per pixel): integer seed (12 % error), three
// Newton steps -> ~1e-7 relative. x > 0; x = 0 gives a huge finite value times the
// zero it multiplies where the reference has 0/0 = NaN -- handled by the caller.
__device__ __forceinline__ cplx<float> rcp_pair_fma(cplx<float> x) {
#ifdef PSFMC_EMU
  return mk<float>(1.0f / x.x, 1.0f / x.y);
#else
  cplx<float> y = mk<float>(__int_as_float(0x7EF311C7 - __float_as_int(x.x)),
                            __int_as_float(0x7EF311C7 - __float_as_int(x.y)));
  const cplx<float> nx = mk<float>(-x.x, -x.y);
#pragma unroll
  for (int it = 0; it < 3; ++it) y = pfma(y, pfma(nx, y, bcast(1.0f)), y);
  return y;
#endif
}

// Reciprocal of a pixel pair for the centroid correction: from the FMA pipe (integer
// seed + three Newton steps, rcp_pair_fma) or from the SFU (two MUFU.RCP, 1 ulp).
template <bool MUFU = true>
__device__ __forceinline__ cplx<float> rcp_pair(cplx<float> x) {
#ifndef PSFMC_EMU
  if (!MUFU) return rcp_pair_fma(x);
#endif
  return mk<float>(fast_rcp(x.x), fast_rcp(x.y));
}
// Which of the eight pixel pairs of a fused-render thread take their reciprocal from
// the SFU (bit i set) and which from the FMA pipe. Measured on the B200 (r2, C1, 2048
// walkers per launch): all SFU 241.8 us (the render phase becomes SFU-bound: XU 31 %,
// MIO-queue stalls), half 240.3, three quarters 239.8, none 239.8 -- a wash, so the
// Newton form stays (it is also the more accurate one: unbiased, ~0.5 ulp).
#ifndef PSFMC_RCP_PATTERN
#define PSFMC_RCP_PATTERN 0x00
#endif

// Two pixels of one row at once (x offsets dx.x, dx.y from the centre), element-wise
// pair arithmetic (packed FFMA2/FMUL2 on the device): same formula as
// sersic_pixel_f32. cu = a01*dy, cv = a11*dy, dy2 = dy*dy are per-row constants.
// Returns acc + value (the sum rides on the last multiply).
template <bool MUFU = true>
__device__ __forceinline__ cplx<float> sersic_pair_f32(const SersicF32 &s, cplx<float> dx,
                                                       float cu, float cv, float dy2,
                                                       cplx<float> acc) {
  const cplx<float> u = pfma(bcast(s.a00), dx, bcast(cu));
  const cplx<float> v = pfma(bcast(s.a10), dx, bcast(cv));
  const cplx<float> sq = pfma(v, v, pmul(u, u));
  const cplx<float> r2 = pfma(dx, dx, bcast(dy2));
  // (sq - sq) = 0 for finite sq, NaN for an infinite one (see SersicF32): rides on the
  // multiply by p as its addend
  const cplx<float> e = pfma(bcast(s.p), mk<float>(fast_lg2(sq.x), fast_lg2(sq.y)), sq - sq);
  const cplx<float> t = mk<float>(fast_ex2(e.x), fast_ex2(e.y));
  const cplx<float> arg = pfma(bcast(-s.c1), t, bcast(s.c0));
  const cplx<float> sb = mk<float>(fast_ex2(arg.x), fast_ex2(arg.y));
  const cplx<float> gu = pmul(bcast(s.kq), t);
  const cplx<float> g = mk<float>(fminf(gu.x, 1.0e18f), fminf(gu.y, 1.0e18f));
  return pfma(sb, pfma(pmul(g, g), rcp_pair<MUFU>(r2), bcast(1.0f)), acc);    // + sb * (1 + q)
}

// Two arbitrary pixels at once (offsets (dx.x, dy.x) and (dx.y, dy.y) from the centre).
__device__ __forceinline__ cplx<float> sersic_pair2_f32(const SersicF32 &s, cplx<float> dx,
                                                        cplx<float> dy, cplx<float> acc) {
  const cplx<float> u = pfma(bcast(s.a00), dx, pmul(bcast(s.a01), dy));
  const cplx<float> v = pfma(bcast(s.a10), dx, pmul(bcast(s.a11), dy));
  const cplx<float> sq = pfma(v, v, pmul(u, u));
  const cplx<float> r2 = pfma(dx, dx, pmul(dy, dy));
  const cplx<float> e = pfma(bcast(s.p), mk<float>(fast_lg2(sq.x), fast_lg2(sq.y)), sq - sq);
  const cplx<float> t = mk<float>(fast_ex2(e.x), fast_ex2(e.y));
  const cplx<float> arg = pfma(bcast(-s.c1), t, bcast(s.c0));
  const cplx<float> sb = mk<float>(fast_ex2(arg.x), fast_ex2(arg.y));
  const cplx<float> gu = pmul(bcast(s.kq), t);
  const cplx<float> g = mk<float>(fminf(gu.x, 1.0e18f), fminf(gu.y, 1.0e18f));
  return pfma(sb, pfma(pmul(g, g), rcp_pair<false>(r2), bcast(1.0f)), acc);    // + sb * (1 + q)
}

}  // namespace psfmc
