// 512 x 512 frames (float32) as 4 x 4 interleaved 128 x 128 sub-images.
//
// A 512 x 512 packed frame is 2 MB of complex64: it fits neither one SM's shared memory
// (the 128 x 128 fused kernel) nor a portable thread-block cluster (the 256 x 256 cluster
// kernel). The staged row / column kernels that served it in round 1 run at a quarter of
// the fused kernel's efficiency (shared-memory Stockham passes, index arithmetic). This
// path keeps ALL transform work in the fused kernel's register-resident passes by
// splitting the transform once, in both dimensions, by decimation:
//
//   pixel (4 y' + ry, 4 x' + rx) belongs to sub-image s = 4 ry + rx, pixel (y', x');
//   with F_s = DFT_128x128(sub-image s),
//     Z[ky + 128 qy][kx + 128 qx] = sum_{ry,rx} W4^(ry qy + rx qx)
//                                   * W512^(ry ky + rx kx) * F_{ry,rx}[ky][kx]
//   -- per (ky, kx) a twiddle and a 4 x 4 point DFT over the sixteen sub-spectra.
//
//   K1  fused_lnlike_kernel<.., MODE_FWD>: per job (walker, sub-image): render the
//       sub-image (every 4th pixel of every 4th row), forward rows + columns in shared
//       memory, column spectrum -> sub_spec[job][ky][kx]          (2 MB per walker out)
//   K2  tiled_combine_kernel: per (ky, kx) and its mirror (-ky, -kx): twiddles, 4 x 4 DFT
//       -> the sixteen frequencies (ky + 128 qy, kx + 128 qx) of the full spectrum,
//       mirror-pair product with the PSF / PSF-variance spectra (see mirror_pair in
//       kernels_fused.cuh), inverse 4 x 4 DFT, conjugate twiddles, in place
//                                                                (2 MB in, 2 MB out)
//   K3  fused_lnlike_kernel<.., MODE_INV>: per job: inverse columns + rows of the
//       sub-spectrum, chi-square terms against the sub-image's observation, one float64
//       partial per job                                           (2 MB per walker in)
//   finalize_kernel sums a walker's sixteen partials in fixed order.
//
// Algorithmic HBM traffic: 8 MB per walker, like the staged path (32 N
// bytes); what changes is where the FLOPs run.
#pragma once
#include "kernels_fused.cuh"

namespace psfmc {

#define PSFMC_TILED_N 512
#define PSFMC_TILED_SUBS 16
#define PSFMC_TILED_KX 65          // kx = 0..64: every other column is a mirror
#define PSFMC_TILED_THREADS 96     // one thread per kx (65 of 96 lanes active)

// 4-point DFT over v[0..3] (stride 1 in the array handed in), forward kernel W4 = -i;
// CONJ: the conjugate kernel (+i)
template <bool CONJ>
__device__ __forceinline__ void tiled_dft4(cplx<float> &v0, cplx<float> &v1, cplx<float> &v2,
                                           cplx<float> &v3) {
  dft4<float, CONJ>(v0, v1, v2, v3);
}

struct TiledParams {
  cplx<float> *sub_spec;      // [walkers][16][128][128]
  const float4 *spec;         // [K][16 q][128 ky][65 kx]: (S+, S-) at (ky + 128 qy, kx + 128 qx)
  const float2 *tw512;        // [512]: W512^m
  const int *psf_sel;         // [walkers]
  int n_walkers;
};

// grid = (128 ky, walkers), block = 96 (thread = kx 0..64)
__global__ void __launch_bounds__(PSFMC_TILED_THREADS)
tiled_combine_kernel(const TiledParams P) {
  constexpr int M = PSFMC_FUSED_N;
  const int kx = threadIdx.x, ky = blockIdx.x;
  const long long wb = blockIdx.y;
  if (kx >= PSFMC_TILED_KX) return;
  // one thread per unordered pair {k, -k}: kx = 1..63 takes every ky; the self-mirrored
  // columns kx = 0 and 64 take ky = 0..64 (ky = 65..127 are the mirrors of 63..1)
  const bool xself = (kx == 0 || kx == 64);
  if (xself && ky > 64) return;
  const bool self = xself && (ky == 0 || ky == 64);   // k = -k: one point, not a pair
  const int my = (M - ky) & (M - 1), mx = (M - kx) & (M - 1);
  int sel = P.psf_sel[wb];
  if (sel < 0) sel = 0;   // flagged invalid, overwritten in finalize
  cplx<float> *base = P.sub_spec + (size_t)wb * PSFMC_TILED_SUBS * M * M;
  cplx<float> *pu = base + (size_t)ky * M + kx, *pm = base + (size_t)my * M + mx;

  // sub-spectra of the point and of its mirror: g[ry][rx]
  cplx<float> gu[4][4], gm[4][4];
#pragma unroll
  for (int ry = 0; ry < 4; ++ry)
#pragma unroll
    for (int rx = 0; rx < 4; ++rx) {
      gu[ry][rx] = pu[(size_t)(4 * ry + rx) * M * M];
      gm[ry][rx] = pm[(size_t)(4 * ry + rx) * M * M];
    }
  // Twiddles W512^(r k). The mirror point is transformed with the CONJUGATE 4-point
  // kernel and the exponent k' + 128 c, c = 4 - (k != 0): its output slot q then holds
  // frequency (c - q) mod 4, i.e. exactly the mirror of the point's slot q -- partners
  // meet in the same slot without any index arithmetic.
  const int ey = my + 128 * (ky != 0 ? 3 : 4), ex = mx + 128 * (kx != 0 ? 3 : 4);
  cplx<float> wuy[4], wux[4], wmy[4], wmx[4];
#pragma unroll
  for (int r = 1; r < 4; ++r) {
    const float2 a = __ldg(P.tw512 + ((r * ky) & 511)), b = __ldg(P.tw512 + ((r * kx) & 511));
    const float2 c = __ldg(P.tw512 + ((r * ey) & 511)), d = __ldg(P.tw512 + ((r * ex) & 511));
    wuy[r] = mk<float>(a.x, a.y);
    wux[r] = mk<float>(b.x, b.y);
    wmy[r] = mk<float>(c.x, c.y);
    wmx[r] = mk<float>(d.x, d.y);
  }
  // forward: along x (twiddle, 4-point DFT over rx), then along y
#pragma unroll
  for (int ry = 0; ry < 4; ++ry) {
#pragma unroll
    for (int rx = 1; rx < 4; ++rx) {
      gu[ry][rx] = gu[ry][rx] * wux[rx];
      gm[ry][rx] = gm[ry][rx] * wmx[rx];
    }
    tiled_dft4<false>(gu[ry][0], gu[ry][1], gu[ry][2], gu[ry][3]);
    tiled_dft4<true>(gm[ry][0], gm[ry][1], gm[ry][2], gm[ry][3]);
  }
#pragma unroll
  for (int qx = 0; qx < 4; ++qx) {
#pragma unroll
    for (int ry = 1; ry < 4; ++ry) {
      gu[ry][qx] = gu[ry][qx] * wuy[ry];
      gm[ry][qx] = gm[ry][qx] * wmy[ry];
    }
    tiled_dft4<false>(gu[0][qx], gu[1][qx], gu[2][qx], gu[3][qx]);
    tiled_dft4<true>(gm[0][qx], gm[1][qx], gm[2][qx], gm[3][qx]);
  }
  // gu[qy][qx] = Z[ky + 128 qy][kx + 128 qx], gm[qy][qx] = Z at the mirror frequency
  const float4 *sp = P.spec + ((size_t)sel * 16 * M + ky) * PSFMC_TILED_KX + kx;
#pragma unroll
  for (int qy = 0; qy < 4; ++qy)
#pragma unroll
    for (int qx = 0; qx < 4; ++qx) {
      const float4 s4 = __ldg(sp + (size_t)(4 * qy + qx) * M * PSFMC_TILED_KX);
      if (self) {
        // the "mirror" array is the point itself (loaded twice): keep the point's half
        cplx<float> partner = gm[qy][qx];
        mirror_pair(gu[qy][qx], partner, s4);
      } else {
        mirror_pair(gu[qy][qx], gm[qy][qx], s4);
      }
    }
  // inverse: along y (4-point DFT with the opposite kernel, conjugate twiddle), then x
#pragma unroll
  for (int qx = 0; qx < 4; ++qx) {
    tiled_dft4<true>(gu[0][qx], gu[1][qx], gu[2][qx], gu[3][qx]);
    tiled_dft4<false>(gm[0][qx], gm[1][qx], gm[2][qx], gm[3][qx]);
#pragma unroll
    for (int ry = 1; ry < 4; ++ry) {
      gu[ry][qx] = cmul_conj(gu[ry][qx], wuy[ry]);
      gm[ry][qx] = cmul_conj(gm[ry][qx], wmy[ry]);
    }
  }
#pragma unroll
  for (int ry = 0; ry < 4; ++ry) {
    tiled_dft4<true>(gu[ry][0], gu[ry][1], gu[ry][2], gu[ry][3]);
    tiled_dft4<false>(gm[ry][0], gm[ry][1], gm[ry][2], gm[ry][3]);
#pragma unroll
    for (int rx = 1; rx < 4; ++rx) {
      gu[ry][rx] = cmul_conj(gu[ry][rx], wux[rx]);
      gm[ry][rx] = cmul_conj(gm[ry][rx], wmx[rx]);
    }
  }
#pragma unroll
  for (int ry = 0; ry < 4; ++ry)
#pragma unroll
    for (int rx = 0; rx < 4; ++rx) {
      pu[(size_t)(4 * ry + rx) * M * M] = gu[ry][rx];
      if (!self) pm[(size_t)(4 * ry + rx) * M * M] = gm[ry][rx];
    }
}

// -------------------------------------------------------------- host side --

template <typename T>
inline bool tiled_path_available(const StagedPlan &plan) {
  return sizeof(T) == 4 && plan.fr.H == PSFMC_TILED_N && plan.fr.W == PSFMC_TILED_N &&
         !plan.fr.padded;
}

inline int tiled_prepare_device() {
#ifndef PSFMC_EMU
  if (cudaFuncSetAttribute(fused_lnlike_kernel<false, PSFMC_MODE_FWD>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                           PSFMC_FUSED_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(fused_lnlike_kernel<false, PSFMC_MODE_INV>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                           PSFMC_FUSED_SMEM) != cudaSuccess)
    return 1;
#endif
  return 0;
}

// Spectra for tiled_combine_kernel from the float64 spectra [K][2*Wc][H] (column-major,
// Wc = 257 retained columns; P = PSF, V = PSF variance, normalisation 1/(H W) and
// ifftshift sign folded in): spec[K][q = 4 qy + qx][ky][kx] = (S+, S-) = (P +- v V) / 2 at
// (Ky, Kx) = (ky + 128 qy, kx + 128 qx); columns beyond 256 by Hermitian symmetry. The
// 1/(128 * 128) of the sub-images' inverse transforms is NOT part of 1/(H W): the full
// inverse is 1/(512 * 512) = 1/(128 * 128) * 1/16, and the sub-transforms are
// unnormalised, so nothing is missing.
inline void tiled_spectrum_layout(const cplx<double> *spec64, int n_psf, const double *vscale,
                                  float4 *spec) {
  constexpr int N = PSFMC_TILED_N, Wc = N / 2 + 1, M = PSFMC_FUSED_N;
  for (int k = 0; k < n_psf; ++k) {
    const cplx<double> *src = spec64 + (size_t)k * 2 * Wc * N;
    auto at = [&](int chan, int Ky, int Kx) {
      cplx<double> v;
      if (Kx <= N / 2) {
        v = src[((size_t)chan * Wc + Kx) * N + Ky];
      } else {
        v = src[((size_t)chan * Wc + (N - Kx)) * N + ((N - Ky) & (N - 1))];
        v.y = -v.y;
      }
      return v;
    };
    for (int q = 0; q < 16; ++q)
      for (int ky = 0; ky < M; ++ky)
        for (int kx = 0; kx < PSFMC_TILED_KX; ++kx) {
          const int Ky = ky + M * (q >> 2), Kx = kx + M * (q & 3);
          cplx<double> pp = at(0, Ky, Kx), vv = at(1, Ky, Kx);
          const bool self = (Kx == 0 || Kx == N / 2) && (Ky == 0 || Ky == N / 2);
          if (self) pp.y = vv.y = 0.0;
          float4 o;
          o.x = (float)(0.5 * (pp.x + vscale[k] * vv.x));
          o.y = (float)(0.5 * (pp.y + vscale[k] * vv.y));
          o.z = (float)(0.5 * (pp.x - vscale[k] * vv.x));
          o.w = (float)(0.5 * (pp.y - vscale[k] * vv.y));
          spec[(((size_t)k * 16 + q) * M + ky) * PSFMC_TILED_KX + kx] = o;
        }
  }
}

struct TiledBuffers {
  float *rconst = nullptr;
  const float4 *spec = nullptr;        // tiled_spectrum_layout
  const float2 *tw512 = nullptr;
  const float2 *ow = nullptr;          // [16][128][128] sub-image order
  const unsigned short *maskw = nullptr;   // [16][128][8]
  const unsigned *skip_tab = nullptr;  // [16]
  cplx<float> *sub_spec = nullptr;     // [chunk][16][128][128]
  double *partials = nullptr;          // [B][16]
  double lnl_const = 0.0;
  long long chunk = 1;                 // walkers whose sub-spectra are in flight at once
  int n_sms = 148;
};

// theta -> lnL for n_batch walkers. Returns the number of kernels launched.
template <typename T>
inline int launch_tiled_lnlike(const StagedPlan &plan, const StagedBuffers<T> &buf,
                               const TiledBuffers &tb, const Program &prog_h,
                               const double *theta, long long n_batch, long long ld,
                               double *lnl, cudaStream_t stream,
                               cudaEvent_t ev_begin = nullptr, cudaEvent_t ev_end = nullptr) {
  if (n_batch <= 0) return 0;
  const int ncomp = prog_h.n_components;
  launch_prepare(*buf.prog_host, theta, n_batch, ld, plan.fr.Hr, plan.fr.Wr, ncomp, buf.derived,
                 buf.psf_sel, buf.wscale, tb.rconst, stream);
  int launches = 1;
  if (ev_begin) cudaEventRecord(ev_begin, stream);
  FoldParams F = {};
  for (long long start = 0; start < n_batch; start += tb.chunk) {
    const long long nb = n_batch - start < tb.chunk ? n_batch - start : tb.chunk;
    FusedParams P = {};
    P.rconst = tb.rconst + start * ncomp * PSFMC_RC_STRIDE;
    P.derived = buf.derived + start * ncomp * PSFMC_DERIVED_STRIDE;
    P.wscale = buf.wscale + start;
    P.psf_sel = buf.psf_sel + start;
    P.vscale_inv = buf.vscale_inv;
    P.ow = tb.ow;
    P.maskw = tb.maskw;
    P.lnl_const = tb.lnl_const;
    P.skip_tab = tb.skip_tab;
    P.sub_spec = tb.sub_spec;
    P.partials = tb.partials + start * PSFMC_TILED_SUBS;
    P.n_batch = nb * PSFMC_TILED_SUBS;   // jobs
    P.ncomp = ncomp;
    for (int c = 0; c < ncomp; ++c)
      P.kind_bits |= (unsigned long long)(prog_h.kind[c] & 3) << (2 * c);
    const unsigned grid = (unsigned)(P.n_batch < tb.n_sms ? P.n_batch : tb.n_sms);
    launch_kernel(fused_lnlike_kernel<false, PSFMC_MODE_FWD>, dim3(grid),
                  dim3(PSFMC_FUSED_THREADS), (size_t)PSFMC_FUSED_SMEM, stream, P, F);
    TiledParams Q;
    Q.sub_spec = tb.sub_spec;
    Q.spec = tb.spec;
    Q.tw512 = tb.tw512;
    Q.psf_sel = buf.psf_sel + start;
    Q.n_walkers = (int)nb;
    launch_kernel(tiled_combine_kernel, dim3(PSFMC_FUSED_N, (unsigned)nb),
                  dim3(PSFMC_TILED_THREADS), 0, stream, Q);
    launch_kernel(fused_lnlike_kernel<false, PSFMC_MODE_INV>, dim3(grid),
                  dim3(PSFMC_FUSED_THREADS), (size_t)PSFMC_FUSED_SMEM, stream, P, F);
    launches += 3;
  }
  if (ev_end) cudaEventRecord(ev_end, stream);
  {
    const int block = 128;
    const unsigned grid = (unsigned)((n_batch + block - 1) / block);
    launch_kernel(finalize_kernel, dim3(grid), dim3(block), 0, stream,
                  (const double *)tb.partials, PSFMC_TILED_SUBS, (const int *)buf.psf_sel,
                  n_batch, lnl);
    ++launches;
  }
  return launches;
}

}  // namespace psfmc
