// Fused single-kernel shared-memory path: 128 x 128 frames, float32 render + FFT,
// float64 chi-square accumulation. One persistent CTA per SM walks over the
// walkers of the batch; the whole packed frame z = raw + i*raw^2 (128 KB of
// complex64) lives in shared memory from render to reduction, so per walker the
// kernel touches HBM/L2 only for its parameters, the PSF spectra (128 KB, L2
// resident, shared by all walkers), the observation (128 KB, L2 resident) and
// one double of output.
//
// Length-128 transforms are split 16 x 8: a radix-16 butterfly in registers, the
// inter-stage twiddles W128^(n2*k1), one exchange through shared memory, a radix-8
// butterfly in registers (index maps n = n2 + 8*n1, k = k1 + 16*k2).
//
//   rows fwd  (per warp, 4 rows at a time, no CTA barrier): render 16 px/thread
//             -> radix-16 -> twiddle -> exchange (warp-private) -> radix-8 -> split
//             the packed row spectrum into the row spectra of the two real images
//             (A: raw, B: raw^2; partners kx <-> -kx are held by the same thread)
//             -> 128 columns: c = kx (A, kx=1..63), c = 64+kx (B), c = 0 and c = 64
//             carry the real DC/Nyquist columns of A and B packed in pairs
//   cols      radix-16 -> exchange inside a 4-warp column group -> radix-8 ->
//             multiply by the PSF / PSF-variance spectrum (registers) -> inverse
//             radix-8 -> exchange -> inverse radix-16
//   rows inv  rebuild Y = A' + i B' (Hermitian extension), inverse radix-8 ->
//             exchange -> inverse radix-16 -> Re = convolved model, Im = model
//             variance -> residual, composite IVM, masked chi-square terms in
//             registers -> float64 warp + CTA reduction -> lnL
//
// Two CTA-wide barriers and two 128-thread named barriers per walker; warps drift
// apart between them, which overlaps the SFU-bound render of one warp with the
// shared-memory-bound butterflies of another.
//
// Reference arithmetic: see render.cuh (components) and kernels_staged.cuh
// (convolution / likelihood); this file only re-schedules it.
#pragma once
#include "pipeline.cuh"
#include "twiddle128.cuh"

namespace psfmc {

#define PSFMC_FUSED_N 128
#define PSFMC_FUSED_THREADS 512
#define PSFMC_FUSED_SMEM (PSFMC_FUSED_N * PSFMC_FUSED_N * 8)

struct FusedParams {
  const float *rconst;       // [B][ncomp][PSFMC_RC_STRIDE]
  const double *derived;     // [B][ncomp][PSFMC_DERIVED_STRIDE] (point-source taps)
  const double *wscale;      // [B]
  const int *psf_sel;        // [B]
  const double *vscale_inv;  // [K]
  const cplx<float> *spec;   // [K][ky=128][c=128], see fused_spectrum_layout()
  const cplx<float> *specx;  // [K][2][ky=128]: P[ky][kx=64], V[ky][kx=64]
  const float2 *ow;          // [128*128]: (obs, bad ? -ovar : +ovar)
  double *lnl;               // [B]
  long long n_batch;
  int ncomp;
  signed char kind[PSFMC_MAX_COMPONENTS];
};

__device__ __forceinline__ void group_barrier(int id, int nthreads) {
#ifdef PSFMC_EMU
  emu::named_barrier(id, nthreads);
#else
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
#endif
}

__device__ __forceinline__ cplx<float> tw128(int n2, int k1) {
  const int idx = n2 * 16 + k1;
  return mk<float>(c_tw128[idx][0], c_tw128[idx][1]);
}

// 16-point DFT in registers, natural order in and out: two radix-4 layers
// (n = 4*n1 + n2, k = k1 + 4*k2) with the W16 twiddles in between.
template <bool INV>
__device__ __forceinline__ void dft16(cplx<float> *v) {
  const float C = 0.92387953251128675613f, S = 0.38268343236508977173f;
  const float h = 0.70710678118654752440f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4<float, INV>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
  // v[4*k1 + n2] *= W16^(n2*k1)  (conjugated for the inverse)
  const float sg = INV ? -1.0f : 1.0f;
  auto rot = [&](cplx<float> &a, float c, float s) {   // a *= (c, -sg*s)
    float x = a.x * c + sg * a.y * s, y = a.y * c - sg * a.x * s;
    a.x = x;
    a.y = y;
  };
  rot(v[4 * 1 + 1], C, S);        // e = 1
  rot(v[4 * 1 + 2], h, h);        // e = 2
  rot(v[4 * 1 + 3], S, C);        // e = 3
  rot(v[4 * 2 + 1], h, h);        // e = 2
  v[4 * 2 + 2] = INV ? mul_pos_i(v[4 * 2 + 2]) : mul_neg_i(v[4 * 2 + 2]);  // e = 4
  rot(v[4 * 2 + 3], -h, h);       // e = 6
  rot(v[4 * 3 + 1], S, C);        // e = 3
  rot(v[4 * 3 + 2], -h, h);       // e = 6
  rot(v[4 * 3 + 3], -C, -S);      // e = 9
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
    dft4<float, INV>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  // X[k1 + 4*k2] sits in v[4*k1 + k2]: transpose to natural order
  cplx<float> t[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) t[i] = v[i];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = t[4 * (k & 3) + (k >> 2)];
}

// position (in complex elements) of exchange element (k1, n2) inside a row:
// rotate n2 by k1 inside each block of 8 so that both the radix-16 side (fixed k1,
// lanes over n2) and the radix-8 side (fixed n2, lanes over k1) are free of bank
// conflicts; `swz` = 8 * (row parity) keeps two rows of a half-warp apart.
__device__ __forceinline__ int xpos(int k1, int n2, int swz) {
  return (k1 * 8 + ((n2 + k1) & 7)) ^ swz;
}

// Raw model at the 16 pixels x = l + 8*j of row y -> packed z = raw + i*wsc*raw^2.
__device__ __forceinline__ void fused_render16(const FusedParams &P, long long b, int y,
                                               int l, float wsc, cplx<float> *v) {
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
  const float fy = (float)y;
  for (int c = 0; c < P.ncomp; ++c) {
    const int kind = P.kind[c];
    const float *rc = P.rconst + (b * P.ncomp + c) * PSFMC_RC_STRIDE;
    if (kind == PSFMC_SKY) {
      const float adu = __ldg(rc);
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] += adu;
    } else if (kind == PSFMC_SERSIC) {
      SersicF32 s;
      const float4 q0 = __ldg(reinterpret_cast<const float4 *>(rc));
      const float4 q1 = __ldg(reinterpret_cast<const float4 *>(rc) + 1);
      const float4 q2 = __ldg(reinterpret_cast<const float4 *>(rc) + 2);
      s.xi = q0.x; s.xf = q0.y; s.yi = q0.z; s.yf = q0.w;
      s.a00 = q1.x; s.a01 = q1.y; s.a10 = q1.z; s.a11 = q1.w;
      s.p = q2.x; s.c0 = q2.y; s.c1 = q2.z; s.kq = q2.w;
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] += sersic_pixel_f32(s, (float)(l + 8 * j), fy);
    } else {  // point source: at most 7 x 7 pixels of the frame, float64 taps
      const double *d = P.derived + (b * P.ncomp + c) * PSFMC_DERIVED_STRIDE;
      const int ymin = (int)__ldg(d + D_PS_YMIN), ymax = (int)__ldg(d + D_PS_YMAX);
      if (y >= ymin && y <= ymax) {
        const int xmin = (int)__ldg(d + D_PS_XMIN), xmax = (int)__ldg(d + D_PS_XMAX);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int x = l + 8 * j;
          if (x >= xmin && x <= xmax) acc[j] += (float)point_pixel(d, x, y);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = mk<float>(acc[j], acc[j] * acc[j] * wsc);
}

// Special (packed DC/Nyquist) columns: U = FFT(p + i q) of two real sequences that
// are multiplied by different spectra S0 (for p) and S64 (for q):
//   U'[ky] = U[ky] (S0+S64)/2 + conj(U[-ky]) (S0-S64)/2, spectra taken at ky.
__device__ __forceinline__ void special_pair(cplx<float> &u, cplx<float> &um,
                                             cplx<float> s0u, cplx<float> s64u,
                                             cplx<float> s0m, cplx<float> s64m) {
  const cplx<float> pu = mk<float>(0.5f * (s0u.x + s64u.x), 0.5f * (s0u.y + s64u.y));
  const cplx<float> du = mk<float>(0.5f * (s0u.x - s64u.x), 0.5f * (s0u.y - s64u.y));
  const cplx<float> pm = mk<float>(0.5f * (s0m.x + s64m.x), 0.5f * (s0m.y + s64m.y));
  const cplx<float> dm = mk<float>(0.5f * (s0m.x - s64m.x), 0.5f * (s0m.y - s64m.y));
  const cplx<float> nu = u * pu + cconj(um) * du;
  const cplx<float> nm = um * pm + cconj(u) * dm;
  u = nu;
  um = nm;
}

__global__ void __launch_bounds__(PSFMC_FUSED_THREADS, 1)
fused_lnlike_kernel(const FusedParams P) {
  PSFMC_DYN_SMEM(smem_raw);
  cplx<float> *tile = reinterpret_cast<cplx<float> *>(smem_raw);
  __shared__ double red_s[PSFMC_FUSED_THREADS / 32];
  __shared__ int cnt_s;
  constexpr int N = PSFMC_FUSED_N;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) cnt_s = 0;
  __syncthreads();

  // row-phase role: 4 rows per warp, 8 threads per row
  const int rr = lane >> 3, l = lane & 7;
  const bool l0 = (l == 0);
  const int kA = l, kB = l0 ? 8 : 16 - l;
  // column-phase role: column c, two of the eight n2 residues
  const int cg = w & 3, m = w >> 2;
  const int c = cg * 32 + lane;
  const bool special = (c == 0) || (c == 64);
  // radix-8 side of the columns: four k1 values closed under k1 -> -k1 (mod 16)
  const int ck1[4] = {m == 0 ? 0 : m, m == 0 ? 8 : 16 - m, m == 0 ? 4 : 8 - m,
                      m == 0 ? 12 : 8 + m};

  for (long long b = blockIdx.x; b < P.n_batch; b += gridDim.x) {
    int sel = P.psf_sel[b];
    const bool invalid = sel < 0;
    if (invalid) sel = 0;
    const double wscale_b = P.wscale[b];
    const float wsc = (float)wscale_b;
    const float unscale = (float)(P.vscale_inv[sel] / wscale_b);
    const cplx<float> *sp = P.spec + (size_t)sel * N * N;

    // per-thread twiddles of the row phases: W128^(l*k1)
    cplx<float> tw[16];
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) tw[k1] = tw128(l, k1);

    // ---------------------------------------------------------- rows fwd --
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
      const int y = it * 64 + w * 4 + rr;
      cplx<float> *row = tile + y * N;
      const int swz = (y & 1) << 3;
      {
        cplx<float> v[16];
        fused_render16(P, b, y, l, wsc, v);
        dft16<false>(v);
#pragma unroll
        for (int k1 = 1; k1 < 16; ++k1) v[k1] = v[k1] * tw[k1];
        __syncwarp();   // the previous walker's last reads of this row are done
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) row[xpos(k1, l, swz)] = v[k1];
      }
      __syncwarp();
      cplx<float> a[8], bb[8];
#pragma unroll
      for (int n2 = 0; n2 < 8; ++n2) {
        a[n2] = row[xpos(kA, n2, swz)];
        bb[n2] = row[xpos(kB, n2, swz)];
      }
      __syncwarp();
      dft8<float, false>(a);    // a[k2]  = Z[kA + 16 k2]
      dft8<float, false>(bb);   // bb[k2] = Z[kB + 16 k2]
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) {
        {  // kx = kA + 16 k2, partner -kx
          const cplx<float> zk = a[k2];
          const cplx<float> zp = l0 ? a[(8 - k2) & 7] : bb[7 - k2];
          cplx<float> oa = mk<float>(0.5f * (zk.x + zp.x), 0.5f * (zk.y - zp.y));
          cplx<float> ob = mk<float>(0.5f * (zk.y + zp.y), -0.5f * (zk.x - zp.x));
          if (k2 == 0 && l0) {   // real DC / Nyquist columns, packed in pairs
            oa = mk<float>(a[0].x, a[4].x);
            ob = mk<float>(a[0].y, a[4].y);
          }
          row[(kA + 16 * k2) ^ swz] = oa;
          row[(64 + kA + 16 * k2) ^ swz] = ob;
        }
        {  // kx = kB + 16 k2
          const cplx<float> zk = bb[k2];
          const cplx<float> zp = l0 ? bb[7 - k2] : a[7 - k2];
          row[(kB + 16 * k2) ^ swz] =
              mk<float>(0.5f * (zk.x + zp.x), 0.5f * (zk.y - zp.y));
          row[(64 + kB + 16 * k2) ^ swz] =
              mk<float>(0.5f * (zk.y + zp.y), -0.5f * (zk.x - zp.x));
        }
      }
    }

    // spectrum values of the first half of the column multiply (latency is hidden
    // behind the first column pass)
    cplx<float> sa[8], sb[8];
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
      sa[k2] = sp[(ck1[0] + 16 * k2) * N + c];
      sb[k2] = sp[(ck1[1] + 16 * k2) * N + c];
    }
    __syncthreads();

    // ------------------------------------------------- columns: radix-16 --
#pragma unroll 1
    for (int hh = 0; hh < 2; ++hh) {
      const int n2 = m + 4 * hh;
      cplx<float> *col = tile + n2 * N + (c ^ ((n2 & 1) << 3));
      cplx<float> v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = col[8 * j * N];
      dft16<false>(v);
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) v[k1] = v[k1] * tw128(n2, k1);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) col[8 * k1 * N] = v[k1];
    }
    group_barrier(1 + cg, 128);

    // --------------- columns: radix-8, spectrum multiply, inverse radix-8 --
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int k1a = ck1[2 * half], k1b = ck1[2 * half + 1];
      cplx<float> a[8], bb[8];
#pragma unroll
      for (int n2 = 0; n2 < 8; ++n2) {
        const int cc = c ^ ((n2 & 1) << 3);
        a[n2] = tile[(n2 + 8 * k1a) * N + cc];
        bb[n2] = tile[(n2 + 8 * k1b) * N + cc];
      }
      dft8<float, false>(a);    // a[k2]  = U[k1a + 16 k2][c]
      dft8<float, false>(bb);
      if (!special) {
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
          a[k2] = a[k2] * sa[k2];
          bb[k2] = bb[k2] * sb[k2];
        }
      } else {
        const cplx<float> *sx = P.specx + ((size_t)sel * 2 + (c == 64 ? 1 : 0)) * N;
        if (m == 0 && half == 0) {   // k1a = 0: ky <-> (128 - ky); k1b = 8: k2 <-> 7-k2
          special_pair(a[0], a[0], sa[0], sx[0], sa[0], sx[0]);
          special_pair(a[4], a[4], sa[4], sx[64], sa[4], sx[64]);
#pragma unroll
          for (int k2 = 1; k2 < 4; ++k2)
            special_pair(a[k2], a[8 - k2], sa[k2], sx[16 * k2], sa[8 - k2],
                         sx[16 * (8 - k2)]);
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2)
            special_pair(bb[k2], bb[7 - k2], sb[k2], sx[8 + 16 * k2], sb[7 - k2],
                         sx[8 + 16 * (7 - k2)]);
        } else {                     // a[k2] <-> bb[7 - k2]
#pragma unroll
          for (int k2 = 0; k2 < 8; ++k2)
            special_pair(a[k2], bb[7 - k2], sa[k2], sx[k1a + 16 * k2], sb[7 - k2],
                         sx[k1b + 16 * (7 - k2)]);
        }
      }
      if (half == 0) {   // prefetch the second half's spectrum values
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
          sa[k2] = sp[(ck1[2] + 16 * k2) * N + c];
          sb[k2] = sp[(ck1[3] + 16 * k2) * N + c];
        }
      }
      dft8<float, true>(a);     // a[n2]: inverse over k2
      dft8<float, true>(bb);
#pragma unroll
      for (int n2 = 0; n2 < 8; ++n2) {
        const int cc = c ^ ((n2 & 1) << 3);
        tile[(n2 + 8 * k1a) * N + cc] = a[n2];
        tile[(n2 + 8 * k1b) * N + cc] = bb[n2];
      }
    }
    group_barrier(1 + cg, 128);

    // ----------------------------------------- columns: inverse radix-16 --
#pragma unroll 1
    for (int hh = 0; hh < 2; ++hh) {
      const int n2 = m + 4 * hh;
      cplx<float> *col = tile + n2 * N + (c ^ ((n2 & 1) << 3));
      cplx<float> v[16];
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) v[k1] = col[8 * k1 * N];
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) v[k1] = v[k1] * cconj(tw128(n2, k1));
      dft16<true>(v);
#pragma unroll
      for (int j = 0; j < 16; ++j) col[8 * j * N] = v[j];
    }
    __syncthreads();

    // ------------------------------------------------ rows inv + epilogue --
    double acc = 0.0;
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
      const int y = it * 64 + w * 4 + rr;
      cplx<float> *row = tile + y * N;
      const int swz = (y & 1) << 3;
      cplx<float> a[8], bb[8];
      {
        cplx<float> yd1[4], ym1[4], yd2[4], ym2[4];
        cplx<float> a10, b10;
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
          const cplx<float> a1 = row[(kA + 16 * k2) ^ swz];
          const cplx<float> b1 = row[(64 + kA + 16 * k2) ^ swz];
          const cplx<float> a2 = row[(kB + 16 * k2) ^ swz];
          const cplx<float> b2 = row[(64 + kB + 16 * k2) ^ swz];
          if (k2 == 0) {
            a10 = a1;
            b10 = b1;
          }
          // Y[kx] = A' + i B',  Y[-kx] = conj(A') + i conj(B')
          yd1[k2] = mk<float>(a1.x - b1.y, a1.y + b1.x);
          ym1[k2] = mk<float>(a1.x + b1.y, b1.x - a1.y);
          yd2[k2] = mk<float>(a2.x - b2.y, a2.y + b2.x);
          ym2[k2] = mk<float>(a2.x + b2.y, b2.x - a2.y);
        }
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
          a[k2] = yd1[k2];
          bb[k2] = yd2[k2];
          a[4 + k2] = l0 ? ym1[(4 - k2) & 3] : ym2[3 - k2];
          bb[4 + k2] = l0 ? ym2[3 - k2] : ym1[3 - k2];
        }
        if (l0) {   // packed DC / Nyquist columns
          a[0] = mk<float>(a10.x, b10.x);
          a[4] = mk<float>(a10.y, b10.y);
        }
      }
      __syncwarp();
      dft8<float, true>(a);     // a[n2]
      dft8<float, true>(bb);
#pragma unroll
      for (int n2 = 0; n2 < 8; ++n2) {
        row[xpos(kA, n2, swz)] = a[n2];
        row[xpos(kB, n2, swz)] = bb[n2];
      }
      __syncwarp();
      cplx<float> v[16];
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) v[k1] = row[xpos(k1, l, swz)];
      // observation + signed variance of this thread's 16 pixels
      float2 o[16];
      const float2 *owr = P.ow + y * N + l;
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = __ldg(owr + 8 * j);
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) v[k1] = v[k1] * cconj(tw[k1]);
      dft16<true>(v);           // v[j] = (convolved model, scaled model variance)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float resid, ivm;
        const double t = Epilogue<float>::term(v[j].x, v[j].y * unscale, o[j].x,
                                               fabsf(o[j].y), &resid, &ivm);
        if (__float_as_int(o[j].y) >= 0) acc += t;
      }
    }
    // a[4 + k2] for l0 uses ym1[8 - (4 + k2)] = ym1[4 - k2], k2 = 1..3 (k2 = 0 is
    // overwritten by the packed Nyquist value above)

    // float64 reduction: warp shuffles, then the last warp to arrive sums the
    // per-warp partials in fixed order (deterministic) and writes lnL.
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      volatile double *red = red_s;
      red[w] = acc;
      __threadfence_block();
      const int prev = atomicAdd(&cnt_s, 1);
      if (prev == PSFMC_FUSED_THREADS / 32 - 1) {
        __threadfence_block();
        double tot = 0.0;
        for (int k = 0; k < PSFMC_FUSED_THREADS / 32; ++k) tot += red[k];
        double val = -0.5 * tot;
        if (!isfinite(val) || invalid) val = -INFINITY;
        P.lnl[b] = val;
        cnt_s = 0;
      }
    }
  }
}

// -------------------------------------------------------------- host side --

template <typename T>
inline bool fused_path_available(const StagedPlan &plan, const Program &) {
  return sizeof(T) == 4 && plan.fr.H == PSFMC_FUSED_N && plan.fr.W == PSFMC_FUSED_N;
}

template <typename T>
inline int fused_prepare_device(const StagedPlan &) {
#ifndef PSFMC_EMU
  if (cudaFuncSetAttribute(fused_lnlike_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           PSFMC_FUSED_SMEM) != cudaSuccess)
    return 1;
#endif
  return 0;
}

// Re-layout of the float64 spectra [K][2*Wc][H] (column-major, see
// kernels_staged.cuh) for the fused kernel:
//   spec [K][ky][c]: c in 1..63 -> P[ky][kx=c]; c in 65..127 -> V[ky][kx=c-64];
//                    c = 0 -> P[ky][0]; c = 64 -> V[ky][0]
//   specx[K][0][ky] = P[ky][64], specx[K][1][ky] = V[ky][64]
// `vscale[k]` multiplies the V channel (power of two, undone in the epilogue).
inline void fused_spectrum_layout(const cplx<double> *spec64, int n_psf,
                                  const double *vscale, cplx<float> *spec,
                                  cplx<float> *specx) {
  constexpr int N = PSFMC_FUSED_N, Wc = N / 2 + 1;
  for (int k = 0; k < n_psf; ++k) {
    const cplx<double> *src = spec64 + (size_t)k * 2 * Wc * N;
    auto at = [&](int chan, int kx, int ky) {
      const cplx<double> &s = src[((size_t)chan * Wc + kx) * N + ky];
      const double f = chan ? vscale[k] : 1.0;
      cplx<float> o;
      o.x = (float)(s.x * f);
      o.y = (float)(s.y * f);
      return o;
    };
    for (int ky = 0; ky < N; ++ky) {
      for (int c = 0; c < N; ++c) {
        const int chan = c >= 64 ? 1 : 0, kx = c & 63;
        spec[((size_t)k * N + ky) * N + c] = at(chan, kx, ky);
      }
      specx[((size_t)k * 2 + 0) * N + ky] = at(0, 64, ky);
      specx[((size_t)k * 2 + 1) * N + ky] = at(1, 64, ky);
    }
  }
}

struct FusedBuffers {
  float *rconst = nullptr;
  const cplx<float> *spec = nullptr, *specx = nullptr;
  const float2 *ow = nullptr;
  int n_sms = 148;
};

// theta -> lnL for n_batch walkers: prepare kernel + one persistent fused kernel.
// Returns the number of kernels launched.
template <typename T>
inline int launch_fused_lnlike(const StagedPlan &plan, const StagedBuffers<T> &buf,
                               const FusedBuffers &fb, const Program &prog_h,
                               const double *theta, long long n_batch, long long ld,
                               double *lnl, cudaStream_t stream) {
  if (n_batch <= 0) return 0;
  const int ncomp = prog_h.n_components;
  {
    long long nthreads = n_batch * (ncomp > 0 ? ncomp : 1);
    int block = 128;
    unsigned grid = (unsigned)((nthreads + block - 1) / block);
    launch_kernel(prepare_kernel, dim3(grid), dim3(block), 0, stream, buf.prog, theta,
                  n_batch, ld, plan.fr.H, plan.fr.W, buf.derived, buf.psf_sel, buf.wscale,
                  fb.rconst);
  }
  FusedParams P;
  P.rconst = fb.rconst;
  P.derived = buf.derived;
  P.wscale = buf.wscale;
  P.psf_sel = buf.psf_sel;
  P.vscale_inv = buf.vscale_inv;
  P.spec = fb.spec;
  P.specx = fb.specx;
  P.ow = fb.ow;
  P.lnl = lnl;
  P.n_batch = n_batch;
  P.ncomp = ncomp;
  for (int c = 0; c < PSFMC_MAX_COMPONENTS; ++c)
    P.kind[c] = (signed char)(c < ncomp ? prog_h.kind[c] : 0);
  unsigned grid = (unsigned)(n_batch < fb.n_sms ? n_batch : fb.n_sms);
  launch_kernel(fused_lnlike_kernel, dim3(grid), dim3(PSFMC_FUSED_THREADS),
                (size_t)PSFMC_FUSED_SMEM, stream, P);
  return 2;
}

}  // namespace psfmc
