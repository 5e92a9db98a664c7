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
  const cplx<float> *specx;  // [K][2 (P,V)][ky=128]: (S0-S64)/2 of the packed DC/Nyquist
                             // columns (S0 = kx 0, S64 = kx 64); (S0+S64)/2 is in spec
  const float2 *ow;          // [128*128]: (obs, bad ? -ovar : +ovar)
  double *lnl;               // [B]
  long long n_batch;
  int ncomp;
  unsigned skip_quads;       // bit q: rows 4q .. 4q+3 hold no good pixel (mask, bad pixels,
                             // padding): their inverse row transform and epilogue are skipped
  signed char kind[PSFMC_MAX_COMPONENTS];
};

// padded frames (see Frame in common.cuh): observation frame and fold bounds
struct FoldParams {
  int Hr, Wr, fy_hi, fy_lo, fx_hi, fx_lo;
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
  // v[4*k1 + n2] *= W16^(n2*k1) = (cos, -sin)(pi e / 8), conjugated for the inverse
  v[4 * 1 + 1] = crot<float, INV>(v[4 * 1 + 1], C, S);     // e = 1
  v[4 * 1 + 2] = crot<float, INV>(v[4 * 1 + 2], h, h);     // e = 2
  v[4 * 1 + 3] = crot<float, INV>(v[4 * 1 + 3], S, C);     // e = 3
  v[4 * 2 + 1] = crot<float, INV>(v[4 * 2 + 1], h, h);     // e = 2
  v[4 * 2 + 2] = INV ? mul_pos_i(v[4 * 2 + 2]) : mul_neg_i(v[4 * 2 + 2]);  // e = 4
  v[4 * 2 + 3] = crot<float, INV>(v[4 * 2 + 3], -h, h);    // e = 6
  v[4 * 3 + 1] = crot<float, INV>(v[4 * 3 + 1], S, C);     // e = 3
  v[4 * 3 + 2] = crot<float, INV>(v[4 * 3 + 2], -h, h);    // e = 6
  v[4 * 3 + 3] = crot<float, INV>(v[4 * 3 + 3], -C, -S);   // e = 9
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

// Shared-memory tile access by BYTE offset (32-bit address arithmetic; XOR swizzles
// act directly on the address). Exchange layout of a row between its radix-16 and
// radix-8 sides: element (k1, n2) of row y sits at complex position
//     8 * (k1 ^ s) + (n2 ^ (k1 & 7)),   s = y & 1,
// which is free of bank conflicts on both sides (fixed k1 / lanes over n2, and fixed
// n2 / lanes over k1) and keeps the two rows of a half-warp on different banks.
// All its fields are bit-disjoint, so the byte offset is (thread constant) XOR
// (compile-time constant): one LOP3 per access.
// The tile is 1024-byte aligned, so a row base plus a thread constant below 1024
// can be XOR-ed with compile-time constants on the final address.
#ifdef PSFMC_EMU
typedef uintptr_t smem_addr_t;
#else
typedef unsigned smem_addr_t;
#endif

__device__ __forceinline__ cplx<float> lds64(smem_addr_t addr) {
#ifdef PSFMC_EMU
  return *reinterpret_cast<const cplx<float> *>(addr);
#else
  cplx<float> v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
#endif
}
__device__ __forceinline__ void sts64(smem_addr_t addr, cplx<float> v) {
#ifdef PSFMC_EMU
  *reinterpret_cast<cplx<float> *>(addr) = v;
#else
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
#endif
}
__device__ __forceinline__ smem_addr_t smem_base(unsigned char *ptr) {
#ifdef PSFMC_EMU
  return (smem_addr_t)ptr;
#else
  return (smem_addr_t)__cvta_generic_to_shared(ptr);
#endif
}

// Raw model at the 16 pixels x = l + XS*j of row y -> packed z = raw + i*wsc*raw^2
// (XS = 8: 128-wide rows, XS = 16: 256-wide rows of the cluster kernel).
// Pixels are rendered in pairs (j, j+1) with element-wise pair arithmetic.
template <int XS, bool STAGED>
__device__ __forceinline__ void fused_render16(const FusedParams &P, const float *rc0,
                                               const double *der0, int y, int l, float wsc,
                                               cplx<float> *v) {
  // rc0 / der0: this walker's render constants / float64 constants; STAGED = they were
  // copied to shared memory beforehand (plain loads), else read-only global loads
  cplx<float> acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = mk<float>(0.0f, 0.0f);
  for (int c = 0; c < P.ncomp; ++c) {
    const int kind = P.kind[c];
    const float *rc = rc0 + c * PSFMC_RC_STRIDE;
    if (kind == PSFMC_SKY) {
      const cplx<float> adu = bcast(STAGED ? *rc : __ldg(rc));
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = acc[i] + adu;
    } else if (kind == PSFMC_SERSIC) {
      SersicF32 s;
      const float4 *rc4 = reinterpret_cast<const float4 *>(rc);
      const float4 q0 = STAGED ? rc4[0] : __ldg(rc4);
      const float4 q1 = STAGED ? rc4[1] : __ldg(rc4 + 1);
      const float4 q2 = STAGED ? rc4[2] : __ldg(rc4 + 2);
      s.xi = q0.x; s.xf = q0.y; s.yi = q0.z; s.yf = q0.w;
      s.a00 = q1.x; s.a01 = q1.y; s.a10 = q1.z; s.a11 = q1.w;
      s.p = q2.x; s.c0 = q2.y; s.c1 = q2.z; s.kq = q2.w;
      const float dy = ((float)y - s.yi) - s.yf;
      const float cu = s.a01 * dy, cv = s.a11 * dy, dy2 = dy * dy;
      // (x - xi) is an exact small integer; subtracting the fraction LAST keeps dx
      // accurate to an ulp of dx itself next to the centre (where the profile and its
      // centroid correction vary as a power of the distance)
      const cplx<float> dxi = bcast((float)l - s.xi), nxf = bcast(-s.xf);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const cplx<float> dx =
            (dxi + mk<float>((float)(2 * XS * i), (float)(2 * XS * i + XS))) + nxf;
        acc[i] = acc[i] + sersic_pair_f32(s, dx, cu, cv, dy2);
      }
    } else {  // point source: at most 7 x 7 pixels of the frame, float64 taps
      const double *d = der0 + c * PSFMC_DERIVED_STRIDE;
      const int ymin = (int)(STAGED ? d[D_PS_YMIN] : __ldg(d + D_PS_YMIN));
      const int ymax = (int)(STAGED ? d[D_PS_YMAX] : __ldg(d + D_PS_YMAX));
      if (y >= ymin && y <= ymax) {
        const int xmin = (int)(STAGED ? d[D_PS_XMIN] : __ldg(d + D_PS_XMIN));
        const int xmax = (int)(STAGED ? d[D_PS_XMAX] : __ldg(d + D_PS_XMAX));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int x = l + 2 * XS * i;
          if (x >= xmin && x <= xmax) acc[i].x += (float)point_pixel(d, x, y);
          if (x + XS >= xmin && x + XS <= xmax) acc[i].y += (float)point_pixel(d, x + XS, y);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const cplx<float> sq = pmul(pmul(acc[i], acc[i]), bcast(wsc));
    v[2 * i] = mk<float>(acc[i].x, sq.x);
    v[2 * i + 1] = mk<float>(acc[i].y, sq.y);
  }
}

__device__ __forceinline__ cplx<float> ldc2(const cplx<float> *p) {
#ifdef PSFMC_EMU
  return *p;
#else
  const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
  return mk<float>(v.x, v.y);
#endif
}

// Special (packed DC/Nyquist) columns: U = FFT(p + i q) of two real sequences that
// are multiplied by different spectra S0 (for p) and S64 (for q):
//   U'[ky] = U[ky] (S0+S64)/2 + conj(U[-ky]) (S0-S64)/2, spectra taken at ky.
// pu / pm: half-sums at ky / -ky (they sit in the `spec` table at c = 0 and c = 64),
// du / dm: half-differences (from `specx`), all prefetched by the caller.
__device__ __forceinline__ void special_pair(cplx<float> &u, cplx<float> &um,
                                             cplx<float> pu, cplx<float> du,
                                             cplx<float> pm, cplx<float> dm) {
  const cplx<float> nu = u * pu + cmul_conj(du, um);
  const cplx<float> nm = um * pm + cmul_conj(dm, u);
  u = nu;
  um = nm;
}

// Per-thread constants of the row passes (4 rows per warp, 8 threads per row).
struct RowRole {
  int w, rr, l;
  bool l0;
  unsigned t16;     // radix-16 side of the exchange layout: (8 l) ^ (64 s)
  unsigned qa, qb;  // radix-8 side, for k1 = kA / kB: 64 (k ^ s) + 8 (k & 7)
  unsigned fa, fb;  // column layout: 8 (k ^ 8 s)
};

// render + forward row transform + real-pair split of row batch `it` of walker b
// (rc0 / der0: the walker's constants, STAGED = in shared memory, see fused_render16)
// PADDED: the observation frame is P.Hr x P.Wr in the corner of the 128 x 128 transform
// frame; nothing is rendered outside it.
template <bool STAGED, bool PADDED = false>
__device__ __forceinline__ void fused_rows_forward(const FusedParams &P, smem_addr_t tile,
                                                   const RowRole &R, smem_addr_t twl,
                                                   const float *rc0, const double *der0,
                                                   int it, float wsc,
                                                   const FoldParams *F = nullptr) {
  const int y = it * 64 + R.w * 4 + R.rr;
  const smem_addr_t rb = tile + (unsigned)y * (PSFMC_FUSED_N * 8);
  if (PADDED && y - R.rr >= F->Hr) {
    // all four rows of the warp lie outside the observation frame: their row spectra
    // are zero (the tile still holds the previous walker's values there)
    __syncwarp();
    const cplx<float> zero = mk<float>(0.0f, 0.0f);
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
      sts64(rb + R.fa + 8 * 16 * k2, zero);
      sts64(rb + R.fa + 8 * (64 + 16 * k2), zero);
      sts64(rb + R.fb + 8 * 16 * k2, zero);
      sts64(rb + R.fb + 8 * (64 + 16 * k2), zero);
    }
    return;
  }
  {
    cplx<float> v[16];
    if (!PADDED || y < F->Hr) {
      fused_render16<8, STAGED>(P, rc0, der0, y, R.l, wsc, v);
      if (PADDED) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (R.l + 8 * j >= F->Wr) v[j] = mk<float>(0.0f, 0.0f);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = mk<float>(0.0f, 0.0f);
    }
    dft16<false>(v);
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) v[k1] = v[k1] * lds64(twl + 64 * k1);
    __syncwarp();   // every lane is done reading this row (previous walker)
    const smem_addr_t rt = rb + R.t16;
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) sts64(rt ^ (unsigned)(64 * k1 + 8 * (k1 & 7)), v[k1]);
  }
  __syncwarp();
  cplx<float> a[8], bb[8];
  {
    const smem_addr_t ra = rb + R.qa, rq = rb + R.qb;
#pragma unroll
    for (int n2 = 0; n2 < 8; ++n2) {
      a[n2] = lds64(ra ^ (unsigned)(8 * n2));
      bb[n2] = lds64(rq ^ (unsigned)(8 * n2));
    }
  }
  __syncwarp();
  dft8<float, false>(a);    // a[k2]  = Z[kA + 16 k2]
  dft8<float, false>(bb);   // bb[k2] = Z[kB + 16 k2]
  const bool l0 = R.l0;
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) {
    // row spectra of the two real images (the factor 1/2 of the split is folded
    // into the PSF spectra):  A = Z[kx] + conj Z[-kx],  B = -i (Z[kx] - conj Z[-kx])
    {  // kx = kA + 16 k2, partner -kx
      const cplx<float> zk = a[k2];
      const cplx<float> zp = l0 ? a[(8 - k2) & 7] : bb[7 - k2];
      cplx<float> oa = zk + mk<float>(zp.x, -zp.y);
      cplx<float> ob = mk<float>(zk.y, -zk.x) + mk<float>(zp.y, zp.x);
      if (k2 == 0 && l0) {   // real DC / Nyquist columns, packed in pairs
        oa = mk<float>(a[0].x, a[4].x);
        ob = mk<float>(a[0].y, a[4].y);
      }
      sts64(rb + R.fa + 8 * 16 * k2, oa);
      sts64(rb + R.fa + 8 * (64 + 16 * k2), ob);
    }
    {  // kx = kB + 16 k2
      const cplx<float> zk = bb[k2];
      const cplx<float> zp = l0 ? bb[7 - k2] : a[7 - k2];
      sts64(rb + R.fb + 8 * 16 * k2, zk + mk<float>(zp.x, -zp.y));
      sts64(rb + R.fb + 8 * (64 + 16 * k2), mk<float>(zk.y, -zk.x) + mk<float>(zp.y, zp.x));
    }
  }
}

// Hermitian rebuild + inverse row transform + chi-square terms of row batch `it`;
// returns this thread's float64 partial sum over its 16 pixels. PREFETCH: issue the
// observation loads before the transform (needs 32 registers for its duration).
template <bool PREFETCH, bool PADDED = false>
__device__ __forceinline__ double fused_rows_inverse(const FusedParams &P, smem_addr_t tile,
                                                     const RowRole &R, smem_addr_t twl,
                                                     int it, float unscale,
                                                     const FoldParams *F = nullptr) {
  const int y = it * 64 + R.w * 4 + R.rr;
  if (PADDED && y - R.rr >= F->Hr) return 0.0;   // the warp's four rows are padding
  // ... or hold no unmasked pixel: nothing of them enters the sum (models.py:233-236)
  if ((P.skip_quads >> (it * 16 + R.w)) & 1u) return 0.0;
  const smem_addr_t rb = tile + (unsigned)y * (PSFMC_FUSED_N * 8);
  const bool l0 = R.l0;
  // observation + signed variance of this thread's 16 pixels: issued first, used
  // last (L2 latency hidden behind the whole inverse transform)
  float2 o[16];
  const float2 *owr = P.ow + y * PSFMC_FUSED_N + R.l;
  if (PREFETCH) {
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = __ldg(owr + 8 * j);
  }
  cplx<float> a[8], bb[8];
  {
    cplx<float> yd1[4], ym1[4], yd2[4], ym2[4];
    cplx<float> a10, b10;
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
      const cplx<float> a1 = lds64(rb + R.fa + 8 * 16 * k2);
      const cplx<float> b1 = lds64(rb + R.fa + 8 * (64 + 16 * k2));
      const cplx<float> a2 = lds64(rb + R.fb + 8 * 16 * k2);
      const cplx<float> b2 = lds64(rb + R.fb + 8 * (64 + 16 * k2));
      if (k2 == 0) {
        a10 = a1;
        b10 = b1;
      }
      // Y[kx] = A' + i B',  Y[-kx] = conj(A') + i conj(B')
      yd1[k2] = a1 + mk<float>(-b1.y, b1.x);
      ym1[k2] = mk<float>(a1.x, -a1.y) + mk<float>(b1.y, b1.x);
      yd2[k2] = a2 + mk<float>(-b2.y, b2.x);
      ym2[k2] = mk<float>(a2.x, -a2.y) + mk<float>(b2.y, b2.x);
    }
    // a[k2] = Y[kA + 16 k2], bb[k2] = Y[kB + 16 k2]; the upper halves are mirrored
    // entries of the other set (of the same set for l = 0, where kA = 0, kB = 8)
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
  {
    const smem_addr_t ra = rb + R.qa, rq = rb + R.qb;
#pragma unroll
    for (int n2 = 0; n2 < 8; ++n2) {
      sts64(ra ^ (unsigned)(8 * n2), a[n2]);
      sts64(rq ^ (unsigned)(8 * n2), bb[n2]);
    }
  }
  __syncwarp();
  cplx<float> v[16];
  {
    const smem_addr_t rt = rb + R.t16;
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) v[k1] = lds64(rt ^ (unsigned)(64 * k1 + 8 * (k1 & 7)));
  }
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) v[k1] = cmul_conj(v[k1], lds64(twl + 64 * k1));
  dft16<true>(v);           // v[j] = (convolved model, scaled model variance)
  if (PADDED) {
    // fold the linear convolution back modulo Wr (see Frame): through the row's own
    // 1 KB of the tile (free once every lane has taken its exchange values), pixel x at
    // position x
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; ++j) sts64(rb + 8u * (R.l + 8 * j), v[j]);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int x = R.l + 8 * j;
      if (x <= F->fx_hi) v[j] = v[j] + lds64(rb + 8u * (x + F->Wr));
      if (x >= F->fx_lo && x < F->Wr)
        v[j] = v[j] + lds64(rb + 8u * (PSFMC_FUSED_N + x - F->Wr));
    }
    __syncwarp();
  }
  if (!PREFETCH) {
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = __ldg(owr + 8 * j);
  }
  double acc = 0.0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float resid, ivm;
    const double t = Epilogue<float>::term(v[j].x, v[j].y * unscale, o[j].x,
                                           fabsf(o[j].y), &resid, &ivm);
    if (__float_as_int(o[j].y) >= 0) acc += t;
  }
  return acc;
}

template <bool PADDED>
__global__ void __launch_bounds__(PSFMC_FUSED_THREADS, 1)
fused_lnlike_kernel(const FusedParams P, const FoldParams F) {
  PSFMC_DYN_SMEM(smem_raw);
  const smem_addr_t tile = smem_base(smem_raw);
  __shared__ double red_s[PSFMC_FUSED_THREADS / 32];
  __shared__ int cnt_s;
  constexpr int N = PSFMC_FUSED_N;
  constexpr unsigned ROWB = N * 8;   // bytes per tile row
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) cnt_s = 0;

  // row-pass role
  RowRole R;
  R.w = w;
  R.rr = lane >> 3;
  R.l = lane & 7;
  R.l0 = (R.l == 0);
  {
    const unsigned s = R.rr & 1;
    const unsigned kA = R.l, kB = R.l0 ? 8 : 16 - R.l;
    R.t16 = (8u * R.l) ^ (64u * s);
    R.qa = 64u * (kA ^ s) + 8u * (kA & 7);
    R.qb = 64u * (kB ^ s) + 8u * (kB & 7);
    R.fa = 8u * (kA ^ (8u * s));
    R.fb = 8u * (kB ^ (8u * s));
  }
  // twiddles of the row passes W128^(l*k1) in shared memory as [k1][l]: the eight
  // threads of a row read eight consecutive entries (no bank conflicts, 4 rows
  // broadcast); this thread's entries are twl + 64 * k1
  __shared__ __align__(16) float tw_s[128][2];
  if (tid < 128) {
    const int k1 = tid >> 3, ll = tid & 7;
    tw_s[tid][0] = c_tw128[ll * 16 + k1][0];
    tw_s[tid][1] = c_tw128[ll * 16 + k1][1];
  }
  const smem_addr_t twl = smem_base(reinterpret_cast<unsigned char *>(&tw_s[0][0])) + 8u * R.l;
  // constants of the walker the next forward pass renders, staged one walker ahead
  // (during the column passes) so that the render does not wait for L2
  __shared__ __align__(16) float rc_s[PSFMC_MAX_COMPONENTS * PSFMC_RC_STRIDE];
  __shared__ double der_s[PSFMC_MAX_COMPONENTS * PSFMC_DERIVED_STRIDE];
  __shared__ float wsc_s;
  auto stage_params = [&](long long bs) {
    if (bs >= P.n_batch) return;
    if (tid < P.ncomp * PSFMC_RC_STRIDE)
      rc_s[tid] = __ldg(P.rconst + bs * P.ncomp * PSFMC_RC_STRIDE + tid);
    for (int k = tid; k < P.ncomp * PSFMC_DERIVED_STRIDE; k += PSFMC_FUSED_THREADS)
      der_s[k] = __ldg(P.derived + bs * P.ncomp * PSFMC_DERIVED_STRIDE + k);
    if (tid == PSFMC_FUSED_THREADS - 1) wsc_s = (float)P.wscale[bs];
  };
  stage_params(blockIdx.x);
  __syncthreads();

  // column-pass role: column c, two of the eight n2 residues. The four warps of a
  // column group (which meet at the named barriers) sit on four DIFFERENT
  // schedulers (warp w runs on scheduler w & 3), so every scheduler hosts one warp
  // of each group and its warps are not phase-locked to each other.
  const int cg = w >> 2, m = w & 3;
  const int c = cg * 32 + lane;
  const bool special = (c == 0) || (c == 64);
  // radix-8 side of the columns: four k1 values closed under k1 -> -k1 (mod 16)
  const int ck1[4] = {m == 0 ? 0 : m, m == 0 ? 8 : 16 - m, m == 0 ? 4 : 8 - m,
                      m == 0 ? 12 : 8 + m};
  // byte offsets of column c in even / odd rows (row swizzle = 8 * parity)
  const unsigned cev = 8u * c, cod = 8u * (c ^ 8);

  // Order of the row work between two column passes. Every warp owns two row
  // batches; per batch the inverse pass of walker b must precede the forward pass
  // of the next walker (same rows). Half of the warps of each scheduler interleave
  // (inv 0, fwd 0, inv 1, fwd 1), the other half run (inv 0, inv 1, fwd 0, fwd 1),
  // so the SFU-bound render of some warps overlaps the FMA/LSU-bound inverse
  // transforms of the others.
  const bool interleave = ((w >> 2) & 1) != 0;

  // The first pass through the loop (b < 0) only renders and forward-transforms the
  // CTA's first walker; every later pass runs the column passes of walker b, then
  // its inverse rows interleaved with the forward rows of the CTA's next walker.
#pragma unroll 1
  for (long long b = (long long)blockIdx.x - (long long)gridDim.x; b < P.n_batch;
       b += gridDim.x) {
    const bool cur = b >= 0;
    int sel = cur ? P.psf_sel[b] : 0;
    const bool invalid = sel < 0;
    if (invalid) sel = 0;
    const double wscale_b = cur ? P.wscale[b] : 1.0;
    const float unscale = (float)(P.vscale_inv[sel] / wscale_b);
    const cplx<float> *sp = P.spec + (size_t)sel * N * N;
    if (cur) {

    __syncthreads();
    // every warp is past the forward rows of walker b: stage walker b + grid
    stage_params(b + gridDim.x);

    // ------------------------------------------------- columns: radix-16 --
    // both residues n2 = m and m + 4 of this thread in flight at once (ILP)
    {
      const smem_addr_t cb0 = tile + (unsigned)m * ROWB + ((m & 1) ? cod : cev);
      const smem_addr_t cb1 = cb0 + 4 * ROWB;
      cplx<float> v0[16], v1[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v0[j] = lds64(cb0 + 8 * j * ROWB);
#pragma unroll
      for (int j = 0; j < 16; ++j) v1[j] = lds64(cb1 + 8 * j * ROWB);
      dft16<false>(v0);
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) v0[k1] = v0[k1] * tw128(m, k1);
      dft16<false>(v1);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) sts64(cb0 + 8 * k1 * ROWB, v0[k1]);
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) v1[k1] = v1[k1] * tw128(m + 4, k1);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) sts64(cb1 + 8 * k1 * ROWB, v1[k1]);
    }
    // spectrum values of the first half of the column multiply, issued before the
    // group barrier so that their L2 latency overlaps the wait
    cplx<float> sa[8], sb[8];
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
      sa[k2] = sp[(ck1[0] + 16 * k2) * N + c];
      sb[k2] = sp[(ck1[1] + 16 * k2) * N + c];
    }
    // half-difference table of the two special columns (one lane in 8 of 16 warps;
    // 1 KB per PSF, L1-resident across walkers)
    const cplx<float> *sx = P.specx + ((size_t)sel * 2 + (c == 64 ? 1 : 0)) * N;
    group_barrier(1 + cg, 128);

    // --------------- columns: radix-8, spectrum multiply, inverse radix-8 --
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int k1a = ck1[2 * half], k1b = ck1[2 * half + 1];
      const smem_addr_t ba = tile + 8u * k1a * ROWB, bq = tile + 8u * k1b * ROWB;
      cplx<float> a[8], bb[8];
#pragma unroll
      for (int n2 = 0; n2 < 8; ++n2) {
        const unsigned cc = (unsigned)n2 * ROWB + ((n2 & 1) ? cod : cev);
        a[n2] = lds64(ba + cc);
        bb[n2] = lds64(bq + cc);
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
        if (m == 0 && half == 0) {   // k1a = 0: ky <-> (128 - ky); k1b = 8: k2 <-> 7-k2
          special_pair(a[0], a[0], sa[0], ldc2(sx), sa[0], ldc2(sx));
          special_pair(a[4], a[4], sa[4], ldc2(sx + 64), sa[4], ldc2(sx + 64));
#pragma unroll
          for (int k2 = 1; k2 < 4; ++k2)
            special_pair(a[k2], a[8 - k2], sa[k2], ldc2(sx + 16 * k2), sa[8 - k2],
                         ldc2(sx + 16 * (8 - k2)));
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2)
            special_pair(bb[k2], bb[7 - k2], sb[k2], ldc2(sx + 8 + 16 * k2), sb[7 - k2],
                         ldc2(sx + 8 + 16 * (7 - k2)));
        } else {                     // a[k2] <-> bb[7 - k2]
#pragma unroll
          for (int k2 = 0; k2 < 8; ++k2)
            special_pair(a[k2], bb[7 - k2], sa[k2], ldc2(sx + k1a + 16 * k2), sb[7 - k2],
                         ldc2(sx + k1b + 16 * (7 - k2)));
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
        const unsigned cc = (unsigned)n2 * ROWB + ((n2 & 1) ? cod : cev);
        sts64(ba + cc, a[n2]);
        sts64(bq + cc, bb[n2]);
      }
    }
    group_barrier(1 + cg, 128);

    // ----------------------------------------- columns: inverse radix-16 --
    {
      const smem_addr_t cb0 = tile + (unsigned)m * ROWB + ((m & 1) ? cod : cev);
      const smem_addr_t cb1 = cb0 + 4 * ROWB;
      cplx<float> v0[16], v1[16];
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) v0[k1] = lds64(cb0 + 8 * k1 * ROWB);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) v1[k1] = lds64(cb1 + 8 * k1 * ROWB);
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) v0[k1] = cmul_conj(v0[k1], tw128(m, k1));
      dft16<true>(v0);
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) v1[k1] = cmul_conj(v1[k1], tw128(m + 4, k1));
#pragma unroll
      for (int j = 0; j < 16; ++j) sts64(cb0 + 8 * j * ROWB, v0[j]);
      dft16<true>(v1);
#pragma unroll
      for (int j = 0; j < 16; ++j) sts64(cb1 + 8 * j * ROWB, v1[j]);
    }
    __syncthreads();
    if (PADDED) {
      // fold the linear convolution back modulo Hr (see Frame): row p receives rows
      // p + Hr (p <= fy_hi) and 128 + p - Hr (p >= fy_lo); the sources lie at or beyond
      // row Hr, the targets below it. Column c of row y sits at c ^ (8 * (y & 1)).
      for (int e = tid; e < F.Hr * N; e += PSFMC_FUSED_THREADS) {
        const int p = e >> 7, cc = e & (N - 1);
        const bool hi = p <= F.fy_hi, lo = p >= F.fy_lo;
        if (hi || lo) {
          const smem_addr_t dst = tile + (unsigned)p * ROWB + 8u * (cc ^ (8 * (p & 1)));
          cplx<float> acc = lds64(dst);
          if (hi) {
            const int q = p + F.Hr;
            acc = acc + lds64(tile + (unsigned)q * ROWB + 8u * (cc ^ (8 * (q & 1))));
          }
          if (lo) {
            const int q = N + p - F.Hr;
            acc = acc + lds64(tile + (unsigned)q * ROWB + 8u * (cc ^ (8 * (q & 1))));
          }
          sts64(dst, acc);
        }
      }
      __syncthreads();
    }
    }  // if (cur)

    // ------ rows: inverse + chi-square of walker b, render + forward of the next --
    const long long bn = b + gridDim.x;
    const bool has_next = bn < P.n_batch;
    const float wsc_next = has_next ? wsc_s : 0.0f;
    double acc = 0.0;
#pragma unroll 1
    for (int step = 0; step < 4; ++step) {
      // interleave: I0 F0 I1 F1 ; otherwise: I0 I1 F0 F1
      const bool fwd = interleave ? (step & 1) : (step >= 2);
      const int it = interleave ? (step >> 1) : (step & 1);
      if (!fwd) {
        if (cur) acc += fused_rows_inverse<true, PADDED>(P, tile, R, twl, it, unscale, &F);
      } else if (has_next) {
        fused_rows_forward<true, PADDED>(P, tile, R, twl, rc_s, der_s, it, wsc_next, &F);
      }
      if (cur && ((interleave && step == 2) || (!interleave && step == 1))) {
        // both inverse batches of this warp are done: float64 reduction. Warp
        // shuffles, then the last warp to arrive sums the per-warp partials in
        // fixed order (deterministic) and writes lnL.
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
          acc += __shfl_down_sync(0xffffffffu, acc, off);
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
  }
}

// --------------------------------------------------------------------------
// 1024-thread variant: the same passes with ONE 16-point unit per thread and pass
// (32 warps, 64 registers): twice the resident warps per scheduler to hide the
// latencies the 512-thread kernel is bound by. Rows: warp w owns rows 4w..4w+3.
// Columns: thread (c, m), m = 0..7: radix-16 of residue n2 = m, then the radix-8
// pair k1 in {m, 16 - m} ({0, 8} for m = 0), closed under k1 -> -k1. The eight warps of
// a column group are spread over the four schedulers (see fused_lnlike_kernel).
#define PSFMC_FUSED_THREADS_WIDE 1024

// Thread roles are recomputed from an opaque copy of the thread index at the start of
// every phase, so that the compiler cannot keep them alive across phases (64
// registers per thread leave no room for loop-invariant luggage).
__device__ __forceinline__ int opaque_tid() {
  int t = threadIdx.x;
#ifndef PSFMC_EMU
  asm volatile("" : "+r"(t));
#endif
  return t;
}

__device__ __forceinline__ RowRole make_row_role(int tid) {
  RowRole R;
  const int lane = tid & 31;
  R.w = tid >> 5;
  R.rr = lane >> 3;
  R.l = lane & 7;
  R.l0 = (R.l == 0);
  const unsigned s = R.rr & 1;
  const unsigned kA = R.l, kB = R.l0 ? 8 : 16 - R.l;
  R.t16 = (8u * R.l) ^ (64u * s);
  R.qa = 64u * (kA ^ s) + 8u * (kA & 7);
  R.qb = 64u * (kB ^ s) + 8u * (kB & 7);
  R.fa = 8u * (kA ^ (8u * s));
  R.fb = 8u * (kB ^ (8u * s));
  return R;
}

__global__ void __launch_bounds__(PSFMC_FUSED_THREADS_WIDE, 1)
fused_lnlike_kernel_wide(const FusedParams P) {
  PSFMC_DYN_SMEM(smem_raw);
  const smem_addr_t tile = smem_base(smem_raw);
  constexpr int NW = PSFMC_FUSED_THREADS_WIDE / 32;
  __shared__ double red_s[NW];
  __shared__ int cnt_s;
  constexpr int N = PSFMC_FUSED_N;
  constexpr unsigned ROWB = N * 8;
  __shared__ __align__(16) float tw_s[128][2];
  {
    const int tid = threadIdx.x;
    if (tid == 0) cnt_s = 0;
    if (tid < 128) {
      const int k1 = tid >> 3, ll = tid & 7;
      tw_s[tid][0] = c_tw128[ll * 16 + k1][0];
      tw_s[tid][1] = c_tw128[ll * 16 + k1][1];
    }
  }
  const smem_addr_t tw_base = smem_base(reinterpret_cast<unsigned char *>(&tw_s[0][0]));
  __syncthreads();

#pragma unroll 1
  for (long long b = (long long)blockIdx.x - (long long)gridDim.x; b < P.n_batch;
       b += gridDim.x) {
    const bool cur = b >= 0;
    if (cur) {
      int sel = P.psf_sel[b];
      const bool invalid = sel < 0;
      if (invalid) sel = 0;
      const cplx<float> *sp = P.spec + (size_t)sel * N * N;
      __syncthreads();
      // ---- columns: radix-16 of residue n2 = m
      {
        const int tid = opaque_tid();
        const int w = tid >> 5, m = w & 7, cg = w >> 3, c = cg * 32 + (tid & 31);
        const smem_addr_t cb = tile + (unsigned)m * ROWB + 8u * ((m & 1) ? (c ^ 8) : c);
        cplx<float> v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = lds64(cb + 8 * j * ROWB);
        dft16<false>(v);
#pragma unroll
        for (int k1 = 1; k1 < 16; ++k1) v[k1] = v[k1] * tw128(m, k1);
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) sts64(cb + 8 * k1 * ROWB, v[k1]);
        group_barrier(1 + cg, 256);
      }

      // ---- columns: radix-8 pair, spectrum multiply, inverse radix-8
      {
        const int tid = opaque_tid();
        const int w = tid >> 5, m = w & 7, cg = w >> 3, c = cg * 32 + (tid & 31);
        const bool special = (c == 0) || (c == 64);
        const int k1a = m, k1b = (m == 0) ? 8 : 16 - m;
        const unsigned cev = 8u * c, cod = 8u * (c ^ 8);
        const smem_addr_t ba = tile + 8u * k1a * ROWB, bq = tile + 8u * k1b * ROWB;
        const cplx<float> *spa = sp + k1a * N + c, *spb = sp + k1b * N + c;
        cplx<float> a[8], bb[8];
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) {
          const unsigned cc = (unsigned)n2 * ROWB + ((n2 & 1) ? cod : cev);
          a[n2] = lds64(ba + cc);
          bb[n2] = lds64(bq + cc);
        }
        dft8<float, false>(a);    // a[k2]  = U[k1a + 16 k2][c]
        dft8<float, false>(bb);
        if (!special) {
#pragma unroll
          for (int k2 = 0; k2 < 8; ++k2) a[k2] = a[k2] * ldc2(spa + 16 * k2 * N);
#pragma unroll
          for (int k2 = 0; k2 < 8; ++k2) bb[k2] = bb[k2] * ldc2(spb + 16 * k2 * N);
        } else {
          const cplx<float> *sx = P.specx + ((size_t)sel * 2 + (c == 64 ? 1 : 0)) * N;
          if (m == 0) {   // k1a = 0: ky <-> (128 - ky); k1b = 8: k2 <-> 7 - k2
            special_pair(a[0], a[0], ldc2(spa), ldc2(sx), ldc2(spa), ldc2(sx));
            special_pair(a[4], a[4], ldc2(spa + 64 * N), ldc2(sx + 64), ldc2(spa + 64 * N),
                         ldc2(sx + 64));
#pragma unroll
            for (int k2 = 1; k2 < 4; ++k2)
              special_pair(a[k2], a[8 - k2], ldc2(spa + 16 * k2 * N), ldc2(sx + 16 * k2),
                           ldc2(spa + 16 * (8 - k2) * N), ldc2(sx + 16 * (8 - k2)));
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2)
              special_pair(bb[k2], bb[7 - k2], ldc2(spb + 16 * k2 * N),
                           ldc2(sx + 8 + 16 * k2), ldc2(spb + 16 * (7 - k2) * N),
                           ldc2(sx + 8 + 16 * (7 - k2)));
          } else {        // a[k2] <-> bb[7 - k2]
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2)
              special_pair(a[k2], bb[7 - k2], ldc2(spa + 16 * k2 * N),
                           ldc2(sx + k1a + 16 * k2), ldc2(spb + 16 * (7 - k2) * N),
                           ldc2(sx + k1b + 16 * (7 - k2)));
          }
        }
        dft8<float, true>(a);     // a[n2]: inverse over k2
        dft8<float, true>(bb);
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) {
          const unsigned cc = (unsigned)n2 * ROWB + ((n2 & 1) ? cod : cev);
          sts64(ba + cc, a[n2]);
          sts64(bq + cc, bb[n2]);
        }
        group_barrier(1 + cg, 256);
      }

      // ---- columns: inverse radix-16 of residue n2 = m
      {
        const int tid = opaque_tid();
        const int w = tid >> 5, m = w & 7, cg = w >> 3, c = cg * 32 + (tid & 31);
        const smem_addr_t cb = tile + (unsigned)m * ROWB + 8u * ((m & 1) ? (c ^ 8) : c);
        cplx<float> v[16];
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) v[k1] = lds64(cb + 8 * k1 * ROWB);
#pragma unroll
        for (int k1 = 1; k1 < 16; ++k1) v[k1] = cmul_conj(v[k1], tw128(m, k1));
        dft16<true>(v);
#pragma unroll
        for (int j = 0; j < 16; ++j) sts64(cb + 8 * j * ROWB, v[j]);
      }
      __syncthreads();

      // ---- rows: inverse + chi-square of walker b
      {
        const int tid = opaque_tid();
        const RowRole R = make_row_role(tid);
        const float unscale = (float)(P.vscale_inv[sel] / P.wscale[b]);
        double acc = fused_rows_inverse<false>(P, tile, R, tw_base + 8u * R.l, 0, unscale);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
        if ((tid & 31) == 0) {
          volatile double *red = red_s;
          red[tid >> 5] = acc;
          __threadfence_block();
          const int prev = atomicAdd(&cnt_s, 1);
          if (prev == NW - 1) {
            __threadfence_block();
            double tot = 0.0;
            for (int k = 0; k < NW; ++k) tot += red[k];
            double val = -0.5 * tot;
            if (!isfinite(val) || invalid) val = -INFINITY;
            P.lnl[b] = val;
            cnt_s = 0;
          }
        }
      }
    }
    // ---- rows: render + forward of the CTA's next walker
    const long long bn = b + gridDim.x;
    if (bn < P.n_batch) {
      const RowRole R = make_row_role(opaque_tid());
      fused_rows_forward<false>(P, tile, R, tw_base + 8u * R.l,
                                P.rconst + bn * P.ncomp * PSFMC_RC_STRIDE,
                                P.derived + bn * P.ncomp * PSFMC_DERIVED_STRIDE, 0,
                                (float)P.wscale[bn]);
    }
  }
}

// -------------------------------------------------------------- host side --

// (the TRANSFORM frame must be 128 x 128: the observation frame itself, or any smaller
// frame whose padded transform frame is -- every frame with height + psf_height - 1 <=
// 128 and width + psf_width - 1 <= 128 that is not a power of two)
template <typename T>
inline bool fused_path_available(const StagedPlan &plan, const Program &) {
  return sizeof(T) == 4 && plan.fr.H == PSFMC_FUSED_N && plan.fr.W == PSFMC_FUSED_N;
}

template <typename T>
inline int fused_prepare_device(const StagedPlan &) {
#ifndef PSFMC_EMU
  if (cudaFuncSetAttribute(fused_lnlike_kernel<false>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                           PSFMC_FUSED_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(fused_lnlike_kernel<true>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                           PSFMC_FUSED_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(fused_lnlike_kernel_wide,
                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                           PSFMC_FUSED_SMEM) != cudaSuccess)
    return 1;
#endif
  return 0;
}

// Re-layout of the float64 spectra [K][2*Wc][H] (column-major, see
// kernels_staged.cuh) for the fused kernel:
//   spec [K][ky][c]: c in 1..63 -> P[ky][kx=c]; c in 65..127 -> V[ky][kx=c-64];
//                    c = 0 -> (P[ky][0] + P[ky][64]) / 2; c = 64 -> same for V
//   specx[K][t][ky] = (S[ky][0] - S[ky][64]) / 2 for S = P (t = 0) and S = V (t = 1)
// `vscale[k]` multiplies the V channel (power of two, undone in the epilogue).
inline void fused_spectrum_layout(const cplx<double> *spec64, int n_psf,
                                  const double *vscale, cplx<float> *spec,
                                  cplx<float> *specx) {
  constexpr int N = PSFMC_FUSED_N, Wc = N / 2 + 1;
  for (int k = 0; k < n_psf; ++k) {
    const cplx<double> *src = spec64 + (size_t)k * 2 * Wc * N;
    // `split`: the factor 1/2 of the real-pair split of the row spectra, folded in
    // for the regular columns; the packed DC/Nyquist columns carry exact values
    auto at = [&](int chan, int kx, int ky, double split) {
      const cplx<double> &s = src[((size_t)chan * Wc + kx) * N + ky];
      const double f = (chan ? vscale[k] : 1.0) * split;
      cplx<float> o;
      o.x = (float)(s.x * f);
      o.y = (float)(s.y * f);
      return o;
    };
    for (int ky = 0; ky < N; ++ky) {
      for (int c = 0; c < N; ++c) {
        const int chan = c >= 64 ? 1 : 0, kx = c & 63;
        if (kx != 0) spec[((size_t)k * N + ky) * N + c] = at(chan, kx, ky, 0.5);
      }
      for (int t = 0; t < 2; ++t) {
        const cplx<double> &s0 = src[((size_t)t * Wc + 0) * N + ky];
        const cplx<double> &s64 = src[((size_t)t * Wc + 64) * N + ky];
        const double f = t ? vscale[k] : 1.0;
        cplx<float> &sum = spec[((size_t)k * N + ky) * N + 64 * t];
        cplx<float> &dif = specx[((size_t)k * 2 + t) * N + ky];
        sum.x = (float)(0.5 * f * (s0.x + s64.x));
        sum.y = (float)(0.5 * f * (s0.y + s64.y));
        dif.x = (float)(0.5 * f * (s0.x - s64.x));
        dif.y = (float)(0.5 * f * (s0.y - s64.y));
      }
    }
  }
}

struct FusedBuffers {
  float *rconst = nullptr;
  const cplx<float> *spec = nullptr, *specx = nullptr;
  const float2 *ow = nullptr;
  int n_sms = 148;
  bool wide = false;   // 1024-thread variant
  unsigned skip_quads = 0;   // see FusedParams
};

// theta -> lnL for n_batch walkers: prepare kernel + one persistent fused kernel.
// Returns the number of kernels launched.
template <typename T>
inline int launch_fused_lnlike(const StagedPlan &plan, const StagedBuffers<T> &buf,
                               const FusedBuffers &fb, const Program &prog_h,
                               const double *theta, long long n_batch, long long ld,
                               double *lnl, cudaStream_t stream,
                               cudaEvent_t ev_begin = nullptr, cudaEvent_t ev_end = nullptr) {
  if (n_batch <= 0) return 0;
  const int ncomp = prog_h.n_components;
  launch_prepare(*buf.prog_host, theta, n_batch, ld, plan.fr.Hr, plan.fr.Wr, ncomp, buf.derived,
                 buf.psf_sel, buf.wscale, fb.rconst, stream);
  FusedParams P;
  FoldParams F;
  F.Hr = plan.fr.Hr;
  F.Wr = plan.fr.Wr;
  F.fy_hi = plan.fr.fy_hi;
  F.fy_lo = plan.fr.fy_lo;
  F.fx_hi = plan.fr.fx_hi;
  F.fx_lo = plan.fr.fx_lo;
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
  P.skip_quads = fb.skip_quads;
  for (int c = 0; c < PSFMC_MAX_COMPONENTS; ++c)
    P.kind[c] = (signed char)(c < ncomp ? prog_h.kind[c] : 0);
  unsigned grid = (unsigned)(n_batch < fb.n_sms ? n_batch : fb.n_sms);
  if (ev_begin) cudaEventRecord(ev_begin, stream);
  if (fb.wide && !plan.fr.padded)
    launch_kernel(fused_lnlike_kernel_wide, dim3(grid), dim3(PSFMC_FUSED_THREADS_WIDE),
                  (size_t)PSFMC_FUSED_SMEM, stream, P);
  else if (plan.fr.padded)
    launch_kernel(fused_lnlike_kernel<true>, dim3(grid), dim3(PSFMC_FUSED_THREADS),
                  (size_t)PSFMC_FUSED_SMEM, stream, P, F);
  else
    launch_kernel(fused_lnlike_kernel<false>, dim3(grid), dim3(PSFMC_FUSED_THREADS),
                  (size_t)PSFMC_FUSED_SMEM, stream, P, F);
  if (ev_end) cudaEventRecord(ev_end, stream);
  return 2;
}

}  // namespace psfmc
