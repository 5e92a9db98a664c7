// Shared-memory Stockham FFT building blocks (radix 8/4/2), templated on the real
// type and the transform direction. Used by the staged row/column path.
//
// Convention: forward = exp(-2 pi i jk/L) (numpy.fft.fft), inverse = exp(+...)
// WITHOUT the 1/L factor (the engine folds 1/(H*W) into the PSF spectra).
#pragma once
#include "common.cuh"

namespace psfmc {

template <typename T>
struct Consts;
template <>
struct Consts<float> {
  static __device__ __forceinline__ float rsqrt2() { return 0.70710678118654752440f; }
};
template <>
struct Consts<double> {
  static __device__ __forceinline__ double rsqrt2() { return 0.70710678118654752440; }
};

template <typename T, bool INV>
__device__ __forceinline__ void dft2(cplx<T> &a, cplx<T> &b) {
  cplx<T> s = a + b, d = a - b;
  a = s;
  b = d;
}

// 4-point DFT, natural order in and out.
template <typename T, bool INV>
__device__ __forceinline__ void dft4(cplx<T> &c0, cplx<T> &c1, cplx<T> &c2, cplx<T> &c3) {
  cplx<T> s0 = c0 + c2, s1 = c0 - c2, s2 = c1 + c3, d = c1 - c3;
  cplx<T> s3 = INV ? mul_pos_i(d) : mul_neg_i(d);
  c0 = s0 + s2;
  c2 = s0 - s2;
  c1 = s1 + s3;
  c3 = s1 - s3;
}

// 8-point DFT, natural order in and out (decimation in frequency: one radix-2
// layer, the W8 twiddles, two 4-point DFTs).
template <typename T, bool INV>
__device__ __forceinline__ void dft8(cplx<T> *v) {
  const T h = Consts<T>::rsqrt2();
  cplx<T> a0 = v[0] + v[4], b0 = v[0] - v[4];
  cplx<T> a1 = v[1] + v[5], b1 = v[1] - v[5];
  cplx<T> a2 = v[2] + v[6], b2 = v[2] - v[6];
  cplx<T> a3 = v[3] + v[7], b3 = v[3] - v[7];
  // W8^1, W8^2, W8^3 (conjugated for the inverse)
  b1 = crot<T, INV>(b1, h, h);
  b2 = INV ? mul_pos_i(b2) : mul_neg_i(b2);
  b3 = crot<T, INV>(b3, -h, h);
  dft4<T, INV>(a0, a1, a2, a3);
  dft4<T, INV>(b0, b1, b2, b3);
  v[0] = a0; v[1] = b0; v[2] = a1; v[3] = b1;
  v[4] = a2; v[5] = b2; v[6] = a3; v[7] = b3;
}

template <typename T, bool INV>
__device__ __forceinline__ cplx<T> twiddle(const cplx<T> *tw, int idx) {
  cplx<T> w = tw[idx];
  return INV ? cconj(w) : w;
}

// Twiddle table of fft_line_smem, ordered by pass so that the lanes of a warp read
// consecutive entries (no bank conflicts): for every pass after the first, with
// sub-transform length Ns and radix R, entry [(r - 1) * Ns + k] = exp(-2 pi i r k /
// (R Ns)), r = 1..R-1, k = 0..Ns-1; passes are stored back to back (fewer than L
// entries in total). Filled on the host by fill_twiddles (pipeline.cuh).
PSFMC_HD inline int twiddle_table_entries(int logL) {
  const int n8 = logL / 3, rem = logL - 3 * n8;
  int total = 0, Ns = 1;
  for (int s = 0; s < n8; ++s) {
    if (s > 0) total += 7 * Ns;
    Ns <<= 3;
  }
  if (rem == 2) total += 3 * Ns;
  if (rem == 1) total += Ns;
  return total;
}

// Bank-conflict-free exchange layout between the passes of fft_line_smem: element i
// of a line sits at i ^ ((i >> 3) & 15). The strided stores of a Stockham pass
// (lane stride 8 elements) and its contiguous loads are then both conflict-free for
// 8- and 16-byte elements. Only the intermediate passes use it: the first pass reads
// and the last pass writes the plain layout, so callers never see the swizzle.
__device__ __forceinline__ int fft_swz(int i, bool on) {
  return on ? (i ^ ((i >> 3) & 15)) : i;
}

// In-place length-L FFT of one contiguous shared-memory line, executed by the
// L/8 threads with local indices tl = 0..L/8-1 (each owns 8 points per stage).
// ALL threads of the CTA must call this together (it contains __syncthreads).
// `tw` is the pass-ordered twiddle table described above. L is a power of two >= 16.
// LOGL != 0 fixes the length at compile time (L = 2^LOGL: all index arithmetic folds
// into immediates and the pass loop unrolls); LOGL = 0 takes L, logL at run time.
template <typename T, bool INV, int LOGL = 0>
__device__ __forceinline__ void fft_line_smem(cplx<T> *line, int L_rt, int logL_rt, int tl,
                                              const cplx<T> *tw) {
  const int logL = LOGL ? LOGL : logL_rt;
  const int L = LOGL ? (1 << LOGL) : L_rt;
  const int n8 = logL / 3, rem = logL - 3 * n8;
  const int L8 = L >> 3;
  int Ns = 1, logNs = 0;
  const cplx<T> *twp = tw;     // table of the current pass
  cplx<T> v[8];
#pragma unroll
  for (int s = 0; s < n8; ++s) {
    const bool in_swz = s > 0, out_swz = !(s == n8 - 1 && rem == 0);
    const int j = tl;
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = line[fft_swz(j + r * L8, in_swz)];
    const int k = j & (Ns - 1);
    if (Ns > 1) {
#pragma unroll
      for (int r = 1; r < 8; ++r) v[r] = v[r] * twiddle<T, INV>(twp, (r - 1) * Ns + k);
      twp += 7 * Ns;
    }
    dft8<T, INV>(v);
    __syncthreads();
    const int j0 = ((j - k) << 3) + k;
#pragma unroll
    for (int r = 0; r < 8; ++r) line[fft_swz(j0 + r * Ns, out_swz)] = v[r];
    __syncthreads();
    Ns <<= 3;
    logNs += 3;
  }
  // remainder pass (radix 4 or 2): always the last one, after at least one radix-8
  // pass: reads the swizzled layout, writes the plain one
  if (rem == 2) {
    const int L4 = L >> 2;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int j = tl + q * L8;
#pragma unroll
      for (int r = 0; r < 4; ++r) v[q * 4 + r] = line[fft_swz(j + r * L4, true)];
      const int k = j & (Ns - 1);
#pragma unroll
      for (int r = 1; r < 4; ++r)
        v[q * 4 + r] = v[q * 4 + r] * twiddle<T, INV>(twp, (r - 1) * Ns + k);
      dft4<T, INV>(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int j = tl + q * L8;
      const int k = j & (Ns - 1);
      const int j0 = ((j - k) << 2) + k;
#pragma unroll
      for (int r = 0; r < 4; ++r) line[j0 + r * Ns] = v[q * 4 + r];
    }
    __syncthreads();
  } else if (rem == 1) {
    const int L2 = L >> 1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = tl + q * L8;
      v[q * 2] = line[fft_swz(j, true)];
      v[q * 2 + 1] = line[fft_swz(j + L2, true)];
      const int k = j & (Ns - 1);
      v[q * 2 + 1] = v[q * 2 + 1] * twiddle<T, INV>(twp, k);
      dft2<T, INV>(v[q * 2], v[q * 2 + 1]);
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = tl + q * L8;
      const int k = j & (Ns - 1);
      const int j0 = ((j - k) << 1) + k;
      line[j0] = v[q * 2];
      line[j0 + Ns] = v[q * 2 + 1];
    }
    __syncthreads();
  }
}

}  // namespace psfmc
