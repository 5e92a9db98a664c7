// Common device-side definitions for the psfmc_b200 kernels (sm_100a).
//
// The kernel sources also compile as plain C++ under -DPSFMC_EMU against
// tests/emu/cuda_emu.h, a fiber-based SIMT emulator used ONLY by the CPU test
// tier (this repository is developed on a box without a GPU). The product
// library is always built by nvcc; nothing here falls back to the CPU.
#pragma once

#ifdef PSFMC_EMU
#include "cuda_emu.h"
#define PSFMC_DYN_SMEM(name) unsigned char *name = emu::dyn_smem()
#define PSFMC_HD
#else
#include <cuda_runtime.h>
#define PSFMC_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#define PSFMC_HD __host__ __device__
#endif

#include <stdint.h>

#include "../../include/psfmc_b200.h"

namespace psfmc {

// ---------------------------------------------------------------- complex --
template <typename T>
struct cplx {
  T x, y;
};

template <typename T>
__device__ __forceinline__ cplx<T> mk(T x, T y) {
  cplx<T> c;
  c.x = x;
  c.y = y;
  return c;
}
template <typename T>
__device__ __forceinline__ cplx<T> operator+(cplx<T> a, cplx<T> b) {
  return mk<T>(a.x + b.x, a.y + b.y);
}
template <typename T>
__device__ __forceinline__ cplx<T> operator-(cplx<T> a, cplx<T> b) {
  return mk<T>(a.x - b.x, a.y - b.y);
}
template <typename T>
__device__ __forceinline__ cplx<T> operator*(cplx<T> a, cplx<T> b) {
  return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <typename T>
__device__ __forceinline__ cplx<T> cconj(cplx<T> a) {
  return mk<T>(a.x, -a.y);
}
// multiply by -i (forward) / +i (inverse)
template <typename T>
__device__ __forceinline__ cplx<T> mul_neg_i(cplx<T> a) {
  return mk<T>(a.y, -a.x);
}
template <typename T>
__device__ __forceinline__ cplx<T> mul_pos_i(cplx<T> a) {
  return mk<T>(-a.y, a.x);
}

// Generic conjugate multiply a * conj(w) and constant rotations a * (c -+ i s);
// float32 has packed implementations below.
template <typename T>
__device__ __forceinline__ cplx<T> cmul_conj(cplx<T> a, cplx<T> w) {
  return mk<T>(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y);
}
template <typename T, bool INV>
__device__ __forceinline__ cplx<T> crot(cplx<T> a, T c, T s) {   // a * (c - i s), inverse: (c + i s)
  return INV ? mk<T>(a.x * c - a.y * s, a.y * c + a.x * s)
             : mk<T>(a.x * c + a.y * s, a.y * c - a.x * s);
}
template <typename T>
__device__ __forceinline__ cplx<T> cscale(cplx<T> a, T f) {
  return mk<T>(a.x * f, a.y * f);
}

// Element-wise pair arithmetic (a cplx<T> used as a plain pair of values).
template <typename T>
__device__ __forceinline__ cplx<T> pmul(cplx<T> a, cplx<T> b) {
  return mk<T>(a.x * b.x, a.y * b.y);
}
template <typename T>
__device__ __forceinline__ cplx<T> pfma(cplx<T> a, cplx<T> b, cplx<T> c) {
  return mk<T>(a.x * b.x + c.x, a.y * b.y + c.y);
}
template <typename T>
__device__ __forceinline__ cplx<T> bcast(T a) {
  return mk<T>(a, a);
}

#ifndef PSFMC_EMU
// ---- packed float32 arithmetic (sm_100a FADD2 / FMUL2 / FFMA2) ----------------
// A complex64 value is one 64-bit register pair, so complex add/sub is ONE packed
// instruction and a complex multiply is two (ptxas folds the pack/unpack moves,
// swaps and per-half sign flips into operand modifiers). The packed forms issue at
// half the rate of the scalar ones (same FLOP/s) but take half the issue slots,
// which is what this issue-bound path needs (profiles/microbench/pipe_probe.cu).
typedef unsigned long long u64_t;
__device__ __forceinline__ u64_t pk2(float a, float b) {
  u64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ cplx<float> upk2(u64_t v) {
  cplx<float> r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ u64_t add2(u64_t a, u64_t b) {
  u64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64_t mul2(u64_t a, u64_t b) {
  u64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64_t fma2(u64_t a, u64_t b, u64_t c) {
  u64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ cplx<float> operator+(cplx<float> a, cplx<float> b) {
  return upk2(add2(pk2(a.x, a.y), pk2(b.x, b.y)));
}
__device__ __forceinline__ cplx<float> operator-(cplx<float> a, cplx<float> b) {
  return upk2(add2(pk2(a.x, a.y), pk2(-b.x, -b.y)));
}
__device__ __forceinline__ cplx<float> operator*(cplx<float> a, cplx<float> w) {
  const cplx<float> r = upk2(mul2(pk2(a.y, a.y), pk2(w.y, w.x)));   // (a.y w.y, a.y w.x)
  return upk2(fma2(pk2(a.x, a.x), pk2(w.x, w.y), pk2(-r.x, r.y)));
}
template <>
__device__ __forceinline__ cplx<float> cmul_conj<float>(cplx<float> a, cplx<float> w) {
  const cplx<float> r = upk2(mul2(pk2(a.x, a.x), pk2(w.x, w.y)));   // (a.x w.x, a.x w.y)
  return upk2(fma2(pk2(a.y, a.y), pk2(w.y, w.x), pk2(r.x, -r.y)));
}
template <>
__device__ __forceinline__ cplx<float> crot<float, false>(cplx<float> a, float c, float s) {
  const cplx<float> r = upk2(mul2(pk2(a.x, a.x), pk2(c, s)));       // (a.x c, a.x s)
  return upk2(fma2(pk2(a.y, a.y), pk2(s, c), pk2(r.x, -r.y)));
}
template <>
__device__ __forceinline__ cplx<float> crot<float, true>(cplx<float> a, float c, float s) {
  const cplx<float> r = upk2(mul2(pk2(a.y, a.y), pk2(s, c)));       // (a.y s, a.y c)
  return upk2(fma2(pk2(a.x, a.x), pk2(c, s), pk2(-r.x, r.y)));
}
template <>
__device__ __forceinline__ cplx<float> cscale<float>(cplx<float> a, float f) {
  return upk2(mul2(pk2(a.x, a.y), pk2(f, f)));
}
template <>
__device__ __forceinline__ cplx<float> pmul<float>(cplx<float> a, cplx<float> b) {
  return upk2(mul2(pk2(a.x, a.y), pk2(b.x, b.y)));
}
template <>
__device__ __forceinline__ cplx<float> pfma<float>(cplx<float> a, cplx<float> b,
                                                   cplx<float> c) {
  return upk2(fma2(pk2(a.x, a.y), pk2(b.x, b.y), pk2(c.x, c.y)));
}
#endif

// --------------------------------------------------------- derived params --
// Per-walker, per-component constants produced by the prepare kernel (always in
// float64) and consumed by the render code. Stride PSFMC_DERIVED_STRIDE doubles.
#define PSFMC_DERIVED_STRIDE 24
// Sersic (psfMC/ModelComponents/Sersic.py:73-134)
#define D_SER_X0 0
#define D_SER_Y0 1
#define D_SER_A00 2   //  cos/reff
#define D_SER_A01 3   //  sin/reff
#define D_SER_A10 4   // -sin/reff_b
#define D_SER_A11 5   //  cos/reff_b
#define D_SER_P 6     // radius_pow = 0.5/n
#define D_SER_KAPPA 7
#define D_SER_SBEFF 8
// PointSource (psfMC/ModelComponents/PointSource.py:24-81)
#define D_PS_X 0
#define D_PS_Y 1
#define D_PS_FLUX 2
#define D_PS_YMIN 3   // inclusive stamp bounds, integers stored as doubles
#define D_PS_YMAX 4
#define D_PS_XMIN 5
#define D_PS_XMAX 6
#define D_PS_WX 7     // 7 separable stamp weights along x (x = XMIN + i), then
#define D_PS_WY 14    // 7 along y (y = YMIN + i); unused taps are 0
// Sky
#define D_SKY_ADU 0

// Float32 render constants of the fused path, PSFMC_RC_STRIDE floats per
// (walker, component), written by the prepare kernel next to the float64 ones.
#define PSFMC_RC_STRIDE 12

// Compact, device-resident copy of the component program.
struct Program {
  int32_t n_components;
  int32_t kind[PSFMC_MAX_COMPONENTS];
  int32_t flags[PSFMC_MAX_COMPONENTS];
  int32_t theta_index[PSFMC_MAX_COMPONENTS][PSFMC_NSLOTS];
  double value[PSFMC_MAX_COMPONENTS][PSFMC_NSLOTS];
  int32_t psf_theta_index;
  double psf_value;
  int32_t n_psf;
  double mag_zp;
  // Piecewise Chebyshev table of log(kappa) over u = log2(2n) (devmath.cuh), built
  // at engine creation from the device's own Halley iteration; null = not available
  const double *kappa_coef;   // [kappa_nint][PSFMC_KAPPA_DEG]
  int32_t kappa_nint;
  double kappa_u0, kappa_inv_du;
};

#define PSFMC_KAPPA_DEG 16      // Chebyshev coefficients per interval
#define PSFMC_KAPPA_U0 (-3.5)   // log2(a) range of the table: a = 2n in [0.088, 64]
#define PSFMC_KAPPA_U1 6.0
#define PSFMC_KAPPA_NINT 19     // intervals of width 0.5 in log2(a)

// Frame geometry shared by all kernels of the staged path.
struct Frame {
  int32_t H, W;        // rows, columns of the TRANSFORM frame (powers of two)
  int32_t Wc;          // W/2 + 1 retained row frequencies
  int32_t logH, logW;
  // Observation frame Hr x Wr. padded = 0: identical to the transform frame. padded = 1
  // (frames that are not powers of two): the observation sits in the corner of a larger
  // transform frame (H >= Hr + psf_h - 1, W >= Wr + psf_w - 1), rendered pixels outside it
  // are zero, so the transform yields the LINEAR convolution, which is then folded back
  // modulo (Hr, Wr) into the reference's circular one (psfMC/utils.py:25-32): output p of
  // an axis adds c[p + N] for p <= f?_hi and c[M + p - N] for p >= f?_lo (N = real,
  // M = transform length; the bounds follow from the PSF extent and its origin).
  int32_t Hr, Wr;
  int32_t padded;
  int32_t fy_hi, fy_lo, fx_hi, fx_lo;
};

}  // namespace psfmc
