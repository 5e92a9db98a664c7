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
#define PSFMC_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
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
};

// Frame geometry shared by all kernels of the staged path.
struct Frame {
  int32_t H, W;        // rows, columns
  int32_t Wc;          // W/2 + 1 retained row frequencies
  int32_t logH, logW;
};

}  // namespace psfmc
