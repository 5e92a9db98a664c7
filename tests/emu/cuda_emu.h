// TEST INFRASTRUCTURE ONLY -- a tiny SIMT emulator so that the *same* kernel
// sources under psfmc_b200/csrc/ can be compiled with g++ and exercised on the CPU
// (this container has no GPU). It is never part of the product: libpsfmc_b200.so
// is built by nvcc without this header, and nothing in psfmc_b200/ includes it.
//
// Model: one CTA at a time; every CUDA thread is a ucontext fiber. __syncthreads
// and the warp shuffles are cooperative yield points; the scheduler runs fibers
// round-robin and aborts on deadlock (a barrier not reached by all live threads),
// which catches divergent-barrier bugs. No attempt is made to model timing.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __constant__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct double2 { double x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
typedef void *cudaStream_t;

namespace emu {

struct Fiber {
  ucontext_t ctx;
  std::vector<unsigned char> stack;
  uint3 tid;
  int linear;
  bool done = false;
  // barrier bookkeeping
  int wait_kind = 0;        // 0 none, 1 block barrier, 2 warp exchange
  unsigned long long wait_gen = 0;
};

struct State {
  ucontext_t sched;
  std::vector<Fiber> fibers;
  Fiber *cur = nullptr;
  dim3 grid, block;
  uint3 bid;
  std::vector<unsigned char> smem;
  std::function<void()> body;
  // block barrier
  int bar_count = 0;
  unsigned long long bar_gen = 0;
  // per-warp exchange slots
  struct Warp {
    unsigned char slot[32][16];
    int arrived = 0;
    int departed = 0;
    unsigned long long gen = 0;
    unsigned long long gen2 = 0;
    int sync_count = 0;
    unsigned long long sync_gen = 0;
    int live = 0;
  };
  std::vector<Warp> warps;
  struct Named {
    int count = 0;
    unsigned long long gen = 0;
  };
  std::vector<Named> named;
  int live = 0;
  unsigned long long progress = 0;
};

inline State &S() {
  static State st;
  return st;
}

inline void yield_to_sched() { swapcontext(&S().cur->ctx, &S().sched); }

inline void fiber_entry() {
  State &st = S();
  st.body();
  st.cur->done = true;
  st.progress++;
  st.live--;
  st.warps[st.cur->linear / 32].live--;
  swapcontext(&st.cur->ctx, &st.sched);
}

inline void syncthreads() {
  State &st = S();
  unsigned long long gen = st.bar_gen;
  st.bar_count++;
  if (st.bar_count >= st.live) {  // last arrival releases everybody
    st.bar_count = 0;
    st.bar_gen++;
    st.progress++;
    return;
  }
  st.cur->wait_kind = 1;
  while (st.bar_gen == gen) yield_to_sched();
  st.cur->wait_kind = 0;
}

inline void warp_barrier(State::Warp &w, int &count, unsigned long long &gen) {
  State &st = S();
  unsigned long long g = gen;
  count++;
  if (count >= w.live) {
    count = 0;
    gen++;
    st.progress++;
    return;
  }
  st.cur->wait_kind = 2;
  while (gen == g) yield_to_sched();
  st.cur->wait_kind = 0;
}

// __syncwarp: fibers of a warp do not run in lockstep, so this is a real barrier
inline void syncwarp() {
  State &st = S();
  State::Warp &w = st.warps[st.cur->linear / 32];
  warp_barrier(w, w.sync_count, w.sync_gen);
}

// bar.sync id, nthreads: named barrier over `nthreads` threads of the CTA
inline void named_barrier(int id, int nthreads) {
  State &st = S();
  if ((int)st.named.size() <= id) st.named.resize(id + 1);
  State::Named &nb = st.named[id];
  unsigned long long gen = nb.gen;
  nb.count++;
  if (nb.count >= nthreads) {
    nb.count = 0;
    nb.gen++;
    st.progress++;
    return;
  }
  st.cur->wait_kind = 3;
  while (nb.gen == gen) yield_to_sched();
  st.cur->wait_kind = 0;
}

// all live lanes of the warp publish `bytes` of data, then read lane `src`
// (two warp-wide barriers per exchange: publish, then consume).
inline void warp_exchange(const void *mine, void *out, int bytes, int src_lane) {
  State &st = S();
  Fiber *f = st.cur;
  State::Warp &w = st.warps[f->linear / 32];
  int lane = f->linear % 32;
  std::memcpy(w.slot[lane], mine, bytes);
  warp_barrier(w, w.arrived, w.gen);
  std::memcpy(out, w.slot[src_lane & 31], bytes);
  warp_barrier(w, w.departed, w.gen2);
}

template <typename F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F &&body) {
  State &st = S();
  st.grid = grid;
  st.block = block;
  st.body = body;
  int nthreads = block.x * block.y * block.z;
  st.smem.assign(smem_bytes + 2048, 0);
  if ((int)st.fibers.size() < nthreads) st.fibers.resize(nthreads);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        st.bid = uint3{bx, by, bz};
        st.live = nthreads;
        st.bar_count = 0;
        st.named.clear();
        st.warps.assign((nthreads + 31) / 32, State::Warp());
        // poison shared memory so that reads of unwritten smem show up as NaNs
        std::memset(st.smem.data(), 0xFF, st.smem.size());
        for (int t = 0; t < nthreads; ++t) {
          Fiber &f = st.fibers[t];
          if (f.stack.empty()) f.stack.resize(256 * 1024);
          f.done = false;
          f.wait_kind = 0;
          f.linear = t;
          f.tid = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y),
                        (unsigned)(t / (block.x * block.y))};
          st.warps[t / 32].live++;
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack.data();
          f.ctx.uc_stack.ss_size = f.stack.size();
          f.ctx.uc_link = &st.sched;
          makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
        int idle_rounds = 0;
        while (st.live > 0) {
          unsigned long long before = st.progress;
          for (int t = 0; t < nthreads; ++t) {
            Fiber &f = st.fibers[t];
            if (f.done) continue;
            st.cur = &f;
            swapcontext(&st.sched, &f.ctx);
          }
          bool progress = st.progress != before;
          idle_rounds = progress ? 0 : idle_rounds + 1;
          if (idle_rounds > 4) {
            std::fprintf(stderr,
                         "cuda_emu: DEADLOCK in block (%u,%u,%u): %d threads alive, "
                         "%d at __syncthreads -- divergent barrier or shuffle\n",
                         bx, by, bz, st.live, st.bar_count);
            std::abort();
          }
        }
      }
}

inline unsigned char *dyn_smem() {
  uintptr_t p = (uintptr_t)S().smem.data();
  p = (p + 1023) & ~(uintptr_t)1023;   // PSFMC_DYN_SMEM asks for 1024-byte alignment
  return (unsigned char *)p;
}

}  // namespace emu

#define threadIdx (emu::S().cur->tid)
#define blockIdx (emu::S().bid)
#define blockDim (emu::S().block)
#define gridDim (emu::S().grid)
#define __syncthreads() emu::syncthreads()
#define __syncwarp(...) emu::syncwarp()

template <typename T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  T out;
  int lane = emu::S().cur->linear % 32;
  int base = lane & ~(width - 1);
  emu::warp_exchange(&v, &out, sizeof(T), base + (src & (width - 1)));
  return out;
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int mask, int width = 32) {
  T out;
  int lane = emu::S().cur->linear % 32;
  (void)width;
  emu::warp_exchange(&v, &out, sizeof(T), lane ^ mask);
  return out;
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, int delta, int width = 32) {
  T out;
  int lane = emu::S().cur->linear % 32;
  int src = lane + delta;
  if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
  emu::warp_exchange(&v, &out, sizeof(T), src);
  return out;
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, int delta, int width = 32) {
  T out;
  int lane = emu::S().cur->linear % 32;
  int src = lane - delta;
  if (src < (lane & ~(width - 1))) src = lane;
  emu::warp_exchange(&v, &out, sizeof(T), src);
  return out;
}

inline int __all_sync(unsigned, int pred) {
  int all = 1;
  for (int src = 0; src < 32; ++src) {   // one exchange per source lane keeps it simple
    int got = 0;
    int mine = pred ? 1 : 0;
    emu::warp_exchange(&mine, &got, sizeof(int), src);
    int lane_alive = src < emu::S().warps[emu::S().cur->linear / 32].live ? 1 : 1;
    (void)lane_alive;
    all &= got;
  }
  return all;
}
inline int __any_sync(unsigned mask, int pred) { return !__all_sync(mask, !pred); }
template <typename T>
inline T __ldg(const T *p) { return *p; }
inline double atomicAdd(double *p, double v) { double o = *p; *p += v; return o; }
inline float atomicAdd(float *p, float v) { float o = *p; *p += v; return o; }
inline int atomicAdd(int *p, int v) { int o = *p; *p += v; return o; }
inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p += v; return o; }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline void __threadfence() {}
inline void __threadfence_block() {}

inline void sincospi(double x, double *s, double *c) {
  // exact at multiples of 1/2 like the CUDA function
  double r = std::fmod(x, 2.0);
  *s = std::sin(M_PI * r);
  *c = std::cos(M_PI * r);
  double q = r * 2.0;
  if (q == std::floor(q)) {
    int k = ((int)q % 4 + 4) % 4;
    const double sv[4] = {0, 1, 0, -1}, cv[4] = {1, 0, -1, 0};
    *s = sv[k];
    *c = cv[k];
  }
}
inline void sincospif(float x, float *s, float *c) {
  double sd, cd;
  sincospi((double)x, &sd, &cd);
  *s = (float)sd;
  *c = (float)cd;
}
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline float __expf(float x) { return std::exp(x); }
inline float __logf(float x) { return std::log(x); }
inline float __log2f(float x) { return std::log2(x); }
inline float __fdividef(float a, float b) { return a / b; }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
using std::isfinite;
using std::isnan;

// ----------------------------------------------------------------------------
// CUDA runtime shim: just enough of the host API for psfmc_b200/csrc/engine.cu to
// run unchanged on the CPU ("device" memory is host memory, streams are no-ops).
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; };
enum cudaDeviceAttr { cudaDevAttrComputeCapabilityMajor = 75, cudaDevAttrComputeCapabilityMinor = 76, cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
typedef void *cudaEvent_t;
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
template <typename T> inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)std::calloc(n ? n : 1, 1); return *p ? 0 : 2; }
template <typename T> inline cudaError_t cudaMallocHost(T **p, size_t n) { *p = (T *)std::calloc(n ? n : 1, 1); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void *p) { std::free(p); return 0; }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return 0; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return 0; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int) {
  *v = (a == cudaDevAttrComputeCapabilityMajor) ? 10 : (a == cudaDevAttrMultiProcessorCount ? 148 : 0);
  return 0;
}
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) { a->type = cudaMemoryTypeUnregistered; a->device = 0; return 0; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 1.0f; return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
