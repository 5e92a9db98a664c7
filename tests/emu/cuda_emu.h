// TEST INFRASTRUCTURE ONLY -- a tiny SIMT emulator so that the *same* kernel
// sources under psfmc_b200/csrc/ can be compiled with g++ and exercised on the CPU
// (this container has no GPU). It is never part of the product: libpsfmc_b200.so
// is built by nvcc without this header, and nothing in psfmc_b200/ includes it.
//
// Model: one CTA (or one thread-block cluster) at a time; every CUDA thread is a ucontext fiber. __syncthreads
// and the warp shuffles are cooperative yield points; the scheduler runs fibers
// round-robin and aborts on deadlock (a barrier not reached by all live threads),
// which catches divergent-barrier bugs. No attempt is made to model timing.
#pragma once
#include <sys/mman.h>
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
#define __cluster_dims__(...)
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

// Fiber stacks: mmap-ed and committed lazily (a 4-CTA cluster of 512-thread CTAs is
// 2048 fibers).
constexpr size_t kStackBytes = 256 * 1024;

struct Fiber {
  ucontext_t ctx;
  unsigned char *stack = nullptr;
  uint3 tid;
  int linear;
  bool done = false;
  // barrier bookkeeping
  int wait_kind = 0;        // 0 none, 1 block barrier, 2 warp exchange, 3 named, 4 cluster
  unsigned long long wait_gen = 0;
  unsigned long long cluster_gen = 0;   // generation seen at barrier.cluster.arrive
};

// One CTA. A launch runs one CLUSTER of CTAs at a time (cluster size 1 = the classic
// one-CTA-at-a-time model); all fibers of the cluster are scheduled round-robin.
struct State {
  std::vector<Fiber> fibers;
  Fiber *cur = nullptr;
  dim3 grid, block;
  uint3 bid;
  unsigned cta_rank = 0;
  std::vector<unsigned char> smem;
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
};

struct Global {
  ucontext_t sched;
  State *cur = nullptr;                 // CTA of the running fiber
  std::vector<State *> cluster;         // CTAs of the running cluster (by rank)
  std::function<void()> body;
  unsigned long long progress = 0;
  // barrier.cluster
  int cl_count = 0;
  unsigned long long cl_gen = 0;
};

inline Global &G() {
  static Global g;
  return g;
}

inline State &S() { return *G().cur; }

inline void yield_to_sched() { swapcontext(&S().cur->ctx, &G().sched); }

inline void fiber_entry() {
  G().body();
  State &st = S();
  st.cur->done = true;
  G().progress++;
  st.live--;
  st.warps[st.cur->linear / 32].live--;
  swapcontext(&st.cur->ctx, &G().sched);
}

inline void syncthreads() {
  State &st = S();
  unsigned long long gen = st.bar_gen;
  st.bar_count++;
  if (st.bar_count >= st.live) {  // last arrival releases everybody
    st.bar_count = 0;
    st.bar_gen++;
    G().progress++;
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
    G().progress++;
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
  if (id < 0 || id >= (int)st.named.size()) {
    std::fprintf(stderr, "cuda_emu: named barrier id %d out of range\n", id);
    std::abort();
  }
  State::Named &nb = st.named[id];
  unsigned long long gen = nb.gen;
  nb.count++;
  if (nb.count >= nthreads) {
    nb.count = 0;
    nb.gen++;
    G().progress++;
    return;
  }
  st.cur->wait_kind = 3;
  while (nb.gen == gen) yield_to_sched();
  st.cur->wait_kind = 0;
}

// barrier.cluster.arrive / .wait over all live threads of the cluster (split phase:
// work between the two overlaps the other CTAs' arrival)
inline void cluster_arrive() {
  Global &g = G();
  S().cur->cluster_gen = g.cl_gen;
  int live = 0;
  for (State *c : g.cluster) live += c->live;
  g.cl_count++;
  if (g.cl_count >= live) {
    g.cl_count = 0;
    g.cl_gen++;
    g.progress++;
  }
}
inline void cluster_wait() {
  State &st = S();
  st.cur->wait_kind = 4;
  while (G().cl_gen == st.cur->cluster_gen) yield_to_sched();
  st.cur->wait_kind = 0;
}
inline unsigned cluster_ctarank() { return S().cta_rank; }
inline unsigned cluster_nctarank() { return (unsigned)G().cluster.size(); }

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

inline unsigned char *dyn_smem_of(State &st) {
  uintptr_t p = (uintptr_t)st.smem.data();
  p = (p + 1023) & ~(uintptr_t)1023;   // PSFMC_DYN_SMEM asks for 1024-byte alignment
  return (unsigned char *)p;
}
inline unsigned char *dyn_smem() { return dyn_smem_of(S()); }

// mapa: the address of the same dynamic-shared-memory location in CTA `rank` of the
// cluster (static __shared__ variables are plain statics here, shared by all CTAs:
// cluster kernels must keep everything in dynamic shared memory)
inline uintptr_t map_shared_rank(uintptr_t addr, unsigned rank) {
  Global &g = G();
  return addr - (uintptr_t)dyn_smem_of(S()) + (uintptr_t)dyn_smem_of(*g.cluster[rank]);
}

inline State *pooled_state(size_t k) {
  static std::vector<State *> pool;
  while (pool.size() <= k) pool.push_back(new State());
  return pool[k];
}

template <typename F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F &&body, unsigned cluster_x = 1) {
  Global &g = G();
  g.body = body;
  const int nthreads = block.x * block.y * block.z;
  if (cluster_x < 1 || grid.x % cluster_x) {
    std::fprintf(stderr, "cuda_emu: grid.x is not a multiple of the cluster size\n");
    std::abort();
  }
  g.cluster.clear();
  for (unsigned k = 0; k < cluster_x; ++k) g.cluster.push_back(pooled_state(k));
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx0 = 0; bx0 < grid.x; bx0 += cluster_x) {
        g.cl_count = 0;
        for (unsigned k = 0; k < cluster_x; ++k) {
          State &st = *g.cluster[k];
          st.grid = grid;
          st.block = block;
          st.bid = uint3{bx0 + k, by, bz};
          st.cta_rank = k;
          st.live = nthreads;
          st.bar_count = 0;
          // sixteen hardware barriers, sized once: fibers waiting in named_barrier() hold a
          // reference into this vector, which must therefore never reallocate
          st.named.assign(16, State::Named());
          st.warps.assign((nthreads + 31) / 32, State::Warp());
          // poison shared memory so that reads of unwritten smem show up as NaNs
          st.smem.assign(smem_bytes + 2048, 0xFF);
          if ((int)st.fibers.size() < nthreads) st.fibers.resize(nthreads);
          for (int t = 0; t < nthreads; ++t) {
            Fiber &f = st.fibers[t];
            if (!f.stack) {
              void *mem = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE,
                               MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
              if (mem == MAP_FAILED) {
                std::fprintf(stderr, "cuda_emu: cannot map a fiber stack\n");
                std::abort();
              }
              f.stack = (unsigned char *)mem;
            }
            f.done = false;
            f.wait_kind = 0;
            f.linear = t;
            f.tid = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y),
                          (unsigned)(t / (block.x * block.y))};
            st.warps[t / 32].live++;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = kStackBytes;
            f.ctx.uc_link = &g.sched;
            makecontext(&f.ctx, (void (*)())fiber_entry, 0);
          }
        }
        int idle_rounds = 0;
        for (;;) {
          int live = 0;
          for (State *c : g.cluster) live += c->live;
          if (live <= 0) break;
          unsigned long long before = g.progress;
          for (State *c : g.cluster) {
            for (int t = 0; t < nthreads; ++t) {
              Fiber &f = c->fibers[t];
              if (f.done) continue;
              g.cur = c;
              c->cur = &f;
              swapcontext(&g.sched, &f.ctx);
            }
          }
          bool progress = g.progress != before;
          idle_rounds = progress ? 0 : idle_rounds + 1;
          if (idle_rounds > 4) {
            int kinds[5] = {0, 0, 0, 0, 0};
            for (State *c : g.cluster)
              for (int t = 0; t < nthreads; ++t)
                if (!c->fibers[t].done) kinds[c->fibers[t].wait_kind]++;
            std::fprintf(stderr,
                         "cuda_emu: DEADLOCK in block (%u,%u,%u) (cluster of %u): %d threads "
                         "alive; waiting at __syncthreads %d, warp %d, named %d, cluster %d "
                         "-- divergent barrier or shuffle\n",
                         bx0, by, bz, cluster_x, live, kinds[1], kinds[2], kinds[3], kinds[4]);
            std::abort();
          }
        }
      }
  g.cur = g.cluster[0];
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
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
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
