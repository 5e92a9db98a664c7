// Microbenchmark (evidence for DESIGN.md): issue rate of scalar vs packed FP32
// instructions on sm_100a, alone and interleaved with 64-bit shared-memory loads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipe_probe.cu -o pipe_probe
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 upk(u64 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int MODE>
__global__ void probe(float *out, int iters, float fa, float fb) {
  __shared__ float2 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float2(fa * i, fb);
  __syncthreads();
  float x[8];
  u64 p[8];
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3f + i; p[i] = pk(x[i], x[i] + 0.5f); }
  const u64 pa = pk(fa, fa), pb = pk(fb, fb);
  int idx = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {        // scalar FFMA
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], fa, fb);
    } else if (MODE == 1) { // packed FFMA2
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], pa, pb);
    } else if (MODE == 2) { // scalar FADD
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = x[i] + fb;
    } else if (MODE == 3) { // packed FADD2
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = add2(p[i], pb);
    } else if (MODE == 4) { // 8 FADD2 + 2 LDS.64 (butterfly-like mix)
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = add2(p[i], pb);
      float2 v0 = sm[idx & 1023], v1 = sm[(idx + 32) & 1023];
      p[0] = add2(p[0], pk(v0.x, v0.y));
      p[1] = add2(p[1], pk(v1.x, v1.y));
      idx += 64;
    } else if (MODE == 5) { // 16 scalar FADD + 2 LDS.64
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = x[i] + fb;
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = x[i] + fa;
      float2 v0 = sm[idx & 1023], v1 = sm[(idx + 32) & 1023];
      x[0] += v0.x; x[1] += v0.y; x[2] += v1.x; x[3] += v1.y;
      idx += 64;
    } else if (MODE == 6) { // MUFU ex2 alone
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    } else if (MODE == 7) { // 4 FFMA2 + 4 scalar FFMA interleaved
#pragma unroll
      for (int i = 0; i < 4; ++i) { p[i] = fma2(p[i], pa, pb); x[i] = fmaf(x[i], fa, fb); }
    }
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) { float2 q = upk(p[i]); s += x[i] + q.x + q.y; }
  if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, double flop_per_iter, double inst_per_iter) {
  int dev = 0, sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int block = 512, grid = sms * 4, iters = 1 << 14;
  float *out;
  cudaMalloc(&out, sizeof(float) * grid * block);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    probe<MODE><<<grid, block>>>(out, iters, 1.0000001f, 1e-7f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  double threads = (double)grid * block;
  double tflops = flop_per_iter * iters * threads / (best * 1e-3) / 1e12;
  double warp_inst_per_clk_sm = inst_per_iter * iters * threads / 32.0 / sms / (best * 1e-3 * khz * 1e3);
  printf("%-32s %8.3f ms  %7.2f TFLOP/s  %5.2f warp-inst/clk/SM (at %d MHz nominal)\n", name, best,
         tflops, warp_inst_per_clk_sm, khz / 1000);
  cudaFree(out);
}

int main() {
  run<0>("FFMA scalar x8", 16, 8);
  run<1>("FFMA2 packed x8", 32, 8);
  run<2>("FADD scalar x8", 8, 8);
  run<3>("FADD2 packed x8", 16, 8);
  run<4>("10 FADD2 + 2 LDS.64", 20, 12);
  run<5>("20 FADD + 2 LDS.64", 20, 22);
  run<6>("MUFU.EX2 x8", 8, 8);
  run<7>("4 FFMA2 + 4 FFMA", 24, 8);
  return 0;
}
