// Probe (not part of the product): do CUDA graph conditional IF nodes work on this
// driver / toolkit without relocatable device code? A kernel arms the node with
// cudaGraphSetConditional; the body graph is captured with cudaStreamBeginCaptureToGraph.
// Expected output: "out" advances only on the launches whose flag is 1.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o probe condgraph_probe.cu && ./probe
#include <cuda_runtime.h>
#include <cstdio>
__global__ void setk(cudaGraphConditionalHandle h, const int *flag) {
  if (threadIdx.x == 0) cudaGraphSetConditional(h, *flag != 0);
}
__global__ void body(int *out) { *out += 1; }
int main() {
  int *flag, *out;
  cudaMalloc(&flag, 4); cudaMalloc(&out, 4);
  cudaMemset(out, 0, 4);
  cudaStream_t s, s2; cudaStreamCreate(&s); cudaStreamCreate(&s2);
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  cudaStreamCaptureStatus st; cudaGraph_t g; const cudaGraphNode_t *deps; size_t nd;
  cudaStreamGetCaptureInfo_v2(s, &st, nullptr, &g, &deps, &nd);
  cudaGraphConditionalHandle h;
  cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault);
  setk<<<1, 32, 0, s>>>(h, flag);
  cudaStreamGetCaptureInfo_v2(s, &st, nullptr, &g, &deps, &nd);
  cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional;
  p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeIf; p.conditional.size = 1;
  cudaGraphNode_t cn;
  printf("add %d\n", (int)cudaGraphAddNode(&cn, g, deps, nd, &p));
  cudaGraph_t bg = p.conditional.phGraph_out[0];
  cudaStreamBeginCaptureToGraph(s2, bg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
  body<<<1, 1, 0, s2>>>(out);
  cudaStreamEndCapture(s2, nullptr);
  cudaStreamUpdateCaptureDependencies(s, &cn, 1, cudaStreamSetCaptureDependencies);
  cudaGraph_t gg; printf("end %d\n", (int)cudaStreamEndCapture(s, &gg));
  cudaGraphExec_t ex; printf("inst %d\n", (int)cudaGraphInstantiate(&ex, gg, 0));
  for (int v = 0; v < 4; ++v) {
    int f = v & 1; cudaMemcpy(flag, &f, 4, cudaMemcpyHostToDevice);
    cudaGraphLaunch(ex, s); cudaStreamSynchronize(s);
    int o; cudaMemcpy(&o, out, 4, cudaMemcpyDeviceToHost); printf("flag %d out %d\n", f, o);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
