#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int* counter, cudaGraphConditionalHandle h) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int c = atomicAdd(counter, 1);
    cudaGraphSetConditional(h, c + 1 < 10 ? 1u : 0u);
  }
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main() {
  int* d; CK(cudaMalloc(&d, 4)); CK(cudaMemset(d, 0, 4));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  cudaGraph_t g; CK(cudaGraphCreate(&g, 0));
  cudaGraphConditionalHandle h; CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h;
  p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t node; CK(cudaGraphAddNode(&node, g, nullptr, 0, &p));
  cudaGraph_t bodyg = p.conditional.phGraph_out[0];
  CK(cudaStreamBeginCaptureToGraph(s, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  body<<<4, 32, 0, s>>>(d, h);
  body<<<4, 32, 0, s>>>(d, h);
  cudaGraph_t out; CK(cudaStreamEndCapture(s, &out));
  cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, g, 0));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemsetAsync(d, 0, 4, s));
    cudaEventRecord(e0, s);
    CK(cudaGraphLaunch(ex, s));
    cudaEventRecord(e1, s);
    CK(cudaStreamSynchronize(s));
    int c; CK(cudaMemcpy(&c, d, 4, cudaMemcpyDeviceToHost));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("count %d, %.2f us total, %.2f us per kernel\n", c, ms * 1e3, ms * 1e3 / c);
  }
  return 0;
}
