// Launch overhead of the device-side LM loop: tiny dependent kernels (a) stream-launched, (b) inside a CUDA graph
// WHILE node with a body of B kernels, (c) the same with programmatic dependent launch edges inside the body.
#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int* counter, cudaGraphConditionalHandle h, int total, int last) {
#if __CUDA_ARCH__ >= 900
  cudaGridDependencySynchronize();
#endif
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int c = atomicAdd(counter, 1);
    if (last) cudaGraphSetConditional(h, c + 1 < total ? 1u : 0u);
  }
}
__global__ void plain(int* counter) {
  if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(counter, 1);
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
static int run_while(int B, bool pdl, int total, int grid) {
  int* d; CK(cudaMalloc(&d, 4));
  cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  cudaGraph_t g; CK(cudaGraphCreate(&g, 0));
  cudaGraphConditionalHandle h; CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h;
  p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t node; CK(cudaGraphAddNode(&node, g, nullptr, 0, &p));
  cudaGraph_t bodyg = p.conditional.phGraph_out[0];
  CK(cudaStreamBeginCaptureToGraph(s, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  for (int k = 0; k < B; ++k) {
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(32); cfg.stream = s;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (pdl && k > 0) ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, body, d, h, total, k == B - 1 ? 1 : 0));
  }
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
    if (rep == 2) printf("WHILE graph, body of %2d kernels, grid %3d, pdl %d: count %d, %.2f us total, %.2f us per kernel\n", B, grid, (int)pdl, c, ms * 1e3, ms * 1e3 / c);
  }
  return 0;
}
int main() {
  int* d; CK(cudaMalloc(&d, 4)); CK(cudaMemset(d, 0, 4));
  cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int grid : {4, 148, 296}) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0, s);
      for (int k = 0; k < 28; ++k) plain<<<grid, 32, 0, s>>>(d);
      cudaEventRecord(e1, s);
      CK(cudaStreamSynchronize(s));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2) printf("stream-launched, grid %3d: 28 kernels %.2f us total, %.2f us per kernel\n", grid, ms * 1e3, ms * 1e3 / 28);
    }
  }
  for (int grid : {4, 296}) for (int B : {2, 4, 8, 28}) for (int pdl = 0; pdl < 2; ++pdl) if (run_while(B, pdl, 28 , grid)) return 1;
  return 0;
}
