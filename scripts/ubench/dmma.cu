// fp64 tensor-core (mma.sync.m8n8k4.f64, SASS DMMA) latency and throughput next to DFMA, one SM and full chip.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CHAINS, bool TENSOR> __global__ void k(double* out, long long* cyc, int n) {
  double a = out[threadIdx.x & 31], b = out[32 + (threadIdx.x & 31)];
  double c[CHAINS][2];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { c[i][0] = i; c[i][1] = -i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (TENSOR) dmma(c[i][0], c[i][1], a, b);
      else { c[i][0] = fma(a, b, c[i][0]); c[i][1] = fma(a, b, c[i][1]); }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1];
  out[64 + blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int CHAINS, bool TENSOR> void run(const char* name, int blocks, int threads, double* d, long long* c) {
  const int n = 2048;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<CHAINS, TENSOR><<<blocks, threads>>>(d, c, n); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<CHAINS, TENSOR><<<blocks, threads>>>(d, c, n); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
  const double ops = (double)n * CHAINS * (TENSOR ? 1 : 2);   // warp-instructions per warp
  const double fma_total = (double)blocks * (threads / 32) * n * CHAINS * (TENSOR ? 256.0 : 64.0);
  printf("%-28s blocks %4d thr %4d chains %2d: %7.2f cyc per warp-instr (per warp), %8.1f FMA/clk/SM, %.3f ms, %.2f TFLOP/s\n", name, blocks, threads, CHAINS,
         cy / ops, fma_total / blocks / cy * (blocks > 148 ? blocks / 148.0 : 1.0), ms, 2 * fma_total / (ms * 1e-3) / 1e12);
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 8 * (64 + 1024 * 1024)); cudaMalloc(&c, 64);
  double h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0 + 1e-9 * i; cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  run<1, true>("DMMA latency", 1, 32, d, c);
  run<1, false>("DFMA latency (2 indep)", 1, 32, d, c);
  run<8, true>("DMMA 1 warp 8 chains", 1, 32, d, c);
  run<8, false>("DFMA 1 warp 16 chains", 1, 32, d, c);
  run<8, true>("DMMA 4 warps", 1, 128, d, c);
  run<8, false>("DFMA 4 warps", 1, 128, d, c);
  run<8, true>("DMMA 8 warps", 1, 256, d, c);
  run<8, false>("DFMA 8 warps", 1, 256, d, c);
  run<8, true>("DMMA 16 warps", 1, 512, d, c);
  run<8, false>("DFMA 16 warps", 1, 512, d, c);
  run<8, true>("DMMA chip 148x8 warps", 148, 256, d, c);
  run<8, false>("DFMA chip 148x8 warps", 148, 256, d, c);
  run<8, true>("DMMA chip 148x16 warps", 148, 512, d, c);
  run<8, false>("DFMA chip 148x16 warps", 148, 512, d, c);
  return 0;
}
