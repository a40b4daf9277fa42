// One-shot DMMA latency: LDS operands -> 2 dependent DMMAs -> STS, timed per pass, cold and repeated.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(long long* out, int gap) {
  __shared__ double sh[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = 1.0 + i * 1e-6;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int rep = 0; rep < 8; ++rep) {
    __syncthreads();
    long long t0 = clock64();
    double a = sh[lane + 64 * rep], b = sh[32 + lane + 64 * rep];
    double2 c = *reinterpret_cast<double2*>(&sh[1024 + 2 * lane + 64 * rep + 512 * warp]);
    long long t1 = clock64();
    dmma(c.x, c.y, a, b);
    dmma(c.x, c.y, b, a);
    long long t2 = clock64();
    *reinterpret_cast<double2*>(&sh[1024 + 2 * lane + 64 * rep + 512 * warp]) = c;
    long long t3 = clock64();
    if (lane == 0) { out[(warp * 8 + rep) * 4 + 0] = t1 - t0; out[(warp * 8 + rep) * 4 + 1] = t2 - t1; out[(warp * 8 + rep) * 4 + 2] = t3 - t2; }
    // idle gap between passes (tensor pipe power state?)
    for (int g = 0; g < gap; ++g) __nanosleep(100);
  }
}
int main() {
  long long* d; cudaMalloc(&d, 8 * 8 * 8 * 4);
  for (int gap = 0; gap <= 20; gap += 10) {
    for (int nw = 1; nw <= 8; nw *= 8) {
      k<<<1, 32 * nw>>>(d, gap); cudaDeviceSynchronize();
      long long h[8 * 8 * 4]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("gap %2d x100ns, %d warps: warp0 per pass [lds, 2xdmma, sts]:", gap, nw);
      for (int r = 0; r < 8; ++r) printf(" [%lld %lld %lld]", h[r * 4], h[r * 4 + 1], h[r * 4 + 2]);
      printf("\n");
    }
  }
  return 0;
}
