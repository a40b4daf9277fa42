// SM clock during short kernels: clock64 vs globaltimer over a ~50 us spin, cold and after warm-up.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin(long long* out, long long cycles) {
  unsigned long long g0, g1; long long c0 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  while (clock64() - c0 < cycles) {}
  long long c1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = c1 - c0; out[1] = (long long)(g1 - g0); }
}
int main() {
  long long* d; cudaMalloc(&d, 16); long long h[2];
  for (int rep = 0; rep < 6; ++rep) {
    spin<<<rep < 3 ? 1 : 148, 256>>>(d, 100000); cudaDeviceSynchronize();
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("rep %d grid %3d: %lld cycles in %lld ns -> %.0f MHz\n", rep, rep < 3 ? 1 : 148, h[0], h[1], 1e3 * h[0] / h[1]);
  }
  return 0;
}
