// Dependent-chain latency microbenchmark (one warp): fp64 FMA/ADD/MUL, fp32 FMA, rsqrt(double), F2F, LDS.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(double* out, long long* cyc, int n) {
  double x = out[0], y = out[1];
  float xf = (float)x, yf = (float)y;
  __shared__ double sh[64];
  sh[threadIdx.x] = x; sh[threadIdx.x + 32] = y;
  __syncthreads();
  int idx = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if (OP == 0) x = fma(x, y, y);
    if (OP == 1) x = x + y;
    if (OP == 2) x = x * y;
    if (OP == 3) xf = fmaf(xf, yf, yf);
    if (OP == 4) x = rsqrt(x) + y;
    if (OP == 5) { xf = (float)x; x = (double)xf + 1e-30; }
    if (OP == 6) { idx = (int)sh[idx & 63] & 31; }
    if (OP == 7) x = 1.0 / x + y;
    if (OP == 8) x = sqrt(x) + y;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; }
  out[2 + threadIdx.x] = x + xf + idx;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 1024); cudaMalloc(&c, 64);
  double h[2] = {1.0000001, 0.9999999}; cudaMemcpy(d, h, 16, cudaMemcpyHostToDevice);
  const char* names[] = {"DFMA", "DADD", "DMUL", "FFMA", "rsqrt(double)+DADD", "F2F.F32.F64+F2F.F64.F32+DADD", "LDS.64->idx", "1.0/x(double)+DADD", "sqrt(double)+DADD"};
  const int n = 4096;
  for (int op = 0; op < 9; ++op) {
    for (int rep = 0; rep < 2; ++rep) {
      switch (op) { case 0: k<0><<<1,32>>>(d,c,n); break; case 1: k<1><<<1,32>>>(d,c,n); break; case 2: k<2><<<1,32>>>(d,c,n); break;
        case 3: k<3><<<1,32>>>(d,c,n); break; case 4: k<4><<<1,32>>>(d,c,n); break; case 5: k<5><<<1,32>>>(d,c,n); break;
        case 6: k<6><<<1,32>>>(d,c,n); break; case 7: k<7><<<1,32>>>(d,c,n); break; case 8: k<8><<<1,32>>>(d,c,n); break; }
      cudaDeviceSynchronize();
    }
    long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
    printf("%-32s %7.1f cycles per dependent op\n", names[op], (double)cy / n);
  }
  return 0;
}
