// Standalone timing of the blocked 6x6 Cholesky used by K_B's reduced solve (N = 42), per phase,
// using clock64 on thread 0 and the max arrival time over all threads before each barrier.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#include <cmath>
constexpr int NT = 256;
__global__ void __launch_bounds__(NT) chol(const double* Ain, int N, long long* stamps, long long* arrive, double* Lout) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, nf = N / 6, ld = N + 1;
  double* A = sm; double* Li = A + N * ld;
  for (int e = tid; e < N * N; e += NT) A[(e / N) * ld + e % N] = Ain[e];
  __syncthreads();
  long long t0 = clock64();
  for (int jb = 0; jb < nf; ++jb) {
    const int j0 = 6 * jb, m = N - j0 - 6;
    if (tid == 0) {
      double L[6][6];
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int k = 0; k <= i; ++k) L[i][k] = A[(j0 + i) * ld + j0 + k];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        double d = L[j][j];
        const double id = rsqrt(d);
        L[j][j] = d * id;
        Li[jb * 36 + j * 7] = id;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) L[i][j] *= id;
#pragma unroll
        for (int i = j + 1; i < 6; ++i)
#pragma unroll
          for (int k = j + 1; k <= i; ++k) L[i][k] -= L[i][j] * L[k][j];
      }
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int k = 0; k <= i; ++k) A[(j0 + i) * ld + j0 + k] = L[i][k];
      if (jb == 1) stamps[0] = clock64() - t0;
    }
    if (jb == 1) arrive[tid] = clock64() - t0;
    __syncthreads();
    if (jb == 1 && tid == 0) stamps[1] = clock64() - t0;
    if (tid < m) {
      const int i = j0 + 6 + tid;
      double r[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) r[k] = A[i * ld + j0 + k];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        r[c] *= Li[jb * 36 + c * 7];
#pragma unroll
        for (int k = c + 1; k < 6; ++k) r[k] -= r[c] * A[(j0 + k) * ld + j0 + c];
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) A[i * ld + j0 + c] = r[c];
    }
    if (jb == 1) arrive[NT + tid] = clock64() - t0;
    __syncthreads();
    if (jb == 1 && tid == 0) stamps[2] = clock64() - t0;
    for (int i0 = ty; i0 < m; i0 += 48) {
      for (int k0 = tx; k0 <= i0 + 32 && k0 < m; k0 += 48) {
        double rkv[3][6], res[3][3];
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          const int kk = k0 + 16 * v;
          const double* rk = A + (j0 + 6 + (kk < m ? kk : 0)) * ld + j0;
#pragma unroll
          for (int c = 0; c < 6; ++c) rkv[v][c] = rk[c];
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int ii = i0 + 16 * u;
          const bool rok = ii < m;
          const double* ri = A + (j0 + 6 + (rok ? ii : 0)) * ld + j0;
          const double r0 = ri[0], r1 = ri[1], r2 = ri[2], r3 = ri[3], r4 = ri[4], r5 = ri[5];
#pragma unroll
          for (int v = 0; v < 3; ++v) {
            const int kk = k0 + 16 * v;
            const bool ok = rok && kk <= ii;
            const double sa = r0 * rkv[v][0] + r2 * rkv[v][2] + r4 * rkv[v][4];
            const double sb = r1 * rkv[v][1] + r3 * rkv[v][3] + r5 * rkv[v][5];
            res[u][v] = ok ? A[(j0 + 6 + ii) * ld + j0 + 6 + kk] - (sa + sb) : 0.0;
          }
        }
#pragma unroll
        for (int u = 0; u < 3; ++u)
#pragma unroll
          for (int v = 0; v < 3; ++v) {
            const int ii = i0 + 16 * u, kk = k0 + 16 * v;
            if (ii < m && kk <= ii) A[(j0 + 6 + ii) * ld + j0 + 6 + kk] = res[u][v];
          }
      }
    }
    if (jb == 1) arrive[2 * NT + tid] = clock64() - t0;
    __syncthreads();
    if (jb == 1 && tid == 0) stamps[3] = clock64() - t0;
    if (jb == 0 && tid == 0) stamps[4] = clock64() - t0;
  }
  if (tid == 0) stamps[5] = clock64() - t0;
  __syncthreads();
  for (int e = tid; e < N * N; e += NT) Lout[e] = (e % N <= e / N) ? A[(e / N) * ld + e % N] : 0.0;
}
int main() {
  const int N = 42;
  std::vector<double> B(N * N), A(N * N, 0.0);
  srand(1);
  for (auto& v : B) v = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) { double s = 0; for (int k = 0; k < N; ++k) s += B[i * N + k] * B[j * N + k]; A[i * N + j] = s + (i == j ? N : 0); }
  double* dA; long long *ds, *da; cudaMalloc(&dA, sizeof(double) * N * N); cudaMalloc(&ds, 64 * 8); cudaMalloc(&da, 3 * NT * 8); double* dL; cudaMalloc(&dL, sizeof(double) * N * N);
  cudaMemcpy(dA, A.data(), sizeof(double) * N * N, cudaMemcpyHostToDevice);
  size_t smem = sizeof(double) * (N * (N + 1) + (N / 6) * 36 + N);
  for (int rep = 0; rep < 3; ++rep) { chol<<<1, NT, smem>>>(dA, N, ds, da, dL); cudaDeviceSynchronize(); }
  long long s[8]; std::vector<long long> a(3 * NT);
  cudaMemcpy(s, ds, 64, cudaMemcpyDeviceToHost); cudaMemcpy(a.data(), da, 3 * NT * 8, cudaMemcpyDeviceToHost);
  printf("step0 total %lld | step1: diag done %lld, after sync1 %lld, after sync2 (panel) %lld, after sync3 (trailing) %lld | all 7 steps %lld cycles\n", s[4], s[0], s[1], s[2], s[3], s[5]);
  { std::vector<double> L(N * N); cudaMemcpy(L.data(), dL, sizeof(double) * N * N, cudaMemcpyDeviceToHost); double err = 0; for (int i = 0; i < N; ++i) for (int j = 0; j <= i; ++j) { double t = 0; for (int k = 0; k <= j; ++k) t += L[i * N + k] * L[j * N + k]; err = fmax(err, fabs(t - A[i * N + j])); } printf("  max |L L^T - A| = %.3e\n", err); }
  for (int ph = 0; ph < 3; ++ph) { long long mx = 0, mn = 1ll << 60; int amx = 0; for (int t = 0; t < NT; ++t) { if (a[ph * NT + t] > mx) { mx = a[ph * NT + t]; amx = t; } if (a[ph * NT + t] < mn) mn = a[ph * NT + t]; }
    printf("  arrive before barrier %d: min %lld max %lld (thread %d)\n", ph + 1, mn, mx, amx); }
  return 0;
}
