// Cold vs warm timing of K_B's reduced-system solver (solve_reduced) in isolation: the same device code as
// the product kernel (this file includes k_schur_solve.cu), called several times inside ONE kernel launch
// on a synthetic SPD system; clock64 stamps per phase.  First call = cold instruction cache.
#define PBA_SOLVE_UBENCH 1
#include "../../photobundle_b200/csrc/k_schur_solve.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
using namespace pba;

__global__ void __launch_bounds__(kSchurThreads) k_bench(LmParams lp, const double* S0, int tot, int reps, long long* stamps, double* xout) {
  __shared__ LmState s_st;
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x;
  for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
    reinterpret_cast<int*>(&s_st)[i] = reinterpret_cast<const int*>(lp.st_in)[i];
  __syncthreads();
  for (int r = 0; r < reps; ++r) {
    for (int e = tid; e < tot; e += blockDim.x) lp.S[e] = S0[e];
    __threadfence();
    __syncthreads();
    lp.dbg = reinterpret_cast<unsigned long long*>(stamps + 64 * r);
    if (tid == 0) lp.dbg[2] = gtime();
    if (lp.split) solve_reduced<12>(lp, s_st, sm, lp.n_frames, nullptr, nullptr); else solve_reduced<3>(lp, s_st, sm, lp.n_frames, nullptr, nullptr);
    if (tid == 0) lp.dbg[3] = gtime();
    __syncthreads();
  }
  for (int i = tid; i < 6 * lp.n_frames; i += blockDim.x) xout[i] = s_st.step_c[i];
}

int main(int argc, char** argv) {
  const int nf = argc > 1 ? atoi(argv[1]) : 7, F = nf + 1, N = 6 * nf, ld = reduced_ld(N), reps = 4;
  std::vector<double> B(N * N), M(N * N), rhs(N), S0((size_t)N * ld, 0.0);
  srand(3);
  for (auto& v : B) v = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) { double s = 0; for (int k = 0; k < N; ++k) s += B[i * N + k] * B[j * N + k]; M[i * N + j] = s + (i == j ? N : 0); }
  for (auto& v : rhs) v = rand() / (double)RAND_MAX - 0.5;
  for (int r = 0; r < N; ++r) { for (int c = r; c < N; ++c) S0[(size_t)r * ld + c] = -M[r * N + c]; S0[(size_t)r * ld + N] = -rhs[r]; }
  LmState st; memset(&st, 0, sizeof(st));
  st.n_frames = F; st.fixed_frame = 0; st.n_free = nf; st.radius = 1e4; st.min_diag = 1e-6; st.max_diag = 1e32;
  for (int f = 0; f < kMaxFrames; ++f) st.free_index[f] = (f >= 1 && f < F) ? f - 1 : -1;
  for (int i = 0; i < kMaxD; ++i) st.scale_c[i] = 1.0;
  st.cur = 0; st.eval_buf = 1;
  LmState* d_st; cudaMalloc(&d_st, sizeof(st)); cudaMemcpy(d_st, &st, sizeof(st), cudaMemcpyHostToDevice);
  double *d_S, *d_S0, *d_cams, *d_Ucur, *d_x; long long* d_stamps;
  cudaMalloc(&d_S, sizeof(double) * S0.size()); cudaMalloc(&d_S0, sizeof(double) * S0.size());
  cudaMemcpy(d_S0, S0.data(), sizeof(double) * S0.size(), cudaMemcpyHostToDevice);
  cudaMalloc(&d_cams, sizeof(double) * 2 * F * 6); cudaMemset(d_cams, 0, sizeof(double) * 2 * F * 6);
  cudaMalloc(&d_Ucur, sizeof(double) * F * kUStride); cudaMemset(d_Ucur, 0, sizeof(double) * F * kUStride);
  cudaMalloc(&d_x, sizeof(double) * 6 * F); cudaMalloc(&d_stamps, 8 * 64 * reps); cudaMemset(d_stamps, 0, 8 * 64 * reps);
  LmParams lp; memset(&lp, 0, sizeof(lp));
  lp.st_in = d_st; lp.n_frames = F; lp.split = nf > 7 ? 1 : 0; lp.cams = d_cams; lp.Ucur = d_Ucur; lp.S = d_S;
  const size_t npad = (N + 1) & ~1;
  const size_t smem = sizeof(double) * ((size_t)(N + 8) * ld + nf * 36 + 3 * npad + F * kUStride + 2);
  cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  for (int pass = 0; pass < 2; ++pass) {
    k_bench<<<1, kSchurThreads, smem>>>(lp, d_S0, (int)S0.size(), reps, d_stamps, d_x);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<long long> s(64 * reps); cudaMemcpy(s.data(), d_stamps, 8 * 64 * reps, cudaMemcpyDeviceToHost);
    for (int r = 0; r < reps; ++r)
      printf("N=%d launch %d call %d: copy+assemble %6lld  factor %6lld  backsub %6lld  candidate %6lld  total %6lld cycles\n", N, pass, r,
             s[64 * r + 5] - s[64 * r + 2], s[64 * r + 6] - s[64 * r + 5], s[64 * r + 7] - s[64 * r + 6], s[64 * r + 3] - s[64 * r + 7], s[64 * r + 3] - s[64 * r + 2]);
    { const long long* t = &s[64 * (reps - 1)];
      printf("  block step 1 (thread 0): diag chol %lld  panel(+M wait) %lld  barrier %lld  trailing %lld  barrier %lld\n", t[17] - t[16], t[18] - t[17], t[19] - t[18], t[20] - t[19], t[21] - t[20]);
      printf("  arrival at barrier 1 per warp (rel. to step start):"); for (int w = 0; w < 8; ++w) printf(" %lld", t[32 + w] - t[16]);
      printf("\n  arrival at barrier 2 per warp:"); for (int w = 0; w < 8; ++w) printf(" %lld", t[40 + w] - t[16]); printf("\n"); }
  }
  // check against a host Cholesky solve of (M + damping) x = rhs ; step = -x
  std::vector<double> x(6 * F); cudaMemcpy(x.data(), d_x, sizeof(double) * 6 * F, cudaMemcpyDeviceToHost);
  std::vector<double> A(M); for (int i = 0; i < N; ++i) A[i * N + i] += 1e-6 / 1e4;
  std::vector<double> L(N * N, 0.0), y(N), z(N);
  for (int j = 0; j < N; ++j) { double d = A[j * N + j]; for (int k = 0; k < j; ++k) d -= L[j * N + k] * L[j * N + k]; L[j * N + j] = sqrt(d);
    for (int i = j + 1; i < N; ++i) { double t = A[i * N + j]; for (int k = 0; k < j; ++k) t -= L[i * N + k] * L[j * N + k]; L[i * N + j] = t / L[j * N + j]; } }
  for (int i = 0; i < N; ++i) { double t = rhs[i]; for (int k = 0; k < i; ++k) t -= L[i * N + k] * y[k]; y[i] = t / L[i * N + i]; }
  for (int i = N - 1; i >= 0; --i) { double t = y[i]; for (int k = i + 1; k < N; ++k) t -= L[k * N + i] * z[k]; z[i] = t / L[i * N + i]; }
  double err = 0, mx = 0;
  for (int i = 0; i < N; ++i) { err = fmax(err, fabs(-z[i] - x[6 + i])); mx = fmax(mx, fabs(z[i])); }
  printf("max |step + x_host| = %.3e (max |x| %.3e)\n", err, mx);
  return 0;
}
