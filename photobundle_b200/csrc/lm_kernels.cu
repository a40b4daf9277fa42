// lm_kernels.cu — K2: the Levenberg-Marquardt loop around K1, entirely on the device.
//
// Replaces (Ceres 1.x, not in the reference tree; call site src/photobundle.cc:829 with the
// options of src/photobundle.cc:738-761): TrustRegionMinimizer, LevenbergMarquardtStrategy,
// SchurEliminator + the reduced-camera Cholesky (SPARSE_SCHUR is an exact solve of the damped
// normal equations, so a dense factorisation of the 6F x 6F reduced system is equivalent).
//
// One LM iteration = k_schur -> k_reduce_s -> k_solve -> k_backsub -> K1 (candidate) ->
// k_reduce_u -> k_decide.  All decisions are taken by k_decide on the device (LmState in
// HBM); every kernel returns immediately once st->done is set, so the host can enqueue
// iterations without synchronising.  All reductions are fixed-order (deterministic).
//
// Evaluate-at-candidate fusion: K1 evaluates residuals AND blocks at x+Δ, so an accepted
// step needs no second pass (Ceres evaluates cost at x+Δ, then residuals+Jacobian again).
// The previous blocks are kept (double buffer) because a rejected step re-solves at x.

#include "pba_device.cuh"

#include <cfloat>
#include <cmath>

namespace pba {

__device__ __forceinline__ int utri6(int a, int b) { return a * (13 - a) / 2 + (b - a); }  // a <= b

// ---- k_reduce_u: per-CTA pose-block partials of K1 -> U[buf], E[buf] ---------------------
__global__ void __launch_bounds__(1024) k_reduce_u(const LmParams lp) {
  const LmState* st = lp.st;
  if (st->done) return;
  const int buf = st->eval_buf, F = lp.n_frames, nb = lp.n_k1_ctas;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ double sh[32][33];
  if ((int)blockIdx.x < F) {
    const int f = blockIdx.x;
    double acc = 0.0;
    if (lane < kUStride)
      for (int b = warp; b < nb; b += 32) acc += lp.Upart[((size_t)b * F + f) * kUStride + lane];
    sh[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && lane < kUStride) {
      double t = 0.0;
      for (int w = 0; w < 32; ++w) t += sh[w][lane];
      lp.U[((size_t)buf * F + f) * kUStride + lane] = t;
    }
  } else {
    double c = 0.0, g2 = 0.0, gm = 0.0, x2 = 0.0;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
      const double* e = lp.Epart + (size_t)b * 4;
      c += e[0]; g2 += e[1]; gm = fmax(gm, e[2]); x2 += e[3];
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      c += __shfl_xor_sync(0xffffffffu, c, m);
      g2 += __shfl_xor_sync(0xffffffffu, g2, m);
      gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, m));
      x2 += __shfl_xor_sync(0xffffffffu, x2, m);
    }
    if (lane == 0) { sh[warp][0] = c; sh[warp][1] = g2; sh[warp][2] = gm; sh[warp][3] = x2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      c = 0.0; g2 = 0.0; gm = 0.0; x2 = 0.0;
      for (int w = 0; w < 32; ++w) { c += sh[w][0]; g2 += sh[w][1]; gm = fmax(gm, sh[w][2]); x2 += sh[w][3]; }
      double* E = lp.E + buf * 4;
      E[0] = c; E[1] = g2; E[2] = gm; E[3] = x2;
    }
  }
}

// ---- k_decide: Ceres TrustRegionMinimizer bookkeeping (one warp) --------------------------
__device__ void finish(LmState* st, int type, int code, double a, double b) {
  st->done = 1; st->termination_type = type; st->msg_code = code; st->msg_a = a; st->msg_b = b;
}

__global__ void __launch_bounds__(32) k_decide(const LmParams lp) {
  LmState* st = lp.st;
  if (st->done) return;
  const int lane = threadIdx.x, F = lp.n_frames, buf = st->eval_buf;
  // partial sums of the step under judgement (k_backsub) and gradient norms of the pose columns
  double sg = 0.0, sHs = 0.0, stepsq = 0.0, candsq = 0.0;
  for (int b = lane; b < lp.n_back_ctas; b += 32) {
    const double* q = lp.Bpart + (size_t)b * 4;
    sg += q[0]; sHs += q[1]; stepsq += q[2]; candsq += q[3];
  }
  double gm = 0.0, g2 = 0.0, csq = 0.0;
  for (int i = lane; i < F * 6; i += 32) {
    const int f = i / 6, a = i - f * 6;
    if (st->free_index[f] >= 0) {
      const double g = lp.U[((size_t)buf * F + f) * kUStride + 21 + a];
      gm = fmax(gm, fabs(g)); g2 += g * g;
      const double c = lp.cams[((size_t)buf * F + f) * 6 + a];
      csq += c * c;
    }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    sg += __shfl_xor_sync(0xffffffffu, sg, m);
    sHs += __shfl_xor_sync(0xffffffffu, sHs, m);
    stepsq += __shfl_xor_sync(0xffffffffu, stepsq, m);
    candsq += __shfl_xor_sync(0xffffffffu, candsq, m);
    gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, m));
    g2 += __shfl_xor_sync(0xffffffffu, g2, m);
    csq += __shfl_xor_sync(0xffffffffu, csq, m);
  }
  if (lane != 0) return;
  st->num_evals++;   // one K1 pass was consumed by this decision

  const double* E = lp.E + buf * 4;
  const double cost_e = E[0];
  const double gmax_e = fmax(gm, E[2]), gnorm_e = sqrt(g2 + E[1]);
  IterSummary it;
  memset(&it, 0, sizeof(it));

  if (st->iteration == 0) {
    // IterationZero
    st->x_cost = cost_e; st->initial_cost = cost_e;
    st->x_norm = sqrt(csq + E[3]);
    for (int f = 0; f < F; ++f)
      for (int a = 0; a < 6; ++a)
        st->scale_c[f * 6 + a] = st->jacobi_scaling
            ? 1.0 / (1.0 + sqrt(lp.U[((size_t)buf * F + f) * kUStride + utri6(a, a)])) : 1.0;
    st->gmax = gmax_e; st->gnorm = gnorm_e;
    st->radius = st->initial_radius; st->decrease_factor = 2.0;
    st->cur = buf; st->eval_buf = 1 - buf;
    it.iteration = 0; it.cost = cost_e; it.gradient_max_norm = gmax_e; it.gradient_norm = gnorm_e;
  } else {
    it.iteration = st->iteration;
    it.gradient_max_norm = st->gmax; it.gradient_norm = st->gnorm;
    it.linear_solver_iterations = 1;
    const double mcc = -(st->cam_sg + sg) - 0.5 * (st->cam_sHs + sHs);   // model_cost_change
    const bool valid = st->step_valid && (mcc > 0.0);
    it.step_is_valid = valid ? 1 : 0;
    if (!valid) {
      // HandleInvalidStep; LevenbergMarquardtStrategy::StepIsInvalid == StepRejected(0)
      if (++st->num_invalid >= st->max_invalid) { finish(st, 2, kMsgInvalidSteps, st->max_invalid, 0); return; }
      st->radius = st->radius / st->decrease_factor; st->decrease_factor *= 2.0;
      it.cost = st->x_cost;
      st->num_unsuccessful++;
    } else {
      st->num_invalid = 0;
      it.step_norm = sqrt(st->cam_step_sq + stepsq);
      const double ptol = st->parameter_tolerance;
      if (it.step_norm <= ptol * (st->x_norm + ptol)) {
        finish(st, 0, kMsgParamTol, it.step_norm / (st->x_norm + ptol), ptol); return;
      }
      const double cand_cost = isfinite(cost_e) ? cost_e : DBL_MAX;
      it.cost_change = st->x_cost - cand_cost;
      if (fabs(it.cost_change) <= st->function_tolerance * st->x_cost) {
        finish(st, 0, kMsgFuncTol, fabs(it.cost_change) / st->x_cost, st->function_tolerance); return;
      }
      it.relative_decrease = it.cost_change / mcc;
      if (it.relative_decrease > st->min_relative_decrease) {
        // HandleSuccessfulStep: the candidate's blocks are already in buffer `buf`
        st->cur = buf; st->eval_buf = 1 - buf;
        st->x_cost = cost_e; st->x_norm = sqrt(st->cam_cand_sq + candsq);
        st->gmax = gmax_e; st->gnorm = gnorm_e;
        it.gradient_max_norm = gmax_e; it.gradient_norm = gnorm_e;
        it.step_is_successful = 1; it.cost = cost_e;
        const double q = 2.0 * it.relative_decrease - 1.0;
        st->radius = fmin(st->max_radius, st->radius / fmax(1.0 / 3.0, 1.0 - q * q * q));
        st->decrease_factor = 2.0;
        st->num_successful++;
      } else {
        // HandleUnsuccessfulStep
        it.cost = cand_cost;
        st->radius = st->radius / st->decrease_factor; st->decrease_factor *= 2.0;
        st->num_unsuccessful++;
      }
    }
  }
  // FinalizeIterationAndCheckIfMinimizerCanContinue
  it.trust_region_radius = st->radius;
  lp.trace[st->n_trace++] = it;
  if (it.iteration >= st->max_num_iterations) { finish(st, 1, kMsgMaxIter, it.iteration, 0); return; }
  if (st->gmax <= st->gradient_tolerance) { finish(st, 0, kMsgGradTol, st->gmax, st->gradient_tolerance); return; }
  if (!(st->radius > st->min_radius)) { finish(st, 0, kMsgMinRadius, st->radius, st->min_radius); return; }
  st->iteration++;
}

// ---- k_schur: eliminate the 3x3 point blocks ----------------------------------------------
// Phase A (warp per point): Vs = sp V sp + D_p², its inverse, Ws = sc W sp, Y = Ws Vs^-1 into
// shared memory.  Phase B (thread per 3x3 tile of the D x D reduced matrix, accumulators in
// registers across all of the CTA's points): S -= Y Wsᵀ.  One partial per CTA.
__device__ __forceinline__ void inv_sym3(const double* a /*00 01 02 11 12 22*/, double* inv) {
  const double c00 = a[3] * a[5] - a[4] * a[4];
  const double c01 = a[2] * a[4] - a[1] * a[5];
  const double c02 = a[1] * a[4] - a[2] * a[3];
  const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  const double id = 1.0 / det;
  inv[0] = c00 * id; inv[1] = c01 * id; inv[2] = c02 * id;
  inv[3] = (a[0] * a[5] - a[2] * a[2]) * id;
  inv[4] = (a[1] * a[2] - a[0] * a[4]) * id;
  inv[5] = (a[0] * a[3] - a[1] * a[1]) * id;
}
__device__ __forceinline__ double sym3(const double* s, int a, int b) {
  const int lo = a < b ? a : b, hi = a < b ? b : a;
  return s[lo * (5 - lo) / 2 + hi];  // 00 01 02 11 12 22
}

template <int TPT>
__global__ void __launch_bounds__(kSchurThreads) k_schur(const LmParams lp) {
  const LmState* st = lp.st;
  if (st->done) return;
  const int F = lp.n_frames, D = 6 * F, n = lp.n_points, cur = st->cur;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double radius = st->radius, dmin = st->min_diag, dmax = st->max_diag;
  const bool first = (st->iteration == 1);

  extern __shared__ double sm[];
  double* Yf = sm;                                  // [chunk][D*3]
  double* Wf = Yf + kSchurChunk * D * 3;            // [chunk][D*3]
  double* rh = Wf + kSchurChunk * D * 3;            // [chunk][D]
  __shared__ unsigned s_mask[kSchurChunk];

  const double* Vb = lp.V + (size_t)cur * n * 6;
  const double* gb = lp.gp + (size_t)cur * n * 3;
  const double* Wb = lp.W + (size_t)cur * lp.nnz * 18;

  const int T = D / 3;
  double acc[TPT][9];
#pragma unroll
  for (int k = 0; k < TPT; ++k)
#pragma unroll
    for (int e = 0; e < 9; ++e) acc[k][e] = 0.0;
  double racc = 0.0;

  const int n_chunks = (n + kSchurChunk - 1) / kSchurChunk;
  for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const int p = chunk * kSchurChunk + warp;
    unsigned mask = 0;
    if (p < n) {
      double sp[3], Vs[6], Vi[6], gs[3];
      const double* V = Vb + (size_t)p * 6;
      if (first) {
        sp[0] = st->jacobi_scaling ? 1.0 / (1.0 + sqrt(V[0])) : 1.0;
        sp[1] = st->jacobi_scaling ? 1.0 / (1.0 + sqrt(V[3])) : 1.0;
        sp[2] = st->jacobi_scaling ? 1.0 / (1.0 + sqrt(V[5])) : 1.0;
        if (lane < 3) lp.scale_p[(size_t)p * 3 + lane] = sp[lane];
      } else {
        sp[0] = lp.scale_p[(size_t)p * 3]; sp[1] = lp.scale_p[(size_t)p * 3 + 1]; sp[2] = lp.scale_p[(size_t)p * 3 + 2];
      }
      Vs[0] = sp[0] * V[0] * sp[0]; Vs[1] = sp[0] * V[1] * sp[1]; Vs[2] = sp[0] * V[2] * sp[2];
      Vs[3] = sp[1] * V[3] * sp[1]; Vs[4] = sp[1] * V[4] * sp[2]; Vs[5] = sp[2] * V[5] * sp[2];
      Vs[0] += fmin(fmax(Vs[0], dmin), dmax) / radius;
      Vs[3] += fmin(fmax(Vs[3], dmin), dmax) / radius;
      Vs[5] += fmin(fmax(Vs[5], dmin), dmax) / radius;
      inv_sym3(Vs, Vi);
      if (lane < 6) lp.Vinv[(size_t)p * 6 + lane] = Vi[lane];
      gs[0] = sp[0] * gb[(size_t)p * 3]; gs[1] = sp[1] * gb[(size_t)p * 3 + 1]; gs[2] = sp[2] * gb[(size_t)p * 3 + 2];
      const int o0 = lp.obs_off[p], nobs = lp.obs_off[p + 1] - o0;
      for (int i = 0; i < nobs; ++i) {
        const int f = lp.obs_frame[o0 + i];
        if (st->free_index[f] < 0) continue;
        mask |= 1u << f;
        const int a = lane / 3, b = lane - a * 3;
        double ws = 0.0;
        if (lane < 18) {
          ws = st->scale_c[f * 6 + a] * Wb[(size_t)(o0 + i) * 18 + lane] * sp[b];
          Wf[(warp * D + 6 * f) * 3 + lane] = ws;
        }
        // Y[a][b] = sum_q Ws[a][q] Vi[q][b]: gather the row's three entries with shuffles
        const int base = a * 3;
        const double w0 = __shfl_sync(0xffffffffu, ws, base < 18 ? base : 0);
        const double w1 = __shfl_sync(0xffffffffu, ws, base < 18 ? base + 1 : 0);
        const double w2 = __shfl_sync(0xffffffffu, ws, base < 18 ? base + 2 : 0);
        if (lane < 18) {
          const double y = w0 * sym3(Vi, 0, b) + w1 * sym3(Vi, 1, b) + w2 * sym3(Vi, 2, b);
          Yf[(warp * D + 6 * f) * 3 + lane] = y;
          // rhs[a] -= sum_b Y[a][b] gs[b]: combine the three lanes of a row
          double t = y * gs[b];
          const double t1 = __shfl_down_sync(0x3ffffu, t, 1);
          const double t2 = __shfl_down_sync(0x3ffffu, t, 2);
          if (b == 0) rh[warp * D + 6 * f + a] = -(t + t1 + t2);
        }
      }
    }
    if (lane == 0) s_mask[warp] = mask;
    __syncthreads();
    // Phase B
#pragma unroll
    for (int k = 0; k < TPT; ++k) {
      const int t = threadIdx.x + k * kSchurThreads;
      if (t < T * T) {
        const int tr = t / T, tc = t - tr * T;
        const int fr = tr >> 1, fc = tc >> 1;
#pragma unroll
        for (int w = 0; w < kSchurChunk; ++w) {
          const unsigned m = s_mask[w];
          if (((m >> fr) & 1u) && ((m >> fc) & 1u)) {
            const double* y = Yf + (w * D + 3 * tr) * 3;
            const double* ww = Wf + (w * D + 3 * tc) * 3;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j)
                acc[k][i * 3 + j] -= y[i * 3] * ww[j * 3] + y[i * 3 + 1] * ww[j * 3 + 1] + y[i * 3 + 2] * ww[j * 3 + 2];
          }
        }
      }
    }
    if ((int)threadIdx.x < D) {
      const int f = threadIdx.x / 6;
#pragma unroll
      for (int w = 0; w < kSchurChunk; ++w)
        if ((s_mask[w] >> f) & 1u) racc += rh[w * D + threadIdx.x];
    }
    __syncthreads();
  }
  double* out = lp.Spart + (size_t)blockIdx.x * (D * D + D);
#pragma unroll
  for (int k = 0; k < TPT; ++k) {
    const int t = threadIdx.x + k * kSchurThreads;
    if (t < T * T) {
      const int tr = t / T, tc = t - tr * T;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) out[(3 * tr + i) * D + 3 * tc + j] = acc[k][i * 3 + j];
    }
  }
  if ((int)threadIdx.x < D) out[D * D + threadIdx.x] = racc;
}

// ---- k_reduce_s: sum the per-CTA Schur partials ---------------------------------------------
__global__ void __launch_bounds__(256) k_reduce_s(const LmParams lp) {
  const LmState* st = lp.st;
  if (st->done) return;
  const int D = 6 * lp.n_frames, total = D * D + D, nb = lp.n_schur_ctas;
  const int e = blockIdx.x * 64 + (threadIdx.x & 63), slice = threadIdx.x >> 6;
  __shared__ double sh[4][64];
  double acc = 0.0;
  if (e < total)
    for (int b = slice; b < nb; b += 4) acc += lp.Spart[(size_t)b * total + e];
  sh[slice][threadIdx.x & 63] = acc;
  __syncthreads();
  if (slice == 0 && e < total) {
    const int i = threadIdx.x;
    lp.S[e] = (sh[0][i] + sh[1][i]) + (sh[2][i] + sh[3][i]);
  }
}

// ---- k_solve: reduced camera system, blocked (6x6) Cholesky in fp64, one CTA ------------------
__global__ void __launch_bounds__(256) k_solve(const LmParams lp) {
  LmState* st = lp.st;
  if (st->done) return;
  const int F = lp.n_frames, D = 6 * F, nf = st->n_free, N = 6 * nf;
  const int cur = st->cur, eb = st->eval_buf, tid = threadIdx.x;
  const double radius = st->radius;
  extern __shared__ double sm[];
  const int ld = N + 1;
  double* A = sm;                 // [N][ld] lower triangle
  double* Li = A + N * ld;        // [nf][36] inverse of the diagonal blocks of L
  double* bb = Li + nf * 36;      // [N]
  __shared__ int s_fr[kMaxFrames];  // free index -> frame
  __shared__ int s_ok;
  if (tid == 0) {
    s_ok = 1;
    for (int f = 0; f < F; ++f) if (st->free_index[f] >= 0) s_fr[st->free_index[f]] = f;
  }
  __syncthreads();
  const double* Ub = lp.U + (size_t)cur * F * kUStride;
  // assemble S + Us + Dc², rhs + gs_c
  for (int e = tid; e < N * N; e += blockDim.x) {
    const int r = e / N, c = e - r * N;
    const int fi = r / 6, a = r - fi * 6, fj = c / 6, b = c - fj * 6;
    const int f = s_fr[fi], g = s_fr[fj];
    double v = lp.S[(6 * f + a) * D + 6 * g + b];
    if (fi == fj) {
      const double u = st->scale_c[6 * f + a] * Ub[f * kUStride + (a <= b ? utri6(a, b) : utri6(b, a))] * st->scale_c[6 * f + b];
      v += u;
      if (a == b) v += fmin(fmax(u, st->min_diag), st->max_diag) / radius;
    }
    A[r * ld + c] = v;
  }
  for (int r = tid; r < N; r += blockDim.x) {
    const int fi = r / 6, a = r - fi * 6, f = s_fr[fi];
    bb[r] = lp.S[D * D + 6 * f + a] + st->scale_c[6 * f + a] * Ub[f * kUStride + 21 + a];
  }
  __syncthreads();

  for (int jb = 0; jb < nf; ++jb) {
    const int j0 = 6 * jb, m = N - j0 - 6;
    if (tid == 0) {
      // 6x6 Cholesky of the diagonal block and the inverse of its factor (registers)
      double L[6][6], M[6][6];
      bool ok = true;
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int k = 0; k <= i; ++k) L[i][k] = A[(j0 + i) * ld + j0 + k];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        double d = L[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
        if (!(d > 0.0) || !isfinite(d)) { ok = false; d = 1.0; }
        d = sqrt(d);
        L[j][j] = d;
        const double id = 1.0 / d;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
          double s = L[i][j];
#pragma unroll
          for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
          L[i][j] = s * id;
        }
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) {       // M = L^-1 (lower), column by column
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          if (i < c) { M[i][c] = 0.0; continue; }
          double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
          for (int k = c; k < i; ++k) s -= L[i][k] * M[k][c];
          M[i][c] = s / L[i][i];
        }
      }
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          if (k <= i) A[(j0 + i) * ld + j0 + k] = L[i][k];
          Li[jb * 36 + i * 6 + k] = (k <= i) ? M[i][k] : 0.0;
        }
      if (!ok) s_ok = 0;
    }
    __syncthreads();
    // panel: L_p = A_p * L_dd^-T  -> L_p[i][c] = sum_{k<=c} A_p[i][k] * Linv[c][k]
    if (tid < m) {
      const int i = j0 + 6 + tid;
      double r[6], o[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) r[k] = A[i * ld + j0 + k];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k <= c; ++k) s += r[k] * Li[jb * 36 + c * 6 + k];
        o[c] = s;
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) A[i * ld + j0 + c] = o[c];
    }
    __syncthreads();
    // trailing update (lower triangle)
    for (int e = tid; e < m * m; e += blockDim.x) {
      const int ii = e / m, kk = e - ii * m;
      if (kk <= ii) {
        const double* ri = A + (j0 + 6 + ii) * ld + j0;
        const double* rk = A + (j0 + 6 + kk) * ld + j0;
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) s += ri[c] * rk[c];
        A[(j0 + 6 + ii) * ld + j0 + 6 + kk] -= s;
      }
    }
    __syncthreads();
  }
  // forward substitution L y = b (block-wise with the inverted diagonal blocks)
  for (int jb = 0; jb < nf; ++jb) {
    const int j0 = 6 * jb;
    __shared__ double yj[6];
    if (tid < 6) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) s += Li[jb * 36 + tid * 6 + k] * bb[j0 + k];
      yj[tid] = s;
    }
    __syncthreads();
    if (tid < 6) bb[j0 + tid] = yj[tid];
    const int i = j0 + 6 + tid;
    if (i < N) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) s += A[i * ld + j0 + k] * yj[k];
      bb[i] -= s;
    }
    __syncthreads();
  }
  // backward substitution L^T x = y
  for (int jb = nf - 1; jb >= 0; --jb) {
    const int j0 = 6 * jb;
    __shared__ double xj[6];
    if (tid < 6) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) s += Li[jb * 36 + k * 6 + tid] * bb[j0 + k];   // (L_dd^-T)[tid][k]
      xj[tid] = s;
    }
    __syncthreads();
    if (tid < 6) bb[j0 + tid] = xj[tid];
    if (tid < j0) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) s += A[(j0 + k) * ld + tid] * xj[k];
      bb[tid] -= s;
    }
    __syncthreads();
  }
  // step (scaled space) = -y ; candidate cameras ; camera part of the model cost change
  __shared__ double red[4][kMaxD];
  bool finite_ok = true;
  for (int i = tid; i < D; i += blockDim.x) {
    const int f = i / 6, a = i - f * 6, fi = st->free_index[f];
    const double x = lp.cams[((size_t)cur * F + f) * 6 + a];
    double step = 0.0, cand = x;
    red[0][i] = 0.0; red[1][i] = 0.0; red[2][i] = 0.0; red[3][i] = 0.0;
    if (fi >= 0) {
      step = -bb[6 * fi + a];
      if (!isfinite(step)) finite_ok = false;
      cand = x + step * st->scale_c[i];
      const double gsc = st->scale_c[i] * Ub[f * kUStride + 21 + a];
      double hs = 0.0;  // (Us step)_a
      for (int b = 0; b < 6; ++b) {
        const double u = st->scale_c[6 * f + a] * Ub[f * kUStride + (a <= b ? utri6(a, b) : utri6(b, a))] * st->scale_c[6 * f + b];
        hs += u * (-bb[6 * fi + b]);
      }
      red[0][i] = step * gsc; red[1][i] = step * hs;
      red[2][i] = (x - cand) * (x - cand); red[3][i] = cand * cand;
    }
    st->step_c[i] = step;
    lp.cams[((size_t)eb * F + f) * 6 + a] = cand;
  }
  if (!finite_ok) s_ok = 0;
  __syncthreads();
  if (tid == 0) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int i = 0; i < D; ++i) { a0 += red[0][i]; a1 += red[1][i]; a2 += red[2][i]; a3 += red[3][i]; }
    st->cam_sg = a0; st->cam_sHs = a1; st->cam_step_sq = a2; st->cam_cand_sq = a3;
    st->step_valid = s_ok;
  }
}

// ---- k_backsub: point steps, candidate points, point part of the model cost change ---------
__global__ void __launch_bounds__(kBackThreads) k_backsub(const LmParams lp) {
  const LmState* st = lp.st;
  if (st->done) return;
  const int n = lp.n_points, cur = st->cur, eb = st->eval_buf;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  double sg = 0.0, sHs = 0.0, stepsq = 0.0, candsq = 0.0;
  if (p < n) {
    const double* V = lp.V + ((size_t)cur * n + p) * 6;
    const double* gp = lp.gp + ((size_t)cur * n + p) * 3;
    const double* Wb = lp.W + (size_t)cur * lp.nnz * 18;
    const double sp[3] = {lp.scale_p[(size_t)p * 3], lp.scale_p[(size_t)p * 3 + 1], lp.scale_p[(size_t)p * 3 + 2]};
    double Vi[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) Vi[k] = lp.Vinv[(size_t)p * 6 + k];
    const double gs[3] = {sp[0] * gp[0], sp[1] * gp[1], sp[2] * gp[2]};
    double t[3] = {gs[0], gs[1], gs[2]};
    double wts[3] = {0.0, 0.0, 0.0};   // Ws^T step_c summed over the point's frames
    const int o0 = lp.obs_off[p], o1 = lp.obs_off[p + 1];
    for (int o = o0; o < o1; ++o) {
      const int f = lp.obs_frame[o];
      if (st->free_index[f] < 0) continue;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        const double sc_step = st->scale_c[6 * f + a] * st->step_c[6 * f + a];
#pragma unroll
        for (int b = 0; b < 3; ++b) wts[b] += Wb[(size_t)o * 18 + a * 3 + b] * sc_step;
      }
    }
#pragma unroll
    for (int b = 0; b < 3; ++b) { wts[b] *= sp[b]; t[b] += wts[b]; }   // t = gs - Ws^T y_c, y_c = -step_c
    double s[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
      s[a] = -(sym3(Vi, a, 0) * t[0] + sym3(Vi, a, 1) * t[1] + sym3(Vi, a, 2) * t[2]);
    // model terms: s.gs + (s^T Vs0 s + 2 step_c^T Ws s) with the undamped Vs0
    double q = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      sg += s[a] * gs[a];
#pragma unroll
      for (int b = 0; b < 3; ++b) q += s[a] * (sp[a] * sym3(V, a, b) * sp[b]) * s[b];
      q += 2.0 * wts[a] * s[a];
    }
    sHs = q;
    const double* X = lp.pts + ((size_t)cur * n + p) * 3;
    double* Xc = lp.pts + ((size_t)eb * n + p) * 3;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double c = X[a] + s[a] * sp[a];
      Xc[a] = c;
      stepsq += (X[a] - c) * (X[a] - c);
      candsq += c * c;
    }
  }
  __shared__ double sh[kBackThreads / 32][4];
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    sg += __shfl_xor_sync(0xffffffffu, sg, m);
    sHs += __shfl_xor_sync(0xffffffffu, sHs, m);
    stepsq += __shfl_xor_sync(0xffffffffu, stepsq, m);
    candsq += __shfl_xor_sync(0xffffffffu, candsq, m);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[warp][0] = sg; sh[warp][1] = sHs; sh[warp][2] = stepsq; sh[warp][3] = candsq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int w = 0; w < kBackThreads / 32; ++w) { a0 += sh[w][0]; a1 += sh[w][1]; a2 += sh[w][2]; a3 += sh[w][3]; }
    double* out = lp.Bpart + (size_t)blockIdx.x * 4;
    out[0] = a0; out[1] = a1; out[2] = a2; out[3] = a3;
  }
}

// ---- launchers ----------------------------------------------------------------------------
int schur_grid(int n_points, int sm_count) {
  const int chunks = (n_points + kSchurChunk - 1) / kSchurChunk;
  return chunks < sm_count ? (chunks > 0 ? chunks : 1) : sm_count;
}
int back_grid(int n_points) { return n_points > 0 ? (n_points + kBackThreads - 1) / kBackThreads : 1; }

cudaError_t launch_reduce_u(const LmParams& lp, cudaStream_t s) {
  k_reduce_u<<<lp.n_frames + 1, 1024, 0, s>>>(lp);
  return cudaGetLastError();
}
cudaError_t launch_decide(const LmParams& lp, cudaStream_t s) {
  k_decide<<<1, 32, 0, s>>>(lp);
  return cudaGetLastError();
}
cudaError_t launch_schur(const LmParams& lp, cudaStream_t s) {
  const int D = 6 * lp.n_frames, T = D / 3;
  const size_t smem = sizeof(double) * (size_t)kSchurChunk * (D * 3 * 2 + D);
  const int need = (T * T + kSchurThreads - 1) / kSchurThreads;
  if (need <= 1) {
    k_schur<1><<<lp.n_schur_ctas, kSchurThreads, smem, s>>>(lp);
  } else if (need <= 2) {
    k_schur<2><<<lp.n_schur_ctas, kSchurThreads, smem, s>>>(lp);
  } else {
    static bool cfg = false;
    if (!cfg) { cudaFuncSetAttribute(k_schur<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); cfg = true; }
    k_schur<4><<<lp.n_schur_ctas, kSchurThreads, smem, s>>>(lp);
  }
  return cudaGetLastError();
}
cudaError_t launch_reduce_s(const LmParams& lp, cudaStream_t s) {
  const int D = 6 * lp.n_frames, total = D * D + D;
  k_reduce_s<<<(total + 63) / 64, 256, 0, s>>>(lp);
  return cudaGetLastError();
}
cudaError_t launch_solve(const LmParams& lp, cudaStream_t s) {
  const int N = 6 * lp.n_frames;  // upper bound on 6*n_free
  const size_t smem = sizeof(double) * ((size_t)N * (N + 1) + (size_t)lp.n_frames * 36 + N);
  static bool cfg[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!cfg[dev & 63]) { cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); cfg[dev & 63] = true; }
  k_solve<<<1, 256, smem, s>>>(lp);
  return cudaGetLastError();
}
cudaError_t launch_backsub(const LmParams& lp, cudaStream_t s) {
  k_backsub<<<lp.n_back_ctas, kBackThreads, 0, s>>>(lp);
  return cudaGetLastError();
}

}  // namespace pba
