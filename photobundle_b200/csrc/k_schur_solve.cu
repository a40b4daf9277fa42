// k_schur_solve.cu — K_B: LM decision + Schur elimination of the point blocks + reduced
// camera solve, one kernel per LM iteration.
//
// Replaces (Ceres 1.x, not in the reference tree; call site src/photobundle.cc:829 with the
// options of src/photobundle.cc:738-761): TrustRegionMinimizer bookkeeping,
// LevenbergMarquardtStrategy, SchurEliminator::Eliminate and the reduced-camera Cholesky
// (SPARSE_SCHUR is an exact solve of the damped normal equations, so a dense factorisation of
// the 6F x 6F reduced system is equivalent up to rounding).
//
// Phases inside the kernel:
//  (D) every CTA redundantly takes the trust-region decision for the candidate K_A has just
//      evaluated (accept / reject / converged, new radius) from the complete accumulators —
//      identical inputs, identical arithmetic, so all CTAs agree without a grid barrier;
//  (E) warp per point: Vs = sp V sp + D_p², its inverse, Ws = sc W sp, Y = Ws Vs^-1 into shared
//      memory; thread per 3x3 tile of the D x D reduced matrix (accumulators in registers across
//      all of the CTA's points): S -= Y Wsᵀ; one fp64 atomic per entry per CTA;
//  (S) the last CTA to finish (ticket) assembles S + Us + D_c², factors it with a blocked (6x6)
//      Cholesky in fp64 in shared memory, solves, and writes the camera step, the candidate
//      cameras and the next LmState.  The back-substitution of the points is fused into K_A.
//
// LM state ping-pongs between two LmState structs (st_in is read-only during the kernel).

#include "pba_device.cuh"

#include <cfloat>
#include <cmath>

namespace pba {

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ int utri6(int a, int b) { return a * (13 - a) / 2 + (b - a); }  // a <= b
__device__ __forceinline__ double usym6(const double* u21, int a, int b) { return u21[a <= b ? utri6(a, b) : utri6(b, a)]; }

__device__ __forceinline__ void finish(LmState& st, int type, int code, double a, double b) {
  st.done = 1; st.termination_type = type; st.msg_code = code; st.msg_a = a; st.msg_b = b;
}

// Ceres TrustRegionMinimizer bookkeeping for the candidate just evaluated.  `st` is the CTA's
// private copy of the state; returns whether an IterationSummary was produced.
__device__ bool decide(LmState& st, const double* E, double gm, double g2, double csq, const double* Ubuf, int F,
                       IterSummary& it) {
  st.num_evals++;
  st.took_step = 0;
  const int buf = st.eval_buf;
  const double cost_e = E[0];
  const double gmax_e = fmax(gm, E[2]), gnorm_e = sqrt(g2 + E[1]);
  memset(&it, 0, sizeof(it));
  if (st.iteration == 0) {
    // IterationZero
    st.x_cost = cost_e; st.initial_cost = cost_e;
    st.x_norm = sqrt(csq + E[3]);
    for (int f = 0; f < F; ++f)
      for (int a = 0; a < 6; ++a)
        st.scale_c[f * 6 + a] = st.jacobi_scaling ? 1.0 / (1.0 + sqrt(Ubuf[f * kUStride + utri6(a, a)])) : 1.0;
    st.gmax = gmax_e; st.gnorm = gnorm_e;
    st.radius = st.initial_radius; st.decrease_factor = 2.0;
    st.cur = buf; st.eval_buf = 1 - buf; st.took_step = 1;
    it.iteration = 0; it.cost = cost_e; it.gradient_max_norm = gmax_e; it.gradient_norm = gnorm_e;
  } else {
    it.iteration = st.iteration;
    it.gradient_max_norm = st.gmax; it.gradient_norm = st.gnorm;
    it.linear_solver_iterations = 1;
    const double mcc = -(st.cam_sg + E[4]) - 0.5 * (st.cam_sHs + E[5]);   // model_cost_change
    const bool valid = st.step_valid && (mcc > 0.0);
    it.step_is_valid = valid ? 1 : 0;
    if (!valid) {
      // HandleInvalidStep; LevenbergMarquardtStrategy::StepIsInvalid == StepRejected(0)
      if (++st.num_invalid >= st.max_invalid) { finish(st, 2, kMsgInvalidSteps, st.max_invalid, 0); return false; }
      st.radius = st.radius / st.decrease_factor; st.decrease_factor *= 2.0;
      it.cost = st.x_cost;
      st.num_unsuccessful++;
    } else {
      st.num_invalid = 0;
      it.step_norm = sqrt(st.cam_step_sq + E[6]);
      const double ptol = st.parameter_tolerance;
      if (it.step_norm <= ptol * (st.x_norm + ptol)) {
        finish(st, 0, kMsgParamTol, it.step_norm / (st.x_norm + ptol), ptol); return false;
      }
      const double cand_cost = isfinite(cost_e) ? cost_e : DBL_MAX;
      it.cost_change = st.x_cost - cand_cost;
      if (fabs(it.cost_change) <= st.function_tolerance * st.x_cost) {
        finish(st, 0, kMsgFuncTol, fabs(it.cost_change) / st.x_cost, st.function_tolerance); return false;
      }
      it.relative_decrease = it.cost_change / mcc;
      if (it.relative_decrease > st.min_relative_decrease) {
        // HandleSuccessfulStep: the candidate's blocks are already in buffer `buf`
        st.cur = buf; st.eval_buf = 1 - buf; st.took_step = 1;
        st.x_cost = cost_e; st.x_norm = sqrt(st.cam_cand_sq + E[7]);
        st.gmax = gmax_e; st.gnorm = gnorm_e;
        it.gradient_max_norm = gmax_e; it.gradient_norm = gnorm_e;
        it.step_is_successful = 1; it.cost = cost_e;
        const double q = 2.0 * it.relative_decrease - 1.0;
        st.radius = fmin(st.max_radius, st.radius / fmax(1.0 / 3.0, 1.0 - q * q * q));
        st.decrease_factor = 2.0;
        st.num_successful++;
      } else {
        // HandleUnsuccessfulStep
        it.cost = cand_cost;
        st.radius = st.radius / st.decrease_factor; st.decrease_factor *= 2.0;
        st.num_unsuccessful++;
      }
    }
  }
  // FinalizeIterationAndCheckIfMinimizerCanContinue
  it.trust_region_radius = st.radius;
  st.n_trace++;
  if (it.iteration >= st.max_num_iterations) finish(st, 1, kMsgMaxIter, it.iteration, 0);
  else if (st.gmax <= st.gradient_tolerance) finish(st, 0, kMsgGradTol, st.gmax, st.gradient_tolerance);
  else if (!(st.radius > st.min_radius)) finish(st, 0, kMsgMinRadius, st.radius, st.min_radius);
  else st.iteration++;
  return true;
}

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

// ---- reduced camera system: blocked (6x6) Cholesky + solve, whole CTA -----------------------
// sm: A [N][N+1] | Li [nf][36] | bb [N] | Us [F][27].  Writes the step into st and the candidate
// cameras.  K_B accumulates only the upper block triangle of S (frame pairs g <= f), so the lower
// triangle the factorisation works on is read transposed.  Threads are a 16x16 grid over the
// matrix (no integer divisions in the inner loops).
__device__ void solve_reduced(const LmParams& lp, LmState& st, double* sm, int F, const double* xs) {
  const int D = 6 * F, nf = st.n_free, N = 6 * nf;
  const int cur = st.cur, eb = st.eval_buf, tid = threadIdx.x, nthr = blockDim.x;
  const int ty = tid >> 4, tx = tid & 15;
  const double radius = st.radius;
  const int ld = N + 1;
  double* A = sm;
  double* Li = A + N * ld;
  double* bb = Li + nf * 36;
  double* Us = bb + N;            // [F][27] pose blocks of the accepted point
  __shared__ int s_fr[kMaxFrames];
  __shared__ int s_ok;
  __shared__ double s_v6[6];
  if (tid == 0) {
    s_ok = 1;
    for (int f = 0; f < F; ++f) if (st.free_index[f] >= 0) s_fr[st.free_index[f]] = f;
  }
  // the right-hand side accumulators are requested first (their frame mapping is applied after the barrier);
  // the accepted point's pose blocks come from the evaluation just adopted when there is one (no round trip
  // through the copy that was stored to global memory a moment ago)
  __shared__ double s_rawb[kMaxD];
  const double rb = tid < D ? __ldcg(lp.S + D * D + tid) : 0.0;
  for (int i = tid; i < F * kUStride; i += nthr) Us[i] = (xs && st.took_step) ? xs[i] : lp.Ucur[i];
  if (tid < D) s_rawb[tid] = rb;
  __syncthreads();
  // assemble S + Us + Dc² (lower triangle) and rhs + gs_c; a 3x3 batch of loads is in flight per thread
  for (int r0 = ty; r0 < N; r0 += 48) {
    for (int c0 = tx; c0 < N && c0 <= r0 + 32; c0 += 48) {
      double v[3][3];
#pragma unroll
      for (int u = 0; u < 3; ++u)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int r = r0 + 16 * u, c = c0 + 16 * q;
          v[u][q] = 0.0;
          if (r < N && c <= r) {
            const int f = s_fr[r / 6], a = r % 6, g = s_fr[c / 6], b = c % 6;
            v[u][q] = __ldcg(lp.S + (6 * g + b) * D + 6 * f + a);    // upper block (g <= f), transposed
          }
        }
#pragma unroll
      for (int u = 0; u < 3; ++u)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int r = r0 + 16 * u, c = c0 + 16 * q;
          if (r < N && c <= r) {
            double val = v[u][q];
            const int fi = r / 6, a = r - fi * 6, fj = c / 6, b = c - fj * 6;
            if (fi == fj) {
              const int f = s_fr[fi];
              const double uu = st.scale_c[6 * f + a] * usym6(Us + f * kUStride, a, b) * st.scale_c[6 * f + b];
              val += uu;
              if (a == b) val += fmin(fmax(uu, st.min_diag), st.max_diag) / radius;
            }
            A[r * ld + c] = val;
          }
        }
    }
  }
  for (int r = tid; r < N; r += nthr) {
    const int fi = r / 6, a = r - fi * 6, f = s_fr[fi];
    bb[r] = s_rawb[6 * f + a] + st.scale_c[6 * f + a] * Us[f * kUStride + 21 + a];
  }
  __syncthreads();
  for (int e = tid; e < D * D + D; e += nthr) lp.S[e] = 0.0;   // accumulators start from zero next time
  if (lp.dbg && tid == 0) lp.dbg[5] = gtime();

  for (int jb = 0; jb < nf; ++jb) {
    const int j0 = 6 * jb, m = N - j0 - 6;
    if (tid == 0) {
      // 6x6 Cholesky of the diagonal block, right-looking in registers (short dependency chain);
      // the reciprocal diagonal goes to Li[.][j][j]
      double L[6][6];
      bool ok = true;
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int k = 0; k <= i; ++k) L[i][k] = A[(j0 + i) * ld + j0 + k];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        double d = L[j][j];
        if (!(d > 0.0) || !isfinite(d)) { ok = false; d = 1.0; }
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
      if (!ok) s_ok = 0;
    }
    __syncthreads();
    if (tid < m) {
      // panel: solve L_p L_dd^T = A_p row by row (right-looking forward substitution)
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
    } else if (tid >= 224 && tid < 230) {
      // meanwhile: column c of M = L_dd^-1 (used by the triangular solves), off the critical path
      const int c = tid - 224;
      double M[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        if (i < c) { M[i] = 0.0; continue; }
        double sacc = (i == c) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k)
          if (k >= c && k < i) sacc -= A[(j0 + i) * ld + j0 + k] * M[k];
        M[i] = sacc * Li[jb * 36 + i * 7];
      }
#pragma unroll
      for (int i = 0; i < 6; ++i)
        if (i != c) Li[jb * 36 + i * 6 + c] = M[i];
    }
    __syncthreads();
    // trailing update (lower triangle), 16x16 thread grid; all loads of a 3x3 batch of entries are
    // issued before any store (the panel columns read and the trailing columns written never alias)
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
    __syncthreads();
  }
  if (lp.dbg && tid == 0) lp.dbg[6] = gtime();
  // forward substitution L y = b (block-wise with the inverted diagonal blocks)
  for (int jb = 0; jb < nf; ++jb) {
    const int j0 = 6 * jb;
    if (tid < 6) {
      double sacc = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) sacc += Li[jb * 36 + tid * 6 + k] * bb[j0 + k];
      s_v6[tid] = sacc;
    }
    __syncthreads();
    if (tid < 6) bb[j0 + tid] = s_v6[tid];
    const int i = j0 + 6 + tid;
    if (i < N) {
      double sacc = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) sacc += A[i * ld + j0 + k] * s_v6[k];
      bb[i] -= sacc;
    }
    __syncthreads();
  }
  // backward substitution L^T x = y
  for (int jb = nf - 1; jb >= 0; --jb) {
    const int j0 = 6 * jb;
    if (tid < 6) {
      double sacc = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) sacc += Li[jb * 36 + k * 6 + tid] * bb[j0 + k];   // (L_dd^-T)[tid][k]
      s_v6[tid] = sacc;
    }
    __syncthreads();
    if (tid < 6) bb[j0 + tid] = s_v6[tid];
    if (tid < j0) {
      double sacc = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) sacc += A[(j0 + k) * ld + tid] * s_v6[k];
      bb[tid] -= sacc;
    }
    __syncthreads();
  }
  if (lp.dbg && tid == 0) lp.dbg[7] = gtime();
  // step (scaled space) = -y ; candidate cameras ; camera part of the model cost change
  if (tid < 32) {
    double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
    bool bad = false;
    for (int i = tid; i < D; i += 32) {
      const int f = i / 6, a = i - f * 6, fi = st.free_index[f];
      const double x = lp.cams[((size_t)cur * F + f) * 6 + a];
      double step = 0.0, cand = x;
      if (fi >= 0) {
        step = -bb[6 * fi + a];
        if (!isfinite(step)) bad = true;
        cand = x + step * st.scale_c[i];
        const double gsc = st.scale_c[i] * Us[f * kUStride + 21 + a];
        double hs = 0.0;  // (Us step)_a
#pragma unroll
        for (int b = 0; b < 6; ++b)
          hs += st.scale_c[6 * f + a] * usym6(Us + f * kUStride, a, b) * st.scale_c[6 * f + b] * (-bb[6 * fi + b]);
        r0 += step * gsc; r1 += step * hs; r2 += (x - cand) * (x - cand); r3 += cand * cand;
      }
      st.step_c[i] = step;
      lp.cams[((size_t)eb * F + f) * 6 + a] = cand;
    }
#pragma unroll
    for (int mk = 16; mk > 0; mk >>= 1) {
      r0 += __shfl_xor_sync(0xffffffffu, r0, mk); r1 += __shfl_xor_sync(0xffffffffu, r1, mk);
      r2 += __shfl_xor_sync(0xffffffffu, r2, mk); r3 += __shfl_xor_sync(0xffffffffu, r3, mk);
    }
    const bool any_bad = __any_sync(0xffffffffu, bad);
    if (tid == 0) {
      st.cam_sg = r0; st.cam_sHs = r1; st.cam_step_sq = r2; st.cam_cand_sq = r3;
      st.step_valid = (s_ok && !any_bad) ? 1 : 0;
    }
  }
  __syncthreads();
}

// ---- Schur elimination, fast path: <= 32 pairs of optimised frames (<= 7 free cameras) --------
// Warp per point.  Vs + D_p² = L Lᵀ (3x3 Cholesky); Z_a = Ws_a L^-T (6x3) for each observing free
// frame, so the point's contribution is the symmetric rank-3 update  S_ab -= Z_a Z_bᵀ  and
// rhs_a -= Z_a (L^-1 gs).  Lane l owns the 6x6 block of frame pair l in REGISTERS across all of
// the warp's points (108 FMA per 36 shared-memory loads); warps are merged through shared memory
// and the CTA adds its partial to the global accumulators with one fp64 atomic per entry.
// sm: per warp Z [D][3] + rh [D] | per CTA S_cta [32*36 + D]
__shared__ unsigned long long s_tdbg[4];   // PBA_DEBUG_TIMELINE: elimination sub-phases of this CTA
__device__ void schur_pairs(const LmParams& lp, const LmState& st, double* sm) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int F = lp.n_frames, D = 6 * F, n = lp.n_points, cur = st.cur, nf = st.n_free;
  const int npairs = nf * (nf + 1) / 2;
  const double radius = st.radius, dmin = st.min_diag, dmax = st.max_diag;
  const bool first = (st.iteration == 1);
  double* Zw = sm + warp * (D * 4);                 // [D][3] + rh [D]
  double* rhw = Zw + D * 3;
  double* Scta = sm + 4 * (32 * 36 + 64);          // after the merge buffers (which alias Z/rh)
  __shared__ int s_fr2[kMaxFrames];
  if (tid == 0)
    for (int f = 0; f < F; ++f) if (st.free_index[f] >= 0) s_fr2[st.free_index[f]] = f;
  for (int i = tid; i < 32 * 36 + D; i += blockDim.x) Scta[i] = 0.0;
  // lane's pair (pa <= pb) in free-frame indices
  int pa = 0, pb = 0;
  if (lane < npairs) {
    int rem = lane;
    while (rem >= nf - pa) { rem -= nf - pa; ++pa; }
    pb = pa + rem;
  }
  __syncthreads();
  const double* Vb = lp.V + (size_t)cur * n * 6;
  const double* gb = lp.gp + (size_t)cur * n * 3;
  const double* Wb = lp.W + (size_t)cur * lp.nnz * 18;
  double acc[36];
#pragma unroll
  for (int e = 0; e < 36; ++e) acc[e] = 0.0;
  double racc0 = 0.0, racc1 = 0.0;

  // Per-point inputs are loaded one point ahead: the loads of point p+stride are issued right
  // after point p's phase A has consumed its registers, so their latency hides behind phase B.
  const int pstride = gridDim.x * (kSchurThreads / 32);
  double V[6], g0 = 0, g1 = 0, g2 = 0, sp0 = 1, sp1 = 1, sp2 = 1, w[2][3];
  int rf[2], ra[2];
  // Two-level prefetch: the CSR header of a point is requested one iteration before its data (whose
  // addresses depend on it), and nothing here CONSUMES a loaded value, so neither call blocks.
  int hdr_o0 = 0, hdr_n = 0;
  auto load_header = [&](int q) { hdr_o0 = __ldg(lp.obs_off + q); hdr_n = __ldg(lp.obs_off + q + 1) - hdr_o0; };
  auto load_point = [&](int q, int o0, int nobs) {
#pragma unroll
    for (int k = 0; k < 6; ++k) V[k] = __ldg(Vb + (size_t)q * 6 + k);
    g0 = __ldg(gb + (size_t)q * 3); g1 = __ldg(gb + (size_t)q * 3 + 1); g2 = __ldg(gb + (size_t)q * 3 + 2);
    if (!first) { sp0 = lp.scale_p[(size_t)q * 3]; sp1 = lp.scale_p[(size_t)q * 3 + 1]; sp2 = lp.scale_p[(size_t)q * 3 + 2]; }
    // rows (observation i, pose parameter a): lane + 32k  (nobs <= 8 on this path -> 2 rounds)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int row = lane + 32 * k;
      rf[k] = -1; ra[k] = 0;
      w[k][0] = w[k][1] = w[k][2] = 0.0;
      if (row < nobs * 6) {
        const int i = row / 6;
        ra[k] = row - 6 * i;
        rf[k] = __ldg(lp.obs_frame + o0 + i);           // raw frame; free/fixed is resolved at use
        const double* wp = Wb + (size_t)(o0 + i) * 18 + ra[k] * 3;
        w[k][0] = __ldg(wp); w[k][1] = __ldg(wp + 1); w[k][2] = __ldg(wp + 2);
      }
    }
  };
  const int p_first = blockIdx.x * (kSchurThreads / 32) + warp;
  if (lp.dbg && tid == 0) s_tdbg[0] = gtime();
  if (p_first < n) {
    load_header(p_first);
    load_point(p_first, hdr_o0, hdr_n);
    if (p_first + pstride < n) load_header(p_first + pstride);
  }
  for (int p = p_first; p < n; p += pstride) {
    if (first) {
      sp0 = st.jacobi_scaling ? 1.0 / (1.0 + sqrt(V[0])) : 1.0;
      sp1 = st.jacobi_scaling ? 1.0 / (1.0 + sqrt(V[3])) : 1.0;
      sp2 = st.jacobi_scaling ? 1.0 / (1.0 + sqrt(V[5])) : 1.0;
      if (lane == 0) { lp.scale_p[(size_t)p * 3] = sp0; lp.scale_p[(size_t)p * 3 + 1] = sp1; lp.scale_p[(size_t)p * 3 + 2] = sp2; }
    }
    double a00 = sp0 * V[0] * sp0, a01 = sp0 * V[1] * sp1, a02 = sp0 * V[2] * sp2;
    double a11 = sp1 * V[3] * sp1, a12 = sp1 * V[4] * sp2, a22 = sp2 * V[5] * sp2;
    a00 += fmin(fmax(a00, dmin), dmax) / radius;
    a11 += fmin(fmax(a11, dmin), dmax) / radius;
    a22 += fmin(fmax(a22, dmin), dmax) / radius;
    // 3x3 Cholesky and its inverse
    const double il00 = rsqrt(a00), l10 = a01 * il00, l20 = a02 * il00;
    const double il11 = rsqrt(a11 - l10 * l10), l21 = (a12 - l20 * l10) * il11;
    const double il22 = rsqrt(a22 - l20 * l20 - l21 * l21);
    const double m10 = -l10 * il00 * il11, m20 = -(l20 * il00 + l21 * m10) * il22, m21 = -l21 * il11 * il22;
    if (lane == 0) {
      double* vi = lp.Vinv + (size_t)p * 6;   // (Vs + D²)^-1 = M^T M, used by K_A's back-substitution
      vi[0] = il00 * il00 + m10 * m10 + m20 * m20; vi[1] = m10 * il11 + m20 * m21; vi[2] = m20 * il22;
      vi[3] = il11 * il11 + m21 * m21; vi[4] = m21 * il22; vi[5] = il22 * il22;
    }
    const double gs0 = sp0 * g0, gs1 = sp1 * g1, gs2 = sp2 * g2;
    const double zg0 = gs0 * il00, zg1 = (gs1 - l10 * zg0) * il11, zg2 = (gs2 - l20 * zg0 - l21 * zg1) * il22;
    unsigned mask = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int fi = rf[k] >= 0 ? st.free_index[rf[k]] : -1;
      if (fi >= 0) {
        const double sc = st.scale_c[rf[k] * 6 + ra[k]];
        const double z0 = sc * w[k][0] * sp0 * il00;
        const double z1 = (sc * w[k][1] * sp1 - z0 * l10) * il11;
        const double z2 = (sc * w[k][2] * sp2 - z0 * l20 - z1 * l21) * il22;
        double* zr = Zw + (6 * fi + ra[k]) * 3;
        zr[0] = z0; zr[1] = z1; zr[2] = z2;
        rhw[6 * fi + ra[k]] = -(z0 * zg0 + z1 * zg1 + z2 * zg2);
        mask |= 1u << fi;
      }
    }
    mask = __reduce_or_sync(0xffffffffu, mask);
    __syncwarp();
    if (p + pstride < n) {                            // prefetch (registers of point p are dead now)
      load_point(p + pstride, hdr_o0, hdr_n);
      if (p + 2 * pstride < n) load_header(p + 2 * pstride);
    }
    if (lane < npairs && ((mask >> pa) & 1u) && ((mask >> pb) & 1u)) {
      const double* za = Zw + 18 * pa;
      const double* zb = Zw + 18 * pb;
      double zbv[18];
#pragma unroll
      for (int e = 0; e < 18; ++e) zbv[e] = zb[e];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const double x0 = za[3 * i], x1 = za[3 * i + 1], x2 = za[3 * i + 2];
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[i * 6 + j] -= x0 * zbv[3 * j] + x1 * zbv[3 * j + 1] + x2 * zbv[3 * j + 2];
      }
    }
    if (lane < 6 * nf && ((mask >> (lane / 6)) & 1u)) racc0 += rhw[lane];
    if (lane + 32 < 6 * nf && ((mask >> ((lane + 32) / 6)) & 1u)) racc1 += rhw[lane + 32];
    __syncwarp();
  }
  // merge the warps pairwise through shared memory (fixed tree: deterministic inside the CTA), then
  // one atomic per entry.  Buffers: warp w writes into slot w of Smerge [warps/2][32*36 + 64].
  double* Smerge = sm;   // the Z / rh staging area is dead now
  __syncthreads();
  if (lp.dbg && tid == 0) s_tdbg[1] = gtime();
  for (int half = (kSchurThreads / 32) / 2; half >= 1; half >>= 1) {
    if (warp >= half && warp < 2 * half) {
      double* dst = Smerge + (size_t)(warp - half) * (32 * 36 + 64);
#pragma unroll
      for (int e = 0; e < 36; ++e) dst[e * 32 + lane] = acc[e];
      dst[32 * 36 + lane] = racc0; dst[32 * 36 + 32 + lane] = racc1;
    }
    __syncthreads();
    if (warp < half) {
      const double* src = Smerge + (size_t)warp * (32 * 36 + 64);
#pragma unroll
      for (int e = 0; e < 36; ++e) acc[e] += src[e * 32 + lane];
      racc0 += src[32 * 36 + lane]; racc1 += src[32 * 36 + 32 + lane];
    }
    __syncthreads();
  }
  if (lp.dbg && tid == 0) s_tdbg[2] = gtime();
  if (warp == 0) {
    if (lane < npairs) {
#pragma unroll
      for (int e = 0; e < 36; ++e) Scta[lane * 36 + e] = acc[e];
    }
    if (lane < 6 * nf) Scta[32 * 36 + lane] = racc0;
    if (lane + 32 < 6 * nf) Scta[32 * 36 + lane + 32] = racc1;
  }
  __syncthreads();
  for (int t = tid; t < npairs * 36; t += blockDim.x) {
    const int pr = t / 36, e = t - pr * 36, i = e / 6, j = e - i * 6;
    int qa = 0, rem = pr;
    while (rem >= nf - qa) { rem -= nf - qa; ++qa; }
    const int fa = s_fr2[qa], fb = s_fr2[qa + rem];
    const double v = Scta[t];
    if (v != 0.0) atomicAdd(lp.S + (6 * fa + i) * D + 6 * fb + j, v);
  }
  if (tid < 6 * nf) {
    const double v = Scta[32 * 36 + tid];
    if (v != 0.0) atomicAdd(lp.S + D * D + 6 * s_fr2[tid / 6] + tid % 6, v);
  }
}

// Tail of an LM iteration (one CTA): adopt the candidate's pose blocks if the step was taken, solve
// the reduced camera system, re-zero the accumulators K_A fills next, publish the new state.
__device__ void finish_iteration(const LmParams& lp, LmState& st, double* sm, int F, const double* xs) {
  const int tid = threadIdx.x;
  if (st.took_step)
    for (int i = tid; i < F * kUStride; i += blockDim.x) lp.Ucur[i] = xs ? xs[i] : __ldcg(lp.Xacc + i);
  __syncthreads();
  solve_reduced(lp, st, sm, F, xs);
  if (lp.dbg && tid == 0) lp.dbg[3] = gtime();
  for (int i = tid; i < F * kUStride + kEacc + kMaxRanks; i += blockDim.x) lp.Xacc[i] = 0.0;
  if (tid == 0) *lp.ticket = 0u;
  __syncthreads();
  for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
    reinterpret_cast<int*>(lp.st_out)[i] = reinterpret_cast<const int*>(&st)[i];
  if (lp.dbg && tid == 0) lp.dbg[4] = gtime();
}

// MODE 0: pairs fast path (<= 7 optimised cameras, <= 8 frames); MODE k>0: generic 3x3-tile path, k tiles/thread
template <int MODE>
__global__ void __launch_bounds__(kSchurThreads) k_schur_solve(const LmParams lp) {
  constexpr int TPT = MODE > 0 ? MODE : 1;
  __shared__ LmState s_st;
  __shared__ IterSummary s_it;
  __shared__ int s_push, s_last;
  __shared__ unsigned s_mask[kSchurChunk];
  __shared__ unsigned char s_pair[kMaxFrames * (kMaxFrames + 1) / 2][2];   // upper block pairs (g <= f)
  __shared__ double s_xs[kMaxFrames * kUStride + kEacc + kMaxRanks];        // the evaluation's pose blocks + scalars (summed over ranks)
  __shared__ int s_xok;
  extern __shared__ double sm[];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int F = lp.n_frames, D = 6 * F, n = lp.n_points;
  const unsigned long long t_start = lp.dbg ? gtime() : 0ull;

  // ---- (D) decision, redundantly per CTA ------------------------------------------------
  // Everything the decision reads is requested in ONE round of loads: the state, the evaluation's
  // accumulators (single-GPU / NCCL path) and both buffers of the cameras (which one is the candidate is
  // only known once the state has arrived) - three dependent L2 round trips otherwise.
  const int xn = F * kUStride + kEacc + kMaxRanks;
  const bool xmode = lp.xc.n_ranks > 1;
  constexpr int kXPre = (kMaxFrames * kUStride + kEacc + kMaxRanks + kSchurThreads - 1) / kSchurThreads;
  double x_pre[kXPre];
  if (!xmode) {
#pragma unroll
    for (int k = 0; k < kXPre; ++k) x_pre[k] = (tid + k * kSchurThreads < xn) ? __ldcg(lp.Xacc + tid + k * kSchurThreads) : 0.0;
  }
  double cam_pre[2][(kMaxD + 31) / 32];
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < (kMaxD + 31) / 32; ++k) {
      const int i = lane + 32 * k;
      cam_pre[0][k] = i < F * 6 ? lp.cams[i] : 0.0;
      cam_pre[1][k] = i < F * 6 ? lp.cams[(size_t)F * 6 + i] : 0.0;
    }
  }
  for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
    reinterpret_cast<int*>(&s_st)[i] = reinterpret_cast<const int*>(lp.st_in)[i];
  if (!xmode) {
#pragma unroll
    for (int k = 0; k < kXPre; ++k)
      if (tid + k * kSchurThreads < xn) s_xs[tid + k * kSchurThreads] = x_pre[k];
  }
  __syncthreads();
  if (s_st.done) {
    if (blockIdx.x == 0) {
      for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
        reinterpret_cast<int*>(lp.st_out)[i] = reinterpret_cast<const int*>(&s_st)[i];
      if (lp.cond && tid == 0) cudaGraphSetConditional(lp.cond, 0u);   // leave the while node of the solve graph
    }
    return;
  }
  // the evaluation's accumulators: local (one GPU / NCCL path: already all-reduced) or the sum of every
  // rank's slot in rank order (peer-memory exchange: wait for the flags first)
  if (xmode) {
    if (tid == 0) s_xok = 1;
    __syncthreads();
    const ulonglong2* xb = lp.xc.xa[lp.xc.rank] + (size_t)(s_st.xepoch & 1ull) * lp.xc.n_ranks * lp.xc.xa_n;
    for (int i = tid; i < xn; i += blockDim.x) {
      double acc = 0.0;
      for (int q = 0; q < lp.xc.n_ranks; ++q) {
        double v;
        if (!ll_load(xb + (size_t)q * lp.xc.xa_n + i, s_st.xepoch, v)) s_xok = 0;
        acc += v;
      }
      s_xs[i] = acc;
    }
    __syncthreads();
    if (!s_xok) {   // a peer never arrived: fail the solve instead of hanging the GPU
      if (blockIdx.x == 0) {
        if (tid == 0) { finish(s_st, 2, kMsgXchgTimeout, (double)s_st.xepoch, 0.0); *lp.xc.error = 1; }
        __syncthreads();
        for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
          reinterpret_cast<int*>(lp.st_out)[i] = reinterpret_cast<const int*>(&s_st)[i];
        if (lp.cond && tid == 0) cudaGraphSetConditional(lp.cond, 0u);
      }
      return;
    }
    __syncthreads();
  }
  if (warp == 0) {
    const int buf = s_st.eval_buf;
    double gm = 0.0, g2 = 0.0, csq = 0.0;
#pragma unroll
    for (int k = 0; k < (kMaxD + 31) / 32; ++k) {
      const int i = lane + 32 * k;
      if (i < F * 6) {
        const int f = i / 6, a = i - f * 6;
        if (s_st.free_index[f] >= 0) {
          const double g = s_xs[f * kUStride + 21 + a];
          gm = fmax(gm, fabs(g)); g2 += g * g;
          const double c = buf ? cam_pre[1][k] : cam_pre[0][k];
          csq += c * c;
        }
      }
    }
    double e = lane < kEacc ? s_xs[F * kUStride + lane] : 0.0;
    double gpm = lane < kMaxRanks ? s_xs[F * kUStride + kEacc + lane] : 0.0;   // per-rank max|g_p|
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, m));
      g2 += __shfl_xor_sync(0xffffffffu, g2, m);
      csq += __shfl_xor_sync(0xffffffffu, csq, m);
      gpm = fmax(gpm, __shfl_xor_sync(0xffffffffu, gpm, m));
    }
    double E[kEacc];
#pragma unroll
    for (int k = 0; k < kEacc; ++k) E[k] = __shfl_sync(0xffffffffu, e, k);
    E[2] = gpm;
    if (lane == 0) s_push = decide(s_st, E, gm, g2, csq, s_xs, F, s_it) ? 1 : 0;
  }
  __syncthreads();
  if (blockIdx.x == 0 && tid == 0 && s_push) lp.trace[s_st.n_trace - 1] = s_it;
  if (s_st.done) {
    if (blockIdx.x == 0) {
      for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
        reinterpret_cast<int*>(lp.st_out)[i] = reinterpret_cast<const int*>(&s_st)[i];
      if (lp.cond && tid == 0) cudaGraphSetConditional(lp.cond, 0u);   // leave the while node of the solve graph
    }
    return;
  }

  const unsigned long long t_dec = lp.dbg ? gtime() : 0ull;
  // ---- (E) eliminate the point blocks -----------------------------------------------------
  if (MODE == 0) {
    schur_pairs(lp, s_st, sm);
  } else {
  const int cur = s_st.cur;
  const double radius = s_st.radius, dmin = s_st.min_diag, dmax = s_st.max_diag;
  const bool first = (s_st.iteration == 1);
  double* Yf = sm;                                  // [chunk][D*3]
  double* Wf = Yf + kSchurChunk * D * 3;            // [chunk][D*3]
  double* rh = Wf + kSchurChunk * D * 3;            // [chunk][D]
  const double* Vb = lp.V + (size_t)cur * n * 6;
  const double* gb = lp.gp + (size_t)cur * n * 3;
  const double* Wb = lp.W + (size_t)cur * lp.nnz * 18;

  // each thread owns up to TPT 3x3 tiles of the UPPER block triangle of S: tile = pair*4 + (sr,sc)
  const int npairs = F * (F + 1) / 2, ntiles = npairs * 4;
  if (tid < npairs) {
    int g = 0, rem = tid;
    while (rem >= F - g) { rem -= F - g; ++g; }
    s_pair[tid][0] = (unsigned char)g; s_pair[tid][1] = (unsigned char)(g + rem);
  }
  __syncthreads();
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
      const int o0 = lp.obs_off[p], nobs = lp.obs_off[p + 1] - o0;
      const int my_f = lane < nobs ? lp.obs_frame[o0 + lane] : 0;
      // prefetch the point's W blocks (18 lanes x up to 16 observations in flight)
      double wreg[kMaxFrames];
#pragma unroll
      for (int i = 0; i < kMaxFrames; ++i) wreg[i] = (i < nobs && lane < 18) ? __ldg(Wb + (size_t)(o0 + i) * 18 + lane) : 0.0;
      double sp[3], Vs[6], Vi[6], gs[3];
      double V[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) V[k] = __ldg(Vb + (size_t)p * 6 + k);
      const double gp0 = __ldg(gb + (size_t)p * 3), gp1 = __ldg(gb + (size_t)p * 3 + 1), gp2 = __ldg(gb + (size_t)p * 3 + 2);
      if (first) {
        sp[0] = s_st.jacobi_scaling ? 1.0 / (1.0 + sqrt(V[0])) : 1.0;
        sp[1] = s_st.jacobi_scaling ? 1.0 / (1.0 + sqrt(V[3])) : 1.0;
        sp[2] = s_st.jacobi_scaling ? 1.0 / (1.0 + sqrt(V[5])) : 1.0;
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
      gs[0] = sp[0] * gp0; gs[1] = sp[1] * gp1; gs[2] = sp[2] * gp2;
      const int a = lane / 3, b = lane - a * 3;
      const double spb = b == 0 ? sp[0] : (b == 1 ? sp[1] : sp[2]);
      const double vi0 = sym3(Vi, 0, b), vi1 = sym3(Vi, 1, b), vi2 = sym3(Vi, 2, b);
      const double gsb = b == 0 ? gs[0] : (b == 1 ? gs[1] : gs[2]);
#pragma unroll
      for (int i = 0; i < kMaxFrames; ++i) {
        if (i < nobs) {
          const int f = __shfl_sync(0xffffffffu, my_f, i);
          if (s_st.free_index[f] >= 0) {
            mask |= 1u << f;
            double ws = 0.0;
            if (lane < 18) {
              ws = s_st.scale_c[f * 6 + a] * wreg[i] * spb;
              Wf[(warp * D + 6 * f) * 3 + lane] = ws;
            }
            // Y[a][b] = sum_q Ws[a][q] Vi[q][b]: gather the row's three entries with shuffles
            const int base = lane < 18 ? a * 3 : 0;
            const double w0 = __shfl_sync(0xffffffffu, ws, base);
            const double w1 = __shfl_sync(0xffffffffu, ws, base + 1);
            const double w2 = __shfl_sync(0xffffffffu, ws, base + 2);
            const double y = w0 * vi0 + w1 * vi1 + w2 * vi2;
            // rhs[a] -= sum_b Y[a][b] gs[b]: combine the three lanes of a row
            const double t = y * gsb;
            const double t1 = __shfl_down_sync(0xffffffffu, t, 1);
            const double t2 = __shfl_down_sync(0xffffffffu, t, 2);
            if (lane < 18) {
              Yf[(warp * D + 6 * f) * 3 + lane] = y;
              if (b == 0) rh[warp * D + 6 * f + a] = -(t + t1 + t2);
            }
          }
        }
      }
    }
    if (lane == 0) s_mask[warp] = mask;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TPT; ++k) {
      const int t = tid + k * kSchurThreads;
      if (t < ntiles) {
        const int fr = s_pair[t >> 2][0], fc = s_pair[t >> 2][1];
        const int tr = 2 * fr + ((t >> 1) & 1), tc = 2 * fc + (t & 1);
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
    if (tid < D) {
      const int f = tid / 6;
#pragma unroll
      for (int w = 0; w < kSchurChunk; ++w)
        if ((s_mask[w] >> f) & 1u) racc += rh[w * D + tid];
    }
    __syncthreads();
  }
  // one fp64 atomic per non-zero entry per CTA
#pragma unroll
  for (int k = 0; k < TPT; ++k) {
    const int t = tid + k * kSchurThreads;
    if (t < ntiles) {
      const int tr = 2 * s_pair[t >> 2][0] + ((t >> 1) & 1), tc = 2 * s_pair[t >> 2][1] + (t & 1);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (acc[k][i * 3 + j] != 0.0) atomicAdd(lp.S + (3 * tr + i) * D + 3 * tc + j, acc[k][i * 3 + j]);
    }
  }
  if (tid < D && racc != 0.0) atomicAdd(lp.S + D * D + tid, racc);
  }  // MODE != 0

  // ---- (S) the last CTA solves the reduced camera system -------------------------------------
  // bar.sync orders the CTA's atomics before thread 0's cumulative gpu-scope fence + ticket
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(lp.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __threadfence();
  }
  __syncthreads();
  if (lp.dbg && tid == 0 && s_last) {
    lp.dbg[0] = t_start; lp.dbg[1] = t_dec; lp.dbg[2] = gtime();
    if (MODE == 0) { lp.dbg[12] = s_tdbg[0]; lp.dbg[13] = s_tdbg[1]; lp.dbg[14] = s_tdbg[2]; }
  }
  if (!s_last) return;
  if (lp.split) {
    // multi-GPU: the reduced system is summed across ranks first; k_solve_only finishes the iteration
    if (tid == 0) *lp.ticket = 0u;
    for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
      reinterpret_cast<int*>(lp.st_out)[i] = reinterpret_cast<const int*>(&s_st)[i];
    return;
  }
  if (xmode) {
    // publish this rank's reduced-system contribution (upper block triangle + rhs) to every rank as LL
    // cells, then sum everybody's in rank order straight out of the cells (polling replaces flag + fence)
    const unsigned long long e = s_st.xepoch;
    const int sn = D * D + D;
    const size_t off = ((size_t)(e & 1ull) * lp.xc.n_ranks + lp.xc.rank) * lp.xc.s_n;
    constexpr int kB = 4;                    // elements per thread in flight
    for (int i0 = tid; i0 < sn; i0 += kB * blockDim.x) {
      double v[kB];
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        const int i = i0 + u * blockDim.x;
        v[u] = (i < sn) ? __ldcg(lp.S + i) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        const int i = i0 + u * blockDim.x;
        const bool need = i < sn && (i >= D * D || (i / D) / 6 <= (i % D) / 6);
        if (need)
          for (int q = 0; q < lp.xc.n_ranks; ++q) ll_store(lp.xc.s[q] + off + i, v[u], e);
      }
    }
    if (lp.dbg && tid == 0) lp.dbg[8] = gtime();
    const ulonglong2* sb = lp.xc.s[lp.xc.rank] + (size_t)(e & 1ull) * lp.xc.n_ranks * lp.xc.s_n;
    for (int i0 = tid; i0 < sn; i0 += kB * blockDim.x) {
      double acc[kB];
      bool need[kB];
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        const int i = i0 + u * blockDim.x;
        need[u] = i < sn && (i >= D * D || (i / D) / 6 <= (i % D) / 6);
        acc[u] = 0.0;
      }
      for (int q = 0; q < lp.xc.n_ranks; ++q) {
        double v[kB];
        bool got[kB];
#pragma unroll
        for (int u = 0; u < kB; ++u) {     // first try: all loads in flight
          const int i = i0 + u * blockDim.x;
          v[u] = 0.0;
          got[u] = !need[u] || ll_try_load(sb + (size_t)q * lp.xc.s_n + i, e, v[u]);
        }
#pragma unroll
        for (int u = 0; u < kB; ++u) {
          const int i = i0 + u * blockDim.x;
          if (!got[u] && !ll_load(sb + (size_t)q * lp.xc.s_n + i, e, v[u])) s_xok = 0;
          acc[u] += v[u];
        }
      }
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        const int i = i0 + u * blockDim.x;
        if (need[u]) lp.S[i] = acc[u];
      }
    }
    __threadfence();
    __syncthreads();
    if (lp.dbg && tid == 0) { lp.dbg[9] = lp.dbg[8]; lp.dbg[10] = lp.dbg[8]; }
    if (!s_xok) {
      if (tid == 0) { finish(s_st, 2, kMsgXchgTimeout, (double)e, 1.0); *lp.xc.error = 1; *lp.ticket = 0u; }
      __syncthreads();
      for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
        reinterpret_cast<int*>(lp.st_out)[i] = reinterpret_cast<const int*>(&s_st)[i];
      if (lp.cond && tid == 0) cudaGraphSetConditional(lp.cond, 0u);
      return;
    }
    if (tid == 0) s_st.xepoch = e + 1;
    __syncthreads();
    if (lp.dbg && tid == 0) lp.dbg[11] = gtime();
  }
  finish_iteration(lp, s_st, sm, F, s_xs);
}

// split mode: one CTA, after the all-reduce of S
__global__ void __launch_bounds__(kSchurThreads) k_solve_only(const LmParams lp) {
  __shared__ LmState s_st;
  extern __shared__ double sm[];
  const int tid = threadIdx.x;
  for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
    reinterpret_cast<int*>(&s_st)[i] = reinterpret_cast<const int*>(lp.st_out)[i];
  __syncthreads();
  if (s_st.done) return;
  finish_iteration(lp, s_st, sm, lp.n_frames, nullptr);
}

// Device-side barrier across the ranks of a window (start of a solve): raise my flag in every rank's
// buffer, wait for everybody's.  Keeps the ranks' timelines aligned so that a rank that was called a
// little earlier does not spend its first iteration waiting inside the LM loop.
__global__ void k_rendezvous(const Xchg xc, unsigned long long epoch) {
  const int q = threadIdx.x;
  if (q < xc.n_ranks) {
    st_release_sys(xc.fr[q] + xc.rank, epoch);
    if (!xchg_wait(xc.fr[xc.rank] + q, epoch)) *xc.error = 1;
  }
}

cudaError_t launch_rendezvous(const Xchg& xc, unsigned long long epoch, cudaStream_t stream) {
  k_rendezvous<<<1, 32, 0, stream>>>(xc, epoch);
  return cudaGetLastError();
}

// ---- launchers ----------------------------------------------------------------------------
int schur_grid(int n_points, int sm_count) {
  const int chunks = (n_points + kSchurChunk - 1) / kSchurChunk;
  const int cap = sm_count;
  return chunks < cap ? (chunks > 0 ? chunks : 1) : cap;
}

template <int MODE>
static cudaError_t launch_mode(const LmParams& lp, int grid, size_t smem, cudaStream_t s) {
  static bool cfg[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!cfg[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_schur_solve<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return e;
    cfg[dev & 63] = true;
  }
  k_schur_solve<MODE><<<grid, kSchurThreads, smem, s>>>(lp);
  return cudaGetLastError();
}

cudaError_t launch_schur_solve(const LmParams& lp, int grid, int n_free, cudaStream_t s) {
  const int F = lp.n_frames, D = 6 * F, N = D;
  const size_t solve_b = sizeof(double) * ((size_t)N * (N + 1) + (size_t)F * 36 + N + (size_t)F * kUStride);
  if (F <= 8 && n_free <= 7) {
    const size_t stage = (size_t)(kSchurThreads / 32) * D * 4, merge = 4 * (32 * 36 + 64);
    const size_t pairs_b = sizeof(double) * ((stage > merge ? stage : merge) + 32 * 36 + D);
    return launch_mode<0>(lp, grid, pairs_b > solve_b ? pairs_b : solve_b, s);
  }
  const size_t schur_b = sizeof(double) * (size_t)kSchurChunk * (D * 3 * 2 + D);
  const size_t smem = schur_b > solve_b ? schur_b : solve_b;
  const int need = (2 * F * (F + 1) + kSchurThreads - 1) / kSchurThreads;   // tiles of the upper block triangle
  if (need <= 1) return launch_mode<1>(lp, grid, smem, s);
  if (need <= 2) return launch_mode<2>(lp, grid, smem, s);
  return launch_mode<3>(lp, grid, smem, s);
}

cudaError_t launch_solve_only(const LmParams& lp, cudaStream_t s) {
  const int F = lp.n_frames, N = 6 * F;
  const size_t solve_b = sizeof(double) * ((size_t)N * (N + 1) + (size_t)F * 36 + N + (size_t)F * kUStride);
  static bool cfg[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!cfg[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_solve_only, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return e;
    cfg[dev & 63] = true;
  }
  k_solve_only<<<1, kSchurThreads, solve_b, s>>>(lp);
  return cudaGetLastError();
}

}  // namespace pba
