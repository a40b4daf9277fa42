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
//  (E) elimination.  One lane per OBSERVATION (8 or 16 lanes per point, 4 or 2 points per warp): the
//      lanes of a point factor Vs + D_p² = L Lᵀ (3x3) redundantly, each forms its own Z_a = Ws_a L^-T
//      (6x3) and parks it as rows of a CTA-wide matrix Zt [6 n_free + 1][3 x points] in shared memory
//      (the last row holds L^-1 gs), so that the whole batch's contribution to the reduced system is
//      ONE symmetric product  P = Zt Ztᵀ  — which runs on the fp64 TENSOR pipe (mma.sync m8n8k4.f64,
//      SASS DMMA; the accumulator fragments stay in registers across all of the CTA's points); one
//      fp64 atomic per entry per CTA adds the upper triangle of P to the global accumulator;
//  (S) the last CTA to finish (ticket) assembles S = U_s + D_c² - P with the right-hand side as an
//      extra ROW (so the forward substitution falls out of the factorisation), factors it with a
//      blocked (6x6) Cholesky in fp64 in shared memory — every panel thread factors the diagonal block
//      redundantly in registers, two barriers per block step —, back-substitutes on one warp and
//      writes the camera step, the candidate cameras and the next LmState.  The back-substitution
//      of the points is fused into K_A.
//
// LM state ping-pongs between two LmState structs (st_in is read-only during the kernel).

#include "pba_device.cuh"

#include <cfloat>
#include <cmath>

namespace pba {

__device__ __forceinline__ unsigned long long gtime() {
#ifdef PBA_SOLVE_UBENCH
  return (unsigned long long)clock64();   // scripts/ubench/solve_bench.cu: cycles instead of nanoseconds
#define UB_STAMP(slot) do { if (lp.dbg && jb == 1 && tid == 0) lp.dbg[slot] = (unsigned long long)clock64(); } while (0)
#define UB_ARRIVE(k) do { if (lp.dbg && jb == 1 && (tid & 31) == 0) lp.dbg[32 + 8 * (k) + (tid >> 5)] = (unsigned long long)clock64(); } while (0)
#else
#define UB_STAMP(slot) do {} while (0)
#define UB_ARRIVE(k) do {} while (0)
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
#endif
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
    // (the Jacobi scaling of the pose columns, 1 / (1 + sqrt(diag U)), is formed by the caller on parallel lanes)
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


// fp64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  Lane l holds A[l/4][l%4], B[l%4][l/4] and
// C[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// 1/sqrt(d) for a positive, normal d without the library routine's slow-path branch (which splits the basic block
// and keeps the scheduler from overlapping the refinement with independent work): hardware approximation
// (MUFU.RSQ64H, ~2^-22) + one cubic-convergence correction  y0 (1 + e/2 + 3e²/8),  e = 1 - d y0²  (error term
// 5e³/16 < 2^-66).  Zero / negative / non-finite pivots give a non-finite or garbage result; the caller checks d.
__device__ __forceinline__ double rsqrt_pos(double d) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
  const double e = fma(-(d * y0), y0, 1.0);
  return fma(y0, e * fma(0.375, e, 0.5), y0);
}

constexpr int kSchurWarps = kSchurThreads / 32;
constexpr int kMBase = 224;   // thread kMBase keeps the factored diagonal block for the back-substitution

// Reduced system in global memory: S [N][ld] in FREE-camera index space (N = 6 n_free, ld = reduced_ld(N)),
// upper triangle r <= c only; column N is the right-hand-side part.  It accumulates P = sum_p Zt_p Zt_pᵀ
// (K_B's atomics); the solver forms U_s + D_c² - P in a shared-memory copy of the same shape.

// ---- (S) reduced camera system: blocked (6x6) Cholesky of the bordered matrix + solve, whole CTA ------
// Works on Ut = the UPPER triangle (row k holds column k of the Cholesky factor, i.e. Ut = Lᵀ when done)
// with the right-hand side as column N, so the forward substitution falls out of the factorisation.
// Per block step: every thread that owns a column of the panel factors the 6x6 diagonal block redundantly
// in registers and solves its own column (no factor -> barrier -> panel sequence); the trailing update
// Ut22 -= PᵀP (P = the 6 x m panel) runs on the fp64 tensor pipe, one 8x8 tile per DMMA pair.
// The code is deliberately compact: these phases execute once per launch, from a cold instruction cache.
// sm: Ut [N+8][ld] | Ldd [nf][36] (factored diagonal blocks, transposed) | idv [N] (reciprocal pivots) | xb [N] |
//     yv [N] | Us [F][27].
// Writes the step into st and the candidate cameras.  s_cams: both camera buffers [2][kMaxD] (or null).
// ut_filled: Ut already holds -P (multi-GPU: the rank-ordered sum of every rank's contribution was written there).
template <int TPW>
__device__ void solve_reduced(const LmParams& lp, LmState& st, double* sm, int F, const double* xs, const double* s_cams, bool ut_filled, int n_rep,
                              size_t s_off) {
  const int nf = st.n_free, N = 6 * nf, ld = reduced_ld(N);
  const int cur = st.cur, eb = st.eval_buf, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const double radius = st.radius;
  const int Npad = (N + 1) & ~1;
  double* Ut = sm;
  double* Ldd = Ut + (N + 8) * ld;
  double* idv = Ldd + nf * 36;
  double* xb = idv + Npad;
  double* yv = xb + Npad;
  double* Us = yv + Npad;                      // [F][27] pose blocks of the accepted point
  // Everything this phase reads from global memory is requested FIRST, in one round: the accumulator (-P: a
  // linear copy, all loads of a thread in flight at once) and, after a rejected step, the accepted point's pose
  // blocks; the index arithmetic below runs in the shadow of that round trip.
  constexpr int kCopyB = 5;                    // double2 per thread per round: 2560 entries per round of 256 threads
  const int tot = N * ld, tot2 = tot >> 1;     // (ld is even, the copies are 16-byte aligned)
  const bool us_from_xs = xs && st.took_step;
  double2 v[kCopyB];
  double uc[2];
#pragma unroll
  for (int u = 0; u < kCopyB; ++u) {
    const bool in = !ut_filled && tid + u * nthr < tot2;
    v[u] = in ? __ldcg(reinterpret_cast<const double2*>(lp.S + s_off) + tid + u * nthr) : make_double2(0.0, 0.0);
  }
  if (n_rep > 1) {   // the other copies the CTAs spread their atomics over (same round of loads)
#pragma unroll
    for (int u = 0; u < kCopyB; ++u) {
      double2 w[kSReplicas - 1];
      const bool in = !ut_filled && tid + u * nthr < tot2;
#pragma unroll
      for (int r = 1; r < kSReplicas; ++r)
        w[r - 1] = in ? __ldcg(reinterpret_cast<const double2*>(lp.S + s_off + r * lp.s_cap) + tid + u * nthr) : make_double2(0.0, 0.0);
#pragma unroll
      for (int r = 1; r < kSReplicas; ++r) { v[u].x += w[r - 1].x; v[u].y += w[r - 1].y; }
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) uc[u] = (!us_from_xs && tid + u * nthr < F * kUStride) ? __ldcg(lp.Ucur + tid + u * nthr) : 0.0;
  // this warp's 8x8 tiles (tr <= tc) of the ABSOLUTE tile grid over Ut — the same assignment in every block step,
  // so every address below is formed once: C fragment, and the panel-row-relative A / B fragment offsets
  const int Tabs = (N + 8) >> 3, n_tiles = Tabs * (Tabs + 1) / 2;
  const int fr = lane & 3, fc = lane >> 2;
  int offc[TPW], offa[TPW], offb[TPW];
  unsigned tile_rc[TPW];
#pragma unroll
  for (int k = 0; k < TPW; ++k) {
    const int t = warp + kSchurWarps * k;
    int tr = 0, rem = t < n_tiles ? t : 0;
    while (rem >= Tabs - tr) { rem -= Tabs - tr; ++tr; }
    const int tc = tr + rem;
    offc[k] = (8 * tr + fc) * ld + 8 * tc + 2 * fr;
    offa[k] = fr * ld + 8 * tr + fc;
    offb[k] = fr * ld + 8 * tc + fc;
    tile_rc[k] = t < n_tiles ? (unsigned)(tr | (tc << 8)) : 0xffffu;
  }
  __shared__ int s_fr[kMaxFrames];
  __shared__ int s_ok;
  if (tid == 0) {
    s_ok = 1;
#pragma unroll 1
    for (int f = 0; f < F; ++f) if (st.free_index[f] >= 0) s_fr[st.free_index[f]] = f;
  }
  // the accepted point's pose blocks come from the evaluation just adopted when there is one (no round trip
  // through the copy that was stored to global memory a moment ago)
#pragma unroll
  for (int u = 0; u < 2; ++u)
    if (tid + u * nthr < F * kUStride) Us[tid + u * nthr] = us_from_xs ? xs[tid + u * nthr] : uc[u];
  // (the accumulator copies are re-zeroed for the next elimination by the warps that idle during the back-substitution)
#pragma unroll
  for (int u = 0; u < kCopyB; ++u)
    if (!ut_filled && tid + u * nthr < tot2) reinterpret_cast<double2*>(Ut)[tid + u * nthr] = make_double2(-v[u].x, -v[u].y);
  // wide systems: the rest in further rounds, again with all loads of a round in flight together (one load at a time
  // made this loop 9 of the 11 us that assembling a 90 x 91 system took)
  constexpr int kTailB = 6;
  for (int e0 = tid + kCopyB * nthr; !ut_filled && e0 < tot2; e0 += kTailB * nthr) {
    double2 a[kTailB];
#pragma unroll
    for (int u = 0; u < kTailB; ++u)
      a[u] = e0 + u * nthr < tot2 ? __ldcg(reinterpret_cast<const double2*>(lp.S + s_off) + e0 + u * nthr) : make_double2(0.0, 0.0);
    for (int r = 1; r < n_rep; ++r) {
#pragma unroll
      for (int u = 0; u < kTailB; ++u) {
        const double2 b = e0 + u * nthr < tot2 ? __ldcg(reinterpret_cast<const double2*>(lp.S + s_off + r * lp.s_cap) + e0 + u * nthr)
                                                : make_double2(0.0, 0.0);
        a[u].x += b.x; a[u].y += b.y;
      }
    }
#pragma unroll
    for (int u = 0; u < kTailB; ++u)
      if (e0 + u * nthr < tot2) reinterpret_cast<double2*>(Ut)[e0 + u * nthr] = make_double2(-a[u].x, -a[u].y);
  }
  __syncthreads();
  // + U_s + D_c² on the diagonal blocks, + gs_c on the right-hand-side column: one entry per thread
  for (int t = tid; t < nf * kUStride; t += nthr) {
    const int fi = t / kUStride, k = t - fi * kUStride, f = s_fr[fi];
    if (k < 21) {
      const int a = (k >= 6) + (k >= 11) + (k >= 15) + (k >= 18) + (k >= 20), b = a + k - (a * (13 - a)) / 2;
      const double uu = st.scale_c[6 * f + a] * Us[f * kUStride + k] * st.scale_c[6 * f + b];
      double val = Ut[(6 * fi + a) * ld + 6 * fi + b] + uu;
      if (a == b) val += fmin(fmax(uu, st.min_diag), st.max_diag) / radius;
      Ut[(6 * fi + a) * ld + 6 * fi + b] = val;
    } else {
      const int a = k - 21;
      Ut[(6 * fi + a) * ld + N] += st.scale_c[6 * f + a] * Us[f * kUStride + 21 + a];
    }
  }
  __syncthreads();
  if (lp.dbg && tid == 0) lp.dbg[5] = gtime();

  for (int jb = 0; jb < nf; ++jb) {
    const int j0 = 6 * jb, m = N - j0 - 6;     // panel columns i = j0 + 6 + t, t = 0..m (t == m: the rhs column N)
    const bool col_thr = tid <= m, inv_thr = tid == kMBase;
    UB_STAMP(16);
#if defined(UB_TRAIL) && UB_TRAIL == 3   // ubench: no diagonal / panel work (is the trailing phase slow on its own?)
    if (false) {
#else
    if (col_thr || inv_thr) {
#endif
      // 6x6 Cholesky of the diagonal block, right-looking in registers, REDUNDANTLY in every thread that
      // needs it (broadcast shared-memory reads instead of a factor -> barrier -> panel sequence).  The
      // positivity check stays off the dependency chain: a bad pivot poisons the step, which is then rejected.
      double L[6][6], id[6], r[6];
      bool ok = true;
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int k = 0; k <= i; ++k) L[i][k] = Ut[(j0 + k) * ld + j0 + i];
      if (col_thr) {
#pragma unroll
        for (int k = 0; k < 6; ++k) r[k] = Ut[(j0 + k) * ld + j0 + 6 + tid];
      }
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const double d = L[j][j];
        ok = ok && (d > 0.0) && (d < DBL_MAX);
        id[j] = rsqrt_pos(d);
#pragma unroll
        for (int i = j + 1; i < 6; ++i) L[i][j] *= id[j];
#pragma unroll
        for (int i = j + 1; i < 6; ++i)
#pragma unroll
          for (int k = j + 1; k <= i; ++k) L[i][k] -= L[i][j] * L[k][j];
      }
      UB_STAMP(17);
      if (col_thr) {
        // panel: this column of L_dd^-1 A_12 (forward substitution)
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          r[c] *= id[c];
#pragma unroll
          for (int k = c + 1; k < 6; ++k) r[k] -= r[c] * L[k][c];
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) Ut[(j0 + c) * ld + j0 + 6 + tid] = r[c];
        if (tid == 0 && !ok) s_ok = 0;
      } else {
        // the factored diagonal block and its reciprocal pivots (for the back-substitution)
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) {
#pragma unroll
          for (int i = cc + 1; i < 6; ++i) Ldd[jb * 36 + cc * 6 + i] = L[i][cc];   // not in place: the panel threads may still be reading the block
          idv[j0 + cc] = id[cc];
        }
      }
    }
    UB_STAMP(18);
    UB_ARRIVE(0);
    __syncthreads();
    UB_STAMP(19);
    // trailing update Ut22 -= PᵀP on the tensor pipe (P = the 6 x m panel, rows j0..j0+5): per 8x8 tile two
    // DMMAs (k = 6, the second half-empty).  Tiles straddling the factored part get zeros for rows / columns
    // < j0 + 6; padding rows / columns absorb the overhang.
    {
      const int base = j0 + 6, tb = base >> 3;
      const double* pr = Ut + j0 * ld;
      constexpr int CH = TPW < 3 ? TPW : 3;               // tiles whose operands are in flight together
#pragma unroll
      for (int k0 = 0; k0 < TPW; k0 += CH) {
        double a0[CH], a1[CH], b0[CH], b1[CH];
        double2 c[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          const int k = k0 + u;
          if (tile_rc[k] != 0xffffu && (int)(tile_rc[k] & 0xffu) >= tb) {
            a0[u] = pr[offa[k]]; b0[u] = pr[offb[k]];
            a1[u] = fr < 2 ? pr[4 * ld + offa[k]] : 0.0; b1[u] = fr < 2 ? pr[4 * ld + offb[k]] : 0.0;
            c[u] = *reinterpret_cast<const double2*>(Ut + offc[k]);
          }
        }
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          const int k = k0 + u;
          if (tile_rc[k] != 0xffffu && (int)(tile_rc[k] & 0xffu) >= tb) {
            const bool ra = 8 * (int)(tile_rc[k] & 0xffu) + fc >= base, cb = 8 * (int)(tile_rc[k] >> 8) + fc >= base;
            const double x0 = ra ? -a0[u] : 0.0, x1 = ra ? -a1[u] : 0.0, y0 = cb ? b0[u] : 0.0, y1 = cb ? b1[u] : 0.0;
            dmma884(c[u].x, c[u].y, x0, y0);
            dmma884(c[u].x, c[u].y, x1, y1);
            // rows above the trailing part received a zero update: not stored, so that the panel rows other warps
            // are still reading are never written during this phase
            if (ra) *reinterpret_cast<double2*>(Ut + offc[k]) = c[u];
          }
        }
      }
    }
    UB_STAMP(20);
    UB_ARRIVE(1);
    __syncthreads();
    UB_STAMP(21);
  }
  if (lp.dbg && tid == 0) lp.dbg[6] = gtime();
  // column N now holds y = L^-1 rhs; backward substitution Lᵀ x = y on one warp: lane 0 solves the 6x6
  // triangle of a block (the chain runs through one multiply-add + one multiply per unknown; everything it reads
  // that does not depend on the unknowns is loaded first), every lane then removes the block's unknowns from
  // the rows above (two rows per lane in flight)
  // the accumulator this solve consumed must be zero again before it is used next: with the alternating sets of the
  // single-GPU kernel (n_rep > 1) the CTAs of the NEXT launch clear it, off this CTA's critical path; otherwise the
  // warps that idle during the substitution do
  if (warp != 0 && !ut_filled && n_rep == 1) {
    const double2 z = make_double2(0.0, 0.0);
    for (int e0 = tid - 32; e0 < tot2; e0 += nthr - 32) reinterpret_cast<double2*>(lp.S + s_off)[e0] = z;
  }
  if (warp == 0) {
    for (int i = lane; i < N; i += 32) yv[i] = Ut[i * ld + N];
    __syncwarp();
    for (int jb = nf - 1; jb >= 0; --jb) {
      const int j0 = 6 * jb;
      // rows this lane updates afterwards: request them before the triangular solve
      double rw[2][6], yo[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i = lane + 32 * u;
        const double* row = Ut + (i < j0 ? i : 0) * ld + j0;
#pragma unroll
        for (int k = 0; k < 6; ++k) rw[u][k] = row[k];
        yo[u] = yv[i < j0 ? i : 0];
      }
      if (lane == 0) {
        double x[6], Lt[15], yy[6], idd[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) { yy[c] = yv[j0 + c]; idd[c] = idv[j0 + c]; }
        {
          int e = 0;
#pragma unroll
          for (int c = 0; c < 5; ++c)
#pragma unroll
            for (int k = c + 1; k < 6; ++k) Lt[e++] = Ldd[jb * 36 + c * 6 + k];
        }
#pragma unroll
        for (int c = 5; c >= 0; --c) {
          double acc = yy[c];
          const int e0 = c * 5 - (c * (c - 1)) / 2 - (c + 1);   // index of (c, k) in Lt is e0 + k
#pragma unroll
          for (int k = 5; k > c; --k) acc -= Lt[e0 + k] * x[k];   // newest unknown last: short chain
          x[c] = acc * idd[c];
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) xb[j0 + c] = x[c];
      }
      __syncwarp();
      double xv[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) xv[k] = xb[j0 + k];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i = lane + 32 * u;
        if (i < j0) {
          const double sa = rw[u][0] * xv[0] + rw[u][2] * xv[2] + rw[u][4] * xv[4];
          const double sb = rw[u][1] * xv[1] + rw[u][3] * xv[3] + rw[u][5] * xv[5];
          yv[i] = yo[u] - (sa + sb);
        }
      }
      for (int i = lane + 64; i < j0; i += 32) {   // wide systems (more than 11 cameras)
        const double* row = Ut + i * ld + j0;
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k += 2) { sa += row[k] * xv[k]; sb += row[k + 1] * xv[k + 1]; }
        yv[i] -= sa + sb;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (lp.dbg && tid == 0) lp.dbg[7] = gtime();
  // step (scaled space) = -x ; candidate cameras ; camera part of the model cost change
  const double* bb = xb;
  const int D = 6 * F;
  if (tid < 32) {
    double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
    bool bad = false;
    for (int i = tid; i < D; i += 32) {
      const int f = i / 6, a = i - f * 6, fi = st.free_index[f];
      const double x = s_cams ? s_cams[cur * kMaxD + i] : lp.cams[((size_t)cur * F + f) * 6 + a];
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

// ---- (E) Schur elimination of this CTA's points --------------------------------------------------------
// LPP lanes per point (one per observation slot; 8 for windows of <= 8 frames, else 16), TPW 8x8 output
// tiles per warp.  sm: Zt [Dp][LD], Dp = roundup(N + 1, 8) rows (6 fi + a; row N = L^-1 gs), LD = 3 * points
// per batch + 4 (rows 32 bytes apart modulo 128: the DMMA fragment loads are bank-conflict free).
__shared__ unsigned long long s_tdbg[4];   // PBA_DEBUG_TIMELINE: elimination sub-phases of this CTA
// (pre_o0, pre_o1): the CSR header of this lane's point in the CTA's first batch, requested by the caller before
// the decision phase so that its round trip is off the critical path.
// cur / radius: the buffer holding the blocks to eliminate and the trust-region radius to damp them with (the
// accepted point after a decision — or a HYPOTHESIS about the decision, multi-GPU path); first: this is the first
// elimination of the solve (forms the Jacobi scaling of the points); S_dst / Vinv_dst: where P and (Vs + D²)^-1 go.
// bidx / nblk: this CTA's index among the nblk CTAs that share the elimination (the multi-GPU kernel runs two groups).
template <int LPP, int TPW>
__device__ void eliminate(const LmParams& lp, const LmState& st, double* Zt, int pre_o0, int pre_o1, int cur, double radius,
                          bool first, double* S_dst, double* Vinv_dst, int bidx, int nblk) {
  constexpr int PPW = 32 / LPP, PPB = kSchurWarps * PPW, K = 3 * PPB, LD = K + 4;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = lp.n_points, N = 6 * st.n_free, NS = reduced_ld(N);
  const int Dp = (N + 8) & ~7, T = Dp >> 3, n_tiles = T * (T + 1) / 2;
  const double dmin = st.min_diag, dmax = st.max_diag;
  // this CTA's contiguous block of points
  const int per_cta = (n + nblk - 1) / nblk;
  const int p_begin = bidx * per_cta, p_end = min(n, p_begin + per_cta);
  // this warp's output tiles (tr <= tc): shared-memory row offsets of the A and B fragments
  int offa[TPW], offb[TPW];
  unsigned tile_rc[TPW];
#pragma unroll
  for (int k = 0; k < TPW; ++k) {
    const int t = warp + kSchurWarps * k;
    int tr = 0, rem = t < n_tiles ? t : 0;
    while (rem >= T - tr) { rem -= T - tr; ++tr; }
    const int tc = tr + rem;
    offa[k] = (8 * tr + (lane >> 2)) * LD + (lane & 3);
    offb[k] = (8 * tc + (lane >> 2)) * LD + (lane & 3);
    tile_rc[k] = t < n_tiles ? (unsigned)(tr | (tc << 8)) : 0xffffu;
  }
  double acc[TPW][2];
#pragma unroll
  for (int k = 0; k < TPW; ++k) acc[k][0] = acc[k][1] = 0.0;

  const double* Vb = lp.V + (size_t)cur * n * 6;
  const double* gb = lp.gp + (size_t)cur * n * 3;
  const double* Wb = lp.W + (size_t)cur * lp.nnz * 18;
  const int q = warp * PPW + lane / LPP, g = lane % LPP;   // point within the batch, observation slot
  if (lp.dbg && tid == 0) s_tdbg[0] = gtime();
  for (int b0 = p_begin; b0 < p_end; b0 += PPB) {
    const int npb = min(PPB, p_end - b0), p = b0 + q;
    const bool valid = q < npb;
    // first round of loads: CSR header and the point's own blocks
    int o0 = 0, nobs = 0;
    double V[6] = {1.0, 0.0, 0.0, 1.0, 0.0, 1.0}, g0 = 0.0, g1 = 0.0, g2 = 0.0, sp0 = 1.0, sp1 = 1.0, sp2 = 1.0;
    if (valid) {
      if (b0 == p_begin) { o0 = pre_o0; nobs = pre_o1 - pre_o0; }
      else { o0 = __ldg(lp.obs_off + p); nobs = __ldg(lp.obs_off + p + 1) - o0; }
      const double2* v2 = reinterpret_cast<const double2*>(Vb + (size_t)p * 6);
      const double2 va = __ldg(v2), vb = __ldg(v2 + 1), vc = __ldg(v2 + 2);
      V[0] = va.x; V[1] = va.y; V[2] = vb.x; V[3] = vb.y; V[4] = vc.x; V[5] = vc.y;
      g0 = __ldg(gb + (size_t)p * 3); g1 = __ldg(gb + (size_t)p * 3 + 1); g2 = __ldg(gb + (size_t)p * 3 + 2);
      if (!first) { sp0 = lp.scale_p[(size_t)p * 3]; sp1 = lp.scale_p[(size_t)p * 3 + 1]; sp2 = lp.scale_p[(size_t)p * 3 + 2]; }
    }
    // zero the staging matrix (rows of frames that do not observe a point must read as zero)
    if (b0 != p_begin) __syncthreads();                       // the previous batch's product is done with it
    for (int i = tid; i < Dp * (LD / 2); i += kSchurThreads) reinterpret_cast<double2*>(Zt)[i] = make_double2(0.0, 0.0);
    // second round: this lane's observation
    int f = -1;
    double2 w2[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w2[k] = make_double2(0.0, 0.0);
    if (valid && g < nobs) {
      f = __ldg(lp.obs_frame + o0 + g);
      const double2* wp = reinterpret_cast<const double2*>(Wb + (size_t)(o0 + g) * 18);
#pragma unroll
      for (int k = 0; k < 9; ++k) w2[k] = __ldg(wp + k);
    }
    __syncthreads();
    if (valid) {
      if (first) {
        sp0 = st.jacobi_scaling ? 1.0 / (1.0 + sqrt(V[0])) : 1.0;
        sp1 = st.jacobi_scaling ? 1.0 / (1.0 + sqrt(V[3])) : 1.0;
        sp2 = st.jacobi_scaling ? 1.0 / (1.0 + sqrt(V[5])) : 1.0;
        if (g == 0) { lp.scale_p[(size_t)p * 3] = sp0; lp.scale_p[(size_t)p * 3 + 1] = sp1; lp.scale_p[(size_t)p * 3 + 2] = sp2; }
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
      const double gs0 = sp0 * g0, gs1 = sp1 * g1, gs2 = sp2 * g2;
      const double zg0 = gs0 * il00, zg1 = (gs1 - l10 * zg0) * il11, zg2 = (gs2 - l20 * zg0 - l21 * zg1) * il22;
      if (g == 0) {
        const double m10 = -l10 * il00 * il11, m20 = -(l20 * il00 + l21 * m10) * il22, m21 = -l21 * il11 * il22;
        double* vi = Vinv_dst + (size_t)p * 6;   // (Vs + D²)^-1 = M^T M, used by K_A's back-substitution
        vi[0] = il00 * il00 + m10 * m10 + m20 * m20; vi[1] = m10 * il11 + m20 * m21; vi[2] = m20 * il22;
        vi[3] = il11 * il11 + m21 * m21; vi[4] = m21 * il22; vi[5] = il22 * il22;
        double* zr = Zt + N * LD + 3 * q;
        zr[0] = zg0; zr[1] = zg1; zr[2] = zg2;
      }
      const int fi = f >= 0 ? st.free_index[f] : -1;
      if (fi >= 0) {
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          const double sc = st.scale_c[f * 6 + a];
          // W row a = elements 3a .. 3a+2 of the 18-double block (compile-time selects after unrolling)
          const double wa0 = ((3 * a) & 1) ? w2[(3 * a) >> 1].y : w2[(3 * a) >> 1].x;
          const double wa1 = ((3 * a + 1) & 1) ? w2[(3 * a + 1) >> 1].y : w2[(3 * a + 1) >> 1].x;
          const double wa2 = ((3 * a + 2) & 1) ? w2[(3 * a + 2) >> 1].y : w2[(3 * a + 2) >> 1].x;
          const double z0 = sc * wa0 * sp0 * il00;
          const double z1 = (sc * wa1 * sp1 - z0 * l10) * il11;
          const double z2 = (sc * wa2 * sp2 - z0 * l20 - z1 * l21) * il22;
          double* zr = Zt + (6 * fi + a) * LD + 3 * q;
          zr[0] = z0; zr[1] = z1; zr[2] = z2;
        }
      }
    }
    __syncthreads();
    // P += Zt Ztᵀ on the fp64 tensor pipe: 4 columns (k) per DMMA, independent accumulator chains per tile
    const int ksteps = (3 * npb + 3) >> 2;
    for (int ks = 0; ks < ksteps; ++ks) {
#pragma unroll
      for (int k = 0; k < TPW; ++k) {
        if (tile_rc[k] != 0xffffu) {
          const double a = Zt[offa[k] + 4 * ks];
          const double b = Zt[offb[k] + 4 * ks];
          dmma884(acc[k][0], acc[k][1], a, b);
        }
      }
    }
  }
  if (lp.dbg && tid == 0) s_tdbg[1] = gtime();
  // one fp64 atomic per entry of the upper triangle per CTA
#pragma unroll
  for (int k = 0; k < TPW; ++k) {
    if (tile_rc[k] != 0xffffu) {
      const int r = 8 * (int)(tile_rc[k] & 0xffu) + (lane >> 2);
      const int c = 8 * (int)(tile_rc[k] >> 8) + 2 * (lane & 3);
      if (r < N) {
        if (r <= c && c <= N && acc[k][0] != 0.0) atomicAdd(S_dst + r * NS + c, acc[k][0]);
        if (r <= c + 1 && c + 1 <= N && acc[k][1] != 0.0) atomicAdd(S_dst + r * NS + c + 1, acc[k][1]);
      }
    }
  }
  if (lp.dbg && tid == 0) s_tdbg[2] = gtime();
}

// The trust-region decision for the candidate just evaluated, from the complete accumulators xs (pose blocks +
// scalars); warp 0 reduces, lane 0 decides, and at iteration 0 all threads form the Jacobi scaling of the pose
// columns.  Called by every thread of the CTA; ends with the CTA synchronised.
__device__ __forceinline__ void take_decision(LmState& s_st, const double* s_xs, const double* s_cams, int F, IterSummary& s_it,
                                              int& s_push, int& s_it0) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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
          const double c = buf ? s_cams[kMaxD + i] : s_cams[i];
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
    if (lane == 0) {
      s_it0 = s_st.iteration == 0;
      s_push = decide(s_st, E, gm, g2, csq, s_xs, F, s_it) ? 1 : 0;
    }
  }
  __syncthreads();
  if (s_it0) {   // IterationZero: Jacobi column scaling of the pose columns from the initial evaluation
    for (int i = tid; i < 6 * F; i += blockDim.x) {
      const int f = i / 6, a = i - 6 * f;
      s_st.scale_c[i] = s_st.jacobi_scaling ? 1.0 / (1.0 + sqrt(s_xs[f * kUStride + utri6(a, a)])) : 1.0;
    }
    __syncthreads();
  }
}

// Tail of an LM iteration (one CTA): adopt the candidate's pose blocks if the step was taken, solve
// the reduced camera system, re-zero the accumulators K_A fills next, publish the new state.
template <int TPW>
__device__ void finish_iteration(const LmParams& lp, LmState& st, double* sm, int F, const double* xs, const double* s_cams, bool ut_filled, int n_rep = 1,
                                 size_t s_off = 0) {
  const int tid = threadIdx.x;
  if (st.took_step)
    for (int i = tid; i < F * kUStride; i += blockDim.x) lp.Ucur[i] = xs ? xs[i] : __ldcg(lp.Xacc + i);
  __syncthreads();
  solve_reduced<TPW>(lp, st, sm, F, xs, s_cams, ut_filled, n_rep, s_off);
  if (lp.dbg && tid == 0) lp.dbg[3] = gtime();
  for (int i = tid; i < F * kUStride + kEacc + kMaxRanks; i += blockDim.x) lp.Xacc[i] = 0.0;
  if (tid == 0) *lp.ticket = 0u;
  __syncthreads();
  for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
    reinterpret_cast<int*>(lp.st_out)[i] = reinterpret_cast<const int*>(&st)[i];
  if (lp.dbg && tid == 0) lp.dbg[4] = gtime();
  if (lp.stamps && tid == 0) lp.stamps[2 * (st.num_evals - 1) + 1] = globaltimer_ns();
}

// LPP: lanes per point in the elimination (8: windows of <= 8 frames; 16 otherwise); TPW: 8x8 tiles of the
// reduced system per warp (3: <= 24 tiles, i.e. <= 7 optimised cameras; 12: up to 16 cameras)
template <int LPP, int TPW>
__global__ void __launch_bounds__(kSchurThreads, 1) k_schur_solve(const LmParams lp) {
  __shared__ LmState s_st;
  __shared__ IterSummary s_it;
  __shared__ int s_push, s_last, s_it0;
  __shared__ double s_xs[kMaxFrames * kUStride + kEacc + kMaxRanks];        // the evaluation's pose blocks + scalars (summed over ranks)
  __shared__ double s_cams[2 * kMaxD];                                       // both camera buffers (decision, candidate)
  extern __shared__ __align__(16) double sm[];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int F = lp.n_frames;
  const unsigned long long t_start = lp.dbg ? gtime() : 0ull;
  // CSR header of this lane's point in the CTA's first elimination batch (does not depend on the LM state)
  int pre_o0 = 0, pre_o1 = 0;
  {
    const int per_cta = (lp.n_points + gridDim.x - 1) / gridDim.x;
    const int p = blockIdx.x * per_cta + warp * (32 / LPP) + lane / LPP;
    if (p < lp.n_points && p < (blockIdx.x + 1) * per_cta) { pre_o0 = __ldg(lp.obs_off + p); pre_o1 = __ldg(lp.obs_off + p + 1); }
  }

  // ---- (D) decision, redundantly per CTA ------------------------------------------------
  // Everything the decision reads is requested in ONE round of loads: the state, the evaluation's
  // accumulators (single-GPU / NCCL path) and both buffers of the cameras (which one is the candidate is
  // only known once the state has arrived) - three dependent L2 round trips otherwise.
  if (lp.pdl) pdl_wait();   // K_A has finished (everything above is independent of it)
  const unsigned long long t_entry = globaltimer_ns();
  const int xn = F * kUStride + kEacc + kMaxRanks;
  constexpr int kXPre = (kMaxFrames * kUStride + kEacc + kMaxRanks + kSchurThreads - 1) / kSchurThreads;
  double x_pre[kXPre];
#pragma unroll
  for (int k = 0; k < kXPre; ++k) x_pre[k] = (tid + k * kSchurThreads < xn) ? __ldcg(lp.Xacc + tid + k * kSchurThreads) : 0.0;
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
  // (read from the input state, which nothing writes during this kernel: the shared copy's `done` is set by the
  // deciding lane below while slower warps could still be looking at it)
  const int was_done = lp.st_in->done;
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < (kMaxD + 31) / 32; ++k) {
      const int i = lane + 32 * k;
      if (i < F * 6) { s_cams[i] = cam_pre[0][k]; s_cams[kMaxD + i] = cam_pre[1][k]; }
    }
  }
#pragma unroll
  for (int k = 0; k < kXPre; ++k)
    if (tid + k * kSchurThreads < xn) s_xs[tid + k * kSchurThreads] = x_pre[k];
  __syncthreads();
  if (was_done) {
    if (blockIdx.x == 0) {
      for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
        reinterpret_cast<int*>(lp.st_out)[i] = reinterpret_cast<const int*>(&s_st)[i];
      if (lp.cond && tid == 0) cudaGraphSetConditional(lp.cond, 0u);   // leave the while node of the solve graph
    }
    return;
  }
  if (lp.stamps && blockIdx.x == 0 && tid == 0) lp.stamps[2 * s_st.num_evals] = t_entry;
  take_decision(s_st, s_xs, s_cams, F, s_it, s_push, s_it0);
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
  // The accumulator S: kSReplicas copies the CTAs spread their atomics over, in two SETS that alternate from launch to
  // launch - this launch adds into (and its solving CTA reads) set `par`, and clears its slice of the other set, which
  // the previous launch consumed: nobody clears anything on the solving CTA's critical path.  (Split mode: one copy, the
  // all-reduce sums it; wide systems: one copy - summing four 90 x 91 accumulators costs the solving CTA more than the
  // shorter atomic chains save - and the solving CTA clears it.)
  const int n_rep = (lp.split || s_st.n_free > 8) ? 1 : kSReplicas;
  const size_t s_off = n_rep > 1 ? (size_t)(s_st.num_evals & 1) * kSReplicas * lp.s_cap : 0;
  if (n_rep > 1) {
    double2* other = reinterpret_cast<double2*>(lp.S + (size_t)((s_st.num_evals & 1) ^ 1) * kSReplicas * lp.s_cap);
    const int N0 = 6 * s_st.n_free, tot2 = (N0 * reduced_ld(N0)) >> 1;
    const double2 z = make_double2(0.0, 0.0);
    for (int r = 0; r < kSReplicas; ++r)
      for (int e0 = blockIdx.x * kSchurThreads + tid; e0 < tot2; e0 += gridDim.x * kSchurThreads)
        other[(size_t)r * (lp.s_cap >> 1) + e0] = z;
  }
  eliminate<LPP, TPW>(lp, s_st, sm, pre_o0, pre_o1, s_st.cur, s_st.radius, s_st.iteration == 1,
                      lp.S + s_off + (blockIdx.x % n_rep) * lp.s_cap, lp.Vinv, blockIdx.x, gridDim.x);

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
    lp.dbg[12] = s_tdbg[0]; lp.dbg[13] = s_tdbg[1]; lp.dbg[14] = s_tdbg[2];
  }
  if (!s_last) return;
  if (lp.pdl) pdl_launch_dependents();   // every other CTA has exited: the next K_A may start its prologue during the solve
  if (lp.split) {
    // multi-GPU: the reduced system is summed across ranks first; k_solve_only finishes the iteration
    if (tid == 0) *lp.ticket = 0u;
    for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
      reinterpret_cast<int*>(lp.st_out)[i] = reinterpret_cast<const int*>(&s_st)[i];
    return;
  }
  finish_iteration<TPW>(lp, s_st, sm, F, s_xs, s_cams, false, n_rep, s_off);
}

// ---- multi-GPU kernel: speculative elimination, one exchange per LM iteration (pba_device.cuh `Xchg`) -----------
// sum over the ranks, in rank order, of cell i of every rank's slot; all first attempts are in flight together
__device__ __forceinline__ bool ll_sum(const ulonglong2* base, size_t slot_stride, int n_ranks, size_t i, unsigned long long e, double& out) {
  double v[kMaxRanks];
  bool got[kMaxRanks];
#pragma unroll
  for (int q = 0; q < kMaxRanks; ++q) { v[q] = 0.0; got[q] = q >= n_ranks || ll_try_load(base + q * slot_stride + i, e, v[q]); }
  bool ok = true;
  double acc = 0.0;
#pragma unroll
  for (int q = 0; q < kMaxRanks; ++q) {
    if (!got[q] && !ll_load(base + q * slot_stride + i, e, v[q])) ok = false;
    if (q < n_ranks) acc += v[q];
  }
  out = acc;
  return ok;
}

template <int LPP, int TPW>
__global__ void __launch_bounds__(kSchurThreads, 1) k_schur_solve_x(const LmParams lp) {
  __shared__ LmState s_st;
  __shared__ IterSummary s_it;
  __shared__ int s_push, s_last, s_it0, s_xok, s_code;
  __shared__ double s_xs[kMaxFrames * kUStride + kEacc + kMaxRanks];        // the evaluation's pose blocks + scalars: local, then summed over ranks
  __shared__ double s_cams[2 * kMaxD];
  extern __shared__ __align__(16) double sm[];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int F = lp.n_frames, nr = lp.xc.n_ranks;
  const unsigned long long t_start = lp.dbg ? gtime() : 0ull;
  // two groups of G = gridDim.x / 2 CTAs: group 0 eliminates under hypothesis A, group 1 under hypothesis R, each
  // over the same blocks of points
  // (lp.speculate == 0: one group, nothing is eliminated before the decision is known - every iteration takes the
  // "second elimination" path below with all CTAs, at the price of a second exchange)
  const bool spec = lp.speculate != 0;
  const int G = spec ? gridDim.x >> 1 : gridDim.x, hyp = (spec && blockIdx.x >= G) ? 1 : 0, bidx = blockIdx.x - hyp * G;
  int pre_o0 = 0, pre_o1 = 0;
  const int per_cta = (lp.n_points + G - 1) / G;
  {
    const int p = bidx * per_cta + warp * (32 / LPP) + lane / LPP;
    if (p < lp.n_points && p < (bidx + 1) * per_cta) { pre_o0 = __ldg(lp.obs_off + p); pre_o1 = __ldg(lp.obs_off + p + 1); }
  }
  if (lp.pdl) pdl_wait();
  const unsigned long long t_entry = globaltimer_ns();
  const int xn = F * kUStride + kEacc + kMaxRanks;
  for (int i = tid; i < xn; i += blockDim.x) s_xs[i] = __ldcg(lp.Xacc + i);
  for (int i = tid; i < F * 6; i += blockDim.x) { s_cams[i] = lp.cams[i]; s_cams[kMaxD + i] = lp.cams[(size_t)F * 6 + i]; }
  for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
    reinterpret_cast<int*>(&s_st)[i] = reinterpret_cast<const int*>(lp.st_in)[i];
  if (tid == 0) s_xok = 1;
  __syncthreads();
  if (lp.stamps && !s_st.done && blockIdx.x == 0 && tid == 0) lp.stamps[2 * s_st.num_evals] = t_entry;
  if (s_st.done) {   // the state is replicated: every rank leaves here together
    if (blockIdx.x == 0) {
      for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
        reinterpret_cast<int*>(lp.st_out)[i] = reinterpret_cast<const int*>(&s_st)[i];
      if (lp.cond && tid == 0) cudaGraphSetConditional(lp.cond, 0u);
    }
    return;
  }
  const unsigned long long e = s_st.xepoch;
  const bool it0 = s_st.iteration == 0;
  const int N = 6 * s_st.n_free, NS = reduced_ld(N), NC = N + 1, NN = N * NC;
  // the entries of the reduced system that travel: r <= c <= N (upper triangle + rhs column), as a flat list so
  // that the push / sum loops keep several independent loads in flight per thread
  constexpr int kXB = 4;
  __shared__ unsigned short s_ent[kMaxD * (kMaxD + 3) / 2];
  const int n_ent = N * (N + 3) / 2;
  for (int r = tid; r < N; r += blockDim.x) {
    int k = r * NC - r * (r - 1) / 2;          // entries of the rows above: sum_{i<r} (N + 1 - i)
    for (int c = r; c <= N; ++c) s_ent[k++] = (unsigned short)((r << 8) | c);
  }
  __syncthreads();
  double* S_A = lp.S;
  double* S_R = lp.S + lp.s_cap;
  double* S_M = lp.S + 2 * lp.s_cap;
  double* Vinv_A = lp.Vinv2;
  double* Vinv_R = lp.Vinv2 + (size_t)lp.n_points * 6;
  // the two outcomes of the pending decision whose radius is known in advance (same expressions as decide())
  const double rad_A = fmin(s_st.max_radius, s_st.radius / fmax(1.0 / 3.0, 1.0 / 3.0));
  const double rad_R = s_st.radius / s_st.decrease_factor;
  const int buf_A = s_st.eval_buf, buf_R = s_st.cur;
  if (!it0 && spec) {
    if (hyp == 0) eliminate<LPP, TPW>(lp, s_st, sm, pre_o0, pre_o1, buf_A, rad_A, false, S_A, Vinv_A, bidx, G);
    else eliminate<LPP, TPW>(lp, s_st, sm, pre_o0, pre_o1, buf_R, rad_R, false, S_R, Vinv_R, bidx, G);
  }
  // ---- ticket 1: the last CTA exchanges, decides; the others wait for its verdict ------------------------
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(lp.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __threadfence();
  }
  __syncthreads();
  const int p_begin = bidx * per_cta, p_end = min(lp.n_points, p_begin + per_cta);
  bool redo = false;
  if (!s_last) {
    if (tid == 0) {
      const unsigned long long t0 = globaltimer_ns();
      unsigned long long v;
      while (((v = ld_acquire_sys(lp.xc.verdict)) >> 3) != e) {
        if (globaltimer_ns() - t0 > kXchgTimeoutNs) { v = (e << 3) | kXFail; break; }
        __nanosleep(200);
      }
      s_code = (int)(v & 7ull);
    }
    __syncthreads();
    const int code = s_code;
    if (code == kXHitA || code == kXHitR) {   // adopt (Vs + D²)^-1 of the outcome that came true (the group that formed it)
      if ((code == kXHitA) == (hyp == 0)) {
        const double* src = (code == kXHitA ? Vinv_A : Vinv_R);
        for (int i = p_begin * 6 + tid; i < p_end * 6; i += blockDim.x) lp.Vinv[i] = __ldcg(src + i);
      }
      return;
    }
    if (code != kXRedo || hyp != 0) return;   // the second elimination is group 0's
    for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
      reinterpret_cast<int*>(&s_st)[i] = __ldcg(reinterpret_cast<const int*>(lp.st_out) + i);
    __syncthreads();
    redo = true;
  } else {
    if (lp.dbg && tid == 0) { lp.dbg[0] = t_start; lp.dbg[1] = gtime(); }
    // push: evaluation part, then both hypotheses (local accumulators are re-zeroed on the way); kXB entries per
    // thread in flight
    const size_t slot1 = ((size_t)(e & 1ull) * nr + lp.xc.rank) * lp.xc.x1_n;
    for (int i = tid; i < xn; i += blockDim.x)
      for (int q = 0; q < nr; ++q) ll_store(lp.xc.x1[q] + slot1 + i, s_xs[i], e);
    if (!it0 && spec) {
      for (int k0 = tid; k0 < n_ent; k0 += kXB * blockDim.x) {
        double a[kXB], b[kXB];
        int rc[kXB];
#pragma unroll
        for (int u = 0; u < kXB; ++u) {
          const int k = k0 + u * blockDim.x;
          rc[u] = k < n_ent ? s_ent[k] : -1;
          a[u] = rc[u] >= 0 ? __ldcg(S_A + (rc[u] >> 8) * NS + (rc[u] & 255)) : 0.0;
          b[u] = rc[u] >= 0 ? __ldcg(S_R + (rc[u] >> 8) * NS + (rc[u] & 255)) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kXB; ++u) {
          if (rc[u] >= 0) {
            const int r = rc[u] >> 8, c = rc[u] & 255;
            S_A[r * NS + c] = 0.0; S_R[r * NS + c] = 0.0;
            for (int q = 0; q < nr; ++q) {
              ll_store(lp.xc.x1[q] + slot1 + xn + r * NC + c, a[u], e);
              ll_store(lp.xc.x1[q] + slot1 + xn + NN + r * NC + c, b[u], e);
            }
          }
        }
      }
    }
    if (lp.dbg && tid == 0) lp.dbg[8] = gtime();
    // sum everybody's evaluation part in rank order, decide
    const ulonglong2* in1 = lp.xc.x1[lp.xc.rank] + (size_t)(e & 1ull) * nr * lp.xc.x1_n;
    __syncthreads();
    for (int i = tid; i < xn; i += blockDim.x) {
      double v;
      if (!ll_sum(in1, lp.xc.x1_n, nr, i, e, v)) s_xok = 0;
      s_xs[i] = v;
    }
    __syncthreads();
    if (!s_xok) {   // a peer never arrived: fail the solve instead of hanging the GPU
      if (tid == 0) { finish(s_st, 2, kMsgXchgTimeout, (double)e, 0.0); *lp.xc.error = 1; *lp.ticket = 0u; }
      __syncthreads();
      for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
        reinterpret_cast<int*>(lp.st_out)[i] = reinterpret_cast<const int*>(&s_st)[i];
      if (tid == 0) {
        __threadfence();
        st_release_sys(lp.xc.verdict, (e << 3) | kXFail);
        if (lp.cond) cudaGraphSetConditional(lp.cond, 0u);
      }
      return;
    }
    take_decision(s_st, s_xs, s_cams, F, s_it, s_push, s_it0);
    if (tid == 0 && s_push) lp.trace[s_st.n_trace - 1] = s_it;
    if (lp.dbg && tid == 0) lp.dbg[9] = gtime();
    if (s_st.done) {
      if (tid == 0) s_st.n_xchg += 1;
      __syncthreads();
      for (int i = tid; i < xn; i += blockDim.x) lp.Xacc[i] = 0.0;
      for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
        reinterpret_cast<int*>(lp.st_out)[i] = reinterpret_cast<const int*>(&s_st)[i];
      __syncthreads();
      if (tid == 0) {
        *lp.ticket = 0u;
        __threadfence();
        st_release_sys(lp.xc.verdict, (e << 3) | kXDone);
        if (lp.cond) cudaGraphSetConditional(lp.cond, 0u);
      }
      return;
    }
    const bool hitA = spec && !it0 && s_st.took_step && s_st.radius == rad_A;
    const bool hitR = spec && !it0 && !s_st.took_step && s_st.radius == rad_R;
    if (hitA || hitR) {
      if (tid == 0) st_release_sys(lp.xc.verdict, (e << 3) | (hitA ? kXHitA : kXHitR));
      // -P of the outcome that came true: the rank-ordered sum goes straight into the solver's matrix (entries
      // outside the upper triangle keep whatever the staging left there: they never reach a valid entry)
      double* Ut = sm;
      const size_t hoff = xn + (hitA ? 0 : NN);
      for (int k0 = tid; k0 < n_ent; k0 += 2 * blockDim.x) {      // two entries (2 x n_ranks loads) in flight per thread
        const int k1 = k0 + blockDim.x;
        const int rc0 = s_ent[k0], rc1 = k1 < n_ent ? s_ent[k1] : rc0;
        double v0, v1;
        const bool ok0 = ll_sum(in1, lp.xc.x1_n, nr, hoff + (rc0 >> 8) * NC + (rc0 & 255), e, v0);
        const bool ok1 = ll_sum(in1, lp.xc.x1_n, nr, hoff + (rc1 >> 8) * NC + (rc1 & 255), e, v1);
        if (!ok0 || !ok1) s_xok = 0;
        Ut[(rc0 >> 8) * NS + (rc0 & 255)] = -v0;
        if (k1 < n_ent) Ut[(rc1 >> 8) * NS + (rc1 & 255)] = -v1;
      }
      if (tid == 0) { s_st.xepoch = e + 1; s_st.n_xchg += 1; if (lp.dbg) { lp.dbg[2] = gtime(); lp.dbg[10] = hitA ? 1 : 2; } }
      __syncthreads();
      if (!s_xok && tid == 0) *lp.xc.error = 1;
      if (lp.pdl) pdl_launch_dependents();
      finish_iteration<TPW>(lp, s_st, sm, F, s_xs, s_cams, true);
      if (hitA == (hyp == 0)) {   // this CTA's own share of (Vs + D²)^-1, off the critical path
        const double* src = hitA ? Vinv_A : Vinv_R;
        for (int i = p_begin * 6 + tid; i < p_end * 6; i += blockDim.x) lp.Vinv[i] = __ldcg(src + i);
      }
      return;
    }
    // neither outcome holds (iteration 0: the Jacobi scaling needs the global pose blocks first; or an accepted
    // step whose radius did not triple): publish the decided state and the summed evaluation, everybody eliminates again
    for (int i = tid; i < xn; i += blockDim.x) lp.Xacc[i] = s_xs[i];
    for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
      reinterpret_cast<int*>(lp.st_out)[i] = reinterpret_cast<const int*>(&s_st)[i];
    __syncthreads();
    if (tid == 0) { __threadfence(); st_release_sys(lp.xc.verdict, (e << 3) | kXRedo); }
    redo = hyp == 0;   // the deciding CTA takes part in the second elimination only if it belongs to group 0
    if (!redo) return;
  }
  if (!redo) return;
  // ---- second elimination with the decided radius, second exchange ---------------------------------------
  eliminate<LPP, TPW>(lp, s_st, sm, pre_o0, pre_o1, s_st.cur, s_st.radius, s_st.iteration == 1, S_M, lp.Vinv, bidx, G);
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(lp.ticket, 1u) == gridDim.x + G - 1) ? 1 : 0;   // ticket 1 counted every CTA, ticket 2 counts group 0
    __threadfence();
  }
  __syncthreads();
  if (!s_last) return;
  {
    const size_t slot2 = ((size_t)(e & 1ull) * nr + lp.xc.rank) * lp.xc.x2_n;
    for (int k0 = tid; k0 < n_ent; k0 += kXB * blockDim.x) {
      double a[kXB];
      int rc[kXB];
#pragma unroll
      for (int u = 0; u < kXB; ++u) {
        const int k = k0 + u * blockDim.x;
        rc[u] = k < n_ent ? s_ent[k] : -1;
        a[u] = rc[u] >= 0 ? __ldcg(S_M + (rc[u] >> 8) * NS + (rc[u] & 255)) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kXB; ++u) {
        if (rc[u] >= 0) {
          const int r = rc[u] >> 8, c = rc[u] & 255;
          S_M[r * NS + c] = 0.0;
          for (int q = 0; q < nr; ++q) ll_store(lp.xc.x2[q] + slot2 + r * NC + c, a[u], e);
        }
      }
    }
    const ulonglong2* in2 = lp.xc.x2[lp.xc.rank] + (size_t)(e & 1ull) * nr * lp.xc.x2_n;
    double* Ut = sm;
    __syncthreads();
    for (int k0 = tid; k0 < n_ent; k0 += 2 * blockDim.x) {
      const int k1 = k0 + blockDim.x;
      const int rc0 = s_ent[k0], rc1 = k1 < n_ent ? s_ent[k1] : rc0;
      double v0, v1;
      const bool ok0 = ll_sum(in2, lp.xc.x2_n, nr, (size_t)(rc0 >> 8) * NC + (rc0 & 255), e, v0);
      const bool ok1 = ll_sum(in2, lp.xc.x2_n, nr, (size_t)(rc1 >> 8) * NC + (rc1 & 255), e, v1);
      if (!ok0 || !ok1) s_xok = 0;
      Ut[(rc0 >> 8) * NS + (rc0 & 255)] = -v0;
      if (k1 < n_ent) Ut[(rc1 >> 8) * NS + (rc1 & 255)] = -v1;
    }
    if (tid == 0) { s_st.xepoch = e + 1; s_st.n_xchg += 2; s_st.n_respec += spec ? 1 : 0; if (lp.dbg) { lp.dbg[2] = gtime(); lp.dbg[10] = 0; } }
    __syncthreads();
    if (!s_xok && tid == 0) *lp.xc.error = 1;
    if (lp.pdl) pdl_launch_dependents();
    finish_iteration<TPW>(lp, s_st, sm, F, nullptr, s_cams, true);
  }
}

// split mode: one CTA, after the all-reduce of S
template <int TPW>
__global__ void __launch_bounds__(kSchurThreads) k_solve_only(const LmParams lp) {
  __shared__ LmState s_st;
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x;
  for (int i = tid; i < (int)(sizeof(LmState) / 4); i += blockDim.x)
    reinterpret_cast<int*>(&s_st)[i] = reinterpret_cast<const int*>(lp.st_out)[i];
  __syncthreads();
  if (s_st.done) return;
  finish_iteration<TPW>(lp, s_st, sm, lp.n_frames, nullptr, nullptr, false);
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

// End of a solve: the accepted x must sit in buffer 0 of the cameras / points (what pba_get_* and the next solve read).
// Decided on the device so that the host can enqueue its result copies behind the loop without a round trip.
__global__ void k_publish(const LmState* __restrict__ st, double* cams, double* pts, int n_cam, int n_pts) {
  if (st->cur == 0) return;
  const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = t0; i < n_cam; i += stride) cams[i] = cams[n_cam + i];
  for (int i = t0; i < n_pts; i += stride) pts[i] = pts[n_pts + i];
}

cudaError_t launch_publish(const LmState* st, double* cams, double* pts, int n_frames, int n_points, cudaStream_t stream) {
  const int n_pts = 3 * n_points;
  int grid = (n_pts + 255) / 256;
  grid = grid < 1 ? 1 : (grid > 64 ? 64 : grid);
  k_publish<<<grid, 256, 0, stream>>>(st, cams, pts, 6 * n_frames, n_pts);
  return cudaGetLastError();
}

// ---- launchers ----------------------------------------------------------------------------
// One contiguous block of points per CTA (at least 32 = a full elimination batch for windows of <= 8 frames, so that
// small windows do not pay one round of atomics per handful of points), at most one CTA per SM.
int schur_grid(int n_points, int sm_count) {
  int per_cta = (n_points + sm_count - 1) / sm_count;
  if (per_cta < 32) per_cta = 32;   // a full batch of the elimination: fewer CTAs, fewer rounds of atomics
  const int g = (n_points + per_cta - 1) / per_cta;
  return g > 0 ? g : 1;
}

int schur_grid_x(int n_points, int sm_count) { return 2 * schur_grid(n_points, sm_count / 2); }

static size_t solve_smem_doubles(int n_free, int F) {
  const size_t N = 6 * (size_t)n_free, ld = reduced_ld((int)N), npad = (N + 1) & ~(size_t)1;
  return (N + 8) * ld + (size_t)n_free * 36 + 3 * npad + (size_t)F * kUStride + 2;
}

template <int LPP, int TPW>
static cudaError_t launch_mode(const LmParams& lp, int grid, int n_free, cudaStream_t s) {
  static bool cfg[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!cfg[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_schur_solve<LPP, TPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    if (e != cudaSuccess) return e;
    cfg[dev & 63] = true;
  }
  constexpr int PPB = kSchurWarps * (32 / LPP), LD = 3 * PPB + 4;
  const int N = 6 * n_free, Dp = (N + 8) & ~7;
  const size_t elim_d = (size_t)Dp * LD, solve_d = solve_smem_doubles(n_free, lp.n_frames);
  const size_t smem = sizeof(double) * (elim_d > solve_d ? elim_d : solve_d);
  if (lp.pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kSchurThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_schur_solve<LPP, TPW>, lp);
  }
  k_schur_solve<LPP, TPW><<<grid, kSchurThreads, smem, s>>>(lp);
  return cudaGetLastError();
}

template <int LPP, int TPW>
static cudaError_t launch_mode_x(const LmParams& lp, int grid, int n_free, cudaStream_t s) {
  static bool cfg[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!cfg[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_schur_solve_x<LPP, TPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    if (e != cudaSuccess) return e;
    cfg[dev & 63] = true;
  }
  constexpr int PPB = kSchurWarps * (32 / LPP), LD = 3 * PPB + 4;
  const int N = 6 * n_free, Dp = (N + 8) & ~7;
  const size_t elim_d = (size_t)Dp * LD, solve_d = solve_smem_doubles(n_free, lp.n_frames);
  const size_t smem = sizeof(double) * (elim_d > solve_d ? elim_d : solve_d);
  // every CTA must be resident at once (the CTAs wait for the deciding CTA's verdict): one CTA per SM, grid <= SMs;
  // `grid` = 2 G, two groups of G CTAs (schur_grid_x)
  if (lp.pdl) {
    cudaLaunchConfig_t cfg2 = {};
    cfg2.gridDim = dim3(grid); cfg2.blockDim = dim3(kSchurThreads); cfg2.dynamicSmemBytes = smem; cfg2.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg2.attrs = at; cfg2.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg2, k_schur_solve_x<LPP, TPW>, lp);
  }
  k_schur_solve_x<LPP, TPW><<<grid, kSchurThreads, smem, s>>>(lp);
  return cudaGetLastError();
}

cudaError_t launch_schur_solve(const LmParams& lp, int grid, int n_free, cudaStream_t s) {
  const int T = (6 * n_free + 8) >> 3, tiles = T * (T + 1) / 2;
  if (lp.xc.n_ranks > 1) {
    if (lp.n_frames <= 8) return tiles <= 3 * kSchurWarps ? launch_mode_x<8, 3>(lp, grid, n_free, s) : launch_mode_x<8, 12>(lp, grid, n_free, s);
    return tiles <= 3 * kSchurWarps ? launch_mode_x<16, 3>(lp, grid, n_free, s) : launch_mode_x<16, 12>(lp, grid, n_free, s);
  }
  if (lp.n_frames <= 8) return tiles <= 3 * kSchurWarps ? launch_mode<8, 3>(lp, grid, n_free, s) : launch_mode<8, 12>(lp, grid, n_free, s);
  return tiles <= 3 * kSchurWarps ? launch_mode<16, 3>(lp, grid, n_free, s) : launch_mode<16, 12>(lp, grid, n_free, s);
}

template <int TPW>
static cudaError_t launch_solve_only_t(const LmParams& lp, int n_free, cudaStream_t s) {
  static bool cfg[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!cfg[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_solve_only<TPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    if (e != cudaSuccess) return e;
    cfg[dev & 63] = true;
  }
  k_solve_only<TPW><<<1, kSchurThreads, sizeof(double) * solve_smem_doubles(n_free, lp.n_frames), s>>>(lp);
  return cudaGetLastError();
}

cudaError_t launch_solve_only(const LmParams& lp, int n_free, cudaStream_t s) {
  const int T = (6 * n_free + 8) >> 3, tiles = T * (T + 1) / 2;
  return tiles <= 3 * kSchurWarps ? launch_solve_only_t<3>(lp, n_free, s) : launch_solve_only_t<12>(lp, n_free, s);
}

}  // namespace pba
