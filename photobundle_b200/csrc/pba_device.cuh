// pba_device.cuh — device-side data layout shared by the kernels (sm_100a only).
//
// HBM layout of one window. "x2" = double-buffered: buffer st.cur holds the
// accepted point x, buffer st.eval_buf the candidate x+Δ being evaluated.
//   frames   u8  [F][rows][pitch]            pitch = roundup(cols,16) bytes   (Intensity)
//         or f32 [F][C][rows][pitch]         pitch = roundup(cols,16) floats  (generic)
//   cams     f64 x2 [F][6]                   world->camera [angle-axis, t]
//   points   f64 x2 [n][3]
//   desc     f32 [n][C*P]                    reference descriptors (exact: the reference
//                                            widens float channel values to double,
//                                            src/photobundle.cc:466-479)
//   obs_off  i32 [n+1], obs_frame i32 [nnz]  CSR visibility, window-local frame index
//   V,gp,W   f64 x2 [n][6] [n][3] [nnz][18]  point blocks / cross blocks at x
//   Xacc     f64 [F][27] + [8] + [8]         candidate's pose blocks (21 upper-tri U + 6 g_c),
//                                            scalars {cost, Σg_p², -, Σ|X|², s·g, sᵀHs, |Δ|², |x+Δ|²}
//                                            and one max|g_p| slot per rank; accumulated by K_A with
//                                            fp64 atomics (one per CTA per entry); this is the buffer
//                                            that is all-reduced when a window spans several GPUs
//   Ucur     f64 [F][27]                     pose blocks of the accepted point
//   S        f64 [N][ld]                     reduced-system accumulator P = sum Zt Ztᵀ in free-camera index space
//                                            (N = 6 n_free, ld = reduced_ld(N), upper triangle, column N = rhs part)
//   state    LmState x2                      ping-pong: K_B reads one, writes the other
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pba {

constexpr int kMaxFrames = 16;
constexpr int kMaxD = 6 * kMaxFrames;
constexpr int kObsBatch = 8;          // observations whose geometry / footprints are staged together
constexpr int kStageSlots = 8;        // (observation, channel) footprints staged together
constexpr int kPoseConst = 36;        // doubles per frame, see pose_consts()
constexpr int kUStride = 27;          // 21 upper-tri U + 6 g_c
constexpr int kSchurThreads = 256;
constexpr int kEacc = 8;
constexpr int kMaxRanks = 8;         // max|g_p| slots in Xacc (one per rank)

struct Frames {
  const uint8_t* u8;   // non-null: Intensity planes, gradients formed in-kernel
  const float* f32;    // non-null: generic fp32 channel planes
  int rows, cols, pitch, n_channels;
  size_t plane;        // elements per plane (rows * pitch)
};

// Same layout as pba_iteration_summary (include/pba_b200.h).
struct IterSummary {
  int32_t iteration, step_is_valid, step_is_nonmonotonic, step_is_successful;
  double cost, cost_change, gradient_max_norm, gradient_norm, step_norm, relative_decrease,
      trust_region_radius, eta, step_size;
  int32_t ls_f, ls_g, ls_it, linear_solver_iterations;
  double iteration_time_in_seconds, step_solver_time_in_seconds, cumulative_time_in_seconds;
};

enum MsgCode {
  kMsgNone = 0, kMsgGradTol = 1, kMsgParamTol = 2, kMsgFuncTol = 3, kMsgMaxIter = 4,
  kMsgMinRadius = 5, kMsgInvalidSteps = 6, kMsgXchgTimeout = 7
};

// Levenberg-Marquardt state machine, resident in HBM (Ceres TrustRegionMinimizer +
// LevenbergMarquardtStrategy semantics).
struct LmState {
  // options
  int max_num_iterations, max_invalid, jacobi_scaling;
  double function_tolerance, gradient_tolerance, parameter_tolerance;
  double initial_radius, max_radius, min_radius, min_relative_decrease, min_diag, max_diag;
  // problem
  int n_frames, fixed_frame, n_free, n_points, nnz;
  int free_index[kMaxFrames];     // frame -> index among optimised frames, -1 otherwise
  // dynamic
  int iteration;                  // iteration whose step is being computed / judged
  int cur, eval_buf;              // buffer of accepted x / buffer the next evaluation writes
  int done, termination_type, msg_code;
  int took_step;                  // the last decision accepted a candidate (or was iteration 0)
  double msg_a, msg_b;
  double radius, decrease_factor;
  int num_invalid, num_successful, num_unsuccessful, n_trace, num_evals;
  double x_cost, x_norm, initial_cost;
  double gmax, gnorm;
  // step under evaluation (written by the solve)
  int step_valid;
  double cam_sg, cam_sHs, cam_step_sq, cam_cand_sq;
  double scale_c[kMaxD];          // Jacobi column scaling of the pose columns (iteration 0)
  double step_c[kMaxD];           // trust-region step of the pose columns, scaled space
  unsigned long long xepoch;      // multi-GPU: epoch of the exchange this state waits for / publishes (monotone across solves)
  int n_xchg, n_respec;           // multi-GPU: exchanges so far, and how many LM iterations needed the second one (mis-speculation)
};

// Multi-GPU exchange over NVLink peer memory.  Points are sharded over the ranks; per LM iteration
// ONE exchange carries everything the replicated reduced solve needs.  K_B eliminates its shard under the two
// outcomes of the pending decision that are known in advance — (A) the candidate is accepted and the radius triples,
// (R) it is rejected and the radius shrinks by the current decrease factor — and the last CTA pushes
//   x1 slot = [pose blocks + cost scalars of the evaluation | P under A | P under R]
// into slot [own rank] of EVERY rank's buffer, sums everybody's evaluation part in rank order (bit-identical on all
// ranks), takes the decision, and then sums only the chosen hypothesis.  When neither hypothesis holds (iteration 0,
// or an accepted step with another radius) all CTAs — which have been waiting for the verdict — eliminate again
// with the true radius and a second exchange (x2 slot = [P]) follows.
// Every rank owns one exchange buffer (cudaMalloc + CUDA IPC) that all peers map:
//   x1  [2 parity][n_ranks][x1_n] cells,  x2  [2 parity][n_ranks][x2_n] cells
// A cell is 16 bytes = two self-validating 8-byte packets {32 data bits, 32-bit epoch tag} (the "LL" idea: an
// aligned 8-byte store is single-copy atomic, so a packet whose tag matches carries valid data and neither a fence
// nor a separate flag round trip is needed).  Slots alternate by epoch parity; the tag tells a fresh cell from the
// one written two epochs ago.
//   fr  u64 [n_ranks]                     rendezvous flags (start of a solve)
struct Xchg {
  int n_ranks, rank;              // n_ranks <= 1: single GPU, everything below unused
  int x1_n, x2_n;                 // cells per slot
  ulonglong2* x1[kMaxRanks];      // rank q's x1 region (peer mapping; [rank] = local)
  ulonglong2* x2[kMaxRanks];
  unsigned long long* fr[kMaxRanks];   // rendezvous flags (start of a solve)
  unsigned long long* verdict;    // local: (epoch << 3) | code, from the deciding CTA to the waiting CTAs of the same kernel
  int* error;                     // local: set when a wait timed out
};
enum XVerdict { kXHitA = 1, kXHitR = 2, kXDone = 3, kXRedo = 4, kXFail = 5 };

// K_A parameters (k_step.cu)
struct StepParams {
  Frames fr;
  int n_frames, fixed_frame, n_points, nnz;
  double fx, fy, cx, cy, huber;
  const LmState* st;         // null: plain evaluation of buffer 0 (pba_eval)
  double* cams;              // x2 [F][6]
  double* pts;               // x2 [n][3]
  const float* desc;         // [n][C*P]
  const int* obs_off;        // [n+1]
  const int* obs_frame;      // [nnz]
  const double* weights;     // [P]
  double* V;                 // x2 [n][6]   upper triangle 00 01 02 11 12 22
  double* gp;                // x2 [n][3]
  double* W;                 // x2 [nnz][18] row-major 6x3
  double* Xacc;              // [F][27] + [8] + [kMaxRanks]
  int rank;                  // which max|g_p| slot this process owns
  const double* scale_p;     // [n][3]   (back-substitution)
  const double* Vinv;        // [n][6]
  double* obs_sqnorm;        // optional [nnz]
  double* residuals;         // optional [nnz][C*P]
  int pdl;                   // launched with programmatic stream serialization: wait for the previous kernel before reading its output
};

// K_B parameters (k_schur_solve.cu)
struct LmParams {
  const LmState* st_in;
  LmState* st_out;
  IterSummary* trace;
  unsigned int* ticket;      // last-CTA detection
  int n_frames, n_points, nnz;
  const int* obs_off;
  const int* obs_frame;
  double* cams;              // x2
  const double* V;           // x2
  const double* gp;          // x2
  const double* W;           // x2
  double* Xacc;              // [F][27] + [8] + [kMaxRanks]  (candidate)
  double* Ucur;              // [F][27]                       (accepted point)
  int split;                 // 1: multi-GPU — the reduced system is all-reduced before a separate solve kernel
  double* scale_p;           // [n][3]
  double* Vinv;              // [n][6]
  double* S;                 // [N][reduced_ld(N)], see k_schur_solve.cu; multi-GPU: three of them (hypotheses A, R, re-elimination), s_cap apart
  size_t s_cap;
  double* Vinv2;             // multi-GPU: [2][n][6], (Vs + D²)^-1 under hypothesis A / R
  unsigned long long* dbg;   // optional: globaltimer stamps of the last CTA {start, decided, schur done, solved, end}
  unsigned long long* stamps;// [capacity][2]: globaltimer at the start / end of the K_B launch that takes decision number num_evals
  unsigned long long cond;   // non-zero: cudaGraphConditionalHandle of the device-side LM loop, cleared when done
  Xchg xc;                   // multi-GPU exchange over peer memory (xc.n_ranks > 1), else split/NCCL or single GPU
  int pdl;                   // launched with programmatic stream serialization (see pdl_wait)
  int speculate;             // multi-GPU: eliminate under both outcomes of the pending decision while the evaluation sums travel (one
                             // exchange per iteration); 0: decide first, then eliminate once with all CTAs (two exchanges) - large shards
};

// ---- programmatic dependent launch (PDL): the LM loop is a chain K_B -> K_A -> K_B ... of dependent kernels; a
// kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start (launch latency, CTA
// scheduling, everything that does not read its predecessor's output) while the predecessor's tail still runs,
// and blocks in pdl_wait() until the predecessor has completed and its writes are visible.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- system-scope flag helpers for the peer-memory exchange ---------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long kXchgTimeoutNs = 4000000000ull;   // a peer that never arrives must not hang the GPU
// LL cells: store / try-load one double tagged with the low 32 bits of the epoch
__device__ __forceinline__ void ll_store(ulonglong2* cell, double v, unsigned long long epoch) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v), t = epoch << 32;
  const unsigned long long p0 = t | (b & 0xffffffffull), p1 = t | (b >> 32);
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(cell), "l"(p0), "l"(p1) : "memory");
}
__device__ __forceinline__ bool ll_try_load(const ulonglong2* cell, unsigned long long epoch, double& v) {
  unsigned long long p0, p1;
  asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(p0), "=l"(p1) : "l"(cell) : "memory");
  const unsigned long long t = epoch & 0xffffffffull;
  v = __longlong_as_double((long long)((p0 & 0xffffffffull) | (p1 << 32)));
  return (p0 >> 32) == t && (p1 >> 32) == t;
}
// Blocking load of one cell; returns false on timeout.
__device__ __forceinline__ bool ll_load(const ulonglong2* cell, unsigned long long epoch, double& v) {
  if (ll_try_load(cell, epoch, v)) return true;
  const unsigned long long t0 = globaltimer_ns();
  while (!ll_try_load(cell, epoch, v))
    if (globaltimer_ns() - t0 > kXchgTimeoutNs) return false;
  return true;
}
// Thread q < n waits until flags[q] >= epoch; returns false on timeout.  Call from >= n threads, then barrier.
__device__ __forceinline__ bool xchg_wait(const unsigned long long* flag, unsigned long long epoch) {
  const unsigned long long t0 = globaltimer_ns();
  while (ld_acquire_sys(flag) < epoch) {
    if (globaltimer_ns() - t0 > kXchgTimeoutNs) return false;
    __nanosleep(64);
  }
  return true;
}

// Row stride of the reduced system S [N][ld] (global accumulator and the solver's shared-memory copy): the
// smallest ld = 4 (mod 16) with ld >= N + 8 — rows 32 bytes apart modulo 128 make the DMMA fragment loads of
// the trailing update bank-conflict free, and 8x8 tiles may overhang the last column.
__host__ __device__ constexpr int reduced_ld(int N) { return ((N + 19) >> 4) * 16 + 4; }
__host__ __device__ constexpr size_t reduced_capacity(int max_frames) { return (size_t)(6 * max_frames) * reduced_ld(6 * max_frames); }

// launchers
cudaError_t launch_k_step(const StepParams& prm, int radius, cudaStream_t stream);
int schur_grid(int n_points, int sm_count);
int schur_grid_x(int n_points, int sm_count);   // multi-GPU kernel: two groups of CTAs (one per hypothesis), all resident at once
cudaError_t launch_schur_solve(const LmParams& lp, int grid, int n_free, cudaStream_t stream);   // lp.xc.n_ranks > 1: the multi-GPU kernel
cudaError_t launch_solve_only(const LmParams& lp, int n_free, cudaStream_t stream);   // split mode, after the all-reduce of S
// Copies of the reduced-system accumulator S that the CTAs of the single-GPU K_B spread their atomics over (CTA b adds
// into copy b % kSReplicas; the solving CTA sums the copies): every entry of S otherwise receives one fp64 atomic from
// EVERY CTA at about the same time, and same-address atomics are performed one after the other in L2.
constexpr int kSReplicas = 4;
cudaError_t launch_publish(const LmState* st, double* cams, double* pts, int n_frames, int n_points, cudaStream_t stream);   // buffer `cur` -> buffer 0
cudaError_t launch_rendezvous(const Xchg& xc, unsigned long long epoch, cudaStream_t stream);   // device-side barrier across the ranks

// descriptor channels (k_prep.cu): type 1 = IntensityAndGradient (3 planes), 2 = BitPlanes (8 planes)
cudaError_t launch_channels(int descriptor_type, const uint8_t* src, int rows, int cols, int spitch, uint8_t* scratch_a,
                            uint8_t* scratch_b, float* dst, int dpitch, size_t dplane, cudaStream_t stream);
cudaError_t launch_saliency(const float* planes, int n_channels, int rows, int cols, int pitch, size_t plane, float* out,
                            cudaStream_t stream);
cudaError_t launch_extract_patches(const float* planes, int n_channels, int rows, int cols, int pitch, size_t plane, int radius,
                                   int n, const int* xy, double* desc, cudaStream_t stream);
// addFrame front end (k_prep.cu)
cudaError_t launch_associate(const uint8_t* img, int rows, int cols, int pitch, int n, const double* xyz, const float* ref_patch,
                             const float* ref_norm, const double* Tc, const double* K, int border, float* score, int* rc,
                             cudaStream_t stream);
cudaError_t launch_candidates(const float* sal, uint8_t* mask, const float* depth, int rows, int cols, int border, int nms,
                              int n_masked, const int* masked_rc, int mask_radius, double min_depth, double max_depth, int capacity,
                              int* count, int* cand_rc, float* cand_sal, cudaStream_t stream);
cudaError_t launch_pyrdown_u8(const uint8_t* src, int srows, int scols, int spitch, uint8_t* dst, int dpitch,
                              cudaStream_t stream);

}  // namespace pba
