// k_step.cu — K_A: back-substitution + per-(point, frame) photometric residual + analytic
// Jacobian + block accumulation, one hand-written sm_100a kernel batched over all
// points x observing frames.
//
// Replaces (reference, /root/reference):
//   DescriptorError::operator()<Jet<double,9>>      src/photobundle.cc:696-727
//   SampleWithDerivative / SampleLinear             src/sample_eigen.h:33-126
//   ceres::Chain<float,2,Jet>::Rule                 src/jet_extras.h:74-111
//   imgradient (central difference, zero borders)   src/imgproc.cc:27-106
//   Calibration::project                            src/calibration.h:33-38
//   ceres::AngleAxisRotatePoint, HuberLoss + Corrector, the J^T J / J^T r products Ceres'
//   SchurEliminator forms per residual block, and SchurEliminator::BackSubstitute.
//
// Work decomposition: one warp per scene point.
//   (B) back-substitution of the point (when a step is pending): Δp = -(V+D²)^-1 (g_p + Wᵀ Δc),
//       candidate X = X + Δp, and the point's share of the model cost change / step norms;
//   (G) lane i forms the geometry of observation i in fp64: Xc = R(w)X + t, (u,v) and the
//       2x9 matrix A = d(u,v)/d[w t X]  (dual numbers are not needed: every one of the
//       (2r+1)^2 Jacobian rows is -w_j*[gx gy]*A);
//   (L) the warp stages each observation's (2r+5)x(2r+5) image footprint into shared memory
//       with ONE coalesced, 4-byte/16-byte aligned load instruction (8 footprints in flight),
//       converting uint8 -> fp32 on the way (DescriptorFrame::Create's cast);
//   (S) lane j samples patch pixel j of FOUR observations at a time (independent dependency
//       chains, ILP 4).  The taps of I come from the footprint, the taps of Gx, Gy are formed in
//       registers from the same footprint (exactly the 0.5*(a-b) the reference precomputes
//       into planes), and all three interpolations use the reference's float/double promotion
//       sequence, so residuals AND sampled gradients are bit-identical to the CPU path given
//       the same (u,v);
//   (R) six fp64 patch sums  s=Σr², G=Σ w²ggᵀ (3), b=Σ w r g (2)  per observation are reduced
//       through a shared-memory transpose (24 sums of a 4-observation group land in 24 lanes),
//       the Huber corrector of the four observations is evaluated in four lanes at once, and
//       the whole 9x9 block of each observation,  rho' Aᵀ G A  and  -rho' Aᵀ b, is expanded by
//       27 lanes in parallel in fp64;
//   the point's V (3x3) and g_p live in registers across its frames, W (6x3) goes out once
//   per observation, pose blocks U (6x6) / g_c are summed per CTA in shared memory and added
//   to the global accumulators with one fp64 atomic per entry per CTA.
// Out-of-image / border observations, multi-channel descriptors and patches with more than
// 32 pixels take the generic per-observation path with the reference's clamp and
// zero-gradient-border rules (src/sample_eigen.h:38-46, src/imgproc.cc:34-43).

#include "pba_device.cuh"

#include <climits>
#include <cfloat>
#include <type_traits>

// CTA geometry.  The bench window is 4000 points = 4000 warps of work; two 14-warp CTAs per SM
// (<= 72 registers, uint8 footprints in shared memory) keep EVERY point resident in a single wave
// on 148 SMs, which matters more than per-warp ILP: with 16 warps/SM the kernel ran 1.7 waves and
// the half-empty second wave cost a full per-point latency (profiles/r01e_kstep_trace.txt).
#ifndef K_STEP_WARPS_U8
#define K_STEP_WARPS_U8 14
#endif
#ifndef K_STEP_WARPS_F32
#define K_STEP_WARPS_F32 8
#endif
#ifndef K_STEP_GROUP
#define K_STEP_GROUP 2      // observations sampled together per warp (ILP)
#endif

namespace pba {

template <bool U8> struct Geo { static constexpr int WARPS = U8 ? K_STEP_WARPS_U8 : K_STEP_WARPS_F32; };

// K_STEP_TRACE: per-warp clock64 stamps at the phase boundaries (scratch builds only).
#ifdef K_STEP_TRACE
__device__ long long g_kstep_trace[4096 * 16];
#define KTRACE(slot) do { if (lane == 0 && p < 4096) g_kstep_trace[p * 16 + (slot)] = clock64(); } while (0)
#else
#define KTRACE(slot) do {} while (0)
#endif

// ---- per-frame pose constants (computed once per CTA) ---------------------------------
//  [0..2] unit axis k (or the raw angle-axis when tiny)   [3..5] t   [6] cos  [7] sin
//  [8] tiny flag   [9..17] R   [18..26] Rj   [27..35] M,  with
//  d(Xc)/dw = -Rj [X]x M ;  M = (w wᵀ + (Rᵀ - I)[w]x)/θ² ; tiny angle: Rj = M = I, R = I+[w]x
// One thread per (frame, matrix element e = 3a+b): the nine threads of a frame evaluate the angle,
// its sine/cosine and the axis redundantly (the latency of that chain is what costs, not its
// throughput) and then one element each of R, Rj and M; thread e = 0 also writes the vector part.
__device__ __forceinline__ double sel3(double x0, double x1, double x2, int i) { return i == 0 ? x0 : (i == 1 ? x1 : x2); }
// element (r, c) of the cross-product matrix [v]x  (selects instead of indexed local arrays: local memory
// means L2 here, the shared-memory carve-out leaves almost no L1)
__device__ __forceinline__ double skew_el(double v0, double v1, double v2, int r, int c) {
  if (r == c) return 0.0;
  const double m = sel3(v0, v1, v2, 3 - r - c);
  return ((c - r + 3) % 3 == 1) ? -m : m;
}
__device__ void pose_consts(const double* cam, double* pc, int e) {
  const double w0 = cam[0], w1 = cam[1], w2 = cam[2];
  const double theta2 = __dadd_rn(__dadd_rn(__dmul_rn(w0, w0), __dmul_rn(w1, w1)), __dmul_rn(w2, w2));
  const int a = e / 3, b = e - 3 * a;
  if (theta2 > DBL_EPSILON) {
    const double theta = sqrt(theta2);
    double s, c;
    sincos(theta, &s, &c);
    const double k0 = w0 / theta, k1 = w1 / theta, k2 = w2 / theta;
    const double ka = sel3(k0, k1, k2, a), kb = sel3(k0, k1, k2, b);
    // column a of R, and R[a][b]:  R = c I + s [k]x + (1 - c) k k^T
    const double R0a = (a == 0 ? c : 0.0) + s * skew_el(k0, k1, k2, 0, a) + (1.0 - c) * k0 * ka;
    const double R1a = (a == 1 ? c : 0.0) + s * skew_el(k0, k1, k2, 1, a) + (1.0 - c) * k1 * ka;
    const double R2a = (a == 2 ? c : 0.0) + s * skew_el(k0, k1, k2, 2, a) + (1.0 - c) * k2 * ka;
    const double Rab = (a == b ? c : 0.0) + s * skew_el(k0, k1, k2, a, b) + (1.0 - c) * ka * kb;
    double acc = sel3(w0, w1, w2, a) * sel3(w0, w1, w2, b);
    acc += (R0a - (a == 0 ? 1.0 : 0.0)) * skew_el(w0, w1, w2, 0, b);
    acc += (R1a - (a == 1 ? 1.0 : 0.0)) * skew_el(w0, w1, w2, 1, b);
    acc += (R2a - (a == 2 ? 1.0 : 0.0)) * skew_el(w0, w1, w2, 2, b);
    pc[9 + e] = Rab; pc[18 + e] = Rab; pc[27 + e] = acc / theta2;
    if (e == 0) {
      const double ti = 1.0 / theta;
      pc[0] = __dmul_rn(w0, ti); pc[1] = __dmul_rn(w1, ti); pc[2] = __dmul_rn(w2, ti);
      pc[3] = cam[3]; pc[4] = cam[4]; pc[5] = cam[5]; pc[6] = c; pc[7] = s; pc[8] = 0.0;
    }
  } else {
    const double id = (a == b) ? 1.0 : 0.0;
    pc[9 + e] = id + skew_el(w0, w1, w2, a, b); pc[18 + e] = id; pc[27 + e] = id;
    if (e == 0) {
      pc[0] = w0; pc[1] = w1; pc[2] = w2; pc[3] = cam[3]; pc[4] = cam[4]; pc[5] = cam[5];
      pc[6] = 1.0; pc[7] = 0.0; pc[8] = 1.0;
    }
  }
}

// ---- slow path: reference sampler semantics tap by tap ----------------------------------
template <bool U8>
__device__ __forceinline__ float px_at(const Frames& fr, int f, int k, int y, int x) {
  if (U8) return (float)fr.u8[(size_t)f * fr.plane + (size_t)y * fr.pitch + x];
  return fr.f32[((size_t)f * fr.n_channels + k) * fr.plane + (size_t)y * fr.pitch + x];
}
template <bool U8>
__device__ __forceinline__ void grad_at(const Frames& fr, int f, int k, int y, int x, float& gx, float& gy) {
  if (y <= 0 || y >= fr.rows - 1 || x <= 0 || x >= fr.cols - 1) { gx = 0.f; gy = 0.f; return; }
  gx = __fmul_rn(0.5f, __fsub_rn(px_at<U8>(fr, f, k, y, x + 1), px_at<U8>(fr, f, k, y, x - 1)));
  gy = __fmul_rn(0.5f, __fsub_rn(px_at<U8>(fr, f, k, y + 1, x), px_at<U8>(fr, f, k, y - 1, x)));
}
__device__ __forceinline__ void init_axis(float s, int size, int& i1, int& i2, float& d) {
  // static_cast<int>(float) of the reference binary = cvttss2si: NaN / out of range -> INT_MIN
  const int ix = (s > -2147483648.0f && s < 2147483648.0f) ? __float2int_rz(s) : INT_MIN;
  if (ix < 0) { i1 = 0; i2 = 0; d = 1.0f; }
  else if (ix > size - 2) { i1 = size - 1; i2 = size - 1; d = 1.0f; }
  else { i1 = ix; i2 = ix + 1; d = __fsub_rn((float)i2, s); }
}
// sample_eigen.h:82-83 with its C++ promotions: dx*a11 in float, (1.0-dx) a double,
// (1-dy) a float, the sum rounded to float once.
__device__ __forceinline__ float bilerp(float dx, float dy, double omdx, float omdy,
                                        float a11, float a12, float a21, float a22) {
  const double top = __dadd_rn((double)__fmul_rn(dx, a11), __dmul_rn(omdx, (double)a12));
  const double bot = __dadd_rn((double)__fmul_rn(dx, a21), __dmul_rn(omdx, (double)a22));
  return __double2float_rn(__dadd_rn(__dmul_rn((double)dy, top), __dmul_rn((double)omdy, bot)));
}

__device__ __forceinline__ double quad(const double* A, int a, int b, double G11, double G12, double G22) {
  const double A0a = A[a], A1a = A[9 + a], A0b = A[b], A1b = A[9 + b];
  return A0a * (G11 * A0b + G12 * A1b) + A1a * (G12 * A0b + G22 * A1b);
}

template <int R> struct Foot {
  static constexpr int SIDE = 2 * R + 1;
  static constexpr int P = SIDE * SIDE;
  static constexpr int ROWS = 2 * R + 5;                 // taps + gradient halo + rounding slack
  static constexpr int NW = (2 * R + 11) / 4;            // aligned 4-element words per row
  static constexpr int W = 4 * NW;                       // floats per staged row
  static constexpr int WORDS = ROWS * NW;
  static constexpr int ROUNDS = (WORDS + 31) / 32;
  static constexpr int FLOATS = ROWS * W;
};


constexpr int kG = K_STEP_GROUP;
constexpr int kRedStride = 6 * kG + 2;   // doubles per lane row of the reduction transpose (6 sums x group + pad, 16 B aligned)

// ---- shared memory carve-up --------------------------------------------------------------
// per CTA : pose consts [F][36] f64 | sstep [F][6] f64 (scale_c*step_c) | E [warps][8] f64 | weights [P] f64
// per warp: geometry [8][20] f64 | reduction transpose [25][6G+2] f64 | patch sums [8][6] f64 |
//           pose-block accumulators [F][27] f64 | ints [8] int4 | frames [16] i32 |
//           footprints [8][ROWS][W] f32
// USHARE: warps that share one pose-block accumulator (1: private, plain adds; 2: pairs, shared-memory fp64 atomics -
// what lets 16-frame windows keep 14 warps per CTA and two CTAs per SM)
template <int R, bool U8, int WARPS = Geo<U8>::WARPS, int USHARE = 1>
__host__ __device__ constexpr size_t k_step_smem_bytes(int n_frames) {
  // footprints: raw uint8 (ROWS x W bytes) on the Intensity path, fp32 otherwise
  constexpr size_t fp_bytes = (size_t)kStageSlots * Foot<R>::FLOATS * (U8 ? 1 : 4);
  return sizeof(double) * ((size_t)n_frames * (kPoseConst + 6) + WARPS * kEacc + ((Foot<R>::P + 1) & ~1)) +
         (size_t)WARPS * (sizeof(double) * (kObsBatch * 20 + (R == 1 ? 27 : 25) * kRedStride + 6 * kObsBatch) +
                          sizeof(int4) * kObsBatch + sizeof(int) * kMaxFrames + ((fp_bytes + 15) / 16) * 16) +
         (size_t)(WARPS / USHARE) * sizeof(double) * ((size_t)n_frames * kUStride + (n_frames & 1));
}

struct Sums { double s, G11, G12, G22, b1, b2; };

__device__ __forceinline__ void warp_reduce(Sums& q) {   // generic path only
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    q.s += __shfl_xor_sync(0xffffffffu, q.s, m);
    q.G11 += __shfl_xor_sync(0xffffffffu, q.G11, m);
    q.G12 += __shfl_xor_sync(0xffffffffu, q.G12, m);
    q.G22 += __shfl_xor_sync(0xffffffffu, q.G22, m);
    q.b1 += __shfl_xor_sync(0xffffffffu, q.b1, m);
    q.b2 += __shfl_xor_sync(0xffffffffu, q.b2, m);
  }
}

// ceres::HuberLoss (rho'' <= 0 -> the Corrector is a plain sqrt(rho') scaling of r and J).
// rsqrt-based: a/sqrt(s) and sqrt(s) = s*rsqrt(s) to ~1 ulp (parity tolerance on cost is 1e-9).
__device__ __forceinline__ void huber_rho(double a, double s, double& rho0, double& rho1) {
  rho0 = s; rho1 = 1.0;
  if (a > 0.0 && s > a * a) {
    const double rs = rsqrt(s);
    rho0 = 2.0 * a * (s * rs) - a * a;
    rho1 = fmax(DBL_MIN, a * rs);
  }
}

// One patch pixel of one observation from the staged footprint (all taps interior):
// I, Gx, Gy exactly as SampleLinear returns them.
template <int R, class T>
__device__ __forceinline__ void sample_fast(const T* __restrict__ fp, int r0, int cb, double u, double v,
                                            double pdx, double pdy, float& I1, float& gx, float& gy) {
  using FT = Foot<R>;
  const float su = __double2float_rn(__dadd_rn(u, pdx));
  const float sv = __double2float_rn(__dadd_rn(v, pdy));
  const int ix = __float2int_rz(su), iy = __float2int_rz(sv);
  const float dx = __fsub_rn((float)(ix + 1), su), dy = __fsub_rn((float)(iy + 1), sv);
  const T* q = fp + (iy - r0) * FT::W + (ix - cb);
  const float a11 = (float)q[0], a12 = (float)q[1], a21 = (float)q[FT::W], a22 = (float)q[FT::W + 1];
  const float l1 = (float)q[-1], r1 = (float)q[2], l2 = (float)q[FT::W - 1], r2 = (float)q[FT::W + 2];
  const float t1 = (float)q[-FT::W], t2 = (float)q[-FT::W + 1], u1 = (float)q[2 * FT::W], u2 = (float)q[2 * FT::W + 1];
  const double omdx = __dsub_rn(1.0, (double)dx);
  const float omdy = __fsub_rn(1.0f, dy);
  I1 = bilerp(dx, dy, omdx, omdy, a11, a12, a21, a22);
  gx = bilerp(dx, dy, omdx, omdy, __fmul_rn(0.5f, __fsub_rn(a12, l1)), __fmul_rn(0.5f, __fsub_rn(r1, a11)),
              __fmul_rn(0.5f, __fsub_rn(a22, l2)), __fmul_rn(0.5f, __fsub_rn(r2, a21)));
  gy = bilerp(dx, dy, omdx, omdy, __fmul_rn(0.5f, __fsub_rn(a21, t1)), __fmul_rn(0.5f, __fsub_rn(a22, t2)),
              __fmul_rn(0.5f, __fsub_rn(u1, a11)), __fmul_rn(0.5f, __fsub_rn(u2, a12)));
}

// ---- exact small-integer -> float/double conversions on the FMA / FP64 pipes --------------------
// The F2F/I2F conversion instructions run on the quarter-rate XU pipe, which is what bounds the
// sampling phase; for the 8-bit taps of an Intensity frame the same values are produced exactly by
// the classic magic-number constructions (no rounding anywhere, so parity is unaffected).
__device__ __forceinline__ float u8_to_f32(int b) {            // b in [0, 255]
  return __fsub_rn(__int_as_float(0x4B000000 | b), 8388608.0f);
}
__device__ __forceinline__ double u8_to_f64(int b) {           // b in [0, 255]
  return __dsub_rn(__hiloint2double(0x43300000, b), 4503599627370496.0);
}
__device__ __forceinline__ float half_diff_f32(int d) {        // 0.5f * d, d in [-255, 255]
  return __fmaf_rn(__int_as_float(0x4B400000 + d), 0.5f, -6291456.0f);
}
__device__ __forceinline__ double diff_f64(int d) {            // (double)d, d in [-255, 255]
  return __dsub_rn(__hiloint2double(0x43300000, d ^ (int)0x80000000), 4503599627370496.0 + 2147483648.0);
}
// bilerp() with the two right-hand taps already in double and the 1-dx weight pre-scaled by the
// caller (wd*da12 == (1-dx)*(double)a12 bit for bit: power-of-two scalings commute with rounding).
__device__ __forceinline__ float bilerp_m(float dx, double dyd, double omdyd, double wd, float fa11, double da12,
                                          float fa21, double da22) {
  const double top = __dadd_rn((double)__fmul_rn(dx, fa11), __dmul_rn(wd, da12));
  const double bot = __dadd_rn((double)__fmul_rn(dx, fa21), __dmul_rn(wd, da22));
  return __double2float_rn(__dadd_rn(__dmul_rn(dyd, top), __dmul_rn(omdyd, bot)));
}
// sample_fast for uint8 footprints: identical results, 8-bit taps kept as integers.
template <int R>
__device__ __forceinline__ void sample_fast_u8(const uint8_t* __restrict__ fp, int r0, int cb, double u, double v,
                                               double pdx, double pdy, float& I1, float& gx, float& gy) {
  using FT = Foot<R>;
  const float su = __double2float_rn(__dadd_rn(u, pdx));
  const float sv = __double2float_rn(__dadd_rn(v, pdy));
  const int ix = __float2int_rz(su), iy = __float2int_rz(sv);
  const float dx = __fsub_rn((float)(ix + 1), su), dy = __fsub_rn((float)(iy + 1), sv);
  const uint8_t* q = fp + (iy - r0) * FT::W + (ix - cb);
  const int a11 = q[0], a12 = q[1], a21 = q[FT::W], a22 = q[FT::W + 1];
  const int l1 = q[-1], r1 = q[2], l2 = q[FT::W - 1], r2 = q[FT::W + 2];
  const int t1 = q[-FT::W], t2 = q[-FT::W + 1], u1 = q[2 * FT::W], u2 = q[2 * FT::W + 1];
  const double omdx = __dsub_rn(1.0, (double)dx);
  const double homdx = __dmul_rn(0.5, omdx);
  const double dyd = (double)dy, omdyd = (double)__fsub_rn(1.0f, dy);
  I1 = bilerp_m(dx, dyd, omdyd, omdx, u8_to_f32(a11), u8_to_f64(a12), u8_to_f32(a21), u8_to_f64(a22));
  gx = bilerp_m(dx, dyd, omdyd, homdx, half_diff_f32(a12 - l1), diff_f64(r1 - a11), half_diff_f32(a22 - l2), diff_f64(r2 - a21));
  gy = bilerp_m(dx, dyd, omdyd, homdx, half_diff_f32(a21 - t1), diff_f64(a22 - t2), half_diff_f32(u1 - a11), diff_f64(u2 - a12));
}

// ---- asynchronous global -> shared copies (LDGSTS) -----------------------------------------------
__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Block expansion of one observation from its loss-scaled patch sums: pose block -> per-warp
// smem accumulator, W -> HBM, V / g_p -> the caller's register accumulator.
__device__ __forceinline__ void emit_blocks(double dG11, double dG12, double dG22, double db1, double db2,
                                            const double* __restrict__ g, int f, bool free_cam, int lane, int e1a, int e1b,
                                            int e2a, int e2b, double* s_U_w, double* __restrict__ outW_o, double& acc_pt,
                                            const bool shared_u = false) {
  const double* A = g + 2;
  if (lane < 27) {
    if (free_cam) {
      const double v1 = lane < 21 ? quad(A, e1a, e1b, dG11, dG12, dG22) : -(A[e1a] * db1 + A[9 + e1a] * db2);
      if (shared_u) atomicAdd(s_U_w + f * kUStride + lane, v1);
      else s_U_w[f * kUStride + lane] += v1;   // warp-private: no atomics
    }
    const double v2 = lane < 24 ? quad(A, e2a, e2b, dG11, dG12, dG22) : -(A[e2a] * db1 + A[9 + e2a] * db2);
    if (lane < 18) outW_o[lane] = free_cam ? v2 : 0.0;
    else acc_pt += v2;
  }
}

// The two block entries lane `lane` owns for one observation, straight-line (no divergence, no stores): the
// caller evaluates a whole batch of observations first - independent chains the scheduler can interleave -
// and only then touches memory.
__device__ __forceinline__ void block_entries(const double* __restrict__ t, const double* __restrict__ g, int lane, int e1a,
                                              int e1b, int e2a, int e2b, double& v1, double& v2) {
  const double* A = g + 2;
  const double dG11 = t[1], dG12 = t[2], dG22 = t[3], db1 = t[4], db2 = t[5];
  const double A0a = A[e1a], A1a = A[9 + e1a], A0b = A[e1b], A1b = A[9 + e1b];
  const double q1 = A0a * (dG11 * A0b + dG12 * A1b) + A1a * (dG12 * A0b + dG22 * A1b);
  const double l1 = -(A0a * db1 + A1a * db2);
  v1 = lane < 21 ? q1 : l1;
  const double B0a = A[e2a], B1a = A[9 + e2a], B0b = A[e2b], B1b = A[9 + e2b];
  const double q2 = B0a * (dG11 * B0b + dG12 * B1b) + B1a * (dG12 * B0b + dG22 * B1b);
  const double l2 = -(B0a * db1 + B1a * db2);
  v2 = lane < 24 ? q2 : l2;
}

// NCH: compile-time channel count (1 = Intensity, the north-star descriptor); 0 = runtime count.
// WARPS: warps per CTA (two CTAs per SM); the default keeps a 4 000-point window in one wave, wide windows
// (per-warp pose-block accumulators grow with the frame count) use fewer so that two CTAs still fit.
template <int R, bool U8, int NCH, int WARPS = Geo<U8>::WARPS, int USHARE = 1>
__global__ void __launch_bounds__(WARPS * 32, 2) k_step(const StepParams prm) {
  static_assert(WARPS % USHARE == 0, "warps per pose-block accumulator");
  using FT = Foot<R>;
  using FPT = typename std::conditional<U8, uint8_t, float>::type;   // footprint element in shared memory
  constexpr int P = FT::P;
  constexpr int PR = (P + 31) / 32;             // pixel rounds per lane
  constexpr bool kQuad = (NCH == 1 && PR == 1); // ILP-4 fast path available
  // 3x3 patches on the uint8 path (the reference's own KITTI configuration, config/kitti_stereo.cfg:22): a patch fills 9
  // of 32 lanes, so three observations are sampled side by side (lane / 9) and a group covers 6 observations
  constexpr bool kPack3 = (R == 1 && U8 && NCH == 1 && kG == 2);
  constexpr bool kAsyncStage = (NCH == 1);      // footprints staged with cp.async ahead of (G2)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int F = prm.n_frames;
  const int C = NCH ? NCH : prm.fr.n_channels;
  const int CP = C * P;
  const bool want_res = prm.residuals != nullptr;

  const LmState* st = prm.st;
#ifdef K_STEP_TRACE
  const long long t_entry = clock64();
#endif
  if (prm.pdl) { pdl_wait(); pdl_launch_dependents(); }   // K_B's tail has finished; the next K_B may queue behind us
  if (st && st->done) return;
#ifdef K_STEP_TRACE
  const int p_tr = blockIdx.x * WARPS + warp;
  { const int p = p_tr; if (lane == 0 && p < 4096) { g_kstep_trace[p * 16 + 0] = t_entry; g_kstep_trace[p * 16 + 7] = clock64(); } }
#endif
  const int buf = st ? st->eval_buf : 0;
  const int cur = st ? st->cur : 0;
  const bool backsub = st && st->iteration > 0;
  const double* cams = prm.cams + (size_t)buf * F * 6;
  double* pts_out = prm.pts + (size_t)buf * prm.n_points * 3;
  const double* pts_cur = prm.pts + (size_t)cur * prm.n_points * 3;
  double* outV = prm.V + (size_t)buf * prm.n_points * 6;
  double* outgp = prm.gp + (size_t)buf * prm.n_points * 3;
  double* outW = prm.W + (size_t)buf * prm.nnz * 18;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_pose = reinterpret_cast<double*>(smem_raw);                       // [F][36]
  double* s_sstep = s_pose + F * kPoseConst;                                  // [F][6]
  double* s_E = s_sstep + F * 6;                                              // [warps][8] (all offsets even: 16 B alignment)
  double* s_wts = s_E + WARPS * kEacc;                                        // [P (+1)] patch weights
  // one contiguous block per warp, every region but the last at a compile-time offset of its base: a single
  // live address register instead of one per region (the kernel sits at the 72-register ceiling)
  constexpr int kFpBytes = ((kStageSlots * FT::FLOATS * (int)sizeof(FPT) + 15) / 16) * 16;
  constexpr int kOffGeo = 0;                                                  // [8][20] f64
  constexpr int kOffRed = kOffGeo + 8 * kObsBatch * 20;                       // [25][6G+2] f64
  constexpr int kOffTot = kOffRed + 8 * (R == 1 ? 27 : 25) * kRedStride;   // (27 rows: three 3x3 patches side by side)
  //                      // [8][6] f64: patch sums of the batch's observations
  constexpr int kOffGi = kOffTot + 8 * 6 * kObsBatch;                         // [8] int4
  constexpr int kOffFrm = kOffGi + 16 * kObsBatch;                            // [16] i32
  constexpr int kOffFp = kOffFrm + 4 * kMaxFrames;                            // [8][ROWS][W] footprints
  constexpr int kOffU = kOffFp + kFpBytes;                                    // [F][27] f64: this warp's pose blocks
  static_assert(kOffGi % 16 == 0 && kOffFp % 16 == 0 && kOffU % 8 == 0, "per-warp shared-memory layout alignment");
  const int ustride = F * kUStride + (F & 1);
  const int warp_bytes = kOffU + (USHARE == 1 ? 8 * ustride : 0);
  unsigned char* s_warp0 = reinterpret_cast<unsigned char*>(s_wts + ((P + 1) & ~1));
  unsigned char* wbase = s_warp0 + warp * warp_bytes;
  double* s_geo_w = reinterpret_cast<double*>(wbase + kOffGeo);
  double* s_red_w = reinterpret_cast<double*>(wbase + kOffRed);
  double* s_tot_w = reinterpret_cast<double*>(wbase + kOffTot);
  int4* s_gi_w = reinterpret_cast<int4*>(wbase + kOffGi);
  int* s_frm_w = reinterpret_cast<int*>(wbase + kOffFrm);
  FPT* s_fp_w = reinterpret_cast<FPT*>(wbase + kOffFp);
  // pose-block accumulator: at the end of the warp's own region, or (USHARE > 1) one per group of warps behind all regions
  double* s_U_w = USHARE == 1 ? reinterpret_cast<double*>(wbase + kOffU)
                              : reinterpret_cast<double*>(s_warp0 + WARPS * warp_bytes) + (warp / USHARE) * ustride;

  if (threadIdx.x < 9 * F) pose_consts(cams + 6 * (threadIdx.x / 9), s_pose + (threadIdx.x / 9) * kPoseConst, threadIdx.x % 9);
#ifdef K_STEP_TRACE
  { const int p = p_tr; KTRACE(5); }
#endif
  if (backsub)
    for (int i = threadIdx.x; i < F * 6; i += blockDim.x)
      s_sstep[i] = (st->free_index[i / 6] >= 0) ? st->scale_c[i] * st->step_c[i] : 0.0;
  for (int i = lane; i < F * kUStride; i += 32) s_U_w[i] = 0.0;
  if (lane < kEacc) s_E[warp * kEacc + lane] = 0.0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) s_wts[i] = __ldg(prm.weights + i);
#ifdef K_STEP_TRACE
  { const int p = p_tr; KTRACE(8); }
#endif
  // (the CTA barrier that publishes s_pose / s_sstep sits inside the first iteration of the point loop, behind
  //  the requests for the first point's inputs, so their L2 round trips overlap with the pose constants)

  // per-warp scalars {cost, sum g_p^2, max|g_p|, sum|X|^2, s.g, s'Hs, |step|^2, |x+step|^2} accumulate in shared
  // memory (s_E, zeroed before the prologue barrier), not in registers that would stay live across every phase
  double* s_E_w = s_E + warp * kEacc;
  // back-substitution inputs, staged asynchronously: W of the point's observations -> the (idle) reduction
  // buffer, {V, Vinv, g_p, scale_p} -> the (idle) patch-sum buffer
  double* s_bsW = s_red_w;       // [nobs][18]
  double* s_bsR = s_tot_w;       // V[6] | Vinv[6] | g_p[3] | scale_p[3]
  bool need_barrier = true;
  for (int p = blockIdx.x * WARPS + warp;; p += gridDim.x * WARPS) {
    const bool valid = p < prm.n_points;
    int o0 = 0, nobs = 0, frm_l = 0;
    double X0 = 0.0, X1 = 0.0, X2 = 0.0;
    double p0c[PR];       // reference descriptor of this point, channel 0 (first use: sampling)
    if (valid) {
      o0 = __ldg(prm.obs_off + p); nobs = __ldg(prm.obs_off + p + 1) - o0;
      __syncwarp();                                        // the previous point is done with the staging buffers
      if (backsub) {
        const char* Wc = reinterpret_cast<const char*>(prm.W + ((size_t)cur * prm.nnz + o0) * 18);
        for (int c = lane; c < nobs * 9; c += 32) cp_async_16(reinterpret_cast<char*>(s_bsW) + 16 * c, Wc + 16 * c);
        if (lane < 3) cp_async_16(reinterpret_cast<char*>(s_bsR) + 16 * lane, reinterpret_cast<const char*>(prm.V + ((size_t)cur * prm.n_points + p) * 6) + 16 * lane);
        else if (lane < 6) cp_async_16(reinterpret_cast<char*>(s_bsR + 6) + 16 * (lane - 3), reinterpret_cast<const char*>(prm.Vinv + (size_t)p * 6) + 16 * (lane - 3));
        else if (lane < 9) cp_async_8(s_bsR + 12 + (lane - 6), prm.gp + ((size_t)cur * prm.n_points + p) * 3 + (lane - 6));
        else if (lane < 12) cp_async_8(s_bsR + 15 + (lane - 9), prm.scale_p + (size_t)p * 3 + (lane - 9));
        cp_async_commit();
      }
      if (lane < nobs) frm_l = __ldg(prm.obs_frame + o0 + lane);
      X0 = pts_cur[3 * p]; X1 = pts_cur[3 * p + 1]; X2 = pts_cur[3 * p + 2];
#pragma unroll
      for (int r = 0; r < PR; ++r) p0c[r] = (double)__ldg(prm.desc + (size_t)p * CP + (kPack3 ? lane % 9 : min(lane + 32 * r, P - 1)));
    }
    if (need_barrier) { KTRACE(12); __syncthreads(); need_barrier = false; }
    if (!valid) break;
    KTRACE(1);
    if (lane < nobs) s_frm_w[lane] = frm_l;
    if (backsub) cp_async_wait_all();
    __syncwarp();
    KTRACE(6);

    // ---- (B) back-substitution (SchurEliminator::BackSubstitute + model cost change) -------
    if (backsub) {
      // lane = b*8 + a accumulates sum_obs W[a][b] * (scale_c*step_c)[f][a]
      const int a = lane & 7, b = lane >> 3;
      const bool act = (a < 6 && b < 3);
      double sp[3], Vi[6], Vv[6], gcv[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) { sp[k] = s_bsR[15 + k]; gcv[k] = s_bsR[12 + k]; }
#pragma unroll
      for (int k = 0; k < 6; ++k) { Vi[k] = s_bsR[6 + k]; Vv[k] = s_bsR[k]; }
      double acc = 0.0;
      if (act)
        for (int k = 0; k < nobs; ++k) acc = fma(s_bsW[k * 18 + a * 3 + b], s_sstep[s_frm_w[k] * 6 + a], acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      double wts[3];
      wts[0] = __shfl_sync(0xffffffffu, acc, 0);
      wts[1] = __shfl_sync(0xffffffffu, acc, 8);
      wts[2] = __shfl_sync(0xffffffffu, acc, 16);
      const double gs[3] = {sp[0] * gcv[0], sp[1] * gcv[1], sp[2] * gcv[2]};
      double t[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) { wts[k] *= sp[k]; t[k] = gs[k] + wts[k]; }   // gs - Ws^T y_c, y_c = -step_c
      const double s0 = -(Vi[0] * t[0] + Vi[1] * t[1] + Vi[2] * t[2]);
      const double s1 = -(Vi[1] * t[0] + Vi[3] * t[1] + Vi[4] * t[2]);
      const double s2 = -(Vi[2] * t[0] + Vi[4] * t[1] + Vi[5] * t[2]);
      const double c0 = X0 + s0 * sp[0], c1 = X1 + s1 * sp[1], c2 = X2 + s2 * sp[2];
      if (lane == 0) {
        // s.gs + s^T Vs0 s + 2 step_c^T Ws s   (undamped Vs0 = sp V sp)
        const double y0 = s0 * sp[0], y1 = s1 * sp[1], y2 = s2 * sp[2];
        s_E_w[4] += s0 * gs[0] + s1 * gs[1] + s2 * gs[2];
        s_E_w[5] += y0 * (Vv[0] * y0 + Vv[1] * y1 + Vv[2] * y2) + y1 * (Vv[1] * y0 + Vv[3] * y1 + Vv[4] * y2) +
                    y2 * (Vv[2] * y0 + Vv[4] * y1 + Vv[5] * y2) + 2.0 * (wts[0] * s0 + wts[1] * s1 + wts[2] * s2);
        s_E_w[6] += (X0 - c0) * (X0 - c0) + (X1 - c1) * (X1 - c1) + (X2 - c2) * (X2 - c2);
        s_E_w[7] += c0 * c0 + c1 * c1 + c2 * c2;
      }
      X0 = c0; X1 = c1; X2 = c2;
      if (lane < 3) pts_out[3 * p + lane] = lane == 0 ? c0 : (lane == 1 ? c1 : c2);
    }
    if (lane == 0) s_E_w[3] += X0 * X0 + X1 * X1 + X2 * X2;
    double cost_p = 0.0;   // this lane's share of the point's cost
    KTRACE(2);

    // ---- per-lane constants.  Formed per point / per batch from an opaque copy of the lane id: as loop
    // invariants they were kept live across the back-substitution, pushed it over the 72-register ceiling
    // and came back from local memory - which, with the shared-memory carve-out at its maximum, means L2.
    int lane_l = lane;
    asm volatile("" : "+r"(lane_l));
    int st_off[FT::ROUNDS];
#pragma unroll
    for (int rd = 0; rd < FT::ROUNDS; ++rd) {
      const int wi = lane_l + 32 * rd;
      const int row = wi / FT::NW, wd = wi - row * FT::NW;
      st_off[rd] = (wi < FT::WORDS) ? row * prm.fr.pitch + 4 * wd : -1;
    }

    double acc_pt = 0.0;  // lanes 18..23: V entries, 24..26: g_p entries (summed over frames)
    for (int ob = 0; ob < nobs; ob += kObsBatch) {
      const int nb = min(kObsBatch, nobs - ob);
      // ---- (G1) warp + project: lanes i, i+8, i+16 <-> observation ob+i (three copies: each forms one
      // column of the Jacobian in (G2); an FP64 instruction costs the same issue slots for 8 lanes as for 24)
      int g_fast_l = 0;
      double Xc0 = 0.0, Xc1 = 0.0, Xc2 = 1.0;
      const double* pc = s_pose;
      const int g_i = lane & 7, g_b = lane >> 3;          // observation within the batch, Jacobian column
      const bool g_act = g_b < 3 && g_i < nb;
      if (g_act) {
        const int g_f = s_frm_w[ob + g_i];
        pc = s_pose + g_f * kPoseConst;
        if (pc[8] == 0.0) {  // ceres::AngleAxisRotatePoint, same operation order (no FMA)
          const double k0 = pc[0], k1 = pc[1], k2 = pc[2], c = pc[6], s = pc[7];
          const double wx0 = __dsub_rn(__dmul_rn(k1, X2), __dmul_rn(k2, X1));
          const double wx1 = __dsub_rn(__dmul_rn(k2, X0), __dmul_rn(k0, X2));
          const double wx2 = __dsub_rn(__dmul_rn(k0, X1), __dmul_rn(k1, X0));
          const double tmp = __dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(k0, X0), __dmul_rn(k1, X1)), __dmul_rn(k2, X2)),
                                       __dsub_rn(1.0, c));
          Xc0 = __dadd_rn(__dadd_rn(__dmul_rn(X0, c), __dmul_rn(wx0, s)), __dmul_rn(k0, tmp));
          Xc1 = __dadd_rn(__dadd_rn(__dmul_rn(X1, c), __dmul_rn(wx1, s)), __dmul_rn(k1, tmp));
          Xc2 = __dadd_rn(__dadd_rn(__dmul_rn(X2, c), __dmul_rn(wx2, s)), __dmul_rn(k2, tmp));
        } else {
          const double a0 = pc[0], a1 = pc[1], a2 = pc[2];
          Xc0 = __dadd_rn(X0, __dsub_rn(__dmul_rn(a1, X2), __dmul_rn(a2, X1)));
          Xc1 = __dadd_rn(X1, __dsub_rn(__dmul_rn(a2, X0), __dmul_rn(a0, X2)));
          Xc2 = __dadd_rn(X2, __dsub_rn(__dmul_rn(a0, X1), __dmul_rn(a1, X0)));
        }
        Xc0 = __dadd_rn(Xc0, pc[3]); Xc1 = __dadd_rn(Xc1, pc[4]); Xc2 = __dadd_rn(Xc2, pc[5]);
        if (g_b == 0) {
          // Calibration::project: u = ((X*fx)/Z) + cx  (IEEE division, T = double path)
          const double u = __dadd_rn(__ddiv_rn(__dmul_rn(Xc0, prm.fx), Xc2), prm.cx);
          const double v = __dadd_rn(__ddiv_rn(__dmul_rn(Xc1, prm.fy), Xc2), prm.cy);
          double* g = s_geo_w + g_i * 20;
          g[0] = u; g[1] = v;
          // footprint origin and fast-path test (all taps and gradient taps interior)
          int4 gi = make_int4(g_f, 0, 0, 0);   // {frame, r0, cb (aligned first column), fast}
          if (fabs(u) < 1.0e8 && fabs(v) < 1.0e8) {
            const int c0 = (int)floor(u) - R - 1, r0 = (int)floor(v) - R - 1;
            const int fast = (c0 >= 0 && c0 + FT::ROWS - 1 <= prm.fr.cols - 1 && r0 >= 0 &&
                              r0 + FT::ROWS - 1 <= prm.fr.rows - 1) ? 1 : 0;
            gi.y = r0; gi.z = c0 & ~3; gi.w = fast;
          }
          s_gi_w[g_i] = gi;
          g_fast_l = gi.w;
        }
      }
      const unsigned fastmask = __ballot_sync(0xffffffffu, g_fast_l != 0) & 0xffu;
      __syncwarp();
      // ---- (L) 1-channel frames: every footprint of the batch is requested NOW with asynchronous
      // copies (global -> shared, no registers held), so the L2 round trips overlap with (G2) ------
      if (kAsyncStage) {
#pragma unroll
        for (int sl = 0; sl < kStageSlots; ++sl) {
          if (sl < nb) {
            const int4 gi = s_gi_w[sl];
            if (gi.w) {
              const unsigned base = (unsigned)(gi.x * (int)prm.fr.plane + gi.y * prm.fr.pitch + gi.z);
#pragma unroll
              for (int rd = 0; rd < FT::ROUNDS; ++rd) {
                if (st_off[rd] >= 0) {
                  if (U8) cp_async_4(reinterpret_cast<uint32_t*>(s_fp_w + sl * FT::FLOATS) + lane + 32 * rd,
                                     prm.fr.u8 + (base + (unsigned)st_off[rd]));
                  else cp_async_16(reinterpret_cast<float4*>(s_fp_w + sl * FT::FLOATS) + lane + 32 * rd,
                                   prm.fr.f32 + (base + (unsigned)st_off[rd]));
                }
              }
            }
          }
        }
        cp_async_commit();
      }
      // ---- (G2) the 2x9 matrix A = d(u,v)/d[w t X] of each observation; lane (i, b) forms column b of the
      // rotation, translation and point parts --------------------------------------------------------
      if (g_act) {
        double* g = s_geo_w + g_i * 20;
        const int b = g_b;
        const double iz = 1.0 / Xc2;
        const double J00 = prm.fx * iz, J02 = -prm.fx * Xc0 * iz * iz;
        const double J11 = prm.fy * iz, J12 = -prm.fy * Xc1 * iz * iz;
        const double* Rj = pc + 18;
        const double* M = pc + 27;
        // column b of D = -Rj [X]x M
        const double m0 = M[b], m1 = M[3 + b], m2 = M[6 + b];
        const double t0 = X1 * m2 - X2 * m1, t1 = X2 * m0 - X0 * m2, t2 = X0 * m1 - X1 * m0;
        const double D0 = -(Rj[0] * t0 + Rj[1] * t1 + Rj[2] * t2);
        const double D1 = -(Rj[3] * t0 + Rj[4] * t1 + Rj[5] * t2);
        const double D2 = -(Rj[6] * t0 + Rj[7] * t1 + Rj[8] * t2);
        const double* Rm = pc + 9;
        g[2 + b] = J00 * D0 + J02 * D2;                      // du/dw
        g[11 + b] = J11 * D1 + J12 * D2;                     // dv/dw
        g[8 + b] = J00 * Rm[b] + J02 * Rm[6 + b];            // du/dX
        g[17 + b] = J11 * Rm[3 + b] + J12 * Rm[6 + b];       // dv/dX
        g[5 + b] = b == 0 ? J00 : (b == 1 ? 0.0 : J02);      // du/dt
        g[14 + b] = b == 0 ? 0.0 : (b == 1 ? J11 : J12);     // dv/dt
      }
      // patch offsets, weights and block-entry indices of this lane: independent integer work placed in the
      // shadow of the asynchronous footprint copies
      double pdx[PR], pdy[PR], wj[PR];
#pragma unroll
      for (int r = 0; r < PR; ++r) {
        // lanes beyond the patch re-do pixel P-1 with weight 0; 3x3 patches (kPack3): lane -> pixel lane % 9 of one of
        // THREE observations sampled side by side, lanes 27..31 carry weight 0
        const int j = kPack3 ? lane_l % 9 : min(lane_l + 32 * r, P - 1);
        const int py = j / FT::SIDE, pxo = j - py * FT::SIDE;
        pdx[r] = (double)(pxo - R); pdy[r] = (double)(py - R);
        wj[r] = (kPack3 ? lane_l < 27 : lane_l + 32 * r < P) ? s_wts[j] : 0.0;
      }
      // lane l < 21 <-> upper-triangle entry (a, b) of the 6x6 pose block, rows of length 6, 5, ... 1
      int e1a = 0, e1b = 0, e2a = 0, e2b = 0;
      if (lane_l < 21) {
        e1a = (lane_l >= 6) + (lane_l >= 11) + (lane_l >= 15) + (lane_l >= 18) + (lane_l >= 20);
        e1b = e1a + lane_l - (e1a * (13 - e1a)) / 2;
      } else if (lane_l < 27) { e1a = lane_l - 21; }
      if (lane_l < 18) { e2a = lane_l / 3; e2b = 6 + lane_l - 3 * (lane_l / 3); }
      else if (lane_l < 24) {          // upper triangle of the 3x3 point block
        const int i3 = lane_l - 18, a3 = (i3 >= 3) + (i3 >= 5);
        e2a = 6 + a3; e2b = 6 + a3 + i3 - (a3 * (7 - a3)) / 2;
      } else if (lane_l < 27) { e2a = 6 + lane_l - 24; }
      if (kAsyncStage) cp_async_wait_all();
      __syncwarp();
      KTRACE(3);

      const int obs_per_stage = NCH == 1 ? kStageSlots : ((C >= kStageSlots) ? 1 : kStageSlots / C);
      unsigned defmask = 0;   // observations of this batch whose corrector + block expansion are deferred
      for (int sb = 0; sb < nb; sb += obs_per_stage) {
        const int ns_obs = min(obs_per_stage, nb - sb);
        const int nslots = NCH == 1 ? ns_obs : ns_obs * C;
        // ---- (L) multi-channel frames: stage (observation, channel) footprints through registers ----
        if (!kAsyncStage) {
          float4 t32[kStageSlots][FT::ROUNDS];
#pragma unroll
          for (int sl = 0; sl < kStageSlots; ++sl) {
            if (sl < nslots) {
              const int i = sb + sl / C;
              const int k = sl - (sl / C) * C;
              const int4 gi = s_gi_w[i];
              if (gi.w) {
                const unsigned base = (unsigned)((gi.x * C + k) * (int)prm.fr.plane + gi.y * prm.fr.pitch + gi.z);
#pragma unroll
                for (int rd = 0; rd < FT::ROUNDS; ++rd)
                  if (st_off[rd] >= 0) t32[sl][rd] = __ldg(reinterpret_cast<const float4*>(prm.fr.f32 + (base + (unsigned)st_off[rd])));
              }
            }
          }
#pragma unroll
          for (int sl = 0; sl < kStageSlots; ++sl) {
            if (sl < nslots) {
              const int i = sb + sl / C;
              if (s_gi_w[i].w) {
#pragma unroll
                for (int rd = 0; rd < FT::ROUNDS; ++rd)
                  if (st_off[rd] >= 0) reinterpret_cast<float4*>(s_fp_w + sl * FT::FLOATS)[lane + 32 * rd] = t32[sl][rd];
              }
            }
          }
          __syncwarp();
        }

        KTRACE(4);
        // ---- (S)+(R): four observations at a time when every one of them is interior -------
        for (int qb = 0; qb < ns_obs; qb += kG) {
          if constexpr (kPack3) {
            const int n6 = min(6, ns_obs - qb);
            if ((((fastmask >> (sb + qb)) & ((1u << n6) - 1u)) == ((1u << n6) - 1u))) {
              const int sub = min(lane / 9, 2), jpx = lane - 9 * (lane / 9);
              double v6[2][6];
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const int oi = 3 * i + sub;                          // observation of the group this lane samples
                const int ii = sb + qb + min(oi, n6 - 1);            // out-of-range slots redo the last one
                const int4 gi = s_gi_w[ii];
                const double* g = s_geo_w + ii * 20;
                float I1, gx, gy;
                sample_fast_u8<R>(s_fp_w + (ii - sb) * FT::FLOATS, gi.y, gi.z, g[0], g[1], pdx[0], pdy[0], I1, gx, gy);
                const double rr = __dmul_rn(wj[0], __dsub_rn(p0c[0], (double)I1));   // photobundle.cc:720
                if (want_res && lane < 27 && oi < n6) prm.residuals[(size_t)(o0 + ob + ii) * CP + jpx] = rr;
                const double hx = wj[0] * (double)gx, hy = wj[0] * (double)gy;
                v6[i][0] = rr * rr; v6[i][1] = hx * hx; v6[i][2] = hx * hy; v6[i][3] = hy * hy; v6[i][4] = rr * hx; v6[i][5] = rr * hy;
              }
              // transpose-reduce: row = lane (27 rows of 12 sums); total (observation 3i + s, sum k) = sum over the 9
              // pixel rows of sub-observation s of column 6i + k
              if (lane < 27) {
                double2* row = reinterpret_cast<double2*>(s_red_w + lane * kRedStride);
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                  for (int k = 0; k < 3; ++k) row[i * 3 + k] = make_double2(v6[i][2 * k], v6[i][2 * k + 1]);
              }
              __syncwarp();
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {
                const int T = lane + 32 * h2;                        // 36 totals: observation T / 6, sum T % 6
                if (T < 6 * n6) {
                  const int oi = T / 6, k = T - 6 * oi, i = oi / 3, s = oi - 3 * i;
                  const double* col = s_red_w + (9 * s) * kRedStride + 6 * i + k;
                  double tot = 0.0;
#pragma unroll
                  for (int l = 0; l < 9; ++l) tot += col[l * kRedStride];
                  s_tot_w[(sb + qb) * 6 + T] = tot;
                }
              }
              defmask |= ((1u << n6) - 1u) << (sb + qb);
              __syncwarp();
              qb += 6 - kG;                                          // (the loop adds kG)
              continue;
            }
          }
          const int nq = min(kG, ns_obs - qb);
          const bool all_fast = kQuad && (((fastmask >> (sb + qb)) & ((1u << nq) - 1u)) == ((1u << nq) - 1u));
          if (kQuad && all_fast) {
            double v6[kG][6];
#pragma unroll
            for (int i = 0; i < kG; ++i) {
              const int ii = sb + qb + min(i, nq - 1);            // out-of-range slots redo the last one
              const int4 gi = s_gi_w[ii];
              const double* g = s_geo_w + ii * 20;
              float I1, gx, gy;
              if constexpr (U8) sample_fast_u8<R>(s_fp_w + (ii - sb) * FT::FLOATS, gi.y, gi.z, g[0], g[1], pdx[0], pdy[0], I1, gx, gy);
              else sample_fast<R, FPT>(s_fp_w + (ii - sb) * FT::FLOATS, gi.y, gi.z, g[0], g[1], pdx[0], pdy[0], I1, gx, gy);
              const double rr = __dmul_rn(wj[0], __dsub_rn(p0c[0], (double)I1));   // photobundle.cc:720
              if (want_res && lane < P && i < nq) prm.residuals[(size_t)(o0 + ob + ii) * CP + lane] = rr;
              const double hx = wj[0] * (double)gx, hy = wj[0] * (double)gy;
              v6[i][0] = rr * rr; v6[i][1] = hx * hx; v6[i][2] = hx * hy; v6[i][3] = hy * hy; v6[i][4] = rr * hx; v6[i][5] = rr * hy;
            }
            KTRACE(9);
            // transpose-reduce the 24 sums of the group through shared memory: lane r <- sum over pixels of value r
            if (lane < P) {
              double2* row = reinterpret_cast<double2*>(s_red_w + lane * kRedStride);
#pragma unroll
              for (int i = 0; i < kG; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k) row[i * 3 + k] = make_double2(v6[i][2 * k], v6[i][2 * k + 1]);
            }
            __syncwarp();
            double tot = 0.0;
            if (lane < 6 * kG) {
#pragma unroll
              for (int l = 0; l < P; ++l) tot += s_red_w[l * kRedStride + lane];
            }
            KTRACE(10);
            // lane 6i+k holds sum k of observation i: park the raw sums; the loss corrector and the block
            // expansion of the whole batch follow the sampling loop (one latency chain per batch, not per group)
            if (lane < 6 * nq) s_tot_w[(sb + qb) * 6 + lane] = tot;
            defmask |= ((1u << nq) - 1u) << (sb + qb);
            __syncwarp();
            KTRACE(11);
          } else {
            // generic path: any channel count / patch size / border handling, one observation at a time
            for (int i = 0; i < nq; ++i) {
              const int ii = sb + qb + i;
              const int o = o0 + ob + ii;
              const int4 gi = s_gi_w[ii];
              const int f = gi.x, r0 = gi.y, cb = gi.z, fast = gi.w;
              const double* g = s_geo_w + ii * 20;
              const double u = g[0], v = g[1];
              Sums q = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
              for (int k = 0; k < C; ++k) {
                const FPT* fp = s_fp_w + (NCH == 1 ? ii - sb : (ii - sb) * C + k) * FT::FLOATS;
#pragma unroll
                for (int r = 0; r < PR; ++r) {
                  const int j = lane + 32 * r;
                  if (j < P) {
                    const float su = __double2float_rn(__dadd_rn(u, pdx[r]));
                    const float sv = __double2float_rn(__dadd_rn(v, pdy[r]));
                    float I1, gx, gy;
                    if (fast) {
                      const int ix = __float2int_rz(su), iy = __float2int_rz(sv);
                      const float dx = __fsub_rn((float)(ix + 1), su), dy = __fsub_rn((float)(iy + 1), sv);
                      const FPT* qq = fp + (iy - r0) * FT::W + (ix - cb);
                      const float a11 = (float)qq[0], a12 = (float)qq[1], a21 = (float)qq[FT::W], a22 = (float)qq[FT::W + 1];
                      const float l1 = (float)qq[-1], r1 = (float)qq[2], l2 = (float)qq[FT::W - 1], r2 = (float)qq[FT::W + 2];
                      const float t1 = (float)qq[-FT::W], t2 = (float)qq[-FT::W + 1], u1 = (float)qq[2 * FT::W], u2 = (float)qq[2 * FT::W + 1];
                      const double omdx = __dsub_rn(1.0, (double)dx);
                      const float omdy = __fsub_rn(1.0f, dy);
                      I1 = bilerp(dx, dy, omdx, omdy, a11, a12, a21, a22);
                      gx = bilerp(dx, dy, omdx, omdy, __fmul_rn(0.5f, __fsub_rn(a12, l1)), __fmul_rn(0.5f, __fsub_rn(r1, a11)),
                                  __fmul_rn(0.5f, __fsub_rn(a22, l2)), __fmul_rn(0.5f, __fsub_rn(r2, a21)));
                      gy = bilerp(dx, dy, omdx, omdy, __fmul_rn(0.5f, __fsub_rn(a21, t1)), __fmul_rn(0.5f, __fsub_rn(a22, t2)),
                                  __fmul_rn(0.5f, __fsub_rn(u1, a11)), __fmul_rn(0.5f, __fsub_rn(u2, a12)));
                    } else {
                      int x1, x2, y1, y2;
                      float dx, dy;
                      init_axis(sv, prm.fr.rows, y1, y2, dy);
                      init_axis(su, prm.fr.cols, x1, x2, dx);
                      const double omdx = __dsub_rn(1.0, (double)dx);
                      const float omdy = __fsub_rn(1.0f, dy);
                      I1 = bilerp(dx, dy, omdx, omdy, px_at<U8>(prm.fr, f, k, y1, x1), px_at<U8>(prm.fr, f, k, y1, x2),
                                  px_at<U8>(prm.fr, f, k, y2, x1), px_at<U8>(prm.fr, f, k, y2, x2));
                      float gx11, gx12, gx21, gx22, gy11, gy12, gy21, gy22;
                      grad_at<U8>(prm.fr, f, k, y1, x1, gx11, gy11);
                      grad_at<U8>(prm.fr, f, k, y1, x2, gx12, gy12);
                      grad_at<U8>(prm.fr, f, k, y2, x1, gx21, gy21);
                      grad_at<U8>(prm.fr, f, k, y2, x2, gx22, gy22);
                      gx = bilerp(dx, dy, omdx, omdy, gx11, gx12, gx21, gx22);
                      gy = bilerp(dx, dy, omdx, omdy, gy11, gy12, gy21, gy22);
                    }
                    const double p0 = (k == 0) ? p0c[r] : (double)prm.desc[(size_t)p * CP + k * P + j];
                    const double rr = __dmul_rn(wj[r], __dsub_rn(p0, (double)I1));   // photobundle.cc:720
                    if (want_res) prm.residuals[(size_t)o * CP + k * P + j] = rr;
                    q.s = fma(rr, rr, q.s);
                    const double hx = wj[r] * (double)gx, hy = wj[r] * (double)gy;
                    q.G11 = fma(hx, hx, q.G11); q.G12 = fma(hx, hy, q.G12); q.G22 = fma(hy, hy, q.G22);
                    q.b1 = fma(rr, hx, q.b1); q.b2 = fma(rr, hy, q.b2);
                  }
                }
              }
              warp_reduce(q);
              double rho0, rho1;
              huber_rho(prm.huber, q.s, rho0, rho1);
              if (lane == 0) {
                cost_p += 0.5 * rho0;
                if (prm.obs_sqnorm) prm.obs_sqnorm[o] = q.s;
              }
              emit_blocks(rho1 * q.G11, rho1 * q.G12, rho1 * q.G22, rho1 * q.b1, rho1 * q.b2, g, f, f != prm.fixed_frame, lane,
                          e1a, e1b, e2a, e2b, s_U_w, outW + (size_t)o * 18, acc_pt, USHARE > 1);
            }
          }
        }
        __syncwarp();
      }
      if (kQuad && defmask) {
        // ---- (H) Huber corrector of every deferred observation at once: lane 6i+k <-> sum k of observation i
#pragma unroll
        for (int r = 0; r < (6 * kObsBatch + 31) / 32; ++r) {
          const int idx = lane + 32 * r;
          const int i_l = idx / 6, k_l = idx - 6 * i_l;
          if (idx < 6 * nb && ((defmask >> i_l) & 1u)) {
            const double s_i = s_tot_w[6 * i_l], raw = s_tot_w[idx];
            double rho0, rho1;
            huber_rho(prm.huber, s_i, rho0, rho1);
            if (k_l == 0) {
              cost_p += 0.5 * rho0;
              if (prm.obs_sqnorm) prm.obs_sqnorm[o0 + ob + i_l] = s_i;
            } else {
              s_tot_w[idx] = rho1 * raw;
            }
          }
        }
        __syncwarp();
        // ---- (E) block expansion.  Full batch (the common case): the entries of all eight observations are
        // evaluated first, branch-free, and stored afterwards; otherwise observation by observation.
        if (nb == kObsBatch && defmask == 0xffu) {
          constexpr int kE = 4;   // observations evaluated together (8 pushes the kernel over its 72 registers)
#pragma unroll 1
          for (int i0 = 0; i0 < kObsBatch; i0 += kE) {
            double v1[kE], v2[kE];
#pragma unroll
            for (int i = 0; i < kE; ++i)
              block_entries(s_tot_w + 6 * (i0 + i), s_geo_w + (i0 + i) * 20, lane, e1a, e1b, e2a, e2b, v1[i], v2[i]);
#pragma unroll
            for (int i = 0; i < kE; ++i) {
              const int f = s_gi_w[i0 + i].x;
              const bool free_cam = f != prm.fixed_frame;
              if (lane < 27 && free_cam) {
                if (USHARE > 1) atomicAdd(s_U_w + f * kUStride + lane, v1[i]);
                else s_U_w[f * kUStride + lane] += v1[i];   // warp-private: no atomics
              }
              if (lane < 18) outW[(size_t)(o0 + ob + i0 + i) * 18 + lane] = free_cam ? v2[i] : 0.0;
              else if (lane < 27) acc_pt += v2[i];
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < kObsBatch; ++i) {
            if (i < nb && ((defmask >> i) & 1u)) {
              const int f = s_gi_w[i].x;
              const double* t = s_tot_w + 6 * i;
              emit_blocks(t[1], t[2], t[3], t[4], t[5], s_geo_w + i * 20, f, f != prm.fixed_frame, lane, e1a, e1b, e2a, e2b,
                          s_U_w, outW + (size_t)(o0 + ob + i) * 18, acc_pt, USHARE > 1);
            }
          }
        }
        __syncwarp();
      }
    }
    KTRACE(13);
    if (lane >= 18 && lane < 24) outV[(size_t)p * 6 + lane - 18] = acc_pt;
    else if (lane >= 24 && lane < 27) outgp[(size_t)p * 3 + lane - 24] = acc_pt;
    const double gq = (lane >= 24 && lane < 27) ? acc_pt : 0.0;
    double g2 = gq * gq, ga = fabs(gq);
#pragma unroll
    for (int m = 1; m <= 2; m <<= 1) {   // lanes 24..27 form an aligned group of four
      g2 += __shfl_xor_sync(0xffffffffu, g2, m);
      ga = fmax(ga, __shfl_xor_sync(0xffffffffu, ga, m));
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) cost_p += __shfl_xor_sync(0xffffffffu, cost_p, m);
    if (lane == 24) {   // lane 24 holds the reduced g_p sums
      s_E_w[0] += cost_p;
      s_E_w[1] += g2;
      s_E_w[2] = fmax(s_E_w[2], ga);
    }
  }
  __syncthreads();
#ifdef K_STEP_TRACE
  const int p = blockIdx.x * WARPS + warp;   // single-wave geometry: one point per warp
#endif
  KTRACE(14);
  // CTA partials (fixed order inside the CTA), then one fp64 atomic per entry per CTA
  for (int i = threadIdx.x; i < F * kUStride; i += blockDim.x) {
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < WARPS / USHARE; ++w)
      acc += USHARE == 1 ? reinterpret_cast<const double*>(s_warp0 + w * warp_bytes + kOffU)[i]
                         : (reinterpret_cast<const double*>(s_warp0 + WARPS * warp_bytes) + w * ustride)[i];
    if (acc != 0.0) atomicAdd(prm.Xacc + i, acc);
  }
  if (threadIdx.x < kEacc) {
    double acc = 0.0;
    if (threadIdx.x == 2) {
      for (int w = 0; w < WARPS; ++w) acc = fmax(acc, s_E[w * kEacc + 2]);
      // non-negative doubles order like their bit patterns; one slot per rank (summed by the all-reduce)
      atomicMax(reinterpret_cast<unsigned long long*>(prm.Xacc + F * kUStride + kEacc + prm.rank),
                (unsigned long long)__double_as_longlong(acc));
    } else {
      for (int w = 0; w < WARPS; ++w) acc += s_E[w * kEacc + threadIdx.x];
      if (acc != 0.0) atomicAdd(prm.Xacc + F * kUStride + threadIdx.x, acc);
    }
  }
  KTRACE(15);
}

#ifdef K_STEP_TRACE
extern "C" void pba_debug_kstep_trace(long long* out, int n) {
  cudaMemcpyFromSymbol(out, g_kstep_trace, sizeof(long long) * n);
}
#endif

// ---- host launcher -----------------------------------------------------------------------
template <int R, bool U8, int NCH, int WARPS = Geo<U8>::WARPS, int USHARE = 1>
static cudaError_t launch_one(const StepParams& prm, cudaStream_t stream) {
  const size_t smem = k_step_smem_bytes<R, U8, WARPS, USHARE>(prm.n_frames);
  static bool configured[64] = {};
  static int sm_count[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_step<R, U8, NCH, WARPS, USHARE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return e;
    // two CTAs per SM need the large shared-memory carve-out
    cudaFuncSetAttribute(k_step<R, U8, NCH, WARPS, USHARE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaDeviceGetAttribute(&sm_count[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    configured[dev & 63] = true;
  }
  // persistent: at most two CTAs per SM, every warp strides over the points
  const int want = (prm.n_points + WARPS - 1) / WARPS, cap = 2 * sm_count[dev & 63];
  const int grid = want < cap ? want : cap;
  if (grid == 0) return cudaSuccess;
  if (prm.pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(WARPS * 32); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_step<R, U8, NCH, WARPS, USHARE>, prm);
  }
  k_step<R, U8, NCH, WARPS, USHARE><<<grid, WARPS * 32, smem, stream>>>(prm);
  return cudaGetLastError();
}


template <int R>
static cudaError_t launch_r(const StepParams& prm, cudaStream_t stream) {
  if (prm.fr.u8) {                                                // uint8 planes are always 1-channel Intensity
    // two CTAs per SM: 228 KB of shared memory, 1 KB reserved per CTA
    constexpr size_t kPerCta = (233472 - 2 * 1024) / 2;
    if constexpr (R == 2)
      if (k_step_smem_bytes<R, true>(prm.n_frames) > kPerCta) {
        // wide windows: the per-warp pose-block accumulators ([F][27] doubles) no longer fit twice 14 warps; pairs of
        // warps share one (shared-memory atomics) so that the geometry of the single-wave case is kept
        if (getenv("PBA_KA_WIDE_11") == nullptr && k_step_smem_bytes<R, true, 14, 2>(prm.n_frames) <= kPerCta)
          return launch_one<R, true, 1, 14, 2>(prm, stream);
        return launch_one<R, true, 1, 11>(prm, stream);
      }
    return launch_one<R, true, 1>(prm, stream);
  }
  if (prm.fr.n_channels == 1) return launch_one<R, false, 1>(prm, stream);
  return launch_one<R, false, 0>(prm, stream);
}

cudaError_t launch_k_step(const StepParams& prm, int radius, cudaStream_t stream) {
  switch (radius) {
    case 1: return launch_r<1>(prm, stream);
    case 2: return launch_r<2>(prm, stream);
    case 3: return launch_r<3>(prm, stream);
    case 4: return launch_r<4>(prm, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace pba
