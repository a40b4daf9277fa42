// k1_eval.cu — K1: per-(point, frame) photometric residual + analytic Jacobian + block
// accumulation, one hand-written sm_100a kernel batched over all points x observing frames.
//
// Replaces (reference, /root/reference):
//   DescriptorError::operator()<Jet<double,9>>      src/photobundle.cc:696-727
//   SampleWithDerivative / SampleLinear             src/sample_eigen.h:33-126
//   ceres::Chain<float,2,Jet>::Rule                 src/jet_extras.h:74-111
//   imgradient (central difference, zero borders)   src/imgproc.cc:27-106
//   Calibration::project                            src/calibration.h:33-38
//   ceres::AngleAxisRotatePoint, HuberLoss + Corrector, and the J^T J / J^T r products
//   Ceres' SchurEliminator forms per residual block.
//
// Work decomposition: one warp per scene point.  For up to 8 observing frames at a time,
//   (G) lane i forms the geometry of observation i in fp64: Xc = R(w)X + t, (u,v) and the
//       2x9 matrix A = d(u,v)/d[w t X]  (dual numbers are not needed: every one of the
//       (2r+1)^2 Jacobian rows is -w_j*[gx gy]*A);
//   (L) the warp stages each observation's (2r+5)x(2r+5) image footprint into shared memory
//       with ONE coalesced, 4-byte/16-byte aligned load instruction (8 footprints in flight),
//       converting uint8 -> fp32 on the way (DescriptorFrame::Create's cast);
//   (S) lane j samples patch pixel j: the fp32 bilinear taps of I come from the footprint,
//       the taps of Gx, Gy are formed in registers from the same footprint (exactly the
//       0.5*(a-b) the reference precomputes into planes), and the three interpolations use
//       the reference's float/double promotion sequence, so residuals are bit-identical
//       to the CPU path given the same (u,v);
//   (R) six patch sums  s=Σr², G=Σ w²ggᵀ (3), b=Σ w r g (2)  are reduced with warp shuffles;
//       the whole 9x9 block of the observation is  rho' Aᵀ G A  and  -rho' Aᵀ b, expanded by
//       27 lanes in parallel in fp64;
//   the point's V (3x3) and g_p live in registers across its frames, W (6x3) goes out once
//   per observation, pose blocks U (6x6) / g_c are summed per CTA in shared memory and
//   written as one partial per CTA (reduced deterministically by the next kernel).
// Out-of-image / border observations take a per-tap slow path with the reference's clamp
// and zero-gradient-border rules (src/sample_eigen.h:38-46, src/imgproc.cc:34-43).

#include "pba_device.cuh"

#include <climits>
#include <cfloat>

namespace pba {

__constant__ signed char c_pair6[21][2] = {
    {0, 0}, {0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {1, 1}, {1, 2}, {1, 3}, {1, 4}, {1, 5},
    {2, 2}, {2, 3}, {2, 4}, {2, 5}, {3, 3}, {3, 4}, {3, 5}, {4, 4}, {4, 5}, {5, 5}};
__constant__ signed char c_pair3[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};

// ---- per-frame pose constants (computed once per CTA) ---------------------------------
//  [0..2] unit axis k (or the raw angle-axis when tiny)   [3..5] t   [6] cos  [7] sin
//  [8] tiny flag   [9..17] R   [18..26] Rj   [27..35] M,  with
//  d(Xc)/dw = -Rj [X]x M ;  M = (w wᵀ + (Rᵀ - I)[w]x)/θ² ; tiny angle: Rj = M = I, R = I+[w]x
__device__ void pose_consts(const double* cam, double* pc) {
  const double w0 = cam[0], w1 = cam[1], w2 = cam[2];
  const double theta2 = __dadd_rn(__dadd_rn(__dmul_rn(w0, w0), __dmul_rn(w1, w1)), __dmul_rn(w2, w2));
  pc[3] = cam[3]; pc[4] = cam[4]; pc[5] = cam[5];
  if (theta2 > DBL_EPSILON) {
    const double theta = sqrt(theta2);
    const double c = cos(theta), s = sin(theta);
    const double ti = 1.0 / theta;
    const double k0 = __dmul_rn(w0, ti), k1 = __dmul_rn(w1, ti), k2 = __dmul_rn(w2, ti);
    pc[0] = k0; pc[1] = k1; pc[2] = k2; pc[6] = c; pc[7] = s; pc[8] = 0.0;
    const double k[3] = {w0 / theta, w1 / theta, w2 / theta};
    const double K[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
    double R[9];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) R[a * 3 + b] = (a == b ? c : 0.0) + s * K[a * 3 + b] + (1.0 - c) * k[a] * k[b];
    const double w[3] = {w0, w1, w2};
    const double Wx[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        double acc = w[a] * w[b];
        for (int q = 0; q < 3; ++q) acc += (R[q * 3 + a] - (q == a ? 1.0 : 0.0)) * Wx[q * 3 + b];
        pc[27 + a * 3 + b] = acc / theta2;
      }
    for (int a = 0; a < 9; ++a) { pc[9 + a] = R[a]; pc[18 + a] = R[a]; }
  } else {
    pc[0] = w0; pc[1] = w1; pc[2] = w2; pc[6] = 1.0; pc[7] = 0.0; pc[8] = 1.0;
    const double Wx[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
    for (int a = 0; a < 9; ++a) {
      const double id = (a == 0 || a == 4 || a == 8) ? 1.0 : 0.0;
      pc[9 + a] = id + Wx[a]; pc[18 + a] = id; pc[27 + a] = id;
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

// Shared memory per CTA:  pose consts [F][36] f64 | per warp: geometry [8][20] f64, pose-block
// accumulators [F][27] f64, extras [4] f64, per-observation ints [8] int4, footprints [8][ROWS][W] f32
template <int R>
__host__ __device__ constexpr size_t k1_smem_bytes(int n_frames) {
  return sizeof(double) * ((size_t)n_frames * kPoseConst) +
         (size_t)kWarpsPerCta * (sizeof(double) * (kObsBatch * 20 + (size_t)n_frames * kUStride + 4) +
                                 sizeof(int4) * kObsBatch +
                                 sizeof(float) * (size_t)kStageSlots * Foot<R>::FLOATS);
}

// NCH: compile-time channel count (1 = Intensity, the north-star descriptor); 0 = runtime count.
template <int R, bool U8, int NCH>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 3) k1_eval(const EvalParams prm) {
  using FT = Foot<R>;
  constexpr int P = FT::P;
  constexpr int PR = (P + 31) / 32;             // pixel rounds per lane
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int F = prm.n_frames;
  const int C = NCH ? NCH : prm.fr.n_channels;
  const int CP = C * P;

  if (prm.st && prm.st->done) return;
  const int buf = prm.st ? prm.st->eval_buf : 0;
  const double* cams = prm.cams + (size_t)buf * F * 6;
  const double* pts = prm.pts + (size_t)buf * prm.n_points * 3;
  double* outV = prm.V + (size_t)buf * prm.n_points * 6;
  double* outgp = prm.gp + (size_t)buf * prm.n_points * 3;
  double* outW = prm.W + (size_t)buf * prm.nnz * 18;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_pose = reinterpret_cast<double*>(smem_raw);                       // [F][36]
  double* s_geo = s_pose + (size_t)F * kPoseConst;                            // [warps][8][20]
  double* s_geo_w = s_geo + warp * (kObsBatch * 20);
  double* s_U = s_geo + kWarpsPerCta * (kObsBatch * 20);                      // [warps][F][27]
  double* s_U_w = s_U + warp * F * kUStride;
  double* s_E = s_U + kWarpsPerCta * F * kUStride;                            // [warps][4]
  int4* s_gi = reinterpret_cast<int4*>(s_E + kWarpsPerCta * 4);               // [warps][8]
  int4* s_gi_w = s_gi + warp * kObsBatch;
  float* s_fp = reinterpret_cast<float*>(s_gi + kWarpsPerCta * kObsBatch);
  float* s_fp_w = s_fp + warp * (kStageSlots * FT::FLOATS);                   // [8][ROWS][W]

  if (threadIdx.x < F) pose_consts(cams + 6 * threadIdx.x, s_pose + threadIdx.x * kPoseConst);
  for (int i = lane; i < F * kUStride; i += 32) s_U_w[i] = 0.0;
  __syncthreads();

  // ---- per-lane constants --------------------------------------------------------------
  double pdx[PR], pdy[PR], wj[PR];
  float wjf[PR];
#pragma unroll
  for (int r = 0; r < PR; ++r) {
    const int j = lane + 32 * r;
    const int py = j / FT::SIDE, pxo = j - py * FT::SIDE;
    pdx[r] = (double)(pxo - R); pdy[r] = (double)(py - R);
    wj[r] = (j < P) ? prm.weights[j] : 0.0;
    wjf[r] = (float)wj[r];
  }
  // staging: lane (+32*round) <-> aligned 4-element word (row, wd) of the footprint
  int st_off[FT::ROUNDS];
#pragma unroll
  for (int rd = 0; rd < FT::ROUNDS; ++rd) {
    const int wi = lane + 32 * rd;
    const int row = wi / FT::NW, wd = wi - row * FT::NW;
    st_off[rd] = (wi < FT::WORDS) ? row * prm.fr.pitch + 4 * wd : -1;
  }
  // block-expansion roles: round 1 (pose block) and round 2 (W | V | g_p)
  int e1a = 0, e1b = 0, e2a = 0, e2b = 0;
  if (lane < 21) { e1a = c_pair6[lane][0]; e1b = c_pair6[lane][1]; }
  else if (lane < 27) { e1a = lane - 21; }
  if (lane < 18) { e2a = lane / 3; e2b = 6 + lane - 3 * (lane / 3); }
  else if (lane < 24) { e2a = 6 + c_pair3[lane - 18][0]; e2b = 6 + c_pair3[lane - 18][1]; }
  else if (lane < 27) { e2a = 6 + lane - 24; }

  const int p = blockIdx.x * kWarpsPerCta + warp;
  double cost_w = 0.0, gsq_w = 0.0, gmax_w = 0.0, xsq_w = 0.0;
  if (p < prm.n_points) {
    const int o0 = prm.obs_off[p], nobs = prm.obs_off[p + 1] - o0;
    const double X0 = pts[3 * p], X1 = pts[3 * p + 1], X2 = pts[3 * p + 2];
    xsq_w = X0 * X0 + X1 * X1 + X2 * X2;
    double acc_pt = 0.0;  // lanes 18..23: V entries, 24..26: g_p entries (summed over frames)
    const int obs_per_stage = NCH == 1 ? kStageSlots : ((C >= kStageSlots) ? 1 : kStageSlots / C);
    // reference descriptor of this point (channel 0 hoisted; further channels read in the loop)
    double p0c[PR];
#pragma unroll
    for (int r = 0; r < PR; ++r) {
      const int j = lane + 32 * r;
      p0c[r] = (j < P) ? (double)prm.desc[(size_t)p * CP + j] : 0.0;
    }

    for (int ob = 0; ob < nobs; ob += kObsBatch) {
      const int nb = min(kObsBatch, nobs - ob);
      // ---- (G) geometry: lane i <-> observation ob+i --------------------------------
      if (lane < nb) {
        const int g_f = prm.obs_frame[o0 + ob + lane];
        const double* pc = s_pose + g_f * kPoseConst;
        double Xc0, Xc1, Xc2;
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
        // Calibration::project: u = ((X*fx)/Z) + cx  (IEEE division, T = double path)
        const double u = __dadd_rn(__ddiv_rn(__dmul_rn(Xc0, prm.fx), Xc2), prm.cx);
        const double v = __dadd_rn(__ddiv_rn(__dmul_rn(Xc1, prm.fy), Xc2), prm.cy);
        double* g = s_geo_w + lane * 20;
        g[0] = u; g[1] = v;
        const double iz = 1.0 / Xc2;
        const double J00 = prm.fx * iz, J02 = -prm.fx * Xc0 * iz * iz;
        const double J11 = prm.fy * iz, J12 = -prm.fy * Xc1 * iz * iz;
        // D = -Rj [X]x M
        const double* Rj = pc + 18;
        const double* M = pc + 27;
        double D[9];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const double m0 = M[b], m1 = M[3 + b], m2 = M[6 + b];
          const double t0 = X1 * m2 - X2 * m1, t1 = X2 * m0 - X0 * m2, t2 = X0 * m1 - X1 * m0;
#pragma unroll
          for (int a = 0; a < 3; ++a) D[a * 3 + b] = -(Rj[a * 3] * t0 + Rj[a * 3 + 1] * t1 + Rj[a * 3 + 2] * t2);
        }
        const double* Rm = pc + 9;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          g[2 + b] = J00 * D[b] + J02 * D[6 + b];            // du/dw
          g[11 + b] = J11 * D[3 + b] + J12 * D[6 + b];       // dv/dw
          g[8 + b] = J00 * Rm[b] + J02 * Rm[6 + b];          // du/dX
          g[17 + b] = J11 * Rm[3 + b] + J12 * Rm[6 + b];     // dv/dX
        }
        g[5] = J00; g[6] = 0.0; g[7] = J02;                  // du/dt
        g[14] = 0.0; g[15] = J11; g[16] = J12;               // dv/dt
        // footprint origin and fast-path test (all taps and gradient taps interior)
        int4 gi = make_int4(g_f, 0, 0, 0);   // {frame, r0, cb (aligned first column), fast | (c0&3)<<1}
        if (fabs(u) < 1.0e8 && fabs(v) < 1.0e8) {
          const int c0 = (int)floor(u) - R - 1, r0 = (int)floor(v) - R - 1;
          const int fast = (c0 >= 0 && c0 + FT::ROWS - 1 <= prm.fr.cols - 1 && r0 >= 0 &&
                            r0 + FT::ROWS - 1 <= prm.fr.rows - 1) ? 1 : 0;
          gi.y = r0; gi.z = c0 & ~3; gi.w = fast;
        }
        s_gi_w[lane] = gi;
      }
      __syncwarp();

      for (int sb = 0; sb < nb; sb += obs_per_stage) {
        const int ns_obs = min(obs_per_stage, nb - sb);
        const int nslots = NCH == 1 ? ns_obs : ns_obs * C;
        // ---- (L) stage footprints: all loads first, then the stores ----------------------
        {
          uint32_t t8[kStageSlots][FT::ROUNDS];
          float4 t32[U8 ? 1 : kStageSlots][U8 ? 1 : FT::ROUNDS];
#pragma unroll
          for (int sl = 0; sl < kStageSlots; ++sl) {
            if (sl < nslots) {
              const int i = NCH == 1 ? sb + sl : sb + sl / C;
              const int k = NCH == 1 ? 0 : sl - (sl / C) * C;
              const int4 gi = s_gi_w[i];
              if (gi.w) {
                const unsigned base = (unsigned)((NCH == 1 ? gi.x : gi.x * C + k) * (int)prm.fr.plane + gi.y * prm.fr.pitch + gi.z);
#pragma unroll
                for (int rd = 0; rd < FT::ROUNDS; ++rd) {
                  if (st_off[rd] >= 0) {
                    if (U8) t8[sl][rd] = __ldg(reinterpret_cast<const uint32_t*>(prm.fr.u8 + (base + (unsigned)st_off[rd])));
                    else t32[U8 ? 0 : sl][U8 ? 0 : rd] = __ldg(reinterpret_cast<const float4*>(prm.fr.f32 + (base + (unsigned)st_off[rd])));
                  }
                }
              }
            }
          }
#pragma unroll
          for (int sl = 0; sl < kStageSlots; ++sl) {
            if (sl < nslots) {
              const int i = NCH == 1 ? sb + sl : sb + sl / C;
              if (s_gi_w[i].w) {
#pragma unroll
                for (int rd = 0; rd < FT::ROUNDS; ++rd) {
                  if (st_off[rd] >= 0) {
                    float4 o;
                    if (U8) {
                      const uint32_t q = t8[sl][rd];
                      o = make_float4((float)(q & 0xffu), (float)((q >> 8) & 0xffu), (float)((q >> 16) & 0xffu), (float)(q >> 24));
                    } else {
                      o = t32[U8 ? 0 : sl][U8 ? 0 : rd];
                    }
                    reinterpret_cast<float4*>(s_fp_w + sl * FT::FLOATS)[lane + 32 * rd] = o;
                  }
                }
              }
            }
          }
        }
        __syncwarp();

        // ---- (S)+(R) per observation: sample, reduce, expand -----------------------------
        for (int si = 0; si < ns_obs; ++si) {
          const int i = sb + si;
          const int o = o0 + ob + i;
          const int4 gi = s_gi_w[i];
          const int f = gi.x, r0 = gi.y, cb = gi.z, fast = gi.w;
          const double* g = s_geo_w + i * 20;
          const double u = g[0], v = g[1];
          double s_sum = 0.0;
          float G11 = 0.f, G12 = 0.f, G22 = 0.f, b1 = 0.f, b2 = 0.f;
          for (int k = 0; k < C; ++k) {
            const float* fp = s_fp_w + (NCH == 1 ? si : si * C + k) * FT::FLOATS;
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
                  const float* q = fp + (iy - r0) * FT::W + (ix - cb);
                  const float a11 = q[0], a12 = q[1], a21 = q[FT::W], a22 = q[FT::W + 1];
                  const float l1 = q[-1], r1 = q[2], l2 = q[FT::W - 1], r2 = q[FT::W + 2];
                  const float t1 = q[-FT::W], t2 = q[-FT::W + 1], u1 = q[2 * FT::W], u2 = q[2 * FT::W + 1];
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
                const double p0 = (NCH == 1 || k == 0) ? p0c[r] : (double)prm.desc[(size_t)p * CP + k * P + j];
                const double rr = __dmul_rn(wj[r], __dsub_rn(p0, (double)I1));   // photobundle.cc:720
                if (prm.residuals) prm.residuals[(size_t)o * CP + k * P + j] = rr;
                s_sum = fma(rr, rr, s_sum);
                const float rf = (float)rr;
                const float hx = wjf[r] * gx, hy = wjf[r] * gy;
                G11 = fmaf(hx, hx, G11); G12 = fmaf(hx, hy, G12); G22 = fmaf(hy, hy, G22);
                b1 = fmaf(rf, hx, b1); b2 = fmaf(rf, hy, b2);
              }
            }
          }
#pragma unroll
          for (int m = 16; m > 0; m >>= 1) {
            s_sum += __shfl_xor_sync(0xffffffffu, s_sum, m);
            G11 += __shfl_xor_sync(0xffffffffu, G11, m);
            G12 += __shfl_xor_sync(0xffffffffu, G12, m);
            G22 += __shfl_xor_sync(0xffffffffu, G22, m);
            b1 += __shfl_xor_sync(0xffffffffu, b1, m);
            b2 += __shfl_xor_sync(0xffffffffu, b2, m);
          }
          // ceres::HuberLoss + Corrector (rho'' <= 0 -> plain sqrt(rho') scaling)
          double rho0 = s_sum, rho1 = 1.0;
          if (prm.huber > 0.0 && s_sum > prm.huber * prm.huber) {
            const double rr = sqrt(s_sum);
            rho0 = 2.0 * prm.huber * rr - prm.huber * prm.huber;
            rho1 = fmax(DBL_MIN, prm.huber / rr);
          }
          if (lane == 0) {
            cost_w += 0.5 * rho0;
            if (prm.obs_sqnorm) prm.obs_sqnorm[o] = s_sum;
          }
          const double dG11 = rho1 * (double)G11, dG12 = rho1 * (double)G12, dG22 = rho1 * (double)G22;
          const double db1 = rho1 * (double)b1, db2 = rho1 * (double)b2;
          const bool free_cam = (f != prm.fixed_frame);
          const double* A = g + 2;
          // round 1: pose block of this observation (each frame at most once per point)
          if (free_cam && lane < 27) {
            const double val = lane < 21 ? quad(A, e1a, e1b, dG11, dG12, dG22) : -(A[e1a] * db1 + A[9 + e1a] * db2);
            s_U_w[f * kUStride + lane] += val;
          }
          // round 2: W (6x3) out, V / g_p into registers
          if (lane < 27) {
            const double val = lane < 24 ? quad(A, e2a, e2b, dG11, dG12, dG22) : -(A[e2a] * db1 + A[9 + e2a] * db2);
            if (lane < 18) outW[(size_t)o * 18 + lane] = free_cam ? val : 0.0;
            else acc_pt += val;
          }
        }
        __syncwarp();
      }
    }
    if (lane >= 18 && lane < 24) outV[(size_t)p * 6 + lane - 18] = acc_pt;
    else if (lane >= 24 && lane < 27) outgp[(size_t)p * 3 + lane - 24] = acc_pt;
    const double gq = (lane >= 24 && lane < 27) ? acc_pt : 0.0;
    double g2 = gq * gq, ga = fabs(gq);
#pragma unroll
    for (int m = 1; m <= 2; m <<= 1) {   // lanes 24..27 form an aligned group of four
      g2 += __shfl_xor_sync(0xffffffffu, g2, m);
      ga = fmax(ga, __shfl_xor_sync(0xffffffffu, ga, m));
    }
    gsq_w = __shfl_sync(0xffffffffu, g2, 24);
    gmax_w = __shfl_sync(0xffffffffu, ga, 24);
  }
  if (lane == 0) {
    s_E[warp * 4 + 0] = cost_w; s_E[warp * 4 + 1] = gsq_w; s_E[warp * 4 + 2] = gmax_w; s_E[warp * 4 + 3] = xsq_w;
  }
  __syncthreads();
  // CTA partial: sum the warps' accumulators in a fixed order (deterministic)
  for (int i = threadIdx.x; i < F * kUStride; i += blockDim.x) {
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < kWarpsPerCta; ++w) acc += s_U[w * F * kUStride + i];
    prm.Upart[(size_t)blockIdx.x * F * kUStride + i] = acc;
  }
  if (threadIdx.x == 0) {
    double c = 0.0, g2 = 0.0, gm = 0.0, x2 = 0.0;
    for (int w = 0; w < kWarpsPerCta; ++w) {
      c += s_E[w * 4]; g2 += s_E[w * 4 + 1]; gm = fmax(gm, s_E[w * 4 + 2]); x2 += s_E[w * 4 + 3];
    }
    double* e = prm.Epart + (size_t)blockIdx.x * 4;
    e[0] = c; e[1] = g2; e[2] = gm; e[3] = x2;
  }
}

// ---- host launcher -----------------------------------------------------------------------
template <int R, bool U8, int NCH>
static cudaError_t launch_one(const EvalParams& prm, cudaStream_t stream) {
  const size_t smem = k1_smem_bytes<R>(prm.n_frames);
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k1_eval<R, U8, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    configured[dev & 63] = true;
  }
  const int grid = (prm.n_points + kWarpsPerCta - 1) / kWarpsPerCta;
  if (grid == 0) return cudaSuccess;
  k1_eval<R, U8, NCH><<<grid, kWarpsPerCta * 32, smem, stream>>>(prm);
  return cudaGetLastError();
}

int k1_grid(int n_points) { return (n_points + kWarpsPerCta - 1) / kWarpsPerCta; }

size_t k1_smem(int radius, int n_frames) {
  switch (radius) {
    case 1: return k1_smem_bytes<1>(n_frames);
    case 2: return k1_smem_bytes<2>(n_frames);
    case 3: return k1_smem_bytes<3>(n_frames);
    default: return k1_smem_bytes<4>(n_frames);
  }
}

template <int R>
static cudaError_t launch_r(const EvalParams& prm, cudaStream_t stream) {
  if (prm.fr.u8) return launch_one<R, true, 1>(prm, stream);      // uint8 planes are always 1-channel Intensity
  if (prm.fr.n_channels == 1) return launch_one<R, false, 1>(prm, stream);
  return launch_one<R, false, 0>(prm, stream);
}

cudaError_t launch_k1(const EvalParams& prm, int radius, cudaStream_t stream) {
  switch (radius) {
    case 1: return launch_r<1>(prm, stream);
    case 2: return launch_r<2>(prm, stream);
    case 3: return launch_r<3>(prm, stream);
    case 4: return launch_r<4>(prm, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace pba
