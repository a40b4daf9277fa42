// pba_oracle.cc — CPU oracle for the photometric bundle-adjustment hot path.
//
// TEST INFRASTRUCTURE ONLY (see pba_oracle.h).  PARITY UNPINNED: the reference has
// no tests / golden vectors for this path and cannot be built here; this file is a
// restatement of the cited reference lines plus Ceres 1.x published semantics.
//
// Structure follows the reference, not the GPU design: one residual block per
// (point, observing frame), evaluated with 9-lane forward-mode dual numbers
// (ceres::Jet<double,9> of AutoDiffCostFunction<DescriptorError, DYNAMIC, 6, 3>,
// /root/reference/src/photobundle.cc:692), an fp32 bilinear sampler that injects the
// interpolated central-difference gradient as the derivative
// (src/sample_eigen.h:107-126, src/jet_extras.h:86-111), a block-wise Huber
// corrector, exact Schur elimination of the 3x3 point blocks, dense Cholesky of
// the reduced camera system and Ceres' Levenberg-Marquardt trust-region rules.
//
// Build: g++ -O3 -march=x86-64-v3 -ffp-contract=off -fopenmp (oracle/Makefile).
// -ffp-contract=off matters: the reference is built without FMA
// (CMakeLists.txt:22-25: -msse2 -mssse3 -msse4.1 only), so no a*b+c contraction.

#include "pba_oracle.h"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ----------------------------------------------------------------------------
// Forward-mode dual number: value + N partial derivatives (ceres/jet.h semantics).
// ----------------------------------------------------------------------------
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0.0) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  explicit Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; v[k] = 1.0; }
};

template <int N> inline Jet<N> operator-(const Jet<N>& f) {
  Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i];
  return h;
}
template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i];
  return h;
}
template <int N> inline Jet<N> operator+(const Jet<N>& f, double s) { Jet<N> h = f; h.a = f.a + s; return h; }
template <int N> inline Jet<N> operator+(double s, const Jet<N>& f) { Jet<N> h = f; h.a = f.a + s; return h; }
template <int N> inline Jet<N>& operator+=(Jet<N>& f, const Jet<N>& g) { f = f + g; return f; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i];
  return h;
}
template <int N> inline Jet<N> operator-(const Jet<N>& f, double s) { Jet<N> h = f; h.a = f.a - s; return h; }
template <int N> inline Jet<N> operator-(double s, const Jet<N>& f) {
  Jet<N> h; h.a = s - f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i];
  return h;
}
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a * g.a;
  for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a;
  return h;
}
template <int N> inline Jet<N> operator*(const Jet<N>& f, double s) {
  Jet<N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s;
  return h;
}
template <int N> inline Jet<N> operator*(double s, const Jet<N>& f) { return f * s; }
// ceres/jet.h: the quotient's value is f.a * (1/g.a), not f.a / g.a.
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h;
  const double g_a_inverse = 1.0 / g.a;
  const double f_a_by_g_a = f.a * g_a_inverse;
  h.a = f_a_by_g_a;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse;
  return h;
}
template <int N> inline Jet<N> operator/(double s, const Jet<N>& g) {
  Jet<N> h;
  const double minus_s_g_a_inverse2 = -s / (g.a * g.a);
  h.a = s / g.a;
  for (int i = 0; i < N; ++i) h.v[i] = g.v[i] * minus_s_g_a_inverse2;
  return h;
}
template <int N> inline Jet<N> sqrt(const Jet<N>& f) {
  Jet<N> h; const double t = std::sqrt(f.a); const double two_a_inverse = 1.0 / (2.0 * t);
  h.a = t; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * two_a_inverse;
  return h;
}
template <int N> inline Jet<N> cos(const Jet<N>& f) {
  Jet<N> h; h.a = std::cos(f.a); const double ms = -std::sin(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = ms * f.v[i];
  return h;
}
template <int N> inline Jet<N> sin(const Jet<N>& f) {
  Jet<N> h; h.a = std::sin(f.a); const double c = std::cos(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i];
  return h;
}
inline double sqrt(double x) { return std::sqrt(x); }
inline double cos(double x) { return std::cos(x); }
inline double sin(double x) { return std::sin(x); }

inline double scalar_of(double x) { return x; }
template <int N> inline double scalar_of(const Jet<N>& x) { return x.a; }

template <class T> struct Lift;
template <> struct Lift<double> { static double make(double s) { return s; } };
template <int N> struct Lift<Jet<N>> { static Jet<N> make(double s) { return Jet<N>(s); } };

// ----------------------------------------------------------------------------
// ceres::AngleAxisRotatePoint (ceres/rotation.h; call site src/photobundle.cc:700).
// ----------------------------------------------------------------------------
template <class T>
inline void AngleAxisRotatePoint(const T aa[3], const T pt[3], T result[3]) {
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (scalar_of(theta2) > std::numeric_limits<double>::epsilon()) {
    const T theta = sqrt(theta2);
    const T costheta = cos(theta);
    const T sintheta = sin(theta);
    const T theta_inverse = 1.0 / theta;
    const T w[3] = {aa[0] * theta_inverse, aa[1] * theta_inverse, aa[2] * theta_inverse};
    const T w_cross_pt[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2],
                             w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (1.0 - costheta);
    result[0] = pt[0] * costheta + w_cross_pt[0] * sintheta + w[0] * tmp;
    result[1] = pt[1] * costheta + w_cross_pt[1] * sintheta + w[1] * tmp;
    result[2] = pt[2] * costheta + w_cross_pt[2] * sintheta + w[2] * tmp;
  } else {
    // first-order Taylor branch: R ~ I + [aa]x
    const T w_cross_pt[3] = {aa[1] * pt[2] - aa[2] * pt[1], aa[2] * pt[0] - aa[0] * pt[2],
                             aa[0] * pt[1] - aa[1] * pt[0]};
    result[0] = pt[0] + w_cross_pt[0];
    result[1] = pt[1] + w_cross_pt[1];
    result[2] = pt[2] + w_cross_pt[2];
  }
}

// ----------------------------------------------------------------------------
// Sampler: src/sample_eigen.h:33-102.  All types exactly as in the reference:
// TPixel = float; the literal 1.0 is a double, the literal 1 an int.
// ----------------------------------------------------------------------------
inline int trunc_to_int_x86(float x) {
  // static_cast<int>(float) on x86-64 is cvttss2si: out-of-range and NaN inputs
  // produce INT_MIN ("integer indefinite"). Formally UB in C++; pin the behaviour
  // the reference binary has so that the GPU path can match it.
  if (!(x > -2147483648.0f && x < 2147483648.0f)) return INT_MIN;
  return static_cast<int>(x);
}

inline void LinearInitAxis(float x, int size, int* x1, int* x2, float* dx) {
  const int ix = trunc_to_int_x86(x);
  if (ix < 0) {
    *x1 = 0; *x2 = 0; *dx = 1.0;
  } else if (ix > size - 2) {
    *x1 = size - 1; *x2 = size - 1; *dx = 1.0;
  } else {
    *x1 = ix; *x2 = ix + 1; *dx = *x2 - x;
  }
}

struct PlaneSet {
  const float* I;
  const float* Gx;
  const float* Gy;
  int rows, cols;
};

inline void SampleLinear(const PlaneSet& ps, float y, float x, float* sample) {
  int x1, y1, x2, y2;
  float dx, dy;
  LinearInitAxis(y, ps.rows, &y1, &y2, &dy);
  LinearInitAxis(x, ps.cols, &x1, &x2, &dx);
  const size_t c = (size_t)ps.cols;
  const size_t i11 = y1 * c + x1, i12 = y1 * c + x2, i21 = y2 * c + x1, i22 = y2 * c + x2;
  {
    const float im11 = ps.I[i11], im12 = ps.I[i12], im21 = ps.I[i21], im22 = ps.I[i22];
    sample[0] = (dy * (dx * im11 + (1.0 - dx) * im12) + (1 - dy) * (dx * im21 + (1.0 - dx) * im22));
  }
  {
    const float g11 = ps.Gx[i11], g12 = ps.Gx[i12], g21 = ps.Gx[i21], g22 = ps.Gx[i22];
    sample[1] = (dy * (dx * g11 + (1.0 - dx) * g12) + (1 - dy) * (dx * g21 + (1.0 - dx) * g22));
  }
  {
    const float g11 = ps.Gy[i11], g12 = ps.Gy[i12], g21 = ps.Gy[i21], g22 = ps.Gy[i22];
    sample[2] = (dy * (dx * g11 + (1.0 - dx) * g12) + (1 - dy) * (dx * g21 + (1.0 - dx) * g22));
  }
}

// src/sample_eigen.h:107-126 + ceres::Chain<float,2,T>::Rule (src/jet_extras.h:74-111).
inline double SampleWithDerivative(const PlaneSet& ps, const double& x, const double& y) {
  const float sx = (float)x, sy = (float)y;
  float sample[3];
  SampleLinear(ps, sy, sx, sample);
  return sample[0];  // scalar Chain::Rule returns f
}
template <int N>
inline Jet<N> SampleWithDerivative(const PlaneSet& ps, const Jet<N>& x, const Jet<N>& y) {
  const float sx = (float)x.a, sy = (float)y.a;
  float sample[3];
  SampleLinear(ps, sy, sx, sample);
  Jet<N> f;
  f.a = sample[0];
  const double dfdx = sample[1], dfdy = sample[2];
  for (int i = 0; i < N; ++i) f.v[i] = dfdx * x.v[i] + dfdy * y.v[i];
  return f;
}

// ----------------------------------------------------------------------------
// Problem view with gradients resolved.
// ----------------------------------------------------------------------------
struct View {
  const oracle_problem* pb;
  std::vector<float> gx_store, gy_store;
  const float* gx;
  const float* gy;
  int P, CP;
  size_t plane_elems;
  PlaneSet planes(int frame, int k) const {
    const size_t off = ((size_t)frame * pb->n_channels + k) * plane_elems;
    return PlaneSet{pb->planes + off, gx + off, gy + off, pb->rows, pb->cols};
  }
};

void imgradient(const float* I, int rows, int cols, float* gx, float* gy) {
  // src/imgproc.cc:27-96; scale 0.5 from src/imgproc.h:54-58.
  std::memset(gx, 0, sizeof(float) * (size_t)rows * cols);
  std::memset(gy, 0, sizeof(float) * (size_t)rows * cols);
  for (int y = 1; y < rows - 1; ++y) {
    const float* srow = I + (size_t)y * cols;
    float* ix = gx + (size_t)y * cols;
    float* iy = gy + (size_t)y * cols;
    for (int x = 1; x < cols - 1; ++x) {
      ix[x] = 0.5f * (srow[x + 1] - srow[x - 1]);
      iy[x] = 0.5f * (srow[x + cols] - srow[x - cols]);
    }
  }
}

void make_view(const oracle_problem* pb, View* v) {
  v->pb = pb;
  const int side = 2 * pb->radius + 1;
  v->P = side * side;
  v->CP = v->P * pb->n_channels;
  v->plane_elems = (size_t)pb->rows * pb->cols;
  if (pb->grad_x && pb->grad_y) {
    v->gx = pb->grad_x;
    v->gy = pb->grad_y;
  } else {
    const size_t n = v->plane_elems * pb->n_frames * pb->n_channels;
    v->gx_store.resize(n);
    v->gy_store.resize(n);
    const int np = pb->n_frames * pb->n_channels;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < np; ++i)
      imgradient(pb->planes + i * v->plane_elems, pb->rows, pb->cols,
                 v->gx_store.data() + i * v->plane_elems, v->gy_store.data() + i * v->plane_elems);
    v->gx = v->gx_store.data();
    v->gy = v->gy_store.data();
  }
}

// ----------------------------------------------------------------------------
// DescriptorError::operator()<T> — src/photobundle.cc:696-727.
// ----------------------------------------------------------------------------
template <class T>
inline void descriptor_error(const View& vw, int frame, const T* camera, const T* point,
                             const double* p0, T* residuals) {
  const oracle_problem& pb = *vw.pb;
  T xw[3];
  AngleAxisRotatePoint(camera, point, xw);
  xw[0] = xw[0] + camera[3];
  xw[1] = xw[1] + camera[4];
  xw[2] = xw[2] + camera[5];

  // Calibration::project, src/calibration.h:33-38: u = ((X*fx)/Z) + cx.
  const T u_w = ((xw[0] * pb.fx) / xw[2]) + pb.cx;
  const T v_w = ((xw[1] * pb.fy) / xw[2]) + pb.cy;

  const int r = pb.radius;
  int i = 0;
  for (int k = 0; k < pb.n_channels; ++k) {
    const PlaneSet ps = vw.planes(frame, k);
    int j = 0;
    for (int y = -r; y <= r; ++y) {
      const T v = v_w + (double)y;
      for (int x = -r; x <= r; ++x, ++i, ++j) {
        const T u = u_w + (double)x;
        const T i0 = Lift<T>::make(p0[i]);
        const T i1 = SampleWithDerivative(ps, u, v);
        residuals[i] = pb.weights[j] * (i0 - i1);
      }
    }
  }
}
// double / double for the scalar path is a true division (operator/ above only
// applies to Jets) — exactly the reference's T=double instantiation.

// Closed-form Jacobian of the same residual (cross-check of the dual-number path;
// SURVEY.md App. A.4).
void residual_block_analytic(const View& vw, int frame, const double* cam, const double* X,
                             const double* p0, double* r, double* Jc, double* Jp) {
  const oracle_problem& pb = *vw.pb;
  const double* w = cam;
  const double theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double R[9], dXdw[9];
  if (theta2 > std::numeric_limits<double>::epsilon()) {
    const double th = std::sqrt(theta2), c = std::cos(th), s = std::sin(th);
    const double k[3] = {w[0] / th, w[1] / th, w[2] / th};
    const double K[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
        R[a * 3 + b] = (a == b ? c : 0.0) + s * K[a * 3 + b] + (1.0 - c) * k[a] * k[b];
    // dXc/dw = -R [X]x (w w^T + (R^T - I) [w]x) / theta^2
    const double Wx[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double M[9];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        double acc = w[a] * w[b];
        for (int q = 0; q < 3; ++q) acc += (R[q * 3 + a] - (q == a ? 1.0 : 0.0)) * Wx[q * 3 + b];
        M[a * 3 + b] = acc / theta2;
      }
    const double Xx[9] = {0, -X[2], X[1], X[2], 0, -X[0], -X[1], X[0], 0};
    double XM[9];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        double acc = 0;
        for (int q = 0; q < 3; ++q) acc += Xx[a * 3 + q] * M[q * 3 + b];
        XM[a * 3 + b] = acc;
      }
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        double acc = 0;
        for (int q = 0; q < 3; ++q) acc += R[a * 3 + q] * XM[q * 3 + b];
        dXdw[a * 3 + b] = -acc;
      }
  } else {
    const double Wx[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    for (int a = 0; a < 9; ++a) R[a] = Wx[a];
    R[0] += 1.0; R[4] += 1.0; R[8] += 1.0;
    const double Xx[9] = {0, -X[2], X[1], X[2], 0, -X[0], -X[1], X[0], 0};
    for (int a = 0; a < 9; ++a) dXdw[a] = -Xx[a];
  }
  double Xc[3];
  AngleAxisRotatePoint(cam, X, Xc);
  Xc[0] += cam[3]; Xc[1] += cam[4]; Xc[2] += cam[5];
  const double u_w = ((Xc[0] * pb.fx) / Xc[2]) + pb.cx;
  const double v_w = ((Xc[1] * pb.fy) / Xc[2]) + pb.cy;
  const double iz = 1.0 / Xc[2];
  const double Juv[6] = {pb.fx * iz, 0.0, -pb.fx * Xc[0] * iz * iz,
                         0.0, pb.fy * iz, -pb.fy * Xc[1] * iz * iz};
  double A[18];  // 2 x 9: d(u,v)/d[w t X]
  for (int a = 0; a < 2; ++a) {
    for (int b = 0; b < 3; ++b) {
      double acc = 0, accp = 0;
      for (int q = 0; q < 3; ++q) {
        acc += Juv[a * 3 + q] * dXdw[q * 3 + b];
        accp += Juv[a * 3 + q] * R[q * 3 + b];
      }
      A[a * 9 + b] = acc;
      A[a * 9 + 3 + b] = Juv[a * 3 + b];
      A[a * 9 + 6 + b] = accp;
    }
  }
  const int rad = pb.radius;
  int i = 0;
  for (int k = 0; k < pb.n_channels; ++k) {
    const PlaneSet ps = vw.planes(frame, k);
    int j = 0;
    for (int y = -rad; y <= rad; ++y) {
      const double v = v_w + (double)y;
      for (int x = -rad; x <= rad; ++x, ++i, ++j) {
        const double u = u_w + (double)x;
        float sample[3];
        SampleLinear(ps, (float)v, (float)u, sample);
        const double wj = pb.weights[j];
        r[i] = wj * (p0[i] - (double)sample[0]);
        const double gx = sample[1], gy = sample[2];
        if (Jc)
          for (int q = 0; q < 6; ++q) Jc[i * 6 + q] = -wj * (gx * A[q] + gy * A[9 + q]);
        if (Jp)
          for (int q = 0; q < 3; ++q) Jp[i * 3 + q] = -wj * (gx * A[6 + q] + gy * A[15 + q]);
      }
    }
  }
}

void residual_block_autodiff(const View& vw, int frame, const double* cam, const double* X,
                             const double* p0, double* r, double* Jc, double* Jp) {
  typedef Jet<9> J9;
  J9 jc[6], jp[3];
  for (int q = 0; q < 6; ++q) jc[q] = J9(cam[q], q);
  for (int q = 0; q < 3; ++q) jp[q] = J9(X[q], 6 + q);
  J9 res[512];
  std::vector<J9> big;
  J9* out = res;
  if (vw.CP > 512) { big.resize(vw.CP); out = big.data(); }
  descriptor_error<J9>(vw, frame, jc, jp, p0, out);
  for (int i = 0; i < vw.CP; ++i) {
    r[i] = out[i].a;
    if (Jc) for (int q = 0; q < 6; ++q) Jc[i * 6 + q] = out[i].v[q];
    if (Jp) for (int q = 0; q < 3; ++q) Jp[i * 3 + q] = out[i].v[6 + q];
  }
}

// ceres::HuberLoss::Evaluate (ceres/loss_function.cc).
inline void huber(double a, double s, double rho[3]) {
  const double b = a * a;
  if (s > b) {
    const double r = std::sqrt(s);
    rho[0] = 2.0 * a * r - b;
    rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
    rho[2] = -rho[1] / (2.0 * s);
  } else {
    rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
  }
}

int set_threads(const oracle_problem* pb) {
#ifdef _OPENMP
  int n = pb->num_threads > 0 ? pb->num_threads : omp_get_max_threads();
  return n;
#else
  (void)pb; return 1;
#endif
}

double cost_only(const View& vw, const double* cams, const double* points) {
  const oracle_problem& pb = *vw.pb;
  const int nt = set_threads(&pb);
  // per-thread partial sums combined in thread order: the same thread count always gives the same bits (an OpenMP
  // reduction clause combines in arrival order)
  std::vector<double> part((size_t)nt * 8, 0.0);
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int p = 0; p < pb.n_points; ++p) {
#ifdef _OPENMP
    double& total = part[(size_t)omp_get_thread_num() * 8];
#else
    double& total = part[0];
#endif
    double r[512];
    std::vector<double> big;
    double* rr = r;
    if (vw.CP > 512) { big.resize(vw.CP); rr = big.data(); }
    for (int o = pb.obs_offsets[p]; o < pb.obs_offsets[p + 1]; ++o) {
      const int f = pb.obs_frame[o];
      descriptor_error<double>(vw, f, cams + 6 * f, points + 3 * p, pb.desc + (size_t)p * vw.CP, rr);
      double s = 0.0;
      for (int i = 0; i < vw.CP; ++i) s += rr[i] * rr[i];
      if (pb.huber > 0.0) {
        double rho[3];
        huber(pb.huber, s, rho);
        total += 0.5 * rho[0];
      } else {
        total += 0.5 * s;
      }
    }
  }
  double total = 0.0;
  for (int k = 0; k < nt; ++k) total += part[(size_t)k * 8];
  return total;
}

// Evaluate residuals + Jacobians and reduce to block normal equations
// (ResidualBlock::Evaluate + Corrector semantics).
void evaluate_blocks(const View& vw, const double* cams, const double* points, int use_autodiff,
                     oracle_blocks* out) {
  const oracle_problem& pb = *vw.pb;
  const int F = pb.n_frames, n = pb.n_points, CP = vw.CP;
  const int nnz = pb.obs_offsets[n];
  std::memset(out->U, 0, sizeof(double) * F * 36);
  std::memset(out->gc, 0, sizeof(double) * F * 6);
  std::memset(out->V, 0, sizeof(double) * (size_t)n * 9);
  std::memset(out->gp, 0, sizeof(double) * (size_t)n * 3);
  std::memset(out->W, 0, sizeof(double) * (size_t)nnz * 18);
  const int nt = set_threads(&pb);
  std::vector<double> Ut((size_t)nt * F * 36, 0.0), gct((size_t)nt * F * 6, 0.0), ct(nt, 0.0);
#pragma omp parallel num_threads(nt)
  {
#ifdef _OPENMP
    const int tid = omp_get_thread_num();
#else
    const int tid = 0;
#endif
    double* Um = Ut.data() + (size_t)tid * F * 36;
    double* gcm = gct.data() + (size_t)tid * F * 6;
    std::vector<double> r(CP), Jc((size_t)CP * 6), Jp((size_t)CP * 3);
    double cost = 0.0;
#pragma omp for schedule(static)
    for (int p = 0; p < n; ++p) {
      double* V = out->V + (size_t)p * 9;
      double* gp = out->gp + (size_t)p * 3;
      for (int o = pb.obs_offsets[p]; o < pb.obs_offsets[p + 1]; ++o) {
        const int f = pb.obs_frame[o];
        const double* p0 = pb.desc + (size_t)p * CP;
        if (use_autodiff)
          residual_block_autodiff(vw, f, cams + 6 * f, points + 3 * p, p0, r.data(), Jc.data(), Jp.data());
        else
          residual_block_analytic(vw, f, cams + 6 * f, points + 3 * p, p0, r.data(), Jc.data(), Jp.data());
        double s = 0.0;
        for (int i = 0; i < CP; ++i) s += r[i] * r[i];
        if (out->obs_sqnorm) out->obs_sqnorm[o] = s;
        if (out->residuals) std::memcpy(out->residuals + (size_t)o * CP, r.data(), sizeof(double) * CP);
        double scale = 1.0;
        if (pb.huber > 0.0) {
          double rho[3];
          huber(pb.huber, s, rho);
          cost += 0.5 * rho[0];
          scale = std::sqrt(rho[1]);  // Corrector: rho'' <= 0 for Huber -> plain sqrt(rho') scaling
        } else {
          cost += 0.5 * s;
        }
        if (scale != 1.0) {
          for (int i = 0; i < CP; ++i) r[i] *= scale;
          for (int i = 0; i < CP * 6; ++i) Jc[i] *= scale;
          for (int i = 0; i < CP * 3; ++i) Jp[i] *= scale;
        }
        const bool free_cam = (f != pb.fixed_frame);
        double* W = out->W + (size_t)o * 18;
        for (int i = 0; i < CP; ++i) {
          const double* jc = Jc.data() + i * 6;
          const double* jp = Jp.data() + i * 3;
          for (int a = 0; a < 3; ++a) {
            for (int b = 0; b < 3; ++b) V[a * 3 + b] += jp[a] * jp[b];
            gp[a] += jp[a] * r[i];
          }
          if (free_cam) {
            for (int a = 0; a < 6; ++a) {
              for (int b = 0; b < 6; ++b) Um[f * 36 + a * 6 + b] += jc[a] * jc[b];
              for (int b = 0; b < 3; ++b) W[a * 3 + b] += jc[a] * jp[b];
              gcm[f * 6 + a] += jc[a] * r[i];
            }
          }
        }
      }
    }
    ct[tid] = cost;
  }
  double cost = 0.0;
  for (int t = 0; t < nt; ++t) {
    cost += ct[t];
    for (int i = 0; i < F * 36; ++i) out->U[i] += Ut[(size_t)t * F * 36 + i];
    for (int i = 0; i < F * 6; ++i) out->gc[i] += gct[(size_t)t * F * 6 + i];
  }
  out->cost = cost;
}

// ----------------------------------------------------------------------------
// small dense helpers
// ----------------------------------------------------------------------------
bool cholesky_inplace(double* A, int n) {  // lower, row-major
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  return true;
}
void cholesky_solve(const double* L, int n, double* b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i * n + k] * b[k];
    b[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * b[k];
    b[i] = s / L[i * n + i];
  }
}
bool invert_spd3(const double* A, double* inv) {  // via Cholesky (InvertPSDMatrix, full rank)
  double L[9];
  std::memcpy(L, A, sizeof(L));
  if (!cholesky_inplace(L, 3)) return false;
  for (int c = 0; c < 3; ++c) {
    double e[3] = {0, 0, 0};
    e[c] = 1.0;
    cholesky_solve(L, 3, e);
    for (int r = 0; r < 3; ++r) inv[r * 3 + c] = e[r];
  }
  return true;
}

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ----------------------------------------------------------------------------
// Levenberg-Marquardt step on the block system (LevenbergMarquardtStrategy::ComputeStep
// + SchurComplementSolver, in the Jacobi-scaled space).  Returns false on a linear
// solver failure (=> invalid step).
// ----------------------------------------------------------------------------
struct Workspace {
  std::vector<double> U, gc, V, gp, W;      // blocks at x (unscaled)
  std::vector<double> scale_c, scale_p;     // Jacobi scaling (computed once at iteration 0)
  std::vector<double> step_c, step_p;       // trust-region step in the scaled space
  std::vector<int> free_index;              // frame -> index among free frames (-1: fixed)
  int n_free;
};

bool compute_step(const oracle_problem& pb, Workspace& ws, double radius, double min_diag,
                  double max_diag, double* model_cost_change) {
  const int F = pb.n_frames, n = pb.n_points, nf = ws.n_free, dim = 6 * nf;
  const int nt = set_threads(&pb);
  std::vector<double> S((size_t)dim * dim, 0.0), rhs(dim, 0.0);
  std::vector<double> St((size_t)nt * dim * dim, 0.0), rt((size_t)nt * dim, 0.0);
  std::vector<double> Vinv((size_t)n * 9);
  bool ok = true;

#pragma omp parallel num_threads(nt)
  {
#ifdef _OPENMP
    const int tid = omp_get_thread_num();
#else
    const int tid = 0;
#endif
    double* Sm = St.data() + (size_t)tid * dim * dim;
    double* rm = rt.data() + (size_t)tid * dim;
#pragma omp for schedule(static)
    for (int p = 0; p < n; ++p) {
      const double* sp = ws.scale_p.data() + 3 * p;
      double Vs[9], gs[3];
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) Vs[a * 3 + b] = sp[a] * ws.V[(size_t)p * 9 + a * 3 + b] * sp[b];
        gs[a] = sp[a] * ws.gp[(size_t)p * 3 + a];
      }
      for (int a = 0; a < 3; ++a) {
        const double d = std::min(std::max(Vs[a * 3 + a], min_diag), max_diag);
        Vs[a * 3 + a] += d / radius;
      }
      double* Vi = Vinv.data() + (size_t)p * 9;
      if (!invert_spd3(Vs, Vi)) {
#pragma omp atomic write
        ok = false;
        continue;
      }
      const int o0 = pb.obs_offsets[p], o1 = pb.obs_offsets[p + 1];
      for (int oa = o0; oa < o1; ++oa) {
        const int fa = ws.free_index[pb.obs_frame[oa]];
        if (fa < 0) continue;
        const double* sca = ws.scale_c.data() + 6 * pb.obs_frame[oa];
        double Wa[18], Y[18];
        for (int a = 0; a < 6; ++a)
          for (int b = 0; b < 3; ++b) Wa[a * 3 + b] = sca[a] * ws.W[(size_t)oa * 18 + a * 3 + b] * sp[b];
        for (int a = 0; a < 6; ++a)
          for (int b = 0; b < 3; ++b) {
            double acc = 0;
            for (int q = 0; q < 3; ++q) acc += Wa[a * 3 + q] * Vi[q * 3 + b];
            Y[a * 3 + b] = acc;
          }
        for (int a = 0; a < 6; ++a) {
          double acc = 0;
          for (int q = 0; q < 3; ++q) acc += Y[a * 3 + q] * gs[q];
          rm[6 * fa + a] -= acc;
        }
        for (int ob = o0; ob < o1; ++ob) {
          const int fb = ws.free_index[pb.obs_frame[ob]];
          if (fb < 0) continue;
          const double* scb = ws.scale_c.data() + 6 * pb.obs_frame[ob];
          for (int a = 0; a < 6; ++a)
            for (int b = 0; b < 6; ++b) {
              double acc = 0;
              for (int q = 0; q < 3; ++q)
                acc += Y[a * 3 + q] * (scb[b] * ws.W[(size_t)ob * 18 + b * 3 + q] * sp[q]);
              Sm[(size_t)(6 * fa + a) * dim + 6 * fb + b] -= acc;
            }
        }
      }
    }
  }
  if (!ok) return false;
  for (int t = 0; t < nt; ++t) {
    for (size_t i = 0; i < (size_t)dim * dim; ++i) S[i] += St[(size_t)t * dim * dim + i];
    for (int i = 0; i < dim; ++i) rhs[i] += rt[(size_t)t * dim + i];
  }
  for (int f = 0; f < F; ++f) {
    const int fi = ws.free_index[f];
    if (fi < 0) continue;
    const double* sc = ws.scale_c.data() + 6 * f;
    for (int a = 0; a < 6; ++a) {
      for (int b = 0; b < 6; ++b)
        S[(size_t)(6 * fi + a) * dim + 6 * fi + b] += sc[a] * ws.U[f * 36 + a * 6 + b] * sc[b];
      const double h = sc[a] * ws.U[f * 36 + a * 6 + a] * sc[a];
      const double d = std::min(std::max(h, min_diag), max_diag);
      S[(size_t)(6 * fi + a) * dim + 6 * fi + a] += d / radius;
      rhs[6 * fi + a] += sc[a] * ws.gc[f * 6 + a];
    }
  }
  // Solve S y_c = rhs  (the linear solver solves J y = r; the LM step is -y).
  if (dim > 0) {
    if (!cholesky_inplace(S.data(), dim)) return false;
    cholesky_solve(S.data(), dim, rhs.data());
  }
  ws.step_c.assign((size_t)F * 6, 0.0);
  for (int f = 0; f < F; ++f) {
    const int fi = ws.free_index[f];
    if (fi < 0) continue;
    for (int a = 0; a < 6; ++a) ws.step_c[f * 6 + a] = -rhs[6 * fi + a];
  }
  ws.step_p.assign((size_t)n * 3, 0.0);
  // Back-substitution and model cost change  -s^T g_s - 0.5 s^T H_s s.
  std::vector<double> part((size_t)nt * 8, 0.0);   // {s.g, s'Hs} per thread, combined in thread order (deterministic)
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int p = 0; p < n; ++p) {
#ifdef _OPENMP
    double& sg = part[(size_t)omp_get_thread_num() * 8];
    double& sHs = part[(size_t)omp_get_thread_num() * 8 + 1];
#else
    double& sg = part[0];
    double& sHs = part[1];
#endif
    const double* sp = ws.scale_p.data() + 3 * p;
    double t[3];
    for (int a = 0; a < 3; ++a) t[a] = sp[a] * ws.gp[(size_t)p * 3 + a];
    for (int o = pb.obs_offsets[p]; o < pb.obs_offsets[p + 1]; ++o) {
      const int f = pb.obs_frame[o];
      if (ws.free_index[f] < 0) continue;
      const double* sc = ws.scale_c.data() + 6 * f;
      // y_c = -step_c
      for (int b = 0; b < 3; ++b) {
        double acc = 0;
        for (int a = 0; a < 6; ++a) acc += (sc[a] * ws.W[(size_t)o * 18 + a * 3 + b] * sp[b]) * (-ws.step_c[f * 6 + a]);
        t[b] -= acc;
      }
    }
    const double* Vi = Vinv.data() + (size_t)p * 9;
    double sv[3];
    for (int a = 0; a < 3; ++a) {
      double y = 0;
      for (int b = 0; b < 3; ++b) y += Vi[a * 3 + b] * t[b];
      sv[a] = -y;
      ws.step_p[(size_t)p * 3 + a] = sv[a];
    }
    // model terms for this point
    for (int a = 0; a < 3; ++a) {
      sg += sv[a] * sp[a] * ws.gp[(size_t)p * 3 + a];
      for (int b = 0; b < 3; ++b) sHs += sv[a] * (sp[a] * ws.V[(size_t)p * 9 + a * 3 + b] * sp[b]) * sv[b];
    }
    for (int o = pb.obs_offsets[p]; o < pb.obs_offsets[p + 1]; ++o) {
      const int f = pb.obs_frame[o];
      if (ws.free_index[f] < 0) continue;
      const double* sc = ws.scale_c.data() + 6 * f;
      for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 3; ++b)
          sHs += 2.0 * ws.step_c[f * 6 + a] * (sc[a] * ws.W[(size_t)o * 18 + a * 3 + b] * sp[b]) * sv[b];
    }
  }
  double sg = 0.0, sHs = 0.0;
  for (int k = 0; k < nt; ++k) { sg += part[(size_t)k * 8]; sHs += part[(size_t)k * 8 + 1]; }
  for (int f = 0; f < F; ++f) {
    if (ws.free_index[f] < 0) continue;
    const double* sc = ws.scale_c.data() + 6 * f;
    for (int a = 0; a < 6; ++a) {
      sg += ws.step_c[f * 6 + a] * sc[a] * ws.gc[f * 6 + a];
      for (int b = 0; b < 6; ++b)
        sHs += ws.step_c[f * 6 + a] * (sc[a] * ws.U[f * 36 + a * 6 + b] * sc[b]) * ws.step_c[f * 6 + b];
    }
  }
  *model_cost_change = -sg - 0.5 * sHs;
  for (size_t i = 0; i < ws.step_c.size(); ++i)
    if (!std::isfinite(ws.step_c[i])) return false;
  return true;
}

}  // namespace

// ============================================================================
// C API
// ============================================================================
extern "C" {

void oracle_default_options(oracle_solver_options* o) {
  o->max_num_iterations = 500;
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-6;
  o->parameter_tolerance = 1e-6;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->max_num_consecutive_invalid_steps = 5;
  o->jacobi_scaling = 1;
  o->use_autodiff = 1;
}

void oracle_imgradient(const float* I, int32_t rows, int32_t cols, float* gx, float* gy) {
  imgradient(I, rows, cols, gx, gy);
}

// ---- descriptor channels (DescriptorFrame::Create, src/photobundle.cc:220-248) ----------------------
static inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

int32_t oracle_descriptor_channels(int32_t type) { return type == 0 ? 1 : type == 1 ? 3 : type == 2 ? 8 : -1; }

// the uint8 stages of computeBitPlanes: pre-blur and census transform (exposed for the golden-vector test)
void oracle_bitplanes_stages(const uint8_t* img, int32_t rows, int32_t cols, uint8_t* blur_out, uint8_t* census_out) {
  const size_t n = (size_t)rows * cols;
  uint8_t* blurp = blur_out;
  std::memset(census_out, 0, n);
  const int w3[3] = {70, 116, 70};
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      int acc = 0;
      for (int j = 0; j < 3; ++j) {
        const uint8_t* row = img + (size_t)reflect101(y + j - 1, rows) * cols;
        int h = 0;
        for (int i = 0; i < 3; ++i) h += w3[i] * (int)row[reflect101(x + i - 1, cols)];
        acc += w3[j] * h;
      }
      blurp[(size_t)y * cols + x] = (uint8_t)((acc + 32768) >> 16);
    }
  // (2) censusTransform (src/imgproc.cc:140-220): bit k set when neighbour k >= centre (unsigned), neighbours in
  //     row-major order skipping the centre; first/last row and column are 0
  for (int y = 1; y < rows - 1; ++y)
    for (int x = 1; x < cols - 1; ++x) {
      const uint8_t* s = blurp + (size_t)y * cols + x;
      const unsigned c = s[0];
      census_out[(size_t)y * cols + x] = (uint8_t)((s[-cols - 1] >= c ? 0x01 : 0) | (s[-cols] >= c ? 0x02 : 0) | (s[-cols + 1] >= c ? 0x04 : 0) |
                                               (s[-1] >= c ? 0x08 : 0) | (s[1] >= c ? 0x10 : 0) | (s[cols - 1] >= c ? 0x20 : 0) |
                                               (s[cols] >= c ? 0x40 : 0) | (s[cols + 1] >= c ? 0x80 : 0));
    }
}

// planes: dense [C][rows][cols]
void oracle_build_channels(int32_t type, const uint8_t* img, int32_t rows, int32_t cols, float* planes) {
  const size_t n = (size_t)rows * cols;
  if (type == 0 || type == 1) {
    // image.cast<float>() (:226, :229)
    for (size_t i = 0; i < n; ++i) planes[i] = (float)img[i];
    if (type == 1) {
      // imgradient(uint8) (src/imgproc.cc:27-106): S * (float(a) - float(b)), S = 0.5f, zero first/last row and column
      float* gx = planes + n;
      float* gy = planes + 2 * n;
      std::memset(gx, 0, sizeof(float) * n);
      std::memset(gy, 0, sizeof(float) * n);
      for (int y = 1; y < rows - 1; ++y)
        for (int x = 1; x < cols - 1; ++x) {
          const uint8_t* s = img + (size_t)y * cols + x;
          gx[(size_t)y * cols + x] = 0.5f * ((float)s[1] - (float)s[-1]);
          gy[(size_t)y * cols + x] = 0.5f * ((float)s[cols] - (float)s[-cols]);
        }
    }
    return;
  }
  // BitPlanes: computeBitPlanes(image, sigma_ct = 1, sigma_bp = 1.5) (src/imgproc.cc:222-245, src/imgproc.h:45-46)
  // (1) cv::GaussianBlur(uint8, 3x3, sigma 1): fixed-point kernel {70,116,70}/256 per axis, (sum + 2^15) >> 16,
  //     BORDER_REFLECT_101 (pinned against cv2 in tests/golden/make_bitplanes_golden.py)
  std::vector<uint8_t> blur(n), census(n, 0);
  oracle_bitplanes_stages(img, rows, cols, blur.data(), census.data());
  // (3) ExtractBitPlanesChannel: bit b as 0/1 float, cv::GaussianBlur(float, 5x5, sigma 1.5): separable fp32,
  //     kernel = float(getGaussianKernel(5, 1.5)), BORDER_REFLECT_101
  const float k0 = 0x1.2b1778p-2f, k1 = 0x1.defcep-3f, k2 = 0x1.ebd75p-4f;
  std::vector<float> bit(n), rowp(n);
  for (int b = 0; b < 8; ++b) {
    for (size_t i = 0; i < n; ++i) bit[i] = (float)((census[i] & (1 << b)) >> b);
    for (int y = 0; y < rows; ++y) {
      const float* r = bit.data() + (size_t)y * cols;
      for (int x = 0; x < cols; ++x) {
        const float a0 = r[reflect101(x - 2, cols)], a1 = r[reflect101(x - 1, cols)], a2 = r[x], a3 = r[reflect101(x + 1, cols)],
                    a4 = r[reflect101(x + 2, cols)];
        rowp[(size_t)y * cols + x] = (k0 * a2 + k1 * (a1 + a3)) + k2 * (a0 + a4);
      }
    }
    float* dst = planes + (size_t)b * n;
    for (int y = 0; y < rows; ++y)
      for (int x = 0; x < cols; ++x) {
        const float r0 = rowp[(size_t)reflect101(y - 2, rows) * cols + x], r1 = rowp[(size_t)reflect101(y - 1, rows) * cols + x],
                    r2 = rowp[(size_t)y * cols + x], r3 = rowp[(size_t)reflect101(y + 1, rows) * cols + x],
                    r4 = rowp[(size_t)reflect101(y + 2, rows) * cols + x];
        dst[(size_t)y * cols + x] = (k0 * r2 + k1 * (r1 + r3)) + k2 * (r0 + r4);
      }
  }
}

// DescriptorFrame::computeSaliencyMap (src/photobundle.cc:212-220): sum over channels of |Ix| + |Iy|
void oracle_saliency_map(const float* planes, int32_t n_channels, int32_t rows, int32_t cols, float* out) {
  const size_t n = (size_t)rows * cols;
  std::vector<float> gx(n), gy(n);
  for (int k = 0; k < n_channels; ++k) {
    imgradient(planes + (size_t)k * n, rows, cols, gx.data(), gy.data());
    for (size_t i = 0; i < n; ++i) {
      const float m = std::fabs(gx[i]) + std::fabs(gy[i]);
      out[i] = k == 0 ? m : out[i] + m;
    }
  }
}

// ExtractPatch (src/photobundle.cc:466-479) for every channel, channel-major
void oracle_extract_patches(const float* planes, int32_t n_channels, int32_t rows, int32_t cols, int32_t radius, int32_t n,
                            const int32_t* xy, double* desc) {
  const int side = 2 * radius + 1, P = side * side;
  const int max_cols = cols - radius - 1, max_rows = rows - radius - 1;
  for (int p = 0; p < n; ++p)
    for (int k = 0; k < n_channels; ++k) {
      const float* I = planes + (size_t)k * rows * cols;
      int i = 0;
      for (int r = -radius; r <= radius; ++r) {
        const int r_i = std::max(radius, std::min(xy[2 * p + 1] + r, max_rows));
        for (int c = -radius; c <= radius; ++c, ++i) {
          const int c_i = std::max(radius, std::min(xy[2 * p] + c, max_cols));
          desc[((size_t)p * n_channels + k) * P + i] = (double)I[(size_t)r_i * cols + c_i];
        }
      }
    }
}

void oracle_sample_linear(const float* I, const float* Gx, const float* Gy, int32_t rows,
                          int32_t cols, float y, float x, float* out3) {
  PlaneSet ps{I, Gx, Gy, rows, cols};
  SampleLinear(ps, y, x, out3);
}

void oracle_project(const double* k4, const double* X, double* uv) {
  // the expression of descriptor_error() / residual_jacobian() above (Calibration::project, src/calibration.h:33-38)
  uv[0] = ((X[0] * k4[0]) / X[2]) + k4[2];
  uv[1] = ((X[1] * k4[1]) / X[2]) + k4[3];
}

void oracle_patch_weights(int32_t radius, int32_t do_gaussian, double* w) {
  // src/photobundle.cc:617-644 with s_x = s_y = a = 1.
  const int n = (2 * radius + 1) * (2 * radius + 1);
  if (do_gaussian) {
    double sum = 0.0;
    int i = 0;
    for (int r = -radius; r <= radius; ++r) {
      const double d_r = (r * r) / 1.0;
      for (int c = -radius; c <= radius; ++c, ++i) {
        const double d_c = (c * c) / 1.0;
        const double v = 1.0 * std::exp(-0.5 * (d_r + d_c));
        w[i] = v;
        sum += v;
      }
    }
    for (i = 0; i < n; ++i) w[i] /= sum;
  } else {
    for (int i = 0; i < n; ++i) w[i] = 1.0;
  }
}

void oracle_angle_axis_rotate_point(const double* aa, const double* pt, double* out) {
  AngleAxisRotatePoint<double>(aa, pt, out);
}

void oracle_pose_to_params(const double* T, double* p) {
  // ceres::RotationMatrixToAngleAxis = RotationMatrixToQuaternion + QuaternionToAngleAxis.
  auto R = [&](int i, int j) { return T[j * 4 + i]; };  // column-major 4x4
  double q[4];
  const double trace = R(0, 0) + R(1, 1) + R(2, 2);
  if (trace >= 0.0) {
    double t = std::sqrt(trace + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (R(2, 1) - R(1, 2)) * t;
    q[2] = (R(0, 2) - R(2, 0)) * t;
    q[3] = (R(1, 0) - R(0, 1)) * t;
  } else {
    int i = 0;
    if (R(1, 1) > R(0, 0)) i = 1;
    if (R(2, 2) > R(i, i)) i = 2;
    const int j = (i + 1) % 3;
    const int k = (j + 1) % 3;
    double t = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0);
    q[i + 1] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R(k, j) - R(j, k)) * t;
    q[j + 1] = (R(j, i) + R(i, j)) * t;
    q[k + 1] = (R(k, i) + R(i, k)) * t;
  }
  const double sin_squared_theta = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (sin_squared_theta > 0.0) {
    const double sin_theta = std::sqrt(sin_squared_theta);
    const double cos_theta = q[0];
    const double two_theta = 2.0 * ((cos_theta < 0.0) ? std::atan2(-sin_theta, -cos_theta)
                                                      : std::atan2(sin_theta, cos_theta));
    const double k = two_theta / sin_theta;
    p[0] = q[1] * k; p[1] = q[2] * k; p[2] = q[3] * k;
  } else {
    p[0] = q[1] * 2.0; p[1] = q[2] * 2.0; p[2] = q[3] * 2.0;
  }
  p[3] = T[12]; p[4] = T[13]; p[5] = T[14];
}

void oracle_params_to_pose(const double* p, double* T) {
  // ceres::AngleAxisToRotationMatrix (Rodrigues; first-order form near zero).
  double R[9];  // R[i + 3*j] column-major 3x3
  auto at = [&](int i, int j) -> double& { return R[i + 3 * j]; };
  const double theta2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
  if (theta2 > std::numeric_limits<double>::epsilon()) {
    const double theta = std::sqrt(theta2);
    const double wx = p[0] / theta, wy = p[1] / theta, wz = p[2] / theta;
    const double c = std::cos(theta), s = std::sin(theta);
    at(0, 0) = c + wx * wx * (1.0 - c);
    at(1, 0) = wz * s + wx * wy * (1.0 - c);
    at(2, 0) = -wy * s + wx * wz * (1.0 - c);
    at(0, 1) = wx * wy * (1.0 - c) - wz * s;
    at(1, 1) = c + wy * wy * (1.0 - c);
    at(2, 1) = wx * s + wy * wz * (1.0 - c);
    at(0, 2) = wy * s + wx * wz * (1.0 - c);
    at(1, 2) = -wx * s + wy * wz * (1.0 - c);
    at(2, 2) = c + wz * wz * (1.0 - c);
  } else {
    at(0, 0) = 1.0; at(1, 0) = p[2]; at(2, 0) = -p[1];
    at(0, 1) = -p[2]; at(1, 1) = 1.0; at(2, 1) = p[0];
    at(0, 2) = p[1]; at(1, 2) = -p[0]; at(2, 2) = 1.0;
  }
  for (int i = 0; i < 16; ++i) T[i] = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[j * 4 + i] = at(i, j);
  T[12] = p[3]; T[13] = p[4]; T[14] = p[5]; T[15] = 1.0;
}

void oracle_residual_block(const oracle_problem* pb, int32_t frame, const double* cam6,
                           const double* xyz, const double* desc, int32_t use_autodiff,
                           double* r, double* Jc, double* Jp) {
  View vw;
  make_view(pb, &vw);
  if (use_autodiff == 1)
    residual_block_autodiff(vw, frame, cam6, xyz, desc, r, Jc, Jp);
  else if (use_autodiff == 0)
    residual_block_analytic(vw, frame, cam6, xyz, desc, r, Jc, Jp);
  else
    descriptor_error<double>(vw, frame, cam6, xyz, desc, r);  // T = double path
}

double oracle_cost(const oracle_problem* pb, const double* cams, const double* points) {
  View vw;
  make_view(pb, &vw);
  return cost_only(vw, cams, points);
}

void oracle_evaluate(const oracle_problem* pb, const double* cams, const double* points,
                     int32_t use_autodiff, oracle_blocks* out) {
  View vw;
  make_view(pb, &vw);
  evaluate_blocks(vw, cams, points, use_autodiff, out);
}

int32_t oracle_solve(const oracle_problem* pb, const oracle_solver_options* opt, double* cams,
                     double* points, oracle_summary* sum, oracle_iteration_summary* trace) {
  const double t_start = now_s();
  View vw;
  make_view(pb, &vw);
  const int F = pb->n_frames, n = pb->n_points, nnz = pb->obs_offsets[n];
  std::memset(sum, 0, sizeof(*sum));
  sum->num_residual_blocks = nnz;
  sum->num_residuals = nnz * vw.CP;
  sum->fixed_cost = 0.0;

  Workspace ws;
  ws.U.resize((size_t)F * 36); ws.gc.resize((size_t)F * 6);
  ws.V.resize((size_t)n * 9); ws.gp.resize((size_t)n * 3); ws.W.resize((size_t)nnz * 18);
  ws.free_index.assign(F, -1);
  ws.n_free = 0;
  // A camera is a parameter block of the ceres::Problem only if some residual block
  // uses it (src/photobundle.cc:802); unused cameras are not optimised.
  std::vector<char> used(F, 0);
  for (int o = 0; o < nnz; ++o) used[pb->obs_frame[o]] = 1;
  for (int f = 0; f < F; ++f)
    if (used[f] && f != pb->fixed_frame) ws.free_index[f] = ws.n_free++;

  oracle_blocks blk;
  blk.U = ws.U.data(); blk.gc = ws.gc.data(); blk.V = ws.V.data(); blk.gp = ws.gp.data();
  blk.W = ws.W.data(); blk.obs_sqnorm = nullptr; blk.residuals = nullptr;

  auto x_norm_of = [&](const double* c, const double* p) {
    double s = 0.0;
    for (int f = 0; f < F; ++f)
      if (ws.free_index[f] >= 0)
        for (int a = 0; a < 6; ++a) s += c[f * 6 + a] * c[f * 6 + a];
    for (size_t i = 0; i < (size_t)n * 3; ++i) s += p[i] * p[i];
    return std::sqrt(s);
  };
  auto gradient_norms = [&](double* gmax, double* gnorm) {
    double m = 0.0, s = 0.0;
    for (int f = 0; f < F; ++f)
      if (ws.free_index[f] >= 0)
        for (int a = 0; a < 6; ++a) {
          const double g = ws.gc[f * 6 + a];
          m = std::max(m, std::fabs(g)); s += g * g;
        }
    for (size_t i = 0; i < (size_t)n * 3; ++i) {
      const double g = ws.gp[i];
      m = std::max(m, std::fabs(g)); s += g * g;
    }
    *gmax = m; *gnorm = std::sqrt(s);
  };

  int n_trace = 0;
  auto push = [&](const oracle_iteration_summary& it) {
    if (trace) trace[n_trace] = it;
    ++n_trace;
  };
  auto finish = [&](int type, const char* msg, double x_cost) {
    sum->termination_type = type;
    std::snprintf(sum->message, sizeof(sum->message), "%s", msg);
    sum->final_cost = x_cost;
    sum->num_iterations = n_trace;
    sum->total_time_in_seconds = now_s() - t_start;
    return type;
  };

  // ---- iteration 0 ---------------------------------------------------------
  double t0 = now_s();
  evaluate_blocks(vw, cams, points, opt->use_autodiff, &blk);
  sum->jacobian_time_in_seconds += now_s() - t0;
  sum->num_jacobian_evals++;
  double x_cost = blk.cost;
  sum->initial_cost = x_cost;
  double x_norm = x_norm_of(cams, points);

  ws.scale_c.assign((size_t)F * 6, 1.0);
  ws.scale_p.assign((size_t)n * 3, 1.0);
  if (opt->jacobi_scaling) {
    for (int f = 0; f < F; ++f)
      for (int a = 0; a < 6; ++a) ws.scale_c[f * 6 + a] = 1.0 / (1.0 + std::sqrt(ws.U[f * 36 + a * 7]));
    for (int p = 0; p < n; ++p)
      for (int a = 0; a < 3; ++a) ws.scale_p[(size_t)p * 3 + a] = 1.0 / (1.0 + std::sqrt(ws.V[(size_t)p * 9 + a * 4]));
  }

  oracle_iteration_summary it;
  std::memset(&it, 0, sizeof(it));
  it.iteration = 0;
  it.cost = x_cost;
  it.trust_region_radius = opt->initial_trust_region_radius;
  gradient_norms(&it.gradient_max_norm, &it.gradient_norm);
  it.iteration_time_in_seconds = now_s() - t_start;
  it.cumulative_time_in_seconds = now_s() - t_start;

  double radius = opt->initial_trust_region_radius;
  double decrease_factor = 2.0;
  int num_consecutive_invalid = 0;
  char msg[256];

  if (it.gradient_max_norm <= opt->gradient_tolerance) {
    push(it);
    std::snprintf(msg, sizeof(msg), "Gradient tolerance reached. Gradient max norm: %e <= %e",
                  it.gradient_max_norm, opt->gradient_tolerance);
    return finish(0, msg, x_cost);
  }

  std::vector<double> cand_c((size_t)F * 6), cand_p((size_t)n * 3);

  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    it.trust_region_radius = radius;
    it.cumulative_time_in_seconds = now_s() - t_start;
    push(it);
    if (it.step_is_successful) sum->num_successful_steps++;
    else if (it.iteration > 0) sum->num_unsuccessful_steps++;
    if (it.iteration >= opt->max_num_iterations) {
      std::snprintf(msg, sizeof(msg), "Maximum number of iterations reached. Number of iterations: %d.", it.iteration);
      return finish(1, msg, x_cost);
    }
    if (it.gradient_max_norm <= opt->gradient_tolerance) {
      std::snprintf(msg, sizeof(msg), "Gradient tolerance reached. Gradient max norm: %e <= %e",
                    it.gradient_max_norm, opt->gradient_tolerance);
      return finish(0, msg, x_cost);
    }
    if (!(radius > opt->min_trust_region_radius)) {
      std::snprintf(msg, sizeof(msg), "Minimum trust region radius reached. Trust region radius: %e <= %e",
                    radius, opt->min_trust_region_radius);
      return finish(0, msg, x_cost);
    }

    const double iter_start = now_s();
    oracle_iteration_summary prev = it;
    std::memset(&it, 0, sizeof(it));
    it.iteration = prev.iteration + 1;
    it.gradient_max_norm = prev.gradient_max_norm;
    it.gradient_norm = prev.gradient_norm;

    // ComputeTrustRegionStep
    double model_cost_change = 0.0;
    t0 = now_s();
    const bool solved = compute_step(*pb, ws, radius, opt->min_lm_diagonal, opt->max_lm_diagonal, &model_cost_change);
    it.step_solver_time_in_seconds = now_s() - t0;
    sum->linear_solver_time_in_seconds += it.step_solver_time_in_seconds;
    it.linear_solver_iterations = 1;
    it.step_is_valid = solved && (model_cost_change > 0.0);

    if (!it.step_is_valid) {
      // HandleInvalidStep
      if (++num_consecutive_invalid >= opt->max_num_consecutive_invalid_steps) {
        std::snprintf(msg, sizeof(msg), "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps: %d",
                      opt->max_num_consecutive_invalid_steps);
        return finish(2, msg, x_cost);
      }
      radius = radius / decrease_factor;  // StepIsInvalid() == StepRejected(0)
      decrease_factor *= 2.0;
      it.cost = x_cost;
      it.cost_change = 0.0;
      it.step_norm = 0.0;
      it.relative_decrease = 0.0;
      it.iteration_time_in_seconds = now_s() - iter_start;
      continue;
    }
    num_consecutive_invalid = 0;

    // delta = step .* scale; candidate = x + delta
    double step_sq = 0.0;
    for (int f = 0; f < F; ++f)
      for (int a = 0; a < 6; ++a) {
        const double d = ws.step_c[f * 6 + a] * ws.scale_c[f * 6 + a];
        cand_c[f * 6 + a] = cams[f * 6 + a] + (ws.free_index[f] >= 0 ? d : 0.0);
        const double diff = cams[f * 6 + a] - cand_c[f * 6 + a];
        step_sq += diff * diff;
      }
    for (size_t i = 0; i < (size_t)n * 3; ++i) {
      cand_p[i] = points[i] + ws.step_p[i] * ws.scale_p[i];
      const double diff = points[i] - cand_p[i];
      step_sq += diff * diff;
    }
    t0 = now_s();
    double candidate_cost = cost_only(vw, cand_c.data(), cand_p.data());
    sum->cost_time_in_seconds += now_s() - t0;
    sum->num_cost_evals++;
    if (!std::isfinite(candidate_cost)) candidate_cost = std::numeric_limits<double>::max();

    // ParameterToleranceReached
    it.step_norm = std::sqrt(step_sq);
    const double step_size_tolerance = opt->parameter_tolerance * (x_norm + opt->parameter_tolerance);
    if (it.step_norm <= step_size_tolerance) {
      std::snprintf(msg, sizeof(msg), "Parameter tolerance reached. Relative step_norm: %e <= %e.",
                    it.step_norm / (x_norm + opt->parameter_tolerance), opt->parameter_tolerance);
      it.cost = x_cost;
      return finish(0, msg, x_cost);
    }
    // FunctionToleranceReached
    it.cost_change = x_cost - candidate_cost;
    const double absolute_function_tolerance = opt->function_tolerance * x_cost;
    if (std::fabs(it.cost_change) <= absolute_function_tolerance) {
      std::snprintf(msg, sizeof(msg), "Function tolerance reached. |cost_change|/cost: %e <= %e",
                    std::fabs(it.cost_change) / x_cost, opt->function_tolerance);
      it.cost = x_cost;
      return finish(0, msg, x_cost);
    }

    it.relative_decrease = it.cost_change / model_cost_change;
    if (it.relative_decrease > opt->min_relative_decrease) {
      // HandleSuccessfulStep
      std::memcpy(cams, cand_c.data(), sizeof(double) * F * 6);
      std::memcpy(points, cand_p.data(), sizeof(double) * (size_t)n * 3);
      x_norm = x_norm_of(cams, points);
      t0 = now_s();
      evaluate_blocks(vw, cams, points, opt->use_autodiff, &blk);
      sum->jacobian_time_in_seconds += now_s() - t0;
      sum->num_jacobian_evals++;
      x_cost = blk.cost;
      gradient_norms(&it.gradient_max_norm, &it.gradient_norm);
      it.step_is_successful = 1;
      it.cost = x_cost;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * it.relative_decrease - 1.0, 3));
      radius = std::min(opt->max_trust_region_radius, radius);
      decrease_factor = 2.0;
    } else {
      // HandleUnsuccessfulStep
      it.step_is_successful = 0;
      it.cost = candidate_cost;
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
    }
    it.iteration_time_in_seconds = now_s() - iter_start;
  }
}

}  // extern "C"
