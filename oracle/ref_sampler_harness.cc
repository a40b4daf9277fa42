// Compiles the REFERENCE's own sampler — /root/reference/src/sample_eigen.h and
// src/jet_extras.h, included from where they lie — behind a C ABI, so that the
// restated sampler in pba_oracle.cc can be checked against it bit for bit.
// Built only into oracle/_ref/ (git-ignored). Test infrastructure; no reference
// source is copied into this repository.
#include <stdint.h>
#include "sample_eigen.h"   // -I /root/reference/src

namespace {
struct Img {                 // the TImage concept SampleLinear needs
  typedef float Scalar;
  const float* p; int r, c;
  int rows() const { return r; }
  int cols() const { return c; }
  float operator()(int y, int x) const { return p[(long)y * c + x]; }
};
}

extern "C" {

void ref_sample_linear(const float* I, const float* Gx, const float* Gy, int32_t rows,
                       int32_t cols, float y, float x, float* out3) {
  Img i{I, rows, cols}, gx{Gx, rows, cols}, gy{Gy, rows, cols};
  SampleLinear(i, gx, gy, y, x, out3);
}

double ref_sample_with_derivative_double(const float* I, const float* Gx, const float* Gy,
                                         int32_t rows, int32_t cols, double x, double y) {
  Img i{I, rows, cols}, gx{Gx, rows, cols}, gy{Gy, rows, cols};
  return SampleWithDerivative(i, gx, gy, x, y);
}

void ref_sample_with_derivative_jet9(const float* I, const float* Gx, const float* Gy,
                                     int32_t rows, int32_t cols, double xa, const double* xv,
                                     double ya, const double* yv, double* out_a, double* out_v) {
  Img i{I, rows, cols}, gx{Gx, rows, cols}, gy{Gy, rows, cols};
  typedef ceres::Jet<double, 9> J;
  J x, y;
  x.a = xa; y.a = ya;
  for (int k = 0; k < 9; ++k) { x.v[k] = xv[k]; y.v[k] = yv[k]; }
  const J f = SampleWithDerivative(i, gx, gy, x, y);
  *out_a = f.a;
  for (int k = 0; k < 9; ++k) out_v[k] = f.v[k];
}

}
