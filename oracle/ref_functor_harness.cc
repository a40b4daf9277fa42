// The REFERENCE's residual functor, T = double: the body of DescriptorError::operator() (/root/reference/src/
// photobundle.cc:696-727, cut out of that file at build time by the Makefile and included INSIDE the struct below),
// calling the reference's own Calibration::project (src/calibration.h) and SampleWithDerivative (src/sample_eigen.h),
// both included from where they lie.  What is NOT the reference's: the small frame / gradient structs that give the
// body its `_frame->getChannel(k)` / `getChannelGradient(k).Ix()` members (the reference's DescriptorFrame needs real
// Eigen), and ceres::AngleAxisRotatePoint (ref_shim/ceres/rotation.h, restated from the published algorithm: Ceres is
// absent).  Used to check the oracle's residuals against the reference's own code path bit for bit.
// Built only into oracle/_ref/ (git-ignored).  Test infrastructure; no reference source is copied into this repository.
#include <stdint.h>
#include <cstddef>
#include <vector>
#include "ceres/jet.h"          // ref_shim: layout + restated dual-number arithmetic
#include "ceres/rotation.h"     // ref_shim
#include "sample_eigen.h"       // -I /root/reference/src
#include "calibration.h"

namespace {
struct Plane {                  // the TImage concept of the sampler
  typedef float Scalar;
  const float* p; int r, c;
  int rows() const { return r; }
  int cols() const { return c; }
  float operator()(int y, int x) const { return p[(long)y * c + x]; }
};
struct Gradient {
  Plane gx, gy;
  const Plane& Ix() const { return gx; }
  const Plane& Iy() const { return gy; }
};
struct Frame {
  std::vector<Plane> channels;
  std::vector<Gradient> gradients;
  size_t numChannels() const { return channels.size(); }
  const Plane& getChannel(size_t k) const { return channels[k]; }
  const Gradient& getChannelGradient(size_t k) const { return gradients[k]; }
};
struct Functor {
  int _radius;
  const Calibration& _calib;
  const double* _p0;
  const Frame* _frame;
  const double* _patch_weights;
#include REF_FUNCTOR_BODY_INC   // template <class T> inline bool operator()(camera, point, residuals) const { ... }
};
}  // namespace

extern "C" {
// planes: [C][rows][cols] channel values, gx / gy the same shape (the reference's imgradient of each channel)
int32_t ref_residual_block(const float* planes, const float* gx, const float* gy, int32_t n_channels, int32_t rows, int32_t cols,
                           const double* k4, int32_t radius, const double* p0, const double* weights, const double* cam6,
                           const double* xyz, double* residuals) {
  Mat33 K;
  K << k4[0], 0.0, k4[2], 0.0, k4[1], k4[3], 0.0, 0.0, 1.0;
  const Calibration calib(K, 0.1);
  Frame f;
  const size_t plane = (size_t)rows * cols;
  for (int k = 0; k < n_channels; ++k) {
    f.channels.push_back(Plane{planes + k * plane, rows, cols});
    f.gradients.push_back(Gradient{Plane{gx + k * plane, rows, cols}, Plane{gy + k * plane, rows, cols}});
  }
  const Functor fn{radius, calib, p0, &f, weights};
  return fn(cam6, xyz, residuals) ? 1 : 0;
}

// The same functor instantiated with ceres::Jet<double, 9> the way AutoDiffCostFunction<.., DYNAMIC, 6, 3> seeds it
// (camera parameters -> derivative lanes 0..5, point -> lanes 6..8): residuals [CP], Jacobian rows [CP][9].
// The sampler's chain rule is the reference's own (src/jet_extras.h); the Jet algebra and AngleAxisRotatePoint are restated.
int32_t ref_residual_block_jet(const float* planes, const float* gx, const float* gy, int32_t n_channels, int32_t rows, int32_t cols,
                               const double* k4, int32_t radius, const double* p0, const double* weights, const double* cam6,
                               const double* xyz, double* residuals, double* jac9) {
  typedef ceres::Jet<double, 9> J;
  Mat33 K;
  K << k4[0], 0.0, k4[2], 0.0, k4[1], k4[3], 0.0, 0.0, 1.0;
  const Calibration calib(K, 0.1);
  Frame f;
  const size_t plane = (size_t)rows * cols;
  for (int k = 0; k < n_channels; ++k) {
    f.channels.push_back(Plane{planes + k * plane, rows, cols});
    f.gradients.push_back(Gradient{Plane{gx + k * plane, rows, cols}, Plane{gy + k * plane, rows, cols}});
  }
  const Functor fn{radius, calib, p0, &f, weights};
  J cam[6], pt[3];
  for (int k = 0; k < 6; ++k) cam[k] = J(cam6[k], k);
  for (int k = 0; k < 3; ++k) pt[k] = J(xyz[k], 6 + k);
  const int CP = n_channels * (2 * radius + 1) * (2 * radius + 1);
  std::vector<J> r(CP);
  const bool ok = fn(cam, pt, r.data());
  for (int i = 0; i < CP; ++i) {
    residuals[i] = r[i].a;
    for (int k = 0; k < 9; ++k) jac9[i * 9 + k] = r[i].v[k];
  }
  return ok ? 1 : 0;
}
}
