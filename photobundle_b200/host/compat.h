// compat.h — Eigen/Ceres/Boost-free stand-ins for the types that appear in the reference's
// public BA interface (src/types.h, src/calibration.h, <ceres/iteration_callback.h>), so that
// photobundle.h keeps the reference's names and signatures in an image where those libraries
// do not exist.  Where Eigen IS available a maintainer can instead typedef these to the Eigen
// types; layouts match: Mat44 is column-major like Eigen::Matrix<double,4,4>.
#ifndef PBA_HOST_COMPAT_H
#define PBA_HOST_COMPAT_H

#include <array>
#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

struct Vec2 { double v[2]; double& operator[](int i) { return v[i]; } const double& operator[](int i) const { return v[i]; } };
struct Vec3 {
  double v[3];
  Vec3() : v{0, 0, 0} {}
  Vec3(double x, double y, double z) : v{x, y, z} {}
  double& operator[](int i) { return v[i]; }
  const double& operator[](int i) const { return v[i]; }
  const double* data() const { return v; }
  double* data() { return v; }
};

// 3x3 / 4x4 column-major (Eigen default), element access (row, col)
struct Mat33 {
  double m[9];
  Mat33() { for (double& x : m) x = 0; }
  static Mat33 Identity() { Mat33 r; r.m[0] = r.m[4] = r.m[8] = 1; return r; }
  double& operator()(int r, int c) { return m[c * 3 + r]; }
  const double& operator()(int r, int c) const { return m[c * 3 + r]; }
  Vec3 operator*(const Vec3& x) const {
    return Vec3((*this)(0, 0) * x[0] + (*this)(0, 1) * x[1] + (*this)(0, 2) * x[2],
                (*this)(1, 0) * x[0] + (*this)(1, 1) * x[1] + (*this)(1, 2) * x[2],
                (*this)(2, 0) * x[0] + (*this)(2, 1) * x[1] + (*this)(2, 2) * x[2]);
  }
  Mat33 inverse() const;
};

struct Mat44 {
  double m[16];
  Mat44() { for (double& x : m) x = 0; }
  static Mat44 Identity() { Mat44 r; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1; return r; }
  double& operator()(int r, int c) { return m[c * 4 + r]; }
  const double& operator()(int r, int c) const { return m[c * 4 + r]; }
  const double* data() const { return m; }
  double* data() { return m; }
  Mat44 operator*(const Mat44& o) const;
  Vec3 transform(const Vec3& x) const {   // Eigen::Isometry3d * Vec3
    return Vec3((*this)(0, 0) * x[0] + (*this)(0, 1) * x[1] + (*this)(0, 2) * x[2] + (*this)(0, 3),
                (*this)(1, 0) * x[0] + (*this)(1, 1) * x[1] + (*this)(1, 2) * x[2] + (*this)(1, 3),
                (*this)(2, 0) * x[0] + (*this)(2, 1) * x[1] + (*this)(2, 2) * x[2] + (*this)(2, 3));
  }
  Mat44 inverse() const;        // general inverse (Trajectory::push_back uses Matrix::inverse())
  Mat44 rigidInverse() const;   // Eigen::Isometry3d(T).inverse().matrix()
};

template <class T> using EigenAlignedContainer_ = std::vector<T>;
template <class T> using UniquePointer = std::unique_ptr<T>;
typedef EigenAlignedContainer_<Mat44> PoseList;

// src/types.h:58-77
struct ImageSize {
  int rows = 0, cols = 0;
  ImageSize(int r = 0, int c = 0) : rows(r), cols(c) {}
  int numel() const { return rows * cols; }
  int area() const { return numel(); }
  bool empty() const { return 0 == numel(); }
  ImageSize pyrDown() const { return ImageSize((rows + 1) / 2, (cols + 1) / 2); }
};

// src/calibration.h
class Calibration {
 public:
  Calibration() {}
  Calibration(const Mat33& K, double b) : _K(K), _baseline(b) {}
  const double& b() const { return _baseline; }
  const double& fx() const { return _K(0, 0); }
  const double& fy() const { return _K(1, 1); }
  const double& cx() const { return _K(0, 2); }
  const double& cy() const { return _K(1, 2); }
  const Mat33& K() const { return _K; }
  Mat33& K() { return _K; }
  double& baseline() { return _baseline; }
  // src/calibration.h:43  project(Vec3) = normHomog(K * X)
  Vec2 project(const Vec3& X) const {
    const Vec3 p = _K * X;
    Vec2 r; r[0] = p[0] / p[2]; r[1] = p[1] / p[2];
    return r;
  }
  Calibration pyrDown() const {   // src/calibration.h:72-78
    Mat33 K(_K);
    for (double& x : K.m) x *= 0.5;
    K(2, 2) = 1.0;
    return Calibration(K, _baseline * 2);
  }
 private:
  Mat33 _K;
  double _baseline = 0.0;
};

// Field names of ceres::IterationSummary as the reference lists them (src/ceres_cereal.h:12-30).
namespace ceres {
struct IterationSummary {
  int iteration = 0;
  bool step_is_valid = false, step_is_nonmonotonic = false, step_is_successful = false;
  double cost = 0, cost_change = 0, gradient_max_norm = 0, gradient_norm = 0, step_norm = 0;
  double relative_decrease = 0, trust_region_radius = 0, eta = 0, step_size = 0;
  int line_search_function_evaluations = 0, line_search_gradient_evaluations = 0, line_search_iterations = 0;
  int linear_solver_iterations = 0;
  double iteration_time_in_seconds = 0, step_solver_time_in_seconds = 0, cumulative_time_in_seconds = 0;
};
}  // namespace ceres

#endif
