// Stand-in for <ceres/rotation.h>: Ceres is not in the reference tree nor in this image.  AngleAxisRotatePoint is RESTATED
// here from the published algorithm (ceres-solver 1.x, include/ceres/rotation.h: the theta^2 > epsilon branch with
// sin / cos / theta_inverse, else the first-order Taylor branch) - the one piece of the residual functor harness
// (ref_functor_harness.cc) that is not the reference's own source.  Test infrastructure.
#ifndef REF_SHIM_CERES_ROTATION
#define REF_SHIM_CERES_ROTATION
#include <cmath>
#include <limits>
namespace ceres {
template <typename T>
inline void AngleAxisRotatePoint(const T angle_axis[3], const T pt[3], T result[3]) {
  using std::sqrt; using std::cos; using std::sin;   // doubles; Jets find ceres::sqrt / cos / sin by ADL
  const T theta2 = angle_axis[0] * angle_axis[0] + angle_axis[1] * angle_axis[1] + angle_axis[2] * angle_axis[2];
  if (theta2 > T(std::numeric_limits<double>::epsilon())) {
    const T theta = sqrt(theta2);
    const T costheta = cos(theta);
    const T sintheta = sin(theta);
    const T theta_inverse = 1.0 / theta;
    const T w[3] = {angle_axis[0] * theta_inverse, angle_axis[1] * theta_inverse, angle_axis[2] * theta_inverse};
    const T w_cross_pt[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
    result[0] = pt[0] * costheta + w_cross_pt[0] * sintheta + w[0] * tmp;
    result[1] = pt[1] * costheta + w_cross_pt[1] * sintheta + w[1] * tmp;
    result[2] = pt[2] * costheta + w_cross_pt[2] * sintheta + w[2] * tmp;
  } else {
    const T w_cross_pt[3] = {angle_axis[1] * pt[2] - angle_axis[2] * pt[1], angle_axis[2] * pt[0] - angle_axis[0] * pt[2],
                             angle_axis[0] * pt[1] - angle_axis[1] * pt[0]};
    result[0] = pt[0] + w_cross_pt[0];
    result[1] = pt[1] + w_cross_pt[1];
    result[2] = pt[2] + w_cross_pt[2];
  }
}
}  // namespace ceres
#endif
