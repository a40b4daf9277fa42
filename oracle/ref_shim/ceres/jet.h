// Minimal stand-in for <ceres/jet.h>: the data layout of ceres::Jet<T,N> (scalar part
// `a`, derivative part `v`) — all the reference's src/jet_extras.h needs.
#ifndef REF_SHIM_CERES_JET
#define REF_SHIM_CERES_JET
#include "Eigen/Core"
namespace ceres {
template <class T, int N>
struct Jet {
  T a;
  Eigen::Matrix<T, N, 1> v;
  Jet() : a() {}
};
}  // namespace ceres
#endif
