// Stand-in for <ceres/jet.h> (Ceres is absent): the data layout of ceres::Jet<T,N> (scalar part `a`, derivative part `v`)
// - all the reference's src/jet_extras.h needs - plus the dual-number arithmetic the residual functor uses when it is
// instantiated with Jets (ref_functor_harness.cc), RESTATED from the published definitions of ceres-solver 1.x
// include/ceres/jet.h (note the quotient: value f.a * (1 / g.a), derivative (f.v - f.a/g.a * g.v) * (1 / g.a)).
// Test infrastructure.
#ifndef REF_SHIM_CERES_JET
#define REF_SHIM_CERES_JET
#include <cmath>
#include "Eigen/Core"
namespace ceres {
template <class T, int N>
struct Jet {
  T a;
  Eigen::Matrix<T, N, 1> v;
  Jet() : a() {}
  explicit Jet(const T& value) : a(value) {}
  Jet(const T& value, int k) : a(value) { v[k] = T(1.0); }
};
#define PBA_JET template <class T, int N> inline
PBA_JET Jet<T, N> operator-(const Jet<T, N>& f) { Jet<T, N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
PBA_JET Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
PBA_JET Jet<T, N> operator+(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a = f.a + s; return h; }
PBA_JET Jet<T, N> operator+(T s, const Jet<T, N>& f) { Jet<T, N> h = f; h.a = f.a + s; return h; }
PBA_JET Jet<T, N>& operator+=(Jet<T, N>& f, const Jet<T, N>& g) { f = f + g; return f; }
PBA_JET Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
PBA_JET Jet<T, N> operator-(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a = f.a - s; return h; }
PBA_JET Jet<T, N> operator-(T s, const Jet<T, N>& f) { Jet<T, N> h; h.a = s - f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
PBA_JET Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
PBA_JET Jet<T, N> operator*(const Jet<T, N>& f, T s) { Jet<T, N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
PBA_JET Jet<T, N> operator*(T s, const Jet<T, N>& f) { return f * s; }
PBA_JET Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h;
  const T g_a_inverse = T(1.0) / g.a;
  const T f_a_by_g_a = f.a * g_a_inverse;
  h.a = f_a_by_g_a;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse;
  return h;
}
PBA_JET Jet<T, N> operator/(T s, const Jet<T, N>& g) {
  Jet<T, N> h;
  const T minus_s_g_a_inverse2 = -s / (g.a * g.a);
  h.a = s / g.a;
  for (int i = 0; i < N; ++i) h.v[i] = g.v[i] * minus_s_g_a_inverse2;
  return h;
}
PBA_JET bool operator>(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a > g.a; }
PBA_JET Jet<T, N> sqrt(const Jet<T, N>& f) {
  Jet<T, N> h; const T t = std::sqrt(f.a); const T two_a_inverse = T(1.0) / (T(2.0) * t);
  h.a = t; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * two_a_inverse;
  return h;
}
PBA_JET Jet<T, N> cos(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::cos(f.a); const T ms = -std::sin(f.a); for (int i = 0; i < N; ++i) h.v[i] = ms * f.v[i]; return h; }
PBA_JET Jet<T, N> sin(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::sin(f.a); const T c = std::cos(f.a); for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; return h; }
#undef PBA_JET
}  // namespace ceres
#endif
