// Compiles the REFERENCE's own camera model — /root/reference/src/calibration.h (with src/types.h and src/eigen.h),
// included from where it lies — behind a C ABI, so that the restated projection in pba_oracle.cc and the host's
// Calibration / ImageSize (photobundle_b200/host/compat.h) can be checked against it bit for bit.  The second half
// wraps MakePatchWeights (src/photobundle.cc:617-646, a file-static function): the recipe in the Makefile pipes that
// function's text from the reference file into this translation unit at build time (REF_PATCH_WEIGHTS_INC), it is
// never written into this repository.  Built only into oracle/_ref/ (git-ignored).  Test infrastructure.
#include <stdint.h>
#include <cmath>
#include <vector>
#include "calibration.h"   // -I /root/reference/src

namespace {
Calibration make(const double* k4, double b) {
  Mat33 K;
  K << k4[0], 0.0, k4[2], 0.0, k4[1], k4[3], 0.0, 0.0, 1.0;
  return Calibration(K, b);
}
}  // namespace

#ifdef REF_PATCH_WEIGHTS_INC
#include REF_PATCH_WEIGHTS_INC
#endif
// ExtractPatch (src/photobundle.cc:466-479): the reference descriptor of a new point, clamped at the image border
#ifdef REF_EXTRACT_PATCH_INC
#include <algorithm>
#include REF_EXTRACT_PATCH_INC
namespace {
struct ImgF32 {   // a channel as DescriptorFrame keeps it (Image_<float>)
  const float* p; int r, c;
  int rows() const { return r; }
  int cols() const { return c; }
  float operator()(int y, int x) const { return p[(long)y * c + x]; }
};
}  // namespace
#endif
// interp2 + interpolateFixedPatch (src/photobundle.cc:258-310, templates of that file cut out the same way): the bilinear
// lookup of the data association (ZnccPatch_::set) - ZnccPatch_ itself is left out because its sums go through Eigen's
// own reductions, which no shim reproduces.
#ifdef REF_INTERP2_INC
#include REF_INTERP2_INC
namespace {
struct ImgU8 {   // the Image concept interp2 needs; operator() returns uint8_t like Image_<uint8_t>
  const uint8_t* p; int r, c;
  int rows() const { return r; }
  int cols() const { return c; }
  uint8_t operator()(int y, int x) const { return p[(long)y * c + x]; }
};
}  // namespace
#endif

extern "C" {

// the template the residual functor instantiates (src/calibration.h:33-38, called at src/photobundle.cc:706)
void ref_project(const double* k4, const double* X, double* uv) { make(k4, 0.1).project(X, uv[0], uv[1]); }

// Calibration::project(Vec3) = normHomog(K * X) (src/calibration.h:43), used by addFrame (src/photobundle.cc:520)
void ref_project_vec3(const double* k4, const double* X, double* uv) {
  const Vec2 p = make(k4, 0.1).project(Vec3(X[0], X[1], X[2]));
  uv[0] = p[0]; uv[1] = p[1];
}

// one pyramid level down (src/calibration.h:72-78, src/types.h:70-73): out = {fx, fy, cx, cy, baseline, rows, cols}
void ref_pyr_down(const double* k4, double b, int32_t rows, int32_t cols, double* out7) {
  const Calibration c = make(k4, b).pyrDown();
  const ImageSize s = ImageSize(rows, cols).pyrDown();
  out7[0] = c.fx(); out7[1] = c.fy(); out7[2] = c.cx(); out7[3] = c.cy(); out7[4] = c.b();
  out7[5] = s.rows; out7[6] = s.cols;
}

void ref_triangulate(const double* k4, double b, const double* uvd, double* xyz) {
  const Vec_<double, 3> p = make(k4, b).triangulate(uvd);
  xyz[0] = p[0]; xyz[1] = p[1]; xyz[2] = p[2];
}

#ifdef REF_INTERP2_INC
float ref_interp2_u8(const uint8_t* I, int32_t rows, int32_t cols, float x, float y) {
  ImgU8 im{I, rows, cols};
  return interp2(im, x, y);
}
// the 5x5 patch ZnccPatch_<2, float>::set reads (before its mean is removed), uv in double as addFrame passes it
void ref_interp_patch5_u8(const uint8_t* I, int32_t rows, int32_t cols, double u, double v, float* out25) {
  ImgU8 im{I, rows, cols};
  Vec_<float, 25> dst;
  const double uv[2] = {u, v};
  interpolateFixedPatch<2>(dst, im, uv, 0.0f, 0.0f);
  for (int k = 0; k < 25; ++k) out25[k] = dst[k];
}
#endif

#ifdef REF_EXTRACT_PATCH_INC
void ref_extract_patch_f32(const float* I, int32_t rows, int32_t cols, int32_t x, int32_t y, int32_t radius, double* dst) {
  ImgF32 im{I, rows, cols};
  Vec_<int, 2> uv;
  uv[0] = x; uv[1] = y;
  ExtractPatch(dst, im, uv, radius);
}
#endif

#ifdef REF_PATCH_WEIGHTS_INC
int32_t ref_patch_weights(int32_t radius, int32_t do_gaussian, double* w) {
  const std::vector<double> v = MakePatchWeights(radius, do_gaussian != 0);
  for (size_t i = 0; i < v.size(); ++i) w[i] = v[i];
  return (int32_t)v.size();
}
#endif

}
