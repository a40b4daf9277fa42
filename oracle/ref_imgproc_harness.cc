// Compiles the REFERENCE's own image gradient and disparity conversion — imgradient_row / imgradient_ / imgradient
// (:26-106) and disparityToDepth (:274-322) of /root/reference/src/imgproc.cc, cut out of the file at build time by the Makefile because the rest of that translation unit needs OpenCV)
// against the reference's own src/imgproc.h, src/types.h and src/debug.h — behind a C ABI, so that the restated gradient
// in pba_oracle.cc (and through it the on-the-fly gradients of K_A) can be checked against it bit for bit.
// Built only into oracle/_ref/ (git-ignored).  Test infrastructure; no reference source is copied into this repository.
#include <stdint.h>
#include <cstddef>
#include <cstring>
#include <cassert>
#include <type_traits>
#include <smmintrin.h>
#include "opencv2/core/core.hpp"   // ref_shim: names only
#include "debug.h"                  // -I /root/reference/src (FORCE_INLINE)
#include "imgproc.h"
#include REF_IMGRADIENT_INC
#include REF_DISPARITY_INC          // is_aligned + disparityToDepth(const float*, ...) of src/imgproc.cc:274-322

namespace {
template <class T> struct Plane {   // the Image / Mask concept of IsLocalMax_ (src/imgproc.h:175-212)
  typedef T Scalar;
  const T* p; int r, c;
  int rows() const { return r; }
  int cols() const { return c; }
  T operator()(int y, int x) const { return p[(long)y * c + x]; }
};
}  // namespace

extern "C" {
// IsLocalMax_ for every pixel at least `radius` away from the border (elsewhere 0)
void ref_local_maxima(const float* sal, const uint8_t* mask, int32_t rows, int32_t cols, int32_t radius, uint8_t* out) {
  Plane<float> S{sal, rows, cols};
  Plane<uint8_t> M{mask, rows, cols};
  IsLocalMax_<Plane<float>, Plane<uint8_t> > is_max(S, M, radius);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x)
      out[(long)y * cols + x] = (y >= radius && y < rows - radius && x >= radius && x < cols - radius) ? (is_max(y, x) ? 1 : 0) : 0;
}
void ref_imgradient_u8(const uint8_t* I, int32_t rows, int32_t cols, float* gx, float* gy) {
  imgradient(I, ImageSize(rows, cols), gx, gy);
}
void ref_disparity_to_depth(const float* d, int32_t rows, int32_t cols, float Bf, float* z) {
  disparityToDepth(d, ImageSize(rows, cols), Bf, z);
}
void ref_imgradient_f32(const float* I, int32_t rows, int32_t cols, float* gx, float* gy) {
  imgradient(I, ImageSize(rows, cols), gx, gy);
}
}
