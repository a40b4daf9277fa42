// declaration-only stand-in (see ../../README.md): the members of cv::Mat / cv::Mat_ that apps/run_kitti.cc uses
#ifndef PBA_PROOF_CV_CORE
#define PBA_PROOF_CV_CORE
#include <cstdint>
namespace cv {
class Mat {
 public:
  template <class T> T* ptr(int row = 0);
  template <class T> const T* ptr(int row = 0) const;
  bool empty() const;
  int rows = 0, cols = 0;
};
template <class T> class Mat_ : public Mat {
 public:
  Mat_();
};
}  // namespace cv
#endif
