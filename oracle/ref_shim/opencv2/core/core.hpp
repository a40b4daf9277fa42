// Names only: the reference's src/imgproc.h mentions cv::Mat / cv::Mat_ in one declaration (OpenCV itself is not
// installed in this image).  Test infrastructure.
#ifndef REF_SHIM_CV_CORE
#define REF_SHIM_CV_CORE
namespace cv {
class Mat;
template <class T> class Mat_;
}  // namespace cv
#endif
