// declaration-only stand-in: the cv::Mat overload of disparityToDepth (src/imgproc.h:38); the raw-pointer overload
// (src/imgproc.h:36) is the host side's own (../pose_utils.h)
#ifndef PBA_PROOF_IMGPROC_H
#define PBA_PROOF_IMGPROC_H
#include <opencv2/core/core.hpp>
#include "pose_utils.h"
void disparityToDepth(const cv::Mat& disparity, double Bf, cv::Mat_<float>& depth);
#endif
