// pose_utils.h — KITTI pose text I/O with the reference's formats (src/pose_utils.{h,cc}).
#ifndef PBA_HOST_POSE_UTILS_H
#define PBA_HOST_POSE_UTILS_H
#include "compat.h"
PoseList loadPosesKittiFormat(std::string filename);               // src/pose_utils.cc:9-40
bool writePosesKittiFormat(std::string filename, const PoseList&); // src/pose_utils.cc:43-59
PoseList convertPoseToLocal(const PoseList&);                      // src/pose_utils.cc:62-74
// src/imgproc.cc:280-322 semantics with an exact reciprocal (the reference's vector body uses the
// ~12-bit _mm_rcp_ps): z = Bf/d for d > 0.01, else the invalid mark -0.1.
void disparityToDepth(const float* dmap, const ImageSize&, float Bf, float* zmap);
#endif
