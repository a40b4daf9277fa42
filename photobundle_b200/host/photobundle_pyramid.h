// photobundle_pyramid.h — the reference's pyramid front (src/photobundle_pyramid.h:16-47), same constructor and
// addFrame signature.  The reference class is an unfinished sketch; the semantics implemented here
// are stated in photobundle.cc: one point set, every level solves the same window coarse to fine on the device.
#ifndef PHOTOBUNDLE_PHOTOBUNDLE_PYR_H
#define PHOTOBUNDLE_PHOTOBUNDLE_PYR_H

#include "photobundle.h"

class PhotometricBundleAdjustmentPyr {
 public:
  typedef PhotometricBundleAdjustment::Options Options;
  typedef PhotometricBundleAdjustment::Result Result;

  // num_levels: levels of the pyramid; calib / imageSize: at the finest level
  PhotometricBundleAdjustmentPyr(int num_levels, const Calibration& calib, const ImageSize&, const Options& = Options());
  ~PhotometricBundleAdjustmentPyr();

  // image / depth_map at the finest level, T: pose initialisation (frame to frame)
  void addFrame(const uint8_t* image, const float* depth_map, const Mat44& T, Result* = nullptr);

  const PhotometricBundleAdjustment& finest() const { return *_ba; }

 private:
  UniquePointer<PhotometricBundleAdjustment> _ba;
};

#endif
