// declaration-only stand-in for the reference's dataset interface (src/dataset.h:12-70)
#ifndef PBA_PROOF_DATASET_H
#define PBA_PROOF_DATASET_H
#include <string>
#include <opencv2/core/core.hpp>
#include "compat.h"
struct DatasetFrame {
  virtual const cv::Mat& image() const = 0;
  virtual const cv::Mat& disparity() const = 0;
  virtual std::string filename() const = 0;
  virtual ~DatasetFrame() {}
};
class Dataset {
 public:
  virtual ~Dataset() {}
  virtual UniquePointer<DatasetFrame> getFrame(int f_i) const = 0;
  virtual ImageSize imageSize() const = 0;
  virtual Calibration calibration() const = 0;
  virtual std::string name() const = 0;
  static UniquePointer<Dataset> Create(std::string conf_fn);
};
#endif
