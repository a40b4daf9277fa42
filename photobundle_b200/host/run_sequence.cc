// run_sequence.cc — Boost/OpenCV-free counterpart of the reference's apps/run_kitti.cc:17-58 for
// CI: same loop (addFrame per frame with the initial frame-to-frame poses, write the refined
// trajectory in KITTI format), reading a raw sequence file instead of KITTI PNGs + stereo:
//   header: int32 rows, cols, n_frames; double fx, fy, cx, cy, baseline
//   per frame: rows*cols uint8 image, rows*cols float32 depth
// usage: run_sequence <sequence.bin> <init_poses_kitti.txt> <config.cfg|-> <output_poses.txt>
#include <cstdio>
#include <fstream>
#include <vector>

#include "config.h"
#include "photobundle.h"
#include "pose_utils.h"

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: %s sequence.bin init_poses.txt config.cfg|- output.txt\n", argv[0]); return 2; }
  std::ifstream ifs(argv[1], std::ios::binary);
  if (!ifs.is_open()) { fprintf(stderr, "cannot open %s\n", argv[1]); return 1; }
  int32_t rows, cols, n;
  double k[5];
  ifs.read((char*)&rows, 4); ifs.read((char*)&cols, 4); ifs.read((char*)&n, 4); ifs.read((char*)k, sizeof(k));
  Mat33 K = Mat33::Identity();
  K(0, 0) = k[0]; K(1, 1) = k[1]; K(0, 2) = k[2]; K(1, 2) = k[3];
  const PoseList T_init = loadPosesKittiFormat(argv[2]);
  PhotometricBundleAdjustment::Options opt;
  if (std::string(argv[3]) != "-") opt = PhotometricBundleAdjustment::Options(utils::ConfigFile(argv[3]));
  PhotometricBundleAdjustment::Result result;
  PhotometricBundleAdjustment photoba(Calibration(K, k[4]), ImageSize(rows, cols), opt);
  std::vector<uint8_t> I((size_t)rows * cols);
  std::vector<float> Z((size_t)rows * cols);
  for (int f_i = 0; f_i < n && f_i < (int)T_init.size(); ++f_i) {
    ifs.read((char*)I.data(), I.size());
    ifs.read((char*)Z.data(), Z.size() * sizeof(float));
    printf("Frame %05d\n", f_i);
    photoba.addFrame(I.data(), Z.data(), T_init[f_i], &result);
  }
  printf("Writing refined poses to %s\n", argv[4]);
  return writePosesKittiFormat(argv[4], result.poses) ? 0 : 1;
}
