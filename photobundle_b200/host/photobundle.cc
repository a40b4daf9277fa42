// photobundle.cc — host side of the drop-in.  addFrame() keeps the reference's observable behaviour
// (src/photobundle.cc:482-615: which points are re-observed, which pixels become new points, their descriptors)
// with its own structure; optimize() (src/photobundle.cc:764-876) packs the window, keeps the frames of the ring
// buffer resident on the device (one upload per new frame), runs the pyramid levels coarse to fine through the C
// ABI (include/pba_b200.h) with the level hand-over on the device, writes poses/points back, evicts, fills Result.

#include "photobundle.h"

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>

#include <thread>

#include "../../include/pba_b200.h"
#include "config.h"
#include "photobundle_pyramid.h"
#include "pose_utils.h"

// ---------------------------------------------------------------------------- compat math
Mat33 Mat33::inverse() const {
  const Mat33& a = *this;
  Mat33 r;
  const double c00 = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1);
  const double c10 = a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2);
  const double c20 = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0);
  const double det = a(0, 0) * c00 + a(0, 1) * c10 + a(0, 2) * c20;
  const double id = 1.0 / det;
  r(0, 0) = c00 * id; r(0, 1) = (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) * id; r(0, 2) = (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) * id;
  r(1, 0) = c10 * id; r(1, 1) = (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) * id; r(1, 2) = (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) * id;
  r(2, 0) = c20 * id; r(2, 1) = (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) * id; r(2, 2) = (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) * id;
  return r;
}

Mat44 Mat44::operator*(const Mat44& o) const {
  Mat44 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += (*this)(i, k) * o(k, j);
      r(i, j) = s;
    }
  return r;
}

Mat44 Mat44::inverse() const {   // Gauss-Jordan with partial pivoting
  double a[4][8];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) { a[i][j] = (*this)(i, j); a[i][4 + j] = (i == j) ? 1.0 : 0.0; }
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    for (int r = c + 1; r < 4; ++r) if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
    if (piv != c) for (int j = 0; j < 8; ++j) std::swap(a[c][j], a[piv][j]);
    const double d = 1.0 / a[c][c];
    for (int j = 0; j < 8; ++j) a[c][j] *= d;
    for (int r = 0; r < 4; ++r) {
      if (r == c) continue;
      const double f = a[r][c];
      if (f != 0.0) for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j];
    }
  }
  Mat44 r;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r(i, j) = a[i][4 + j];
  return r;
}

Mat44 Mat44::rigidInverse() const {
  Mat44 r = Mat44::Identity();
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r(i, j) = (*this)(j, i);
  for (int i = 0; i < 3; ++i) r(i, 3) = -(r(i, 0) * (*this)(0, 3) + r(i, 1) * (*this)(1, 3) + r(i, 2) * (*this)(2, 3));
  return r;
}

// ---------------------------------------------------------------------------- Trajectory
// World poses keyed by frame id.  push_back() chains a frame-to-frame estimate: T_w,i = T_w,i-1 * T_i^-1
// (behaviour of src/trajectory.cc:7-16); lookups by id throw when the id is unknown (:18-39).
int Trajectory::find(const Id_t id) const {
  // ids arrive in increasing order: the newest are the ones asked for, search from the back
  for (size_t k = _data.size(); k-- > 0;)
    if (_data[k].id == id) return (int)k;
  return -1;
}
void Trajectory::push_back(const Mat44& pose, const Id_t id) {
  if (find(id) >= 0) throw std::runtime_error("duplicate id in trajectory\n");
  PoseWithId e;
  e.id = id;
  e.pose = _data.empty() ? pose.inverse() : back() * pose.inverse();
  _data.push_back(e);
}
const Mat44& Trajectory::atId(const Id_t id) const {
  const int k = find(id);
  if (k < 0) throw std::runtime_error("could not find pose with id");
  return _data[(size_t)k].pose;
}
Mat44& Trajectory::atId(const Id_t id) { return const_cast<Mat44&>(static_cast<const Trajectory*>(this)->atId(id)); }
EigenAlignedContainer_<Mat44> Trajectory::poses() const {
  EigenAlignedContainer_<Mat44> out;
  out.reserve(_data.size());
  for (const PoseWithId& e : _data) out.push_back(e.pose);
  return out;
}
EigenAlignedContainer_<Vec3> Trajectory::cameraPositions() const {
  EigenAlignedContainer_<Vec3> out;
  out.reserve(_data.size());
  for (const PoseWithId& e : _data) out.emplace_back(e.pose(0, 3), e.pose(1, 3), e.pose(2, 3));
  return out;
}

// ---------------------------------------------------------------------------- pose text I/O
// KITTI odometry format: one pose per line, the 12 entries of the top 3x4 of the 4x4 matrix, row by row
// (what src/pose_utils.cc:9-59 reads and writes; the writer's layout — every number followed by one blank — is kept
// so that files compare byte for byte with the reference's output).
PoseList loadPosesKittiFormat(std::string filename) {
  std::ifstream in(filename);
  if (!in.is_open()) throw std::runtime_error("failed to open pose file");
  PoseList poses;
  for (std::string line; std::getline(in, line);) {
    if (line.find_first_not_of(" \t\r") == std::string::npos) continue;
    std::istringstream fields(line);
    Mat44 T = Mat44::Identity();
    for (int k = 0; k < 12; ++k) fields >> T(k / 4, k % 4);
    poses.push_back(T);
  }
  return poses;
}
bool writePosesKittiFormat(std::string filename, const PoseList& poses) {
  std::ofstream out(filename);
  if (!out.is_open()) return false;
  for (const Mat44& T : poses) {
    for (int k = 0; k < 12; ++k) out << T(k / 4, k % 4) << " ";
    out << "\n";
  }
  return true;
}
// world poses -> frame-to-frame poses, the inverse of Trajectory::push_back's chaining (src/pose_utils.cc:62-74)
PoseList convertPoseToLocal(const PoseList& world) {
  if (world.empty()) throw std::runtime_error("no poses");
  PoseList local;
  local.reserve(world.size());
  for (size_t k = 0; k < world.size(); ++k)
    local.push_back(k == 0 ? world[0].rigidInverse() : world[k].rigidInverse() * world[k - 1]);
  return local;
}
// depth = (baseline * focal) / disparity where the disparity is usable (> 0.01), the invalid mark -0.1 elsewhere
// (src/imgproc.cc:280-322; the reference's vector body uses the ~12-bit _mm_rcp_ps, here the reciprocal is exact)
void disparityToDepth(const float* dmap, const ImageSize& sz, float Bf, float* zmap) {
  for (int k = 0, n = sz.numel(); k < n; ++k) {
    const float d = dmap[k];
    zmap[k] = d > 0.01f ? Bf * (1.0f / d) : -0.10f;
  }
}

// ---------------------------------------------------------------------------- pose <-> params
// ceres::RotationMatrixToAngleAxis / AngleAxisToRotationMatrix as used by PoseToParams /
// ParamsToPose (src/photobundle.cc:646-667).
static void PoseToParams(const Mat44& T, double* p) {
  auto R = [&](int i, int j) { return T(i, j); };
  double q[4];
  const double trace = R(0, 0) + R(1, 1) + R(2, 2);
  if (trace >= 0.0) {
    double t = std::sqrt(trace + 1.0);
    q[0] = 0.5 * t; t = 0.5 / t;
    q[1] = (R(2, 1) - R(1, 2)) * t; q[2] = (R(0, 2) - R(2, 0)) * t; q[3] = (R(1, 0) - R(0, 1)) * t;
  } else {
    int i = 0;
    if (R(1, 1) > R(0, 0)) i = 1;
    if (R(2, 2) > R(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0);
    q[i + 1] = 0.5 * t; t = 0.5 / t;
    q[0] = (R(k, j) - R(j, k)) * t; q[j + 1] = (R(j, i) + R(i, j)) * t; q[k + 1] = (R(k, i) + R(i, k)) * t;
  }
  const double s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (s2 > 0.0) {
    const double s = std::sqrt(s2), c = q[0];
    const double two_theta = 2.0 * ((c < 0.0) ? std::atan2(-s, -c) : std::atan2(s, c));
    const double k = two_theta / s;
    p[0] = q[1] * k; p[1] = q[2] * k; p[2] = q[3] * k;
  } else {
    p[0] = q[1] * 2.0; p[1] = q[2] * 2.0; p[2] = q[3] * 2.0;
  }
  p[3] = T(0, 3); p[4] = T(1, 3); p[5] = T(2, 3);
}
static Mat44 ParamsToPose(const double* p) {
  Mat44 T = Mat44::Identity();
  const double theta2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
  if (theta2 > std::numeric_limits<double>::epsilon()) {
    const double theta = std::sqrt(theta2);
    const double wx = p[0] / theta, wy = p[1] / theta, wz = p[2] / theta;
    const double c = std::cos(theta), s = std::sin(theta);
    T(0, 0) = c + wx * wx * (1.0 - c); T(1, 0) = wz * s + wx * wy * (1.0 - c); T(2, 0) = -wy * s + wx * wz * (1.0 - c);
    T(0, 1) = wx * wy * (1.0 - c) - wz * s; T(1, 1) = c + wy * wy * (1.0 - c); T(2, 1) = wx * s + wy * wz * (1.0 - c);
    T(0, 2) = wy * s + wx * wz * (1.0 - c); T(1, 2) = -wx * s + wy * wz * (1.0 - c); T(2, 2) = c + wz * wz * (1.0 - c);
  } else {
    T(0, 0) = 1.0; T(1, 0) = p[2]; T(2, 0) = -p[1];
    T(0, 1) = -p[2]; T(1, 1) = 1.0; T(2, 1) = p[0];
    T(0, 2) = p[1]; T(1, 2) = -p[0]; T(2, 2) = 1.0;
  }
  T(0, 3) = p[3]; T(1, 3) = p[4]; T(2, 3) = p[5];
  return T;
}

// Weights of the patch pixels (behaviour of src/photobundle.cc:617-644): uniform, or an isotropic unit-variance
// Gaussian normalised to sum 1, row-major over the (2r+1)^2 patch.
std::vector<double> MakePatchWeights(int radius, bool do_gaussian) {
  const int side = 2 * radius + 1;
  std::vector<double> w((size_t)side * side, 1.0);
  if (!do_gaussian) return w;
  double total = 0.0;
  for (int k = 0; k < side * side; ++k) {
    const int dy = k / side - radius, dx = k % side - radius;
    w[(size_t)k] = 1.0 * std::exp(-0.5 * ((dy * dy) / 1.0 + (dx * dx) / 1.0));
    total += w[(size_t)k];
  }
  for (double& v : w) v /= total;
  return w;
}

// ---------------------------------------------------------------------------- image helpers
// Bilinear look-up in a uint8 image at (xf, yf) with the reference's border policy and mixed float/double arithmetic
// (src/photobundle.cc:262-294 with T = float): inside -> 4 taps; exactly on the last column / row -> the 1-D
// interpolation along the border; anywhere else -> fillval.  The device front end (pba_associate) reproduces this
// bit for bit, so the expressions keep their promotions.
static inline float interp2_u8(const uint8_t* I, int rows, int cols, float xf, float yf, float fillval = 0.0f) {
  const int last_col = cols - 1, last_row = rows - 1;
  const int xi = (int)std::floor(xf), yi = (int)std::floor(yf);
  xf -= xi; yf -= yi;
  const uint8_t* p = I + (size_t)yi * cols + xi;
  const bool x_in = xi >= 0 && xi < last_col, y_in = yi >= 0 && yi < last_row;
  if (x_in && y_in) {
    const float wx = 1.0 - xf;
    return (1.0 - yf) * ((int)p[0] * wx + (int)p[1] * xf) + yf * ((int)p[cols] * wx + (int)p[cols + 1] * xf);
  }
  if (xi == last_col && y_in) return (xf > 0) ? fillval : (float)((1.0 - yf) * (int)p[0] + yf * (int)p[cols]);
  if (yi == last_row && x_in) return (yf > 0) ? fillval : (float)((1.0 - xf) * (int)p[0] + xf * (int)p[1]);
  if (xi == last_col && yi == last_row) return (xf > 0 || yf > 0) ? fillval : (float)(int)p[0];
  return fillval;
}

// The raw 5x5 lookup ZnccPatch::set starts from, for the test against the reference's own interp2 / interpolateFixedPatch.
void InterpPatch5(const uint8_t* I, int rows, int cols, double px, double py, float* out25) {
  const float x = (float)px, y = (float)py;
  for (int k = 0; k < 25; ++k) out25[k] = interp2_u8(I, rows, cols, (k % 5 - 2) + x, (k / 5 - 2) + y);
}

// Mean-free 5x5 patch and its norm for the zero-normalised cross correlation of the data association
// (ZnccPatch_<2, float>, src/photobundle.cc:315-361).
struct ZnccPatch {
  static constexpr int kSide = 5, kLen = kSide * kSide;
  float data[kLen];
  float norm = 0.f;
  void set(const uint8_t* I, int rows, int cols, double px, double py) {
    const float x = (float)px, y = (float)py;
    for (int k = 0; k < kLen; ++k) data[k] = interp2_u8(I, rows, cols, (k % kSide - 2) + x, (k / kSide - 2) + y);
    float sum = 0.f;
    for (float v : data) sum += v;
    const float mean = sum / (float)kLen;
    float ss = 0.f;
    for (float& v : data) { v -= mean; ss += v * v; }
    norm = std::sqrt(ss);
  }
  float score(const ZnccPatch& o) const {
    const float d = norm * o.norm;
    float dot = 0.f;
    for (int k = 0; k < kLen; ++k) dot += data[k] * o.data[k];
    return d > 1e-6 ? dot / d : -1.0f;
  }
};

// cv::pyrDown for CV_8U on the host (the device has its own, k_pyrdown_u8): separable [1 4 6 4 1] with
// BORDER_REFLECT_101, result (sum + 128) >> 8.  Only used to read the reference descriptors of new points at
// the coarser pyramid levels.
static int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}
static std::vector<uint8_t> PyrDownU8(const std::vector<uint8_t>& src, int rows, int cols) {
  const int drows = (rows + 1) / 2, dcols = (cols + 1) / 2;
  static const int w[5] = {1, 4, 6, 4, 1};
  std::vector<int> h((size_t)rows * dcols);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < dcols; ++x) {
      int s = 0;
      for (int k = 0; k < 5; ++k) s += w[k] * src[(size_t)y * cols + reflect101(2 * x + k - 2, cols)];
      h[(size_t)y * dcols + x] = s;
    }
  std::vector<uint8_t> dst((size_t)drows * dcols);
  for (int y = 0; y < drows; ++y)
    for (int x = 0; x < dcols; ++x) {
      int s = 0;
      for (int k = 0; k < 5; ++k) s += w[k] * h[(size_t)reflect101(2 * y + k - 2, rows) * dcols + x];
      dst[(size_t)y * dcols + x] = (uint8_t)((s + 128) >> 8);
    }
  return dst;
}
// Reference descriptor of a point at a coarser level: the bilinear patch of the reduced reference frame at the
// point's level-0 pixel divided by 2^level (workloads/synthetic.py `pyramid_level` — the level semantics are ours,
// SURVEY App. C #12); channel values are floats widened to double like every descriptor.
static void BilinearPatch(double* dst, const std::vector<uint8_t>& img, int rows, int cols, double x, double y, int radius) {
  int k = 0;
  for (int dy = -radius; dy <= radius; ++dy)
    for (int dx = -radius; dx <= radius; ++dx, ++k) {
      const double xx = std::min(std::max(x + dx, 0.0), cols - 1.0), yy = std::min(std::max(y + dy, 0.0), rows - 1.0);
      const int x0 = std::min((int)xx, cols - 2), y0 = std::min((int)yy, rows - 2);
      const double ax = xx - x0, ay = yy - y0;
      const double v = (1 - ay) * ((1 - ax) * img[(size_t)y0 * cols + x0] + ax * img[(size_t)y0 * cols + x0 + 1]) +
                       ay * ((1 - ax) * img[(size_t)(y0 + 1) * cols + x0] + ax * img[(size_t)(y0 + 1) * cols + x0 + 1]);
      dst[k] = (double)(float)v;
    }
}

struct PhotometricBundleAdjustment::ScenePoint {
  Vec3 X, X_original;
  std::vector<uint32_t> f;      // visibility list, first = reference frame
  ZnccPatch patch;
  std::vector<double> descriptor;                 // finest level
  std::vector<std::vector<double>> coarse_desc;   // levels 1 .. numPyramidLevels-1
  double saliency = 0.0;
  bool was_refined = false;
  int x = 0, y = 0;             // pixel in the reference frame
  ScenePoint(const Vec3& X_, uint32_t f_id) : X(X_), X_original(X_) { f.reserve(8); f.push_back(f_id); }
  uint32_t refFrameId() const { return f.front(); }
  uint32_t lastFrameId() const { return f.back(); }
  size_t numFrames() const { return f.size(); }
};

PhotometricBundleAdjustment::PointView PhotometricBundleAdjustment::scenePoint(size_t i) const {
  const ScenePoint& p = *_scene_points[i];
  return PointView{p.X.data(), &p.f, &p.descriptor, p.x, p.y, p.saliency};
}

// ---------------------------------------------------------------------------- Options
static PhotometricBundleAdjustment::Options::DescriptorType DescriptorTypeFromString(std::string s) {
  std::string l = s;
  std::transform(l.begin(), l.end(), l.begin(), [](unsigned char c) { return std::tolower(c); });
  typedef PhotometricBundleAdjustment::Options::DescriptorType DT;
  if (l == "intensity") return DT::Intensity;
  if (l == "intensityandgradient") return DT::IntensityAndGradient;
  if (l == "bitplanes") return DT::BitPlanes;
  fprintf(stderr, "Unknown descriptorType '%s'\n", s.c_str());
  return DT::Intensity;
}

PhotometricBundleAdjustment::Options::Options(const utils::ConfigFile& cf)   // keys and defaults of src/photobundle.cc:88-103
    : maxNumPoints(cf.get<int>("maxNumPoints", 4096)),
      slidingWindowSize(cf.get<int>("slidingWindowSize", 5)),
      patchRadius(cf.get<int>("patchRadius", 2)),
      maskBlockRadius(cf.get<int>("maskBlockRadius", 1)),
      maxFrameDistance(cf.get<int>("maxFrameDistance", 1)),
      numThreads(cf.get<int>("numThreads", -1)),
      doGaussianWeighting((bool)cf.get<int>("doGaussianWeighting", 0)),
      verbose((bool)cf.get<int>("verbose", 1)),
      minScore(cf.get<double>("minScore", 0.75)),
      robustThreshold(cf.get<double>("robustThreshold", 0.05)),
      minValidDepth(cf.get<double>("minValidDepth", 0.01)),
      maxValidDepth(cf.get<double>("maxValidDepth", 1000.0)),
      nonMaxSuppRadius(cf.get<int>("nonMaxSuppRadius", 1)),
      descriptorType(DescriptorTypeFromString(cf.get<std::string>("descriptorType", "Intensity"))),
      device(cf.get<int>("device", -1)),
      gpuFrontEnd((bool)cf.get<int>("gpuFrontEnd", 0)),
      numPyramidLevels(cf.get<int>("numPyramidLevels", 1)),
      nGpus(cf.get<int>("nGpus", 1)) {}

// ---------------------------------------------------------------------------- ctor / dtor
PhotometricBundleAdjustment::PhotometricBundleAdjustment(const Calibration& calib, const ImageSize& image_size,
                                                         const Options& options)
    : _calib(calib), _image_size(image_size), _options(options) {
  if (_options.slidingWindowSize < 1 || _options.slidingWindowSize > PBA_MAX_FRAMES)
    throw std::runtime_error("slidingWindowSize outside [1, " + std::to_string(PBA_MAX_FRAMES) + "]");
  // DescriptorFrame::Create (src/photobundle.cc:220-248): the channel planes are built on the device
  _desc_type = _options.descriptorType == Options::DescriptorType::Intensity ? PBA_DESC_INTENSITY
             : _options.descriptorType == Options::DescriptorType::IntensityAndGradient ? PBA_DESC_INTENSITY_AND_GRADIENT
                                                                                         : PBA_DESC_BITPLANES;
  _n_channels = pba_descriptor_channels(_desc_type);
  if (_options.numPyramidLevels < 1 || _options.numPyramidLevels > 6) throw std::runtime_error("numPyramidLevels outside [1, 6]");
  if (_options.nGpus < 1 || _options.nGpus > 8) throw std::runtime_error("nGpus outside [1, 8]");
  if (_options.numPyramidLevels > 1 && _desc_type != PBA_DESC_INTENSITY)
    throw std::runtime_error("pyramid levels are defined for the Intensity descriptor");
  _mask.resize((size_t)_image_size.rows * _image_size.cols);
  _saliency_map.resize((size_t)_image_size.rows * _image_size.cols);
  _K_inv = calib.K().inverse();
  // the levels: image size by (n + 1) / 2 (src/types.h:70-73), intrinsics by Calibration::pyrDown (src/calibration.h:72-78)
  _dev.resize((size_t)_options.numPyramidLevels);
  Calibration c = calib;
  ImageSize s = image_size;
  for (int l = 0; l < _options.numPyramidLevels; ++l) {
    _dev[(size_t)l].calib = c; _dev[(size_t)l].size = s; _dev[(size_t)l].levels_down = l;
    _dev[(size_t)l].resident.assign((size_t)_options.slidingWindowSize, -1);
    c = c.pyrDown(); s = s.pyrDown();
  }
}

PhotometricBundleAdjustment::~PhotometricBundleAdjustment() {
  for (DeviceLevel& d : _dev)
    for (pba_handle* h : d.ranks) pba_destroy(h);
}

void* PhotometricBundleAdjustment::Pinned::need(size_t n) {
  if (n <= bytes) return p;
  pba_host_free(p);
  bytes = n + n / 2 + 4096;
  p = pba_host_alloc(bytes);
  if (!p) { bytes = 0; throw std::runtime_error("pba_host_alloc failed"); }
  return p;
}
PhotometricBundleAdjustment::Pinned::~Pinned() { pba_host_free(p); }

static void check_pba(int rc, const char* what) {
  if (rc != PBA_OK) throw std::runtime_error(std::string(what) + ": " + pba_last_error());
}

// The device handle of a level, with room for n_points / n_obs (recreated with more room when a window outgrows it).
pba_handle* PhotometricBundleAdjustment::deviceLevel(int level, int n_points, int n_obs) {
  DeviceLevel& d = _dev[(size_t)level];
  if (d.h && n_points <= d.cap_points && n_obs <= d.cap_obs) return d.h;
  for (pba_handle* h : d.ranks) pba_destroy(h);
  d.ranks.clear(); d.h = nullptr;
  pba_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.rows = d.size.rows; cfg.cols = d.size.cols; cfg.n_channels = _n_channels;
  cfg.patch_radius = _options.patchRadius; cfg.max_frames = _options.slidingWindowSize;
  d.cap_points = std::max(n_points, _options.maxNumPoints * _options.slidingWindowSize);
  d.cap_obs = std::max(n_obs, d.cap_points * std::min(_options.slidingWindowSize, 4));
  cfg.max_points = d.cap_points; cfg.max_observations = d.cap_obs;
  cfg.device = _options.device;
  cfg.fx = d.calib.fx(); cfg.fy = d.calib.fy(); cfg.cx = d.calib.cx(); cfg.cy = d.calib.cy();
  cfg.huber = _options.robustThreshold;
  const int n_gpus = _options.nGpus;
  if (n_gpus > 1 && cfg.device < 0) cfg.device = 0;
  for (int r = 0; r < n_gpus; ++r) {      // one handle per device; the points are sharded at pba_set_points
    pba_handle* h = nullptr;
    check_pba(pba_create(&cfg, &h), "pba_create");
    d.ranks.push_back(h);
    cfg.device += 1;
  }
  d.h = d.ranks[0];
  if (n_gpus > 1) check_pba(pba_comm_init_local(d.ranks.data(), n_gpus), "pba_comm_init_local");
  std::fill(d.resident.begin(), d.resident.end(), (long long)-1);
  return d.h;
}

// ---------------------------------------------------------------------------- addFrame
// Integer-pixel patch of the reference frame, clamped so that the whole patch stays inside the image
// (ExtractPatch, src/photobundle.cc:466-479; the channel is the uint8 image cast to float, widened to double).
static void ExtractPatch(double* dst, const uint8_t* I, int rows, int cols, int ux, int uy, int radius) {
  const int side = 2 * radius + 1;
  for (int k = 0; k < side * side; ++k) {
    const int yy = std::max(radius, std::min(uy + k / side - radius, rows - radius - 1));
    const int xx = std::max(radius, std::min(ux + k % side - radius, cols - radius - 1));
    dst[k] = static_cast<double>(static_cast<float>(I[(size_t)yy * cols + xx]));
  }
}

// PBA_HOST_TIMING=1: wall-clock phases of addFrame / optimize on stderr (developer aid, scripts/seq_timing.py)
namespace {
struct PhaseClock {
  const bool on = getenv("PBA_HOST_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  std::string line;
  void mark(const char* what) {
    if (!on) return;
    const auto n = std::chrono::steady_clock::now();
    char buf[64];
    snprintf(buf, sizeof(buf), " %s %.0f us", what, 1e6 * std::chrono::duration<double>(n - t).count());
    line += buf;
    t = n;
  }
  void flush(const char* who) { if (on) fprintf(stderr, "[pba host] %s:%s\n", who, line.c_str()); }
};
}  // namespace

void PhotometricBundleAdjustment::addFrame(const uint8_t* I_ptr, const float* Z_ptr, const Mat44& T, Result* result) {
  PhaseClock clk;
  _trajectory.push_back(T, (int)_frame_id);
  const Mat44 T_w = _trajectory.back(), T_c = T_w.rigidInverse();
  const int rows = _image_size.rows, cols = _image_size.cols;
  const int radius = _options.patchRadius, descriptor_dim = (2 * radius + 1) * (2 * radius + 1) * _n_channels;
  // border band in which points may live (src/photobundle.cc:498-503): the wider of the mask block, the ZNCC patch
  // and the descriptor patch
  const int border = std::max(_options.maskBlockRadius, std::max(2, radius));
  const int row_end = rows - border - 1, col_end = cols - border - 1, mask_radius = _options.maskBlockRadius;
  const bool multi = _desc_type != PBA_DESC_INTENSITY, on_gpu = _options.gpuFrontEnd;

  Frame frame;
  frame.id = _frame_id;
  frame.image.assign(I_ptr, I_ptr + (size_t)rows * cols);
  if (multi || on_gpu) check_pba(pba_prepare_frame_u8(deviceLevel(0, 0, 0), I_ptr, _desc_type), "pba_prepare_frame_u8");
  pba_handle* const gpu0 = _dev[0].h;
  clk.mark("copy+prepare");

  // ---- (1) which live points are seen again in this frame (src/photobundle.cc:508-542): project with the INITIAL
  // pose, compare the stored 5x5 patch with the one around the projection (ZNCC), block the neighbourhood of a hit
  std::fill(_mask.begin(), _mask.end(), (uint16_t)1);
  int n_reobserved = 0, n_tested = 0;
  std::vector<int32_t> hit_rc;                  // (row, col) of the hits, for the device-side mask
  auto is_live = [&](const ScenePoint& p) { return (int)_frame_id - (int)p.lastFrameId() <= _options.maxFrameDistance; };
  if (on_gpu) {
    std::vector<ScenePoint*> live;
    live.reserve(_scene_points.size());
    for (auto& sp : _scene_points)
      if (is_live(*sp)) live.push_back(sp.get());
    n_tested = (int)live.size();
    const size_t nl = live.size();
    // gathered straight into page-locked memory: 128 bytes per live point cross the link every frame
    double* xyz = static_cast<double*>(_pin_xyz.need(sizeof(double) * 3 * nl));
    float* patches = static_cast<float*>(_pin_patch.need(sizeof(float) * ZnccPatch::kLen * nl));
    float* norms = static_cast<float*>(_pin_norm.need(sizeof(float) * nl));
    float* score = static_cast<float*>(_pin_score.need(sizeof(float) * nl));
    int32_t* rc = static_cast<int32_t*>(_pin_rc.need(sizeof(int32_t) * 2 * nl));
    for (size_t k = 0; k < nl; ++k) {
      const ScenePoint& sp = *live[k];
      for (int a = 0; a < 3; ++a) xyz[3 * k + a] = sp.X[a];
      std::copy(sp.patch.data, sp.patch.data + ZnccPatch::kLen, patches + (size_t)ZnccPatch::kLen * k);
      norms[k] = sp.patch.norm;
    }
    double K_rowmajor[9];
    for (int k = 0; k < 9; ++k) K_rowmajor[k] = _calib.K()(k / 3, k % 3);
    check_pba(pba_associate(gpu0, (int32_t)nl, xyz, patches, norms, T_c.m, K_rowmajor, border, score, rc), "pba_associate");
    for (size_t k = 0; k < live.size(); ++k) {
      if (!(score[k] > _options.minScore)) continue;   // points outside the band carry -2
      ++n_reobserved;
      live[k]->f.push_back(_frame_id);
      hit_rc.push_back(rc[2 * k]); hit_rc.push_back(rc[2 * k + 1]);
    }
  } else {
    for (auto& sp : _scene_points) {
      ScenePoint& pt = *sp;
      if (!is_live(pt)) continue;
      ++n_tested;
      const Vec2 uv = _calib.project(T_c.transform(pt.X));
      const int r = (int)std::round(uv[1]), c = (int)std::round(uv[0]);
      if (!(r >= border && r < row_end && c >= border && c <= col_end)) continue;   // (the reference's band is closed on the right)
      ZnccPatch seen;
      seen.set(I_ptr, rows, cols, uv[0], uv[1]);
      if (!(pt.patch.score(seen) > _options.minScore)) continue;
      ++n_reobserved;
      pt.f.push_back(_frame_id);
      for (int dr = -mask_radius; dr <= mask_radius; ++dr)
        for (int dc = -mask_radius; dc <= mask_radius; ++dc) _mask[(size_t)(r + dr) * cols + c + dc] = 0;
    }
  }

  clk.mark("associate");
  // ---- (2) new points (src/photobundle.cc:545-575): unmasked strict local maxima of the saliency map (sum over the
  // channels of |Ix| + |Iy|, central differences, zero on the image border) that carry a valid depth
  ScenePointPointerList fresh;
  auto lift = [&](int y, int x, float z, float saliency) {
    // X = T_w * (z * K^-1 * [x y 1]^T), the products in the reference's order (src/photobundle.cc:560)
    const double zd = z, pix[3] = {(double)x, (double)y, 1.0};
    Vec3 Xc;
    for (int i = 0; i < 3; ++i) Xc[i] = (zd * _K_inv(i, 0)) * pix[0] + (zd * _K_inv(i, 1)) * pix[1] + (zd * _K_inv(i, 2)) * pix[2];
    UniquePointer<ScenePoint> p(new ScenePoint(T_w.transform(Xc), _frame_id));
    p->patch.set(I_ptr, rows, cols, (double)x, (double)y);
    p->descriptor.resize((size_t)descriptor_dim);
    p->saliency = saliency;
    p->x = x; p->y = y;
    fresh.push_back(std::move(p));
  };
  if (on_gpu) {
    // saliency, mask, non-maximum suppression and the depth test run on the device; candidates come back in scan order
    int32_t room = std::max(4096, (rows * cols) / 16), n_cand = 0;
    std::vector<int32_t> cand_rc;
    std::vector<float> cand_sal;
    for (int attempt = 0; attempt < 2; ++attempt) {
      cand_rc.resize(2 * (size_t)room); cand_sal.resize((size_t)room);
      check_pba(pba_select_candidates(gpu0, nullptr /* depth test below: the map stays on the host */, (int32_t)(hit_rc.size() / 2), hit_rc.data(), mask_radius, _options.nonMaxSuppRadius, border,
                                      _options.minValidDepth, _options.maxValidDepth, room, cand_rc.data(), cand_sal.data(), &n_cand),
                "pba_select_candidates");
      if (n_cand <= room) break;
      room = n_cand;                             // rare: more candidates than room; once more with enough
    }
    // keep the maxNumPoints most salient BEFORE building scene points (step (3) below, on indices: std::nth_element's
    // permutation depends only on the comparison results, so the survivors and their order are those of the host path)
    std::vector<int> keep;
    keep.reserve((size_t)n_cand);
    for (int k = 0; k < n_cand; ++k) {             // the depth test of src/photobundle.cc:552-553, in scan order
      const float z = Z_ptr[(size_t)cand_rc[2 * (size_t)k] * cols + cand_rc[2 * (size_t)k + 1]];
      if ((double)z >= _options.minValidDepth && (double)z <= _options.maxValidDepth) keep.push_back(k);
    }
    if ((int)keep.size() > _options.maxNumPoints) {
      std::nth_element(keep.begin(), keep.begin() + _options.maxNumPoints, keep.end(),
                       [&](int a, int b) { return cand_sal[(size_t)a] > cand_sal[(size_t)b]; });
      keep.resize((size_t)_options.maxNumPoints);
    }
    fresh.reserve(keep.size());
    for (int k : keep) lift(cand_rc[2 * (size_t)k], cand_rc[2 * (size_t)k + 1], Z_ptr[(size_t)cand_rc[2 * (size_t)k] * cols + cand_rc[2 * (size_t)k + 1]], cand_sal[(size_t)k]);
  } else {
    if (multi) {
      check_pba(pba_saliency_map(gpu0, _saliency_map.data()), "pba_saliency_map");
    } else {
      std::fill(_saliency_map.begin(), _saliency_map.end(), 0.0f);
      for (int y = 1; y < rows - 1; ++y) {
        const uint8_t* up = I_ptr + (size_t)(y - 1) * cols, *mid = up + cols, *down = mid + cols;
        for (int x = 1; x < cols - 1; ++x)
          _saliency_map[(size_t)y * cols + x] = std::fabs(0.5f * ((float)mid[x + 1] - (float)mid[x - 1])) + std::fabs(0.5f * ((float)down[x] - (float)up[x]));
      }
    }
    const int nms = _options.nonMaxSuppRadius;
    auto beats_neighbours = [&](int row, int col) -> bool {   // IsLocalMax_ (src/imgproc.h:175-212): ties lose
      if (nms <= 0) return true;
      const float v = _saliency_map[(size_t)row * cols + col];
      if (!_mask[(size_t)row * cols + col] || v < 0.0f) return false;
      for (int dr = -nms; dr <= nms; ++dr)
        for (int dc = -nms; dc <= nms; ++dc)
          if ((dr || dc) && _saliency_map[(size_t)(row + dr) * cols + col + dc] >= v) return false;
      return true;
    };
    for (int y = border; y < row_end; ++y)
      for (int x = border; x < col_end; ++x) {
        const float z = Z_ptr[(size_t)y * cols + x];
        if (z >= _options.minValidDepth && z <= _options.maxValidDepth && beats_neighbours(y, x)) lift(y, x, z, _saliency_map[(size_t)y * cols + x]);
      }
  }
  clk.mark("candidates");
  // ---- (3) keep the maxNumPoints most salient (src/photobundle.cc:578-585; std::nth_element, so which of several
  // equally salient points survives is the standard library's choice, as in the reference)
  if (fresh.size() > (size_t)_options.maxNumPoints) {
    auto cut = fresh.begin() + _options.maxNumPoints;
    std::nth_element(fresh.begin(), cut, fresh.end(),
                     [](const UniquePointer<ScenePoint>& a, const UniquePointer<ScenePoint>& b) { return a->saliency > b->saliency; });
    fresh.erase(cut, fresh.end());
  }
  if (_options.verbose)
    printf("updated %d [%0.2f%%] max %d new %d\n", n_reobserved, 100.0 * n_reobserved / _scene_points.size(), n_tested, (int)fresh.size());
  clk.mark("top-n");
  // ---- (4) reference descriptors of the new points (src/photobundle.cc:597-606), every pyramid level
  if (multi) {
    const int n_new = (int)fresh.size();
    std::vector<int32_t> xy((size_t)2 * n_new);
    for (int k = 0; k < n_new; ++k) { xy[2 * k] = fresh[(size_t)k]->x; xy[2 * k + 1] = fresh[(size_t)k]->y; }
    std::vector<double> dsc((size_t)n_new * descriptor_dim);
    check_pba(pba_extract_descriptors(gpu0, n_new, xy.data(), dsc.data()), "pba_extract_descriptors");
    for (int k = 0; k < n_new; ++k)
      std::copy(dsc.begin() + (size_t)k * descriptor_dim, dsc.begin() + (size_t)(k + 1) * descriptor_dim, fresh[(size_t)k]->descriptor.begin());
  } else {
    for (auto& p : fresh) ExtractPatch(p->descriptor.data(), I_ptr, rows, cols, p->x, p->y, radius);
  }
  if (_options.numPyramidLevels > 1) {
    std::vector<uint8_t> img = frame.image;
    int lr = rows, lc = cols;
    for (int l = 1; l < _options.numPyramidLevels; ++l) {
      img = PyrDownU8(img, lr, lc);
      lr = (lr + 1) / 2; lc = (lc + 1) / 2;
      const double s = (double)(1 << l);
      for (auto& p : fresh) {
        if (p->coarse_desc.empty()) p->coarse_desc.resize((size_t)_options.numPyramidLevels - 1);
        p->coarse_desc[(size_t)l - 1].resize((size_t)descriptor_dim);
        BilinearPatch(p->coarse_desc[(size_t)l - 1].data(), img, lr, lc, p->x / s, p->y / s, radius);
      }
    }
  }
  _scene_points.reserve(_scene_points.size() + fresh.size());
  for (auto& p : fresh) _scene_points.push_back(std::move(p));

  // ring buffer of slidingWindowSize frames (boost::circular_buffer in the reference, :608); a full window is solved
  if ((int)_frame_buffer.size() == _options.slidingWindowSize) _frame_buffer.pop_front();
  _frame_buffer.push_back(std::move(frame));
  clk.mark("descriptors");
  if ((int)_frame_buffer.size() == _options.slidingWindowSize) { optimize(result); clk.mark("optimize"); }
  clk.flush("addFrame");
  ++_frame_id;
}

// ---------------------------------------------------------------------------- optimize
// Frames live in device slots id % slidingWindowSize: a new frame replaces the one that left the window, everything
// else stays resident (the reference's ring buffer, on the device).  The window-local frame index of the C ABI is
// that slot.
void PhotometricBundleAdjustment::uploadWindowFrames(int level) {
  DeviceLevel& d = _dev[(size_t)level];
  const int W = _options.slidingWindowSize, rows = _image_size.rows, cols = _image_size.cols;
  for (const Frame& f : _frame_buffer) {
    const size_t slot = (size_t)(f.id % (uint32_t)W);
    if (d.resident[slot] == (long long)f.id) continue;
    for (pba_handle* h : d.ranks)   // frames are replicated on every device of the level
      check_pba(pba_set_frame_u8_ex(h, (int32_t)slot, f.image.data(), rows, cols, d.levels_down, _desc_type), "pba_set_frame_u8_ex");
    d.resident[slot] = f.id;
  }
}

struct PhotometricBundleAdjustment::SolveOutcome {
  pba_summary summary;
  std::vector<pba_iteration_summary> iters;
};

void PhotometricBundleAdjustment::optimize(Result* result) {
  PhaseClock clk;
  const auto t0 = std::chrono::steady_clock::now();
  const uint32_t first_id = _frame_buffer.front().id, last_id = _frame_buffer.back().id;
  const int W = _options.slidingWindowSize, L = _options.numPyramidLevels;
  const std::vector<double> patch_weights = MakePatchWeights(_options.patchRadius, _options.doGaussianWeighting);
  auto slot_of = [&](uint32_t id) { return (int)(id % (uint32_t)W); };

  // camera parameters: the INVERSE world pose of every window frame as [angle-axis, t] (src/photobundle.cc:774-778)
  std::vector<double> cams((size_t)W * 6, 0.0);
  for (uint32_t id = first_id; id <= last_id; ++id) PoseToParams(_trajectory.atId((int)id).rigidInverse(), &cams[(size_t)slot_of(id) * 6]);

  // residual blocks (src/photobundle.cc:786-806): points seen in >= 3 frames whose reference frame is inside the
  // window, one block per frame of the visibility list that lies in the window (the reference frame included)
  std::vector<ScenePoint*> chosen;
  std::vector<double> xyz;
  std::vector<std::vector<double>> desc((size_t)L);
  std::vector<int32_t> obs_off(1, 0), obs_frame;
  {   // one allocation each instead of growth by doubling (a 10 000-point window packs 2 MB of descriptors)
    const size_t n_all = _scene_points.size(), dd = n_all ? _scene_points[0]->descriptor.size() : 0;
    chosen.reserve(n_all); xyz.reserve(3 * n_all); obs_off.reserve(n_all + 1); obs_frame.reserve(n_all * (size_t)std::min(W, 8));
    for (int l = 0; l < L; ++l) desc[(size_t)l].reserve(n_all * dd);
  }
  for (auto& sp : _scene_points) {
    ScenePoint& pt = *sp;
    if (pt.numFrames() < 3 || pt.refFrameId() < first_id) continue;
    size_t before = obs_frame.size();
    for (uint32_t id : pt.f)
      if (id >= first_id && id <= last_id) obs_frame.push_back(slot_of(id));
    if (obs_frame.size() > before) pt.was_refined = true;
    chosen.push_back(&pt);
    xyz.insert(xyz.end(), pt.X.data(), pt.X.data() + 3);
    desc[0].insert(desc[0].end(), pt.descriptor.begin(), pt.descriptor.end());
    for (int l = 1; l < L; ++l) desc[(size_t)l].insert(desc[(size_t)l].end(), pt.coarse_desc[(size_t)l - 1].begin(), pt.coarse_desc[(size_t)l - 1].end());
    obs_off.push_back((int32_t)obs_frame.size());
  }
  const int n_sel = (int)chosen.size(), nnz = (int)obs_frame.size();
  if (_options.verbose) printf("Using %d points (%d residual blocks) [id start %u]\n", n_sel, nnz, first_id);

  clk.mark("pack");
  SolveOutcome solved;
  pba_summary& summary = solved.summary;
  std::vector<pba_iteration_summary>& iters = solved.iters;
  memset(&summary, 0, sizeof(summary));
  if (n_sel > 0 && nnz > 0) {
    pba_solver_options opt;
    pba_default_solver_options(&opt);       // GetSolverOptions, src/photobundle.cc:738-761
    const DeviceLevel* coarser_level = nullptr;
    pba_handle* coarser = nullptr;
    for (int l = L - 1; l >= 0; --l) {      // coarse to fine; a single level is the reference's optimize()
      deviceLevel(l, n_sel, nnz);
      const DeviceLevel& d = _dev[(size_t)l];
      const size_t R = d.ranks.size();
      for (pba_handle* h : d.ranks) check_pba(pba_begin_batch(h), "pba_begin_batch");   // uploads are enqueued; every buffer lives until pba_solve returns
      uploadWindowFrames(l);
      for (size_t r = 0; r < R; ++r) {      // every rank receives the whole window and keeps its shard of the points
        pba_handle* h = d.ranks[r];
        check_pba(pba_set_poses(h, W, cams.data(), slot_of(first_id) /* first camera constant, :809-816 */), "pba_set_poses");
        check_pba(pba_set_points(h, n_sel, xyz.data(), desc[(size_t)l].data(), obs_off.data(), obs_frame.data(), patch_weights.data()), "pba_set_points");
        if (coarser_level) check_pba(pba_copy_state(h, coarser_level->ranks[r]), "pba_copy_state");   // the coarser level's result, on the device
      }
      if (R == 1) {
        check_pba(pba_solve(d.h, &opt, &summary), "pba_solve");
      } else {                               // one host thread per device: the ranks exchange sums from inside the kernels
        std::vector<pba_summary> sums(R);
        std::vector<std::string> errors(R);
        std::vector<std::thread> workers;
        for (size_t r = 0; r < R; ++r)
          workers.emplace_back([&, r] { if (pba_solve(d.ranks[r], &opt, &sums[r]) != PBA_OK) errors[r] = pba_last_error(); });
        for (std::thread& w : workers) w.join();
        for (size_t r = 0; r < R; ++r)
          if (!errors[r].empty()) throw std::runtime_error("pba_solve (device rank " + std::to_string(r) + "): " + errors[r]);
        summary = sums[0];
      }
      coarser_level = &d;
      coarser = d.h;
    }
    clk.mark("upload+solve");
    check_pba(pba_get_results(coarser, cams.data(), xyz.data()), "pba_get_results");
    int32_t n_it = 0;
    pba_get_iterations(coarser, nullptr, 0, &n_it);
    iters.resize((size_t)n_it);
    if (n_it) pba_get_iterations(coarser, iters.data(), n_it, &n_it);
    for (int k = 0; k < n_sel; ++k)
      for (int a = 0; a < 3; ++a) chosen[(size_t)k]->X[a] = xyz[(size_t)k * 3 + a];
  } else {
    snprintf(summary.message, sizeof(summary.message), "no residual blocks in the window");
  }

  // refined world poses back into the trajectory (src/photobundle.cc:841-844)
  for (uint32_t id = first_id; id <= last_id; ++id)
    _trajectory.atId((int)id) = ParamsToPose(&cams[(size_t)slot_of(id) * 6]).rigidInverse();

  // points anchored at the oldest frame leave the system with it (:851, :888-905)
  ScenePointPointerList leaving = removePointsAtFrame(first_id);
  if (_options.verbose) printf("removing %zu old points\n", leaving.size());
  clk.mark("write-back+evict");
  if (result) fillResult(*result, solved, leaving, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  clk.mark("result");
  clk.flush("optimize");
}

// Result as the reference fills it (src/photobundle.cc:857-875): the WHOLE trajectory, the points that just left
// (refined and as initialised), the solver's figures.
void PhotometricBundleAdjustment::fillResult(Result& out, const SolveOutcome& solved, const ScenePointPointerList& leaving, double seconds) const {
  const pba_summary& summary = solved.summary;
  const std::vector<pba_iteration_summary>& iters = solved.iters;
  out.poses = _trajectory.poses();
  out.refinedPoints.clear(); out.originalPoints.clear();
  for (const auto& p : leaving) { out.refinedPoints.push_back(p->X); out.originalPoints.push_back(p->X_original); }
  out.initialCost = summary.initial_cost; out.finalCost = summary.final_cost; out.fixedCost = summary.fixed_cost;
  out.numSuccessfulStep = summary.num_successful_steps;
  out.numResiduals = summary.num_residuals;
  out.totalTime = seconds;
  out.message = summary.message;
  out.iterationSummary.clear();
  for (const pba_iteration_summary& s : iters) {
    ceres::IterationSummary o;
    o.iteration = s.iteration; o.step_is_valid = s.step_is_valid; o.step_is_nonmonotonic = s.step_is_nonmonotonic;
    o.step_is_successful = s.step_is_successful; o.cost = s.cost; o.cost_change = s.cost_change;
    o.gradient_max_norm = s.gradient_max_norm; o.gradient_norm = s.gradient_norm; o.step_norm = s.step_norm;
    o.relative_decrease = s.relative_decrease; o.trust_region_radius = s.trust_region_radius;
    o.linear_solver_iterations = s.linear_solver_iterations;
    out.iterationSummary.push_back(o);
  }
}

auto PhotometricBundleAdjustment::removePointsAtFrame(uint32_t id) -> ScenePointPointerList {
  auto stays = [id](const UniquePointer<ScenePoint>& p) { return p->refFrameId() > id; };
  auto split = std::stable_partition(_scene_points.begin(), _scene_points.end(), stays);
  ScenePointPointerList gone(std::make_move_iterator(split), std::make_move_iterator(_scene_points.end()));
  _scene_points.erase(split, _scene_points.end());
  return gone;
}

// ---------------------------------------------------------------------------- pyramid front
// The reference's PhotometricBundleAdjustmentPyr (src/photobundle_pyramid.{h,cc}) is an unfinished sketch: one
// independent optimiser per level, none at level 0, a level loop that starts out of range.
// Kept: its interface.  Defined here: ONE point set (created at the finest level), every level solves the same
// window — frames reduced with cv::pyrDown's rule, intrinsics halved per level, descriptors re-read per level —
// coarse to fine, each level starting from the coarser level's poses and points.
PhotometricBundleAdjustmentPyr::PhotometricBundleAdjustmentPyr(int num_levels, const Calibration& calib, const ImageSize& size, const Options& options) {
  if (num_levels < 1) throw std::runtime_error("PhotometricBundleAdjustmentPyr: num_levels < 1");
  Options o = options;
  o.numPyramidLevels = num_levels;
  _ba.reset(new PhotometricBundleAdjustment(calib, size, o));
}
PhotometricBundleAdjustmentPyr::~PhotometricBundleAdjustmentPyr() {}
void PhotometricBundleAdjustmentPyr::addFrame(const uint8_t* image, const float* depth_map, const Mat44& T, Result* result) {
  _ba->addFrame(image, depth_map, T, result);
}
