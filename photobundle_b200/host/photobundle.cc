// photobundle.cc — host side of the drop-in: the reference's addFrame() bookkeeping
// (src/photobundle.cc:482-615) restated without Eigen/Boost, and optimize()
// (src/photobundle.cc:764-876) re-implemented as "pack the window, call the B200 kernels through
// the C ABI, write the poses/points back, evict, fill Result".

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

#include "../../include/pba_b200.h"
#include "config.h"
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
int Trajectory::find(const Id_t id) const {
  for (size_t i = 0; i < _data.size(); ++i) if (_data[i].id == id) return (int)i;
  return -1;
}
void Trajectory::push_back(const Mat44& pose, const Id_t id) {
  if (find(id) >= 0) throw std::runtime_error("duplicate id in trajectory\n");
  const Mat44 T_inv = pose.inverse();
  if (!_data.empty()) _data.push_back({back() * T_inv, id});
  else _data.push_back({T_inv, id});
}
const Mat44& Trajectory::atId(const Id_t id) const {
  const int i = find(id);
  if (i < 0) throw std::runtime_error("could not find pose with id");
  return _data[i].pose;
}
Mat44& Trajectory::atId(const Id_t id) {
  const int i = find(id);
  if (i < 0) throw std::runtime_error("could not find pose with id");
  return _data[i].pose;
}
EigenAlignedContainer_<Mat44> Trajectory::poses() const {
  EigenAlignedContainer_<Mat44> ret(_data.size());
  for (size_t i = 0; i < ret.size(); ++i) ret[i] = _data[i].pose;
  return ret;
}
EigenAlignedContainer_<Vec3> Trajectory::cameraPositions() const {
  EigenAlignedContainer_<Vec3> ret(_data.size());
  for (size_t i = 0; i < ret.size(); ++i) ret[i] = Vec3(_data[i].pose(0, 3), _data[i].pose(1, 3), _data[i].pose(2, 3));
  return ret;
}

// ---------------------------------------------------------------------------- pose utils
PoseList loadPosesKittiFormat(std::string fn) {
  std::ifstream ifs(fn);
  if (!ifs.is_open()) throw std::runtime_error("failed to open pose file");
  PoseList ret;
  std::string line;
  while (std::getline(ifs, line)) {
    if (line.empty()) continue;
    std::stringstream ss(line);
    double vals[12];
    for (int i = 0; i < 12; ++i) ss >> vals[i];
    Mat44 T = Mat44::Identity();
    for (int i = 0, c = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) T(i, j) = vals[c++];
    ret.push_back(T);
  }
  return ret;
}
bool writePosesKittiFormat(std::string fn, const PoseList& T) {
  std::ofstream ofs(fn);
  if (!ofs.is_open()) return false;
  for (size_t i = 0; i < T.size(); ++i) {
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) ofs << (T[i](r, c)) << " ";
    ofs << "\n";
  }
  return true;
}
PoseList convertPoseToLocal(const PoseList& T_w) {
  if (T_w.empty()) throw std::runtime_error("no poses");
  PoseList T_i(T_w.size());
  T_i[0] = T_w[0].rigidInverse();
  for (size_t i = 1; i < T_w.size(); ++i) T_i[i] = T_w[i].rigidInverse() * T_w[i - 1];
  return T_i;
}
void disparityToDepth(const float* dmap, const ImageSize& sz, float Bf, float* zmap) {
  const int N = sz.numel();
  for (int i = 0; i < N; ++i) zmap[i] = dmap[i] > 0.01f ? Bf * (1.0f / dmap[i]) : -0.10f;
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

// src/photobundle.cc:617-644
static std::vector<double> MakePatchWeights(int radius, bool do_gaussian) {
  const int n = (2 * radius + 1) * (2 * radius + 1);
  if (!do_gaussian) return std::vector<double>(n, 1.0);
  std::vector<double> ret(n);
  double sum = 0.0;
  for (int r = -radius, i = 0; r <= radius; ++r) {
    const double d_r = (r * r) / 1.0;
    for (int c = -radius; c <= radius; ++c, ++i) {
      const double d_c = (c * c) / 1.0;
      const double w = 1.0 * std::exp(-0.5 * (d_r + d_c));
      ret[i] = w; sum += w;
    }
  }
  for (int i = 0; i < n; ++i) ret[i] /= sum;
  return ret;
}

// ---------------------------------------------------------------------------- ZNCC patch
// interp2 (src/photobundle.cc:262-294) with T = float on a uint8 image, same mixed precision.
static inline float interp2_u8(const uint8_t* I, int rows, int cols, float xf, float yf, float fillval = 0.0f) {
  const int max_cols = cols - 1, max_rows = rows - 1;
  int xi = (int)std::floor(xf), yi = (int)std::floor(yf);
  xf -= xi; yf -= yi;
  auto at = [&](int y, int x) -> int { return I[(size_t)y * cols + x]; };
  if (xi >= 0 && xi < max_cols && yi >= 0 && yi < max_rows) {
    const float wx = 1.0 - xf;
    return (1.0 - yf) * (at(yi, xi) * wx + at(yi, xi + 1) * xf) + yf * (at(yi + 1, xi) * wx + at(yi + 1, xi + 1) * xf);
  } else {
    if (xi == max_cols && yi < max_rows && yi >= 0) return (xf > 0) ? fillval : (float)((1.0 - yf) * at(yi, xi) + yf * at(yi + 1, xi));
    else if (yi == max_rows && xi < max_cols && xi >= 0) return (yf > 0) ? fillval : (float)((1.0 - xf) * at(yi, xi) + xf * at(yi, xi + 1));
    else if (xi == max_cols && yi == max_rows) return (xf > 0 || yf > 0) ? fillval : (float)at(yi, xi);
    else return fillval;
  }
}

// ZnccPatch_<2, float> (src/photobundle.cc:315-361)
struct ZnccPatch {
  float data[25];
  float norm = 0.f;
  void set(const uint8_t* I, int rows, int cols, double px, double py) {
    const float x = (float)px, y = (float)py;
    int i = 0;
    for (int r = -2; r <= 2; ++r)
      for (int c = -2; c <= 2; ++c) data[i++] = interp2_u8(I, rows, cols, c + x, r + y);
    float sum = 0.f;
    for (int k = 0; k < 25; ++k) sum += data[k];
    const float mean = sum / 25.0f;
    float ss = 0.f;
    for (int k = 0; k < 25; ++k) { data[k] -= mean; ss += data[k] * data[k]; }
    norm = std::sqrt(ss);
  }
  float score(const ZnccPatch& o) const {
    const float d = norm * o.norm;
    float dot = 0.f;
    for (int k = 0; k < 25; ++k) dot += data[k] * o.data[k];
    return d > 1e-6 ? dot / d : -1.0f;
  }
};

struct PhotometricBundleAdjustment::ScenePoint {
  Vec3 X, X_original;
  std::vector<uint32_t> f;      // visibility list, first = reference frame
  ZnccPatch patch;
  std::vector<double> descriptor;
  double saliency = 0.0;
  bool was_refined = false;
  int x = 0, y = 0;             // first projection
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

PhotometricBundleAdjustment::Options::Options(const utils::ConfigFile& cf)   // src/photobundle.cc:88-103
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
      descriptorType(DescriptorTypeFromString(cf.get<std::string>("descriptorType", "Intensity"))) {}

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
  _mask.resize((size_t)_image_size.rows * _image_size.cols);
  _saliency_map.resize((size_t)_image_size.rows * _image_size.cols);
  _K_inv = calib.K().inverse();
}

PhotometricBundleAdjustment::~PhotometricBundleAdjustment() {
  if (_gpu) pba_destroy(_gpu);
}

void PhotometricBundleAdjustment::ensureGpu(int n_points, int n_obs) {
  if (_gpu && n_points <= _gpu_max_points && n_obs <= _gpu_max_obs) return;
  if (_gpu) { pba_destroy(_gpu); _gpu = nullptr; }
  pba_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.rows = _image_size.rows; cfg.cols = _image_size.cols; cfg.n_channels = _n_channels;
  cfg.patch_radius = _options.patchRadius; cfg.max_frames = _options.slidingWindowSize;
  _gpu_max_points = std::max(n_points, _options.maxNumPoints * _options.slidingWindowSize);
  _gpu_max_obs = std::max(n_obs, _gpu_max_points * std::min(_options.slidingWindowSize, 4));
  cfg.max_points = _gpu_max_points; cfg.max_observations = _gpu_max_obs;
  cfg.device = _options.device;
  cfg.fx = _calib.fx(); cfg.fy = _calib.fy(); cfg.cx = _calib.cx(); cfg.cy = _calib.cy();
  cfg.huber = _options.robustThreshold;
  if (pba_create(&cfg, &_gpu) != PBA_OK) throw std::runtime_error(std::string("pba_create: ") + pba_last_error());
}

// ---------------------------------------------------------------------------- addFrame
static inline int PatchSizeFromRadius(int r) { return (2 * r + 1) * (2 * r + 1); }

// ExtractPatch, src/photobundle.cc:466-479 (channel = uint8 image cast to float)
static void ExtractPatch(double* dst, const uint8_t* I, int rows, int cols, int ux, int uy, int radius) {
  const int max_cols = cols - radius - 1, max_rows = rows - radius - 1;
  for (int r = -radius, i = 0; r <= radius; ++r) {
    const int r_i = std::max(radius, std::min(uy + r, max_rows));
    for (int c = -radius; c <= radius; ++c, ++i) {
      const int c_i = std::max(radius, std::min(ux + c, max_cols));
      dst[i] = static_cast<double>(static_cast<float>(I[(size_t)r_i * cols + c_i]));
    }
  }
}

void PhotometricBundleAdjustment::addFrame(const uint8_t* I_ptr, const float* Z_ptr, const Mat44& T, Result* result) {
  _trajectory.push_back(T, (int)_frame_id);
  const Mat44 T_w = _trajectory.back();
  const Mat44 T_c = T_w.rigidInverse();
  const int rows = _image_size.rows, cols = _image_size.cols;

  Frame frame;
  frame.id = _frame_id;
  frame.image.assign(I_ptr, I_ptr + (size_t)rows * cols);

  const int B = std::max(_options.maskBlockRadius, std::max(2, _options.patchRadius));
  const int max_rows = rows - B - 1, max_cols = cols - B - 1, radius = _options.patchRadius,
            patch_length = PatchSizeFromRadius(radius), descriptor_dim = patch_length * _n_channels,
            mask_radius = _options.maskBlockRadius;

  // ---- visibility of the existing points in the new frame (src/photobundle.cc:508-542)
  const bool multi = _desc_type != PBA_DESC_INTENSITY;
  const bool on_gpu = _options.gpuFrontEnd;     // association + candidate selection on the device
  auto check = [&](int rc, const char* what) { if (rc != PBA_OK) throw std::runtime_error(std::string(what) + ": " + pba_last_error()); };
  if (multi || on_gpu) {
    ensureGpu(0, 0);
    check(pba_prepare_frame_u8(_gpu, I_ptr, _desc_type), "pba_prepare_frame_u8");
  }
  std::fill(_mask.begin(), _mask.end(), (uint16_t)1);
  int num_updated = 0, max_num_to_update = 0;
  std::vector<int32_t> masked_rc;               // re-observed pixels (device path)
  if (on_gpu) {
    std::vector<ScenePoint*> live;
    std::vector<double> xyz;
    std::vector<float> ref_patch, ref_norm;
    for (auto& sp : _scene_points)
      if ((int)_frame_id - (int)sp->lastFrameId() <= _options.maxFrameDistance) {
        live.push_back(sp.get());
        for (int k = 0; k < 3; ++k) xyz.push_back(sp->X[k]);
        ref_patch.insert(ref_patch.end(), sp->patch.data, sp->patch.data + 25);
        ref_norm.push_back(sp->patch.norm);
      }
    max_num_to_update = (int)live.size();
    std::vector<float> score(live.size());
    std::vector<int32_t> rc(2 * live.size());
    double Kr[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Kr[3 * i + j] = _calib.K()(i, j);
    check(pba_associate(_gpu, (int32_t)live.size(), xyz.data(), ref_patch.data(), ref_norm.data(), T_c.m, Kr, B, score.data(), rc.data()),
          "pba_associate");
    for (size_t i = 0; i < live.size(); ++i)
      if (score[i] > _options.minScore) {        // not-tested points carry -2
        num_updated++;
        live[i]->f.push_back(_frame_id);
        masked_rc.push_back(rc[2 * i]); masked_rc.push_back(rc[2 * i + 1]);
      }
  } else
  for (size_t i = 0; i < _scene_points.size(); ++i) {
    ScenePoint& pt = *_scene_points[i];
    const int f_dist = (int)_frame_id - (int)pt.lastFrameId();
    if (f_dist <= _options.maxFrameDistance) {
      const Vec2 uv = _calib.project(T_c.transform(pt.X));
      ++max_num_to_update;
      const int r = (int)std::round(uv[1]), c = (int)std::round(uv[0]);
      if (r >= B && r < max_rows && c >= B && c <= max_cols) {
        ZnccPatch other;
        other.set(I_ptr, rows, cols, uv[0], uv[1]);
        const float score = pt.patch.score(other);
        if (score > _options.minScore) {
          num_updated++;
          pt.f.push_back(_frame_id);
          for (int r_i = -mask_radius; r_i <= mask_radius; ++r_i)
            for (int c_i = -mask_radius; c_i <= mask_radius; ++c_i) _mask[(size_t)(r + r_i) * cols + c + c_i] = 0;
        }
      }
    }
  }

  // ---- new points: saliency = sum over channels of |Ix| + |Iy| (imgradient, zero borders), local maxima
  // with valid depth.  Multi-channel descriptors: the new frame's channels, their saliency and (below) the
  // reference descriptors come from the device (pba_prepare_frame_u8 / pba_saliency_map / pba_extract_descriptors).
  if (on_gpu) {
    // candidates come back in scan order with their saliency; the map itself stays on the device
  } else if (multi) {
    check(pba_saliency_map(_gpu, _saliency_map.data()), "pba_saliency_map");
  } else {
  std::fill(_saliency_map.begin(), _saliency_map.end(), 0.0f);
  for (int y = 1; y < rows - 1; ++y)
    for (int x = 1; x < cols - 1; ++x) {
      const float ix = 0.5f * ((float)I_ptr[(size_t)y * cols + x + 1] - (float)I_ptr[(size_t)y * cols + x - 1]);
      const float iy = 0.5f * ((float)I_ptr[(size_t)(y + 1) * cols + x] - (float)I_ptr[(size_t)(y - 1) * cols + x]);
      _saliency_map[(size_t)y * cols + x] = std::fabs(ix) + std::fabs(iy);
    }
  }
  const int nms = _options.nonMaxSuppRadius;
  auto is_local_max = [&](int row, int col) -> bool {   // IsLocalMax_, src/imgproc.h:175-212
    if (nms > 0) {
      const float v = _saliency_map[(size_t)row * cols + col];
      if (!_mask[(size_t)row * cols + col] || v < 0.0f) return false;
      for (int r = -nms; r <= nms; ++r)
        for (int c = -nms; c <= nms; ++c)
          if (!(!r && !c) && _saliency_map[(size_t)(r + row) * cols + c + col] >= v) return false;
    }
    return true;
  };
  ScenePointPointerList new_scene_points;
  auto make_point = [&](int y, int x, float z, float saliency) {
    // X = T_w * (z * K_inv * [x y 1]^T)   (src/photobundle.cc:560)
    const double zd = z, v[3] = {(double)x, (double)y, 1.0};
    Vec3 Xc;
    for (int i = 0; i < 3; ++i)
      Xc[i] = (zd * _K_inv(i, 0)) * v[0] + (zd * _K_inv(i, 1)) * v[1] + (zd * _K_inv(i, 2)) * v[2];
    UniquePointer<ScenePoint> p(new ScenePoint(T_w.transform(Xc), _frame_id));
    p->patch.set(I_ptr, rows, cols, (double)x, (double)y);
    p->descriptor.resize(descriptor_dim);
    p->saliency = saliency;
    p->x = x; p->y = y;
    new_scene_points.push_back(std::move(p));
  };
  if (on_gpu) {
    int32_t cap = std::max(4096, (rows * cols) / 16), n_cand = 0;
    std::vector<int32_t> cand_rc;
    std::vector<float> cand_sal;
    for (int attempt = 0; attempt < 2; ++attempt) {
      cand_rc.resize(2 * (size_t)cap); cand_sal.resize(cap);
      check(pba_select_candidates(_gpu, Z_ptr, (int32_t)(masked_rc.size() / 2), masked_rc.data(), mask_radius, nms, B, _options.minValidDepth,
                                  _options.maxValidDepth, cap, cand_rc.data(), cand_sal.data(), &n_cand), "pba_select_candidates");
      if (n_cand <= cap) break;
      cap = n_cand;                              // rare: more candidates than room; once more with enough
    }
    for (int i = 0; i < n_cand; ++i) {
      const int y = cand_rc[2 * i], x = cand_rc[2 * i + 1];
      make_point(y, x, Z_ptr[(size_t)y * cols + x], cand_sal[i]);
    }
  } else {
  for (int y = B; y < max_rows; ++y) {
    for (int x = B; x < max_cols; ++x) {
      const float z = Z_ptr[(size_t)y * cols + x];
      if (z >= _options.minValidDepth && z <= _options.maxValidDepth) {
        if (is_local_max(y, x)) make_point(y, x, z, _saliency_map[(size_t)y * cols + x]);
      }
    }
  }
  }
  // ---- keep the best N by saliency (src/photobundle.cc:578-585)
  if (new_scene_points.size() > (size_t)_options.maxNumPoints) {
    auto nth = new_scene_points.begin() + _options.maxNumPoints;
    std::nth_element(new_scene_points.begin(), nth, new_scene_points.end(),
                     [&](const UniquePointer<ScenePoint>& a, const UniquePointer<ScenePoint>& b) { return a->saliency > b->saliency; });
    new_scene_points.erase(nth, new_scene_points.end());
  }
  if (_options.verbose)
    printf("updated %d [%0.2f%%] max %d new %d\n", num_updated, 100.0 * num_updated / _scene_points.size(),
           max_num_to_update, (int)new_scene_points.size());
  if (multi) {
    // ExtractPatch of every channel (src/photobundle.cc:466-479, :601-606) on the device
    const int n_new = (int)new_scene_points.size();
    std::vector<int32_t> xy((size_t)2 * n_new);
    for (int i = 0; i < n_new; ++i) { xy[2 * i] = new_scene_points[i]->x; xy[2 * i + 1] = new_scene_points[i]->y; }
    std::vector<double> dsc((size_t)n_new * descriptor_dim);
    check(pba_extract_descriptors(_gpu, n_new, xy.data(), dsc.data()), "pba_extract_descriptors");
    for (int i = 0; i < n_new; ++i)
      std::copy(dsc.begin() + (size_t)i * descriptor_dim, dsc.begin() + (size_t)(i + 1) * descriptor_dim, new_scene_points[i]->descriptor.begin());
  } else {
    for (auto& p : new_scene_points) ExtractPatch(p->descriptor.data(), I_ptr, rows, cols, p->x, p->y, radius);
  }
  _scene_points.reserve(_scene_points.size() + new_scene_points.size());
  std::move(new_scene_points.begin(), new_scene_points.end(), std::back_inserter(_scene_points));

  // boost::circular_buffer(slidingWindowSize)::push_back
  if ((int)_frame_buffer.size() == _options.slidingWindowSize) _frame_buffer.pop_front();
  _frame_buffer.push_back(std::move(frame));
  if ((int)_frame_buffer.size() == _options.slidingWindowSize) optimize(result);
  ++_frame_id;
}

// ---------------------------------------------------------------------------- optimize
void PhotometricBundleAdjustment::optimize(Result* result) {
  const auto t0 = std::chrono::steady_clock::now();
  const uint32_t frame_id_start = _frame_buffer.front().id, frame_id_end = _frame_buffer.back().id;
  const int F = (int)(frame_id_end - frame_id_start + 1);
  const std::vector<double> patch_weights = MakePatchWeights(_options.patchRadius, _options.doGaussianWeighting);
  const int P = (int)patch_weights.size();

  // camera parameters of the window: inverse world pose as [angle-axis, t] (src/photobundle.cc:774-778)
  std::vector<double> cams((size_t)F * 6);
  for (uint32_t id = frame_id_start; id <= frame_id_end; ++id)
    PoseToParams(_trajectory.atId((int)id).rigidInverse(), &cams[(size_t)(id - frame_id_start) * 6]);

  // residual blocks (src/photobundle.cc:786-806)
  std::vector<ScenePoint*> selected;
  std::vector<double> xyz, desc;
  std::vector<int32_t> obs_off(1, 0), obs_frame;
  for (auto& pt : _scene_points) {
    if (pt->numFrames() >= 3 && pt->refFrameId() >= frame_id_start) {
      int n = 0;
      for (uint32_t id : pt->f)
        if (id >= frame_id_start && id <= frame_id_end) { obs_frame.push_back((int32_t)(id - frame_id_start)); ++n; }
      if (n > 0) pt->was_refined = true;
      selected.push_back(pt.get());
      for (int k = 0; k < 3; ++k) xyz.push_back(pt->X[k]);
      desc.insert(desc.end(), pt->descriptor.begin(), pt->descriptor.end());
      obs_off.push_back((int32_t)obs_frame.size());
    }
  }
  const int n_sel = (int)selected.size(), nnz = (int)obs_frame.size();
  if (_options.verbose)
    printf("Using %d points (%d residual blocks) [id start %u]\n", n_sel, nnz, frame_id_start);

  pba_summary summary;
  memset(&summary, 0, sizeof(summary));
  std::vector<pba_iteration_summary> iters;
  if (n_sel > 0 && nnz > 0) {
    ensureGpu(n_sel, nnz);
    std::vector<const uint8_t*> imgs(F);
    for (int f = 0; f < F; ++f) imgs[f] = _frame_buffer[f].image.data();
    auto check = [&](int rc, const char* what) { if (rc != PBA_OK) throw std::runtime_error(std::string(what) + ": " + pba_last_error()); };
    check(pba_set_frames_u8_descriptor(_gpu, F, imgs.data(), _desc_type), "pba_set_frames_u8_descriptor");
    check(pba_set_poses(_gpu, F, cams.data(), 0 /* first camera constant, :809-816 */), "pba_set_poses");
    check(pba_set_points(_gpu, n_sel, xyz.data(), desc.data(), obs_off.data(), obs_frame.data(), patch_weights.data()), "pba_set_points");
    pba_solver_options opt;
    pba_default_solver_options(&opt);       // GetSolverOptions, :738-761
    check(pba_solve(_gpu, &opt, &summary), "pba_solve");
    check(pba_get_poses(_gpu, cams.data()), "pba_get_poses");
    check(pba_get_points(_gpu, xyz.data()), "pba_get_points");
    int32_t n_it = 0;
    pba_get_iterations(_gpu, nullptr, 0, &n_it);
    iters.resize(n_it);
    if (n_it) pba_get_iterations(_gpu, iters.data(), n_it, &n_it);
    for (int i = 0; i < n_sel; ++i)
      for (int k = 0; k < 3; ++k) selected[i]->X[k] = xyz[(size_t)i * 3 + k];
    (void)P;
  } else {
    snprintf(summary.message, sizeof(summary.message), "no residual blocks in the window");
  }

  // put back the refined camera poses (src/photobundle.cc:841-844)
  for (uint32_t id = frame_id_start; id <= frame_id_end; ++id)
    _trajectory.atId((int)id) = ParamsToPose(&cams[(size_t)(id - frame_id_start) * 6]).rigidInverse();

  // all points whose reference frame is the window start leave the system (:851, :888-905)
  ScenePointPointerList points_to_remove = removePointsAtFrame(frame_id_start);
  if (_options.verbose) printf("removing %zu old points\n", points_to_remove.size());

  if (result) {
    result->poses = _trajectory.poses();
    const size_t npts = points_to_remove.size();
    result->refinedPoints.resize(npts);
    result->originalPoints.resize(npts);
    for (size_t i = 0; i < npts; ++i) {
      result->refinedPoints[i] = points_to_remove[i]->X;
      result->originalPoints[i] = points_to_remove[i]->X_original;
    }
    result->initialCost = summary.initial_cost;
    result->finalCost = summary.final_cost;
    result->fixedCost = summary.fixed_cost;
    result->numSuccessfulStep = summary.num_successful_steps;
    result->totalTime = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    result->numResiduals = summary.num_residuals;
    result->message = std::string(summary.message);
    result->iterationSummary.resize(iters.size());
    for (size_t i = 0; i < iters.size(); ++i) {
      ceres::IterationSummary& o = result->iterationSummary[i];
      const pba_iteration_summary& s = iters[i];
      o.iteration = s.iteration; o.step_is_valid = s.step_is_valid; o.step_is_nonmonotonic = s.step_is_nonmonotonic;
      o.step_is_successful = s.step_is_successful; o.cost = s.cost; o.cost_change = s.cost_change;
      o.gradient_max_norm = s.gradient_max_norm; o.gradient_norm = s.gradient_norm; o.step_norm = s.step_norm;
      o.relative_decrease = s.relative_decrease; o.trust_region_radius = s.trust_region_radius;
      o.linear_solver_iterations = s.linear_solver_iterations;
    }
  }
}

auto PhotometricBundleAdjustment::removePointsAtFrame(uint32_t id) -> ScenePointPointerList {
  ScenePointPointerList keep, remove;
  keep.reserve(_scene_points.size());
  for (auto& p : _scene_points) {
    if (p->refFrameId() <= id) remove.push_back(std::move(p));
    else keep.push_back(std::move(p));
  }
  _scene_points.swap(keep);
  return remove;
}
