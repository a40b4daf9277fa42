// host_capi.cc — a C face of the C++ host class so that the Python tests (ctypes) can drive
// PhotometricBundleAdjustment::addFrame end to end.  Exceptions become a status + message,
// exactly like the kernel-level C ABI.
#include <cstring>
#include <string>

#include "photobundle.h"
#include "pose_utils.h"

struct pbah_handle {
  PhotometricBundleAdjustment* ba = nullptr;
  PhotometricBundleAdjustment::Result result;
  int n_optimized = 0;
};
static thread_local std::string g_herr;

struct pbah_options {
  int32_t maxNumPoints, slidingWindowSize, patchRadius, maskBlockRadius, maxFrameDistance, nonMaxSuppRadius;
  int32_t doGaussianWeighting, verbose, device, descriptorType /* 0 Intensity, 1 IntensityAndGradient, 2 BitPlanes */, gpuFrontEnd;
  double minScore, robustThreshold, minValidDepth, maxValidDepth;
  int32_t numPyramidLevels, nGpus;
};

extern "C" {

const char* pbah_last_error(void) { return g_herr.c_str(); }

void pbah_default_options(pbah_options* o) {
  PhotometricBundleAdjustment::Options d;
  o->maxNumPoints = d.maxNumPoints; o->slidingWindowSize = d.slidingWindowSize; o->patchRadius = d.patchRadius;
  o->maskBlockRadius = d.maskBlockRadius; o->maxFrameDistance = d.maxFrameDistance; o->nonMaxSuppRadius = d.nonMaxSuppRadius;
  o->doGaussianWeighting = d.doGaussianWeighting; o->verbose = d.verbose; o->device = d.device; o->descriptorType = (int32_t)d.descriptorType; o->gpuFrontEnd = d.gpuFrontEnd ? 1 : 0;
  o->minScore = d.minScore; o->robustThreshold = d.robustThreshold; o->minValidDepth = d.minValidDepth; o->maxValidDepth = d.maxValidDepth;
  o->numPyramidLevels = d.numPyramidLevels; o->nGpus = d.nGpus;
}

int pbah_create(int32_t rows, int32_t cols, double fx, double fy, double cx, double cy, double baseline,
                const pbah_options* o, pbah_handle** out) {
  try {
    Mat33 K = Mat33::Identity();
    K(0, 0) = fx; K(1, 1) = fy; K(0, 2) = cx; K(1, 2) = cy;
    PhotometricBundleAdjustment::Options opt;
    opt.maxNumPoints = o->maxNumPoints; opt.slidingWindowSize = o->slidingWindowSize; opt.patchRadius = o->patchRadius;
    opt.maskBlockRadius = o->maskBlockRadius; opt.maxFrameDistance = o->maxFrameDistance; opt.nonMaxSuppRadius = o->nonMaxSuppRadius;
    opt.doGaussianWeighting = o->doGaussianWeighting != 0; opt.verbose = o->verbose != 0; opt.device = o->device;
    if (o->descriptorType < 0 || o->descriptorType > 2) throw std::runtime_error("descriptorType outside [0, 2]");
    opt.descriptorType = (PhotometricBundleAdjustment::Options::DescriptorType)o->descriptorType;
    opt.gpuFrontEnd = o->gpuFrontEnd != 0;
    opt.minScore = o->minScore; opt.robustThreshold = o->robustThreshold; opt.minValidDepth = o->minValidDepth; opt.maxValidDepth = o->maxValidDepth;
    opt.numPyramidLevels = o->numPyramidLevels > 0 ? o->numPyramidLevels : 1;
    opt.nGpus = o->nGpus > 0 ? o->nGpus : 1;
    pbah_handle* h = new pbah_handle();
    h->ba = new PhotometricBundleAdjustment(Calibration(K, baseline), ImageSize(rows, cols), opt);
    *out = h;
    return 0;
  } catch (const std::exception& e) { g_herr = e.what(); return -1; }
}

void pbah_destroy(pbah_handle* h) { if (h) { delete h->ba; delete h; } }

// T44: column-major 4x4 frame-to-frame pose. optimized: 1 if this call ran optimize().
int pbah_add_frame(pbah_handle* h, const uint8_t* image, const float* depth, const double* T44, int32_t* optimized) {
  try {
    Mat44 T;
    memcpy(T.m, T44, sizeof(T.m));
    const double before = h->result.totalTime;
    h->result.totalTime = -2.0;
    h->ba->addFrame(image, depth, T, &h->result);
    const bool ran = h->result.totalTime != -2.0;
    if (!ran) h->result.totalTime = before; else h->n_optimized++;
    if (optimized) *optimized = ran ? 1 : 0;
    return 0;
  } catch (const std::exception& e) { g_herr = e.what(); return -1; }
}

void pbah_result_counts(pbah_handle* h, int32_t* n_poses, int32_t* n_points, int32_t* n_iters) {
  *n_poses = (int32_t)h->result.poses.size();
  *n_points = (int32_t)h->result.refinedPoints.size();
  *n_iters = (int32_t)h->result.iterationSummary.size();
}

// scalars: initialCost, finalCost, fixedCost, numSuccessfulStep, numResiduals, totalTime
void pbah_result_get(pbah_handle* h, double* poses16, double* refined3, double* original3, double* scalars6,
                     double* iter_costs, char* message, int32_t message_cap) {
  const auto& r = h->result;
  if (poses16) for (size_t i = 0; i < r.poses.size(); ++i) memcpy(poses16 + 16 * i, r.poses[i].m, sizeof(double) * 16);
  for (size_t i = 0; i < r.refinedPoints.size(); ++i)
    for (int k = 0; k < 3; ++k) {
      if (refined3) refined3[3 * i + k] = r.refinedPoints[i][k];
      if (original3) original3[3 * i + k] = r.originalPoints[i][k];
    }
  if (scalars6) {
    scalars6[0] = r.initialCost; scalars6[1] = r.finalCost; scalars6[2] = r.fixedCost;
    scalars6[3] = r.numSuccessfulStep; scalars6[4] = r.numResiduals; scalars6[5] = r.totalTime;
  }
  if (iter_costs) for (size_t i = 0; i < r.iterationSummary.size(); ++i) iter_costs[i] = r.iterationSummary[i].cost;
  if (message && message_cap > 0) { strncpy(message, r.message.c_str(), message_cap - 1); message[message_cap - 1] = 0; }
}

int32_t pbah_num_scene_points(pbah_handle* h) { return (int32_t)h->ba->numScenePoints(); }

// vis: up to vis_cap frame ids; returns the visibility-list length
int32_t pbah_scene_point(pbah_handle* h, int32_t i, double* X3, int32_t* xy2, uint32_t* vis, int32_t vis_cap, double* desc, int32_t desc_cap) {
  const auto v = h->ba->scenePoint((size_t)i);
  for (int k = 0; k < 3; ++k) X3[k] = v.X[k];
  xy2[0] = v.x; xy2[1] = v.y;
  for (int k = 0; k < (int)v.visibility->size() && k < vis_cap; ++k) vis[k] = (*v.visibility)[k];
  for (int k = 0; k < (int)v.descriptor->size() && k < desc_cap; ++k) desc[k] = (*v.descriptor)[k];
  return (int32_t)v.visibility->size();
}

void pbah_disparity_to_depth(const float* disparity, int32_t rows, int32_t cols, float Bf, float* depth) {
  disparityToDepth(disparity, ImageSize(rows, cols), Bf, depth);
}

// the camera model and patch weights of the host side, for the tests against the reference's own (oracle/_ref/libref_calib.so)
void pbah_project(const double* k4, const double* X, double* uv) {
  Mat33 K = Mat33::Identity();
  K(0, 0) = k4[0]; K(1, 1) = k4[1]; K(0, 2) = k4[2]; K(1, 2) = k4[3];
  const Vec2 p = Calibration(K, 0.1).project(Vec3(X[0], X[1], X[2]));
  uv[0] = p[0]; uv[1] = p[1];
}
void pbah_pyr_down(const double* k4, double b, int32_t rows, int32_t cols, double* out7) {
  Mat33 K = Mat33::Identity();
  K(0, 0) = k4[0]; K(1, 1) = k4[1]; K(0, 2) = k4[2]; K(1, 2) = k4[3];
  const Calibration c = Calibration(K, b).pyrDown();
  const ImageSize s = ImageSize(rows, cols).pyrDown();
  out7[0] = c.fx(); out7[1] = c.fy(); out7[2] = c.cx(); out7[3] = c.cy(); out7[4] = c.b(); out7[5] = s.rows; out7[6] = s.cols;
}
void pbah_interp_patch5_u8(const uint8_t* I, int32_t rows, int32_t cols, double u, double v, float* out25) {
  InterpPatch5(I, rows, cols, u, v, out25);
}
int32_t pbah_patch_weights(int32_t radius, int32_t do_gaussian, double* w) {
  const std::vector<double> v = MakePatchWeights(radius, do_gaussian != 0);
  for (size_t i = 0; i < v.size(); ++i) w[i] = v[i];
  return (int32_t)v.size();
}

int pbah_write_poses_kitti(pbah_handle* h, const char* filename) {
  return writePosesKittiFormat(filename, h->result.poses) ? 0 : -1;
}

int32_t pbah_load_poses_kitti(const char* filename, double* poses16, int32_t cap) {
  try {
    const PoseList p = loadPosesKittiFormat(filename);
    for (size_t i = 0; i < p.size() && (int32_t)i < cap; ++i) memcpy(poses16 + 16 * i, p[i].m, sizeof(double) * 16);
    return (int32_t)p.size();
  } catch (const std::exception& e) { g_herr = e.what(); return -1; }
}

}  // extern "C"
