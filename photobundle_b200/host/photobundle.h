// photobundle.h — the reference's public BA interface (src/photobundle.h:19-196) kept name for
// name: Options, Result, ctor(Calibration, ImageSize, Options), addFrame(image, depth, T, Result*)
// and the protected optimize(Result*), so that apps/run_kitti.cc:32-55 compiles against it
// unchanged.  optimize() no longer builds a ceres::Problem: it packs the window and calls the
// B200 kernels through the C ABI (include/pba_b200.h).  Eigen/Ceres/Boost types are replaced by
// the stand-ins of compat.h (same names, same layout).
#ifndef PHOTOBUNDLE_PHOTOBUNDLE_H
#define PHOTOBUNDLE_PHOTOBUNDLE_H

#include <deque>
#include <iosfwd>
#include <string>
#include <vector>

#include "compat.h"
#include "trajectory.h"

struct pba_handle;
namespace utils { class ConfigFile; }

class PhotometricBundleAdjustment {
 public:
  struct Options {
    int maxNumPoints = 4096;        // src/photobundle.h:29
    int slidingWindowSize = 5;      // :32
    int patchRadius = 2;            // :35
    int maskBlockRadius = 1;        // :39
    int maxFrameDistance = 1;       // :42
    int numThreads = -1;            // :45 (unused: the solve runs on the GPU)
    bool doGaussianWeighting = false;
    bool verbose = true;
    double minScore = 0.75;
    double robustThreshold = 0.05;
    double minValidDepth = 0.01;
    double maxValidDepth = 1000.0;
    int nonMaxSuppRadius = 1;
    enum class DescriptorType { Intensity, IntensityAndGradient, BitPlanes };
    // The reference leaves this uninitialised in the default ctor (src/photobundle.h:77-79, UB);
    // here it defaults to Intensity, the value its ConfigFile ctor falls back to (:102).
    DescriptorType descriptorType = DescriptorType::Intensity;
    int device = -1;                // CUDA ordinal (-1: current)
    bool gpuFrontEnd = false;       // addFrame's data association and new-point selection on the device (added option)
    int numPyramidLevels = 1;       // > 1: every window is solved coarse to fine (PhotometricBundleAdjustmentPyr; added option)
    int nGpus = 1;                  // > 1: every window's points are sharded over devices device .. device + nGpus - 1 (added option)
    Options() {}
    Options(const utils::ConfigFile& cf);
  };

  struct Result {
    EigenAlignedContainer_<Mat44> poses;          // refined world poses (whole trajectory so far)
    EigenAlignedContainer_<Vec3> refinedPoints;   // points that left the window, refined
    EigenAlignedContainer_<Vec3> originalPoints;  // the same points as initialised
    double initialCost = -1.0, finalCost = -1.0, fixedCost = -1.0;
    int numSuccessfulStep = 0, numResiduals = 0;
    double totalTime = -1.0;
    std::string message;
    std::vector<ceres::IterationSummary> iterationSummary;
  };

  PhotometricBundleAdjustment(const Calibration&, const ImageSize&, const Options& = Options());
  ~PhotometricBundleAdjustment();
  PhotometricBundleAdjustment(const PhotometricBundleAdjustment&) = delete;
  PhotometricBundleAdjustment& operator=(const PhotometricBundleAdjustment&) = delete;

  // image: rows x cols uint8, depth_map: rows x cols float, T: frame-to-frame pose initialisation
  // (src/photobundle.h:160).  Both buffers are borrowed for the duration of the call.
  void addFrame(const uint8_t* image, const float* depth_map, const Mat44& T, Result* = nullptr);

  // introspection used by the tests (not part of the reference interface)
  size_t numScenePoints() const { return _scene_points.size(); }
  const Trajectory& trajectory() const { return _trajectory; }
  struct PointView { const double* X; const std::vector<uint32_t>* visibility; const std::vector<double>* descriptor; int x, y; double saliency; };
  PointView scenePoint(size_t i) const;

 protected:
  void optimize(Result*);

 private:
  struct ScenePoint;
  struct Frame { uint32_t id; std::vector<uint8_t> image; };
  typedef std::vector<UniquePointer<ScenePoint>> ScenePointPointerList;
  ScenePointPointerList removePointsAtFrame(uint32_t id);

  uint32_t _frame_id = 0;
  Calibration _calib;
  ImageSize _image_size;
  Options _options;
  Trajectory _trajectory;
  std::deque<Frame> _frame_buffer;          // boost::circular_buffer(slidingWindowSize)
  ScenePointPointerList _scene_points;
  std::vector<uint16_t> _mask;
  std::vector<float> _saliency_map;
  Mat33 _K_inv;
  // one pyramid level on the device ([0] = the finest): its own handle, image size and intrinsics; `resident` = the
  // frame id held by each ring slot (id % slidingWindowSize), -1: none
  struct DeviceLevel {
    pba_handle* h = nullptr;           // rank 0 (the only one when nGpus == 1); also the front end's device
    std::vector<pba_handle*> ranks;    // all nGpus handles of the level, ranks[0] == h
    ImageSize size;
    Calibration calib;
    int levels_down = 0, cap_points = 0, cap_obs = 0;
    std::vector<long long> resident;
  };
  std::vector<DeviceLevel> _dev;
  // grow-only page-locked staging for what addFrame sends to / receives from the device every frame (pba_host_alloc)
  struct Pinned {
    void* p = nullptr; size_t bytes = 0;
    void* need(size_t n);
    ~Pinned();
  };
  Pinned _pin_xyz, _pin_patch, _pin_norm, _pin_score, _pin_rc;
  int _desc_type = 0, _n_channels = 1;   // PBA_DESC_* / channels per pixel of the descriptor
  pba_handle* deviceLevel(int level, int n_points, int n_obs);
  void uploadWindowFrames(int level);
  struct SolveOutcome;                   // the C ABI's summary + iteration trace of the last level solved
  void fillResult(Result& out, const SolveOutcome& solved, const ScenePointPointerList& leaving, double seconds) const;
};

// Per-pixel weights of the patch residuals, row-major over (2 radius + 1)^2 (src/photobundle.cc:617-646 with the default
// s_x = s_y = a = 1): all ones, or a normalised Gaussian.  Exposed for the tests against the reference's own function.
std::vector<double> MakePatchWeights(int radius, bool do_gaussian);
// The 5x5 bilinear lookup of the data association before its mean is removed (interp2 + interpolateFixedPatch<2>,
// src/photobundle.cc:258-310), exposed for the same reason.
void InterpPatch5(const uint8_t* I, int rows, int cols, double px, double py, float* out25);

#endif
