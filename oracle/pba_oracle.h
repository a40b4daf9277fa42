/*
 * pba_oracle.h — C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This library restates, on the CPU, the arithmetic
 * of the reference's photometric bundle-adjustment hot path
 * (PhotometricBundleAdjustment::optimize(), src/photobundle.cc:764-876 and
 * everything it calls).  It exists to check the CUDA path and to be timed as the
 * CPU baseline.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product library
 * (libpba_b200.so) never links, loads or calls it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or expected
 * outputs for this path, and it cannot be compiled here (Eigen, Ceres, Boost
 * and OpenCV are absent).  The in-tree arithmetic (residual,
 * sampler, gradient, projection) is restated from the cited source lines; the
 * Ceres 1.x semantics (AutoDiff Jets, HuberLoss corrector, trust-region LM,
 * Schur elimination) are restated from Ceres' published behaviour (pinned
 * version unknown: CMakeLists.txt:35 `find_package(Ceres REQUIRED)`, API use
 * implies <= 1.14).  The one piece of reference code that CAN be compiled
 * here with tiny shim headers — src/sample_eigen.h + src/jet_extras.h — is
 * compiled from where it lies into oracle/_ref/ and the restated sampler is
 * checked against it bit-for-bit (tests/test_oracle_ref_sampler.py).
 */
#ifndef PBA_ORACLE_H
#define PBA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One sliding-window problem, exactly what optimize() holds when it builds
 * the ceres::Problem (src/photobundle.cc:764-806). All arrays are borrowed. */
typedef struct {
  int32_t rows, cols;       /* image size (ImageSize, src/types.h:58-77)                     */
  int32_t n_frames;         /* frames in the window                                           */
  int32_t n_channels;       /* DescriptorFrame::numChannels(), src/photobundle.cc:187         */
  int32_t radius;           /* Options::patchRadius, P = (2r+1)^2                              */
  int32_t n_points;
  int32_t fixed_frame;      /* window-local index of the constant camera (-1: none)           */
  int32_t num_threads;      /* OpenMP threads (<=0: omp default)                              */
  double fx, fy, cx, cy;    /* Calibration, src/calibration.h:22-25                           */
  double huber;             /* Options::robustThreshold; <= 0 -> no loss (photobundle.cc:797) */
  const float* planes;      /* [n_frames][n_channels][rows][cols] fp32 channel planes          */
  const float* grad_x;      /* same shape: imgradient Ix (may be NULL -> computed internally) */
  const float* grad_y;      /* same shape: imgradient Iy                                      */
  const double* weights;    /* [P] patch weights (MakePatchWeights, photobundle.cc:617-644)   */
  const double* desc;       /* [n_points][n_channels*P] reference descriptors                 */
  const int32_t* obs_offsets; /* [n_points+1] CSR offsets into obs_frame                      */
  const int32_t* obs_frame; /* [nnz] window-local frame index of each observation             */
} oracle_problem;

/* Block normal equations + cost at one linearisation point (what Ceres'
 * evaluator + SchurEliminator see), unscaled, after the loss corrector. */
typedef struct {
  double cost;              /* sum_obs 0.5*rho(s)                                              */
  double* U;                /* [n_frames][6][6]  J_c^T J_c  (zero for the fixed frame)         */
  double* gc;               /* [n_frames][6]     J_c^T r                                       */
  double* V;                /* [n_points][3][3]  J_p^T J_p                                     */
  double* gp;               /* [n_points][3]     J_p^T r                                       */
  double* W;                /* [nnz][6][3]       J_c^T J_p (zero for the fixed frame)          */
  double* obs_sqnorm;       /* [nnz] s = ||r||^2 before the loss (may be NULL)                 */
  double* residuals;        /* [nnz][C*P] raw residuals before the loss (may be NULL)          */
} oracle_blocks;

/* Same fields as ceres::IterationSummary (names listed in the reference at
 * src/ceres_cereal.h:12-30) that a trust-region solve fills. */
typedef struct {
  int32_t iteration;
  int32_t step_is_valid;
  int32_t step_is_nonmonotonic;
  int32_t step_is_successful;
  double cost;
  double cost_change;
  double gradient_max_norm;
  double gradient_norm;
  double step_norm;
  double relative_decrease;
  double trust_region_radius;
  double eta;
  double step_size;
  int32_t line_search_function_evaluations;
  int32_t line_search_gradient_evaluations;
  int32_t line_search_iterations;
  int32_t linear_solver_iterations;
  double iteration_time_in_seconds;
  double step_solver_time_in_seconds;
  double cumulative_time_in_seconds;
} oracle_iteration_summary;

typedef struct {
  int32_t max_num_iterations;     /* 500   src/photobundle.cc:751 */
  double function_tolerance;      /* 1e-6  src/photobundle.cc:756 */
  double gradient_tolerance;      /* 1e-6  src/photobundle.cc:757 */
  double parameter_tolerance;     /* 1e-6  src/photobundle.cc:758 */
  double initial_trust_region_radius; /* 1e4 (Ceres default) */
  double max_trust_region_radius;     /* 1e16 */
  double min_trust_region_radius;     /* 1e-32 */
  double min_relative_decrease;       /* 1e-3 */
  double min_lm_diagonal;             /* 1e-6 */
  double max_lm_diagonal;             /* 1e32 */
  int32_t max_num_consecutive_invalid_steps; /* 5 */
  int32_t jacobi_scaling;             /* 1 */
  int32_t use_autodiff;               /* 1: Jet<double,9> (reference structure); 0: analytic Jacobian */
} oracle_solver_options;

typedef struct {
  double initial_cost, final_cost, fixed_cost;
  int32_t num_successful_steps, num_unsuccessful_steps;
  int32_t num_residuals, num_residual_blocks;
  int32_t num_iterations;        /* entries written to the trace (incl. iteration 0) */
  int32_t termination_type;      /* 0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE */
  int32_t num_jacobian_evals, num_cost_evals;
  double total_time_in_seconds;
  double jacobian_time_in_seconds, cost_time_in_seconds, linear_solver_time_in_seconds;
  char message[256];
} oracle_summary;

void oracle_default_options(oracle_solver_options* o);

/* src/imgproc.cc:27-106 (central difference * 0.5, zero on all four borders). */
void oracle_imgradient(const float* I, int32_t rows, int32_t cols, float* gx, float* gy);

/* Descriptor channels (DescriptorFrame::Create, src/photobundle.cc:220-248): type 0 Intensity (1 plane),
 * 1 IntensityAndGradient (3), 2 BitPlanes (8; computeBitPlanes src/imgproc.cc:222-245 with the two
 * cv::GaussianBlur calls restated, pinned against OpenCV by tests/golden/bitplanes_ref.npz).
 * planes: dense [C][rows][cols]. */
int32_t oracle_descriptor_channels(int32_t type);
void oracle_build_channels(int32_t type, const uint8_t* img, int32_t rows, int32_t cols, float* planes);
void oracle_bitplanes_stages(const uint8_t* img, int32_t rows, int32_t cols, uint8_t* blur_out, uint8_t* census_out);
/* DescriptorFrame::computeSaliencyMap (src/photobundle.cc:212-220) and ExtractPatch (:466-479). */
void oracle_saliency_map(const float* planes, int32_t n_channels, int32_t rows, int32_t cols, float* out);
void oracle_extract_patches(const float* planes, int32_t n_channels, int32_t rows, int32_t cols, int32_t radius, int32_t n,
                            const int32_t* xy, double* desc);

/* src/sample_eigen.h:33-102. out = {I, Gx, Gy} at (x, y). */
void oracle_sample_linear(const float* I, const float* Gx, const float* Gy,
                          int32_t rows, int32_t cols, float y, float x, float* out3);

/* src/photobundle.cc:617-644. */
/* Calibration::project as the residual functor instantiates it (src/calibration.h:33-38): k4 = {fx, fy, cx, cy} */
void oracle_project(const double* k4, const double* X, double* uv);
void oracle_patch_weights(int32_t radius, int32_t do_gaussian, double* w);

/* src/photobundle.cc:646-667 with ceres::RotationMatrixToAngleAxis /
 * AngleAxisToRotationMatrix restated. T is a column-major 4x4. */
void oracle_pose_to_params(const double* T44, double* p6);
void oracle_params_to_pose(const double* p6, double* T44);
void oracle_angle_axis_rotate_point(const double* aa, const double* pt, double* out);

/* One residual block = DescriptorError::operator() (src/photobundle.cc:696-727)
 * evaluated through AutoDiffCostFunction<.., DYNAMIC, 6, 3> (use_autodiff=1) or
 * through the closed-form Jacobian (use_autodiff=0). r: [C*P]; Jc: [C*P][6];
 * Jp: [C*P][3] (row-major, may be NULL). Raw: no loss applied. */
void oracle_residual_block(const oracle_problem* pb, int32_t frame, const double* cam6,
                           const double* xyz, const double* desc, int32_t use_autodiff,
                           double* r, double* Jc, double* Jp);

/* Cost only (T = double path). Returns sum 0.5*rho(||r||^2). */
double oracle_cost(const oracle_problem* pb, const double* cams, const double* points);

/* Residuals + Jacobian -> block normal equations. */
void oracle_evaluate(const oracle_problem* pb, const double* cams, const double* points,
                     int32_t use_autodiff, oracle_blocks* out);

/* ceres::Solve with the reference's options (src/photobundle.cc:738-761, :829).
 * cams [n_frames][6] and points [n_points][3] are updated in place.
 * trace: room for max_num_iterations+1 entries (may be NULL). */
int32_t oracle_solve(const oracle_problem* pb, const oracle_solver_options* opt,
                     double* cams, double* points, oracle_summary* summary,
                     oracle_iteration_summary* trace);

#ifdef __cplusplus
}
#endif
#endif
