/*
 * pba_b200.h — C ABI of the B200-native photometric bundle-adjustment inner loop.
 *
 * This is the drop-in boundary for ONE path of halismai/photobundle:
 * PhotometricBundleAdjustment::optimize() (reference src/photobundle.cc:764-876), i.e.
 * what the reference does between "the window's frames, poses and points are known"
 * and "ceres::Solve returned".  Everything the reference evaluates per LM iteration —
 * DescriptorError::operator() (src/photobundle.cc:696-727), the bilinear sampler
 * (src/sample_eigen.h:33-126), the Jet chain rule (src/jet_extras.h:74-111), the
 * central-difference gradient planes (src/imgproc.cc:27-106), Ceres' Huber corrector,
 * Schur elimination, reduced-camera solve and trust-region bookkeeping — runs in
 * hand-written sm_100a kernels behind these entry points.
 *
 * Plain pointers and sizes only; every function returns 0 on success or a negative
 * pba_status, and pba_last_error() gives the message (the reference throws
 * std::runtime_error / calls Fatal(); the C++ class in photobundle_b200/host translates
 * a non-zero status into std::runtime_error).  There is NO CPU fallback: without a
 * CUDA device every compute entry point fails with PBA_ERR_CUDA.
 *
 * Threading: a handle is not thread-safe (neither is the reference object); each
 * handle owns one CUDA stream; calls return after the stream has drained.
 */
#ifndef PBA_B200_H
#define PBA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBA_MAX_FRAMES 16          /* frames per window (reference default 5; BASELINE 8 / 16) */
#define PBA_MAX_RADIUS 4           /* patch radius (reference default 2; KITTI cfg 1)           */
#define PBA_MAX_CHANNELS 8         /* 1 Intensity, 3 IntensityAndGradient, 8 BitPlanes          */
#define PBA_UNIQUE_ID_BYTES 128    /* sizeof(ncclUniqueId)                                      */

typedef enum {
  PBA_OK = 0,
  PBA_ERR_ARGUMENT = -1,
  PBA_ERR_CUDA = -2,
  PBA_ERR_NCCL = -3,
  PBA_ERR_STATE = -4,
  PBA_ERR_CAPACITY = -5
} pba_status;

typedef struct pba_handle pba_handle;

/* Replaces: PhotometricBundleAdjustment ctor arguments (Calibration, ImageSize, Options;
 * src/photobundle.h:148, :26-85) as far as optimize() reads them. */
typedef struct {
  int32_t rows, cols;          /* ImageSize (src/types.h:58-77)                                  */
  int32_t n_channels;          /* DescriptorFrame::numChannels() (src/photobundle.cc:187)        */
  int32_t patch_radius;        /* Options::patchRadius                                           */
  int32_t max_frames;          /* Options::slidingWindowSize (capacity)                          */
  int32_t max_points;          /* capacity of the point arrays                                   */
  int32_t max_observations;    /* capacity of the observation arrays (<= max_points*max_frames)  */
  int32_t device;              /* CUDA ordinal; -1 = current device                              */
  double fx, fy, cx, cy;       /* Calibration (src/calibration.h:22-25)                          */
  double huber;                /* Options::robustThreshold; <= 0 disables the loss (:797-798)    */
} pba_config;

/* Replaces: GetSolverOptions (src/photobundle.cc:738-761) + the Ceres defaults the
 * reference leaves untouched. pba_default_solver_options() fills
 * exactly those values. */
typedef struct {
  int32_t max_num_iterations;          /* 500  */
  double function_tolerance;           /* 1e-6 */
  double gradient_tolerance;           /* 1e-6 */
  double parameter_tolerance;          /* 1e-6 */
  double initial_trust_region_radius;  /* 1e4  */
  double max_trust_region_radius;      /* 1e16 */
  double min_trust_region_radius;      /* 1e-32 */
  double min_relative_decrease;        /* 1e-3 */
  double min_lm_diagonal;              /* 1e-6 */
  double max_lm_diagonal;              /* 1e32 */
  int32_t max_num_consecutive_invalid_steps; /* 5 */
  int32_t jacobi_scaling;              /* 1 */
} pba_solver_options;

/* Same field names as ceres::IterationSummary, which Result::iterationSummary holds
 * (src/photobundle.h:117; field list src/ceres_cereal.h:12-30). */
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
} pba_iteration_summary;

/* Replaces: the ceres::Solver::Summary fields optimize() copies into Result
 * (src/photobundle.cc:867-874). */
typedef struct {
  double initial_cost, final_cost, fixed_cost;
  int32_t num_successful_steps, num_unsuccessful_steps;
  int32_t num_residuals, num_residual_blocks;
  int32_t num_iterations;          /* entries available from pba_get_iterations()            */
  int32_t termination_type;        /* 0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE             */
  int32_t num_evaluations;         /* residual+Jacobian passes over all observations (K1)    */
  int32_t kernel_launches;         /* CUDA kernels launched by this solve                    */
  int32_t num_collectives;         /* exchanges between the GPUs of a window: NCCL all-reduces, or in-kernel
                                      peer-memory exchanges (0 on one GPU)                  */
  double total_time_in_seconds;    /* host wall time of the call                             */
  double device_time_in_seconds;   /* CUDA-event time on the handle's stream                 */
  double kb_device_time_in_seconds;/* of which inside K_B (decision + Schur + reduced solve): sum over the LM
                                      iterations of the kernel's own start/end stamps (globaltimer)   */
  char message[256];
} pba_summary;

/* Output of one residual + Jacobian + block-accumulation pass (K1), unscaled, after the
 * Huber corrector: what Ceres' evaluator + SchurEliminator would see at the current
 * poses/points.  Any pointer may be NULL. Layouts are row-major and dense. */
typedef struct {
  double cost;              /* sum over observations of 0.5*rho(||r||^2)                     */
  double* U;                /* [n_frames][6][6]  J_c^T J_c (zero for the fixed frame)        */
  double* gc;               /* [n_frames][6]     J_c^T r                                     */
  double* V;                /* [n_points][3][3]  J_p^T J_p                                   */
  double* gp;               /* [n_points][3]     J_p^T r                                     */
  double* W;                /* [n_obs][6][3]     J_c^T J_p (zero for the fixed frame)        */
  double* obs_sqnorm;       /* [n_obs]           ||r||^2 before the loss                     */
  double* residuals;        /* [n_obs][C*P]      raw residuals w_j*(p0_i - I(u+x, v+y))      */
  double device_ms;         /* CUDA-event time of the pass                                   */
} pba_eval_out;

const char* pba_last_error(void);
const char* pba_version(void);
void pba_default_solver_options(pba_solver_options* o);

int pba_create(const pba_config* cfg, pba_handle** out);
void pba_destroy(pba_handle* h);

/* Frames of the window. Replaces DescriptorFrame::Create for the Intensity descriptor
 * (src/photobundle.cc:225-232: uint8 -> float cast) plus ImageGradient::compute
 * (:127-135 -> imgradient): the uint8 plane is uploaded as is and the kernel forms the
 * fp32 intensity and the central-difference gradients (zero-border rule) on the fly.
 * images[f] is a dense row-major rows x cols plane, borrowed for the call. */
int pba_set_frames_u8(pba_handle* h, int32_t n_frames, const uint8_t* const* images);
/* Generic multi-channel descriptors (IntensityAndGradient / BitPlanes): fp32 channel
 * planes as DescriptorFrame holds them; planes[f*n_channels + k]. */
int pba_set_frames_f32(pba_handle* h, int32_t n_frames, const float* const* planes);
/* Pyramid level l of a window (BASELINE config 4): `images` are the LEVEL-0 frames of size
 * src_rows x src_cols; they are uploaded once and reduced `levels_down` times on the device with
 * the cv::pyrDown rule the reference intends at src/photobundle_pyramid.cc:46 (separable
 * [1 4 6 4 1]/16, reflect-101 borders, even samples, size (n+1)/2 per src/types.h:70-73).  The
 * handle must have been created with the level's rows/cols and the level's intrinsics
 * (Calibration::pyrDown, src/calibration.h:72-78). */
int pba_set_frames_u8_pyr(pba_handle* h, int32_t n_frames, const uint8_t* const* images, int32_t src_rows,
                          int32_t src_cols, int32_t levels_down);
/* The same reduction for one image, host to host (rows x cols -> (rows+1)/2 x (cols+1)/2). */
int pba_pyrdown_u8(const uint8_t* src, int32_t rows, int32_t cols, uint8_t* dst, int32_t device);
/* Sliding window: replace the plane(s) in ring slot `slot` only. */
int pba_set_frame_u8(pba_handle* h, int32_t slot, const uint8_t* image);

/* Replaces: camera_params (src/photobundle.cc:774-778): per frame [angle-axis(3), t(3)]
 * of the world->camera transform; fixed_frame = index held constant (:809-816), -1 none. */
int pba_set_poses(pba_handle* h, int32_t n_frames, const double* cam6, int32_t fixed_frame);

/* Replaces: the residual-block loop (src/photobundle.cc:786-806). xyz [n][3] world points
 * (ScenePoint::_X), desc [n][C*P] reference descriptors (channel-major, then row-major),
 * CSR visibility obs_offsets [n+1] / obs_frame [nnz] (window-local frame index), weights
 * [P] (MakePatchWeights, :617-644). With n_ranks > 1 every rank passes the FULL arrays;
 * the library keeps its own contiguous shard (balanced by observation count). */
int pba_set_points(pba_handle* h, int32_t n_points, const double* xyz, const double* desc,
                   const int32_t* obs_offsets, const int32_t* obs_frame, const double* weights);

/* K1 only, at the current poses/points (BASELINE config 2 and the parity tests). */
int pba_eval(pba_handle* h, pba_eval_out* out);
/* K1 launched `iters` times back to back on the handle's stream, CUDA-event timed
 * (inputs resident in HBM). ms_total = elapsed for all launches. */
int pba_eval_timed(pba_handle* h, int32_t iters, double* ms_total);

/* Replaces: ceres::Solve(GetSolverOptions(...), &problem, &summary) (src/photobundle.cc:829).
 * Poses and points are updated on the device; read them with pba_get_*. */
int pba_solve(pba_handle* h, const pba_solver_options* opt, pba_summary* summary);

/* Device-resident snapshot of the current poses and points (the x that pba_solve starts from),
 * and its restoration: lets a caller re-solve the same window without host->device copies
 * (bench.py's HBM-resident `value` leg; also what a pyramid level-to-level hand-over uses). */
int pba_save_state(pba_handle* h);
int pba_restore_state(pba_handle* h);

/* Device-to-device hand-over of the current poses and points from one handle to another with the same frame and
 * point counts (and the same sharding): the coarse-to-fine step of a pyramid (what the reference intends at
 * src/photobundle_pyramid.cc:61-65, T_init of the finer level = the coarser level's result) without a host round
 * trip.  Ordered on dst's stream. */
int pba_copy_state(pba_handle* dst, pba_handle* src);

/* Page-locked host memory for buffers that are passed to pba_set_* / pba_associate / ... on every frame: copies from
 * (and to) such buffers run at the full rate of the host link, pageable memory goes through the driver's staging
 * buffers at a third of it.  Optional - every entry point accepts ordinary memory.  NULL on failure. */
void* pba_host_alloc(size_t bytes);
void pba_host_free(void* p);

/* Diagnostics: how often this handle instantiated its LM-loop graph and how often it re-targeted the instantiated
 * graph in place (cudaGraphExecKernelNodeSetParams) because only sizes had changed since the previous solve - what a
 * sliding window does on every frame. */
int pba_graph_counters(const pba_handle* h, int32_t* builds, int32_t* updates);

/* Uploads without a host round trip per call.  Between pba_begin_batch() and the next pba_solve() (or pba_end_batch())
 * the pba_set_* calls only ENQUEUE their host->device copies; the caller keeps every buffer it passed alive and
 * unchanged until that pba_solve() / pba_end_batch() returns (at most one pba_set_points per batch).  A sliding window
 * changes one frame per solve: pba_set_frame_u8_ex() replaces the frame in ring slot `slot` (on a handle with no window
 * yet the ring has cfg.max_frames slots and unwritten slots read as black frames, so a caller may feed frames one by
 * one from the first addFrame on) - the level-0 uint8 image is reduced `levels_down` times on the device (cv::pyrDown rule) and / or expanded into the channels of
 * `descriptor_type`.  pba_get_results() reads poses and points back with one synchronisation.
 * (what the reference does per frame: DescriptorFrame::Create + ring buffer push, src/photobundle.cc:495, :608) */
int pba_begin_batch(pba_handle* h);
int pba_end_batch(pba_handle* h);
int pba_set_frame_u8_ex(pba_handle* h, int32_t slot, const uint8_t* image, int32_t src_rows, int32_t src_cols, int32_t levels_down,
                        int32_t descriptor_type);
int pba_get_results(pba_handle* h, double* cam6, double* xyz);

int pba_get_poses(pba_handle* h, double* cam6);
int pba_get_points(pba_handle* h, double* xyz);
int pba_get_iterations(pba_handle* h, pba_iteration_summary* out, int32_t capacity, int32_t* n);

/* ---- multi-channel descriptors, built on the device --------------------------------
 * DescriptorFrame::Create (src/photobundle.cc:220-248): Intensity = the uint8 image cast to float (1
 * channel, pba_set_frames_u8); IntensityAndGradient = {I, Ix, Iy} (imgradient, src/imgproc.cc:27-106);
 * BitPlanes = 8 blurred bit planes of the census transform of the pre-blurred image (computeBitPlanes,
 * src/imgproc.cc:222-245).  The handle must have been created with pba_descriptor_channels(type)
 * channels. */
#define PBA_DESC_INTENSITY 0
#define PBA_DESC_INTENSITY_AND_GRADIENT 1
#define PBA_DESC_BITPLANES 2
int pba_descriptor_channels(int32_t descriptor_type);   /* 1, 3, 8; -1 for an unknown type */
/* Upload the window's uint8 images and build every frame's channel planes on the device. */
int pba_set_frames_u8_descriptor(pba_handle* h, int32_t n_frames, const uint8_t* const* images, int32_t descriptor_type);
/* One channel plane of a window frame, dense rows x cols floats (tests / inspection). */
int pba_get_channel_plane(pba_handle* h, int32_t frame, int32_t channel, float* out);
/* The caller side of the path (addFrame, src/photobundle.cc:482-615) needs, for the NEW frame only, the
 * saliency map (DescriptorFrame::computeSaliencyMap, :212-220: sum over channels of |Ix| + |Iy|) and the
 * reference descriptors of the points it selects (ExtractPatch, :466-479).  pba_prepare_frame_u8 builds
 * the channels of one image in a scratch area of the handle; the two calls below read from it. */
int pba_prepare_frame_u8(pba_handle* h, const uint8_t* image, int32_t descriptor_type);
int pba_saliency_map(pba_handle* h, float* out /* rows x cols */);
int pba_extract_descriptors(pba_handle* h, int32_t n, const int32_t* xy /* n x {x, y} */, double* desc /* n x C*P */);

/* ---- addFrame's front end on the device, all on the frame of pba_prepare_frame_u8 ------
 * Data association (src/photobundle.cc:508-542): for each of n live points uv = normHomog(K (T_c X)); when
 * the rounded pixel (row, col) lies inside the border band the point is tested: score = ZNCC
 * (ZnccPatch_<2,float>, :315-361) of its stored mean-free 5x5 patch (ref_patch, ref_norm) against the patch
 * interpolated at uv (interp2, :262-294).  score = -2: not tested.  T_c: column-major 4x4 world->camera,
 * K: row-major 3x3.  Results are bit-identical to the host arithmetic of photobundle_b200/host. */
int pba_associate(pba_handle* h, int32_t n, const double* xyz, const float* ref_patch, const float* ref_norm,
                  const double* T_c, const double* K, int32_t border, float* score, int32_t* row_col);
/* New-point candidates (:545-575): pixels inside the border band with min_depth <= depth <= max_depth that
 * are strict local maxima of the saliency map over (2 nms_radius + 1)^2 (IsLocalMax_, src/imgproc.h:175-212)
 * and not masked by a (2 mask_radius + 1)^2 block around a re-observed point.  Returned in scan order
 * (row-major); *n_out may exceed capacity (then only the first `capacity` of an arbitrary order were
 * stored and the call should be repeated with more room).  `depth` may be NULL: then no depth test is made on the
 * device (the depth map, 4 bytes per pixel, does not cross the host link) and the caller drops the candidates whose
 * depth is out of range - the same set in the same order, since the test is per pixel. */
int pba_select_candidates(pba_handle* h, const float* depth, int32_t n_masked, const int32_t* masked_row_col,
                          int32_t mask_radius, int32_t nms_radius, int32_t border, double min_depth, double max_depth,
                          int32_t capacity, int32_t* cand_row_col, float* cand_saliency, int32_t* n_out);

/* Multi-GPU: one process per GPU; points are sharded, frames/poses replicated, one
 * exchange of the pose blocks and of the reduced camera system per LM iteration.  Rank 0 calls
 * pba_comm_unique_id() and distributes the 128 bytes by any means (torch.distributed,
 * MPI, a file); then every rank calls pba_comm_init(). */
int pba_comm_unique_id(void* id128);
/* The sharding rule (host arithmetic only, no GPU): points [first, last) belong to `rank`. */
int pba_shard_range(int32_t n_points, const int32_t* obs_offsets, int32_t rank, int32_t n_ranks,
                    int32_t* first, int32_t* last);
int pba_comm_init(pba_handle* h, const void* id128, int32_t rank, int32_t n_ranks);
/* The same communicator for `n` handles of ONE process, each on its own device (rank = position in `handles`):
 * no NCCL, no CUDA IPC - peer access is enabled between the devices and the kernels address the peers' exchange
 * buffers directly.  Every member then receives the same pba_set_* calls, and pba_solve() must be called on all
 * members CONCURRENTLY, one host thread per handle (the call blocks while its device waits for the peers; a member
 * whose peers do not show up within 60 s fails with PBA_ERR_STATE).  pba_get_points / pba_get_results gather the
 * shards from the peers and may be called on any member once every pba_solve() has returned.  Destroy the members
 * together.  This is how the C++ class drives Options::nGpus devices behind an unchanged addFrame() (the reference
 * has no counterpart: apps/run_kitti.cc:47 stays as it is). */
int pba_comm_init_local(pba_handle* const* handles, int32_t n);
/* How the per-iteration sums travel between the GPUs of a window: PBA_EXCHANGE_NONE (one GPU),
 * PBA_EXCHANGE_PEER (default: every rank stores its pose blocks / reduced-system contribution
 * straight into all peers' memory over NVLink from inside the kernels and the consumers sum the
 * slots in rank order - no collective call, no extra launch), or PBA_EXCHANGE_NCCL (all-reduce
 * fallback when peer mappings are unavailable or PBA_MGPU_EXCHANGE=nccl). */
/* 1 if the window passed to the last pba_set_points is sharded over the communicator's ranks, 0 if every rank holds
 * (and solves) the whole window.  A window whose points all fit into one wave of K_A on one GPU (<= 28 warps x SM
 * count = 4 144 points on a B200) is NOT sharded: sharding cannot shorten its LM iteration - K_A is that single wave's
 * latency, the reduced solve is replicated anyway - and every exchange over NVLink costs more than it saves (measured:
 * 8 x 4 000 points, 0.69 ms on one GPU, 0.82 / 0.84 / 0.93 ms sharded over 2 / 4 / 8).  Every rank then runs the
 * single-GPU path on the full window; the ranks' results agree to rounding (fp64 atomics are unordered), no exchange
 * takes place and pba_summary.num_collectives is 0.  PBA_MGPU_REPLICATE=0 / 1 forces sharding / replication. */
int pba_comm_sharded(const pba_handle* h);
/* With the peer-memory exchange, 1 if K_B eliminates the point blocks under both outcomes of the pending trust-region
 * decision while the evaluation sums travel (ONE exchange per LM iteration, a second one only when neither outcome
 * holds); 0 if the shard is large enough that doubling the elimination costs more than a second exchange: then the
 * decision is exchanged first and every CTA eliminates once (TWO exchanges per iteration).  Valid after
 * pba_set_points; PBA_MGPU_SPECULATE=0/1 overrides the size rule. */
int pba_comm_speculates(const pba_handle* h);
#define PBA_EXCHANGE_NONE 0
#define PBA_EXCHANGE_PEER 1
#define PBA_EXCHANGE_NCCL 2
int pba_comm_exchange_kind(const pba_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* PBA_B200_H */
