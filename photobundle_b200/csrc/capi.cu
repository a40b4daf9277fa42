// capi.cu — the extern "C" shim (include/pba_b200.h) over the sm_100a kernels.
// Host side of the drop-in boundary: owns device memory, one CUDA stream per handle, and
// enqueues the LM loop.  No CPU fallback: every compute entry point needs a CUDA device.

#include "../../include/pba_b200.h"
#include "pba_device.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <nccl.h>   // types/enums only: the library is dlopen'ed when a communicator is requested

using namespace pba;

// ---- NCCL, loaded lazily (single-GPU users never need it) -----------------------------------
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static const char* load_nccl() {
  if (g_nccl.lib) return nullptr;
  // if the host process (e.g. PyTorch) already carries an NCCL, dlopen returns that same copy
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
  if (!g_nccl.lib) return "libnccl.so.2 not found";
#define PBA_NCCL_SYM(field, sym) \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, sym); if (!g_nccl.field) return "missing NCCL symbol " sym;
  PBA_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  PBA_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  PBA_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  PBA_NCCL_SYM(AllReduce, "ncclAllReduce")
  PBA_NCCL_SYM(Broadcast, "ncclBroadcast")
  PBA_NCCL_SYM(AllGather, "ncclAllGather")
  PBA_NCCL_SYM(GroupStart, "ncclGroupStart")
  PBA_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  PBA_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef PBA_NCCL_SYM
  return nullptr;
}

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define NCCL_TRY(expr)                                                                     \
  do {                                                                                     \
    ncclResult_t r__ = (expr);                                                             \
    if (r__ != ncclSuccess)                                                                \
      return fail(PBA_ERR_NCCL, "%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(r__), __FILE__, __LINE__); \
  } while (0)

#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return fail(PBA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// A communicator whose ranks are handles of ONE process (pba_comm_init_local): the exchange buffers of the peers are
// addressed directly (cudaDeviceEnablePeerAccess), and the threads that call pba_solve meet at a host barrier before
// the first kernel that waits on a peer is launched (allocations made while a peer spins could otherwise block).
struct LocalGroup {
  std::mutex m;
  std::condition_variable cv;
  std::vector<pba_handle*> members;
  int arrived = 0;
  unsigned long long generation = 0;
  bool arrive_and_wait(int seconds) {
    std::unique_lock<std::mutex> lk(m);
    const unsigned long long gen = generation;
    if (++arrived == (int)members.size()) { arrived = 0; ++generation; cv.notify_all(); return true; }
    const bool ok = cv.wait_for(lk, std::chrono::seconds(seconds), [&] { return generation != gen; });
    if (!ok) --arrived;
    return ok;
  }
};

constexpr int kSCopies = 2 * kSReplicas > 3 ? 2 * kSReplicas : 3;   // one GPU: two alternating sets of kSReplicas copies; multi-GPU: three accumulators

struct pba_handle {
  pba_config cfg;
  int device = 0, sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int P = 0, CP = 0, pitch = 0;
  size_t plane = 0;
  // device buffers
  uint8_t* d_u8 = nullptr;
  uint8_t* d_stage_u8 = nullptr;   // dense [F][rows][cols] landing zone of the host->device copies (re-pitched on the device)
  float* d_f32 = nullptr;
  // descriptor construction (pba_set_frames_u8_descriptor / pba_prepare_frame_u8)
  uint8_t *d_scr_a = nullptr, *d_scr_b = nullptr;   // pitched uint8 scratch planes (pre-blur, census)
  uint8_t* d_new_u8 = nullptr;     // the new frame of pba_prepare_frame_u8 (pitched) ...
  float* d_new_planes = nullptr;   // ... and its channel planes [C][rows][pitch]
  int new_channels = 0, new_active = 0;   // planes allocated / channels of the prepared frame
  float* d_sal = nullptr;          // saliency map, dense
  int* d_xy = nullptr;  double* d_desc_out = nullptr;  int xy_cap = 0;
  // addFrame front end (pba_associate / pba_select_candidates)
  double* d_as_xyz = nullptr;  float *d_as_patch = nullptr, *d_as_norm = nullptr, *d_as_score = nullptr;  int* d_as_rc = nullptr;
  double* d_as_mats = nullptr;   // T_c (16) | K (9)
  int as_cap = 0;
  uint8_t* d_mask = nullptr;  float* d_depth = nullptr;  int *d_cand_rc = nullptr, *d_count = nullptr, *d_masked_rc = nullptr;
  float* d_cand_sal = nullptr;  int cand_cap = 0, masked_cap = 0;
  float* p_desc = nullptr;         // pinned staging of pba_set_points (descriptors as float, local offsets)
  int32_t* p_off = nullptr;
  bool frames_are_u8 = false;
  double *d_cams = nullptr, *d_pts = nullptr, *d_weights = nullptr;
  float* d_desc = nullptr;
  int *d_obs_off = nullptr, *d_obs_frame = nullptr;
  double *d_V = nullptr, *d_gp = nullptr, *d_W = nullptr;
  double *d_Xacc = nullptr, *d_Ucur = nullptr;
  double *d_scale_p = nullptr, *d_Vinv = nullptr, *d_S = nullptr;
  double* d_Vinv2 = nullptr;       // multi-GPU: (Vs + D²)^-1 under the two hypotheses of the pending decision, [2][n][6]
  unsigned int* d_ticket = nullptr;
  unsigned long long* d_dbg = nullptr;   // PBA_DEBUG_TIMELINE=1: per-iteration K_B timeline
  unsigned long long* d_stamps = nullptr;   // [trace_cap][2] start/end of every K_B launch (pba_summary.kb_device_time_in_seconds)
  double *d_obs_sqnorm = nullptr, *d_residuals = nullptr;
  double *d_save_cams = nullptr, *d_save_pts = nullptr;
  uint8_t* d_pyr_scratch = nullptr;  size_t pyr_scratch_bytes = 0;   // pba_set_frames_u8_pyr ping-pong planes
  double* d_pts_full = nullptr;      // multi-GPU: all shards gathered (pba_get_points)
  bool have_saved = false;
  bool defer_sync = false;         // between pba_begin_batch and the next pba_solve / pba_end_batch: uploads do not block
  LmState* d_state = nullptr;
  IterSummary* d_trace = nullptr;
  int trace_cap = 0;
  LmState* h_state = nullptr;  // pinned
  // pinned landing zone of everything a solve reports, filled by copies enqueued behind the LM loop (one synchronisation):
  // iteration trace | K_B stamps | exchange error word | cameras | points (this rank's)
  unsigned char* h_post = nullptr;
  size_t post_trace = 0, post_stamps = 0, post_err = 0, post_cams = 0, post_pts = 0, post_bytes = 0;
  bool results_cached = false;   // h_post holds the poses / points of the last solve and nothing has changed them since
  // problem
  int n_frames = 0, fixed_frame = -1, n_points = 0, nnz = 0;
  bool have_frames = false, have_poses = false, have_points = false;
  std::vector<int> frame_used;
  int max_obs_frame = -1;          // largest frame index any observation refers to (re-checked against n_frames before a launch)
  std::vector<IterSummary> trace;
  // multi-GPU: points sharded by contiguous block, frames/poses replicated
  int rank = 0, n_ranks = 1;
  // A window whose points all fit into ONE wave of K_A on one GPU is not sharded even when a communicator exists:
  // sharding cannot shorten its iteration (K_A is that one wave's latency, the reduced solve is replicated anyway) and
  // every exchange costs more than it saves - each rank then solves the whole window (`replicated`).  eff_ranks /
  // eff_xchg are what the solve path looks at: the communicator's values, or 1 / false for a replicated window.
  bool replicated = false;
  int eff_ranks = 1;
  bool eff_xchg = false;
  ncclComm_t comm = nullptr;
  std::shared_ptr<struct LocalGroup> group;   // set by pba_comm_init_local: the ranks are handles of this process
  int n_points_total = 0, nnz_total = 0;
  std::vector<int> shard_begin;   // [n_ranks+1] first global point of each rank
  // per-iteration exchange over NVLink peer memory (CUDA IPC), see pba_device.cuh `Xchg`; when it cannot
  // be set up (no peer access, PBA_MGPU_EXCHANGE=nccl) the NCCL all-reduce path is used instead
  bool use_xchg = false;
  void* d_xchg = nullptr;                       // this rank's exchange buffer
  void* peer_xchg[kMaxRanks] = {};              // every rank's buffer as mapped here ([rank] = d_xchg)
  Xchg xc = {};
  unsigned long long epoch_next = 1;
  // single-GPU LM loop as one CUDA graph: a WHILE conditional node whose body is two LM iterations
  // (K_B, K_A, K_B, K_A); K_B clears the condition when the minimizer terminates
  cudaGraph_t lm_graph = nullptr;
  cudaGraphExec_t lm_exec = nullptr;
  std::vector<unsigned char> lm_key;   // kernel parameters the instantiated graph was built from
  std::vector<cudaGraphNode_t> lm_body_nodes;   // the body's four kernel nodes in launch order (parameter updates in place)
  unsigned long long lm_cond = 0;      // the WHILE node's condition handle
  int lm_updates = 0, lm_builds = 0;   // how often the instantiated graph was updated in place / rebuilt
};

static void drop_lm_graph(pba_handle* h) {
  if (h->lm_exec) cudaGraphExecDestroy(h->lm_exec);
  if (h->lm_graph) cudaGraphDestroy(h->lm_graph);
  h->lm_exec = nullptr; h->lm_graph = nullptr; h->lm_key.clear(); h->lm_body_nodes.clear();
}

static void free_all(pba_handle* h) {
  if (getenv("PBA_DEBUG_GRAPH")) fprintf(stderr, "[pba graph] LM loop graph: %d builds, %d in-place parameter updates\n", h->lm_builds, h->lm_updates);
  drop_lm_graph(h);
  cudaFree(h->d_u8); cudaFree(h->d_stage_u8); cudaFree(h->d_f32);
  cudaFree(h->d_scr_a); cudaFree(h->d_scr_b); cudaFree(h->d_new_u8); cudaFree(h->d_new_planes); cudaFree(h->d_sal);
  cudaFree(h->d_xy); cudaFree(h->d_desc_out);
  cudaFree(h->d_as_xyz); cudaFree(h->d_as_patch); cudaFree(h->d_as_norm); cudaFree(h->d_as_score); cudaFree(h->d_as_rc);
  cudaFree(h->d_as_mats); cudaFree(h->d_mask); cudaFree(h->d_depth); cudaFree(h->d_cand_rc); cudaFree(h->d_count);
  cudaFree(h->d_masked_rc); cudaFree(h->d_cand_sal);
  if (h->p_desc) cudaFreeHost(h->p_desc);
  if (h->p_off) cudaFreeHost(h->p_off); cudaFree(h->d_cams); cudaFree(h->d_pts); cudaFree(h->d_weights);
  cudaFree(h->d_desc); cudaFree(h->d_obs_off); cudaFree(h->d_obs_frame); cudaFree(h->d_V); cudaFree(h->d_gp);
  cudaFree(h->d_W); cudaFree(h->d_Xacc); cudaFree(h->d_Ucur);
  for (int q = 0; q < kMaxRanks; ++q)
    if (!h->group && h->peer_xchg[q] && h->peer_xchg[q] != h->d_xchg) cudaIpcCloseMemHandle(h->peer_xchg[q]);
  if (h->group) {   // leave the group: the remaining members must not be solved again (their peers' buffers are gone)
    std::lock_guard<std::mutex> lk(h->group->m);
    auto& mem = h->group->members;
    mem.erase(std::remove(mem.begin(), mem.end(), h), mem.end());
  }
  cudaFree(h->d_xchg);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  cudaFree(h->d_scale_p); cudaFree(h->d_Vinv); cudaFree(h->d_S); cudaFree(h->d_Vinv2); cudaFree(h->d_ticket); cudaFree(h->d_dbg); cudaFree(h->d_stamps);
  cudaFree(h->d_save_cams); cudaFree(h->d_save_pts); cudaFree(h->d_pyr_scratch); cudaFree(h->d_pts_full);
  cudaFree(h->d_obs_sqnorm); cudaFree(h->d_residuals); cudaFree(h->d_state); cudaFree(h->d_trace);
  if (h->h_state) cudaFreeHost(h->h_state);
  if (h->h_post) cudaFreeHost(h->h_post);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
}


// The exchange buffer of one rank: 16-byte LL cells for exchange 1 (evaluation + P under both hypotheses) and exchange 2
// (P again after a mis-speculation), double-buffered by epoch parity, one region per source rank; then the arrival
// flags, the verdict word and the error word.
struct XchgLayout { size_t x1_n, x2_n, n_cells, bytes; };
static XchgLayout xchg_layout(int max_frames, int n) {
  const size_t F = (size_t)max_frames, D = 6 * F;
  const size_t xa_n = F * kUStride + kEacc + kMaxRanks, nn = D * (D + 1);
  XchgLayout l;
  l.x1_n = xa_n + 2 * nn; l.x2_n = nn;
  l.n_cells = 2 * (size_t)n * (l.x1_n + l.x2_n);
  l.bytes = sizeof(ulonglong2) * l.n_cells + sizeof(unsigned long long) * ((size_t)n + 2) + 64;
  return l;
}

// h->peer_xchg[0..n) are mapped: fill the kernel-side descriptor and switch the handle to the peer-memory exchange
static int adopt_xchg(pba_handle* h, const XchgLayout& lay) {
  const int n = h->n_ranks;
  Xchg& x = h->xc;
  memset(&x, 0, sizeof(x));
  x.n_ranks = n; x.rank = h->rank; x.x1_n = (int)lay.x1_n; x.x2_n = (int)lay.x2_n;
  for (int q = 0; q < n; ++q) {
    ulonglong2* base = static_cast<ulonglong2*>(h->peer_xchg[q]);
    x.x1[q] = base;
    x.x2[q] = base + 2 * n * lay.x1_n;
    x.fr[q] = reinterpret_cast<unsigned long long*>(base + lay.n_cells);
  }
  unsigned long long* fl_own = reinterpret_cast<unsigned long long*>(static_cast<ulonglong2*>(h->d_xchg) + lay.n_cells);
  x.verdict = fl_own + n;
  x.error = reinterpret_cast<int*>(fl_own + n + 1);
  if (!h->d_Vinv2) CUDA_TRY(cudaMalloc(&h->d_Vinv2, sizeof(double) * 2 * (size_t)h->cfg.max_points * 6));
  h->use_xchg = true;
  h->epoch_next = 1;
  return PBA_OK;
}

// Peer-memory exchange buffers: allocate, export through CUDA IPC, all-gather the handles with NCCL, map
// every peer.  All ranks agree (all-reduce of a success flag) on whether the exchange is usable; when
// it is not, the NCCL all-reduce path stays in place.
static int setup_xchg(pba_handle* h) {
  h->use_xchg = false;
  const int n = h->n_ranks;
  const char* mode = getenv("PBA_MGPU_EXCHANGE");
  int want = !(mode && strcmp(mode, "nccl") == 0);
  const XchgLayout lay = xchg_layout(h->cfg.max_frames, n);
  const size_t x1_n = lay.x1_n, x2_n = lay.x2_n, n_cells = lay.n_cells, bytes = lay.bytes;
  if (h->d_xchg) { cudaFree(h->d_xchg); h->d_xchg = nullptr; }
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  int ok = want;
  if (ok && cudaMalloc(&h->d_xchg, bytes) != cudaSuccess) { ok = 0; h->d_xchg = nullptr; }
  if (ok && cudaMemsetAsync(h->d_xchg, 0, bytes, h->stream) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&mine, h->d_xchg) != cudaSuccess) ok = 0;
  cudaGetLastError();
  // gather {ok, handle} of every rank
  struct Rec { int ok; int pad[3]; cudaIpcMemHandle_t hnd; };
  static_assert(sizeof(Rec) == 16 + sizeof(cudaIpcMemHandle_t), "Rec layout");
  Rec* d_rec = nullptr;
  CUDA_TRY(cudaMalloc(&d_rec, sizeof(Rec) * n));
  Rec me; memset(&me, 0, sizeof(me)); me.ok = ok; me.hnd = mine;
  CUDA_TRY(cudaMemcpyAsync(d_rec + h->rank, &me, sizeof(Rec), cudaMemcpyHostToDevice, h->stream));
  NCCL_TRY(g_nccl.AllGather(d_rec + h->rank, d_rec, sizeof(Rec), ncclChar, h->comm, h->stream));
  std::vector<Rec> all(n);
  CUDA_TRY(cudaMemcpyAsync(all.data(), d_rec, sizeof(Rec) * n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  for (int q = 0; q < n; ++q) ok = ok && all[q].ok;
  for (int q = 0; q < kMaxRanks; ++q) h->peer_xchg[q] = nullptr;
  if (ok) {
    for (int q = 0; q < n && ok; ++q) {
      if (q == h->rank) { h->peer_xchg[q] = h->d_xchg; continue; }
      if (cudaIpcOpenMemHandle(&h->peer_xchg[q], all[q].hnd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        h->peer_xchg[q] = nullptr; ok = 0; cudaGetLastError();
      }
    }
  }
  // second agreement round: did every rank map every peer?
  me.ok = ok;
  CUDA_TRY(cudaMemcpyAsync(d_rec + h->rank, &me, sizeof(Rec), cudaMemcpyHostToDevice, h->stream));
  NCCL_TRY(g_nccl.AllGather(d_rec + h->rank, d_rec, sizeof(Rec), ncclChar, h->comm, h->stream));
  CUDA_TRY(cudaMemcpyAsync(all.data(), d_rec, sizeof(Rec) * n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  cudaFree(d_rec);
  for (int q = 0; q < n; ++q) ok = ok && all[q].ok;
  if (!ok) {
    for (int q = 0; q < n; ++q)
      if (h->peer_xchg[q] && h->peer_xchg[q] != h->d_xchg) { cudaIpcCloseMemHandle(h->peer_xchg[q]); h->peer_xchg[q] = nullptr; }
    if (want) fprintf(stderr, "[pba_b200] rank %d: peer-memory exchange unavailable, using NCCL all-reduce\n", h->rank);
    return PBA_OK;
  }
  (void)x1_n; (void)x2_n; (void)n_cells;
  return adopt_xchg(h, lay);
}

// End of an upload call: block until the borrowed host buffers have been consumed - unless the caller has opened a
// batch (pba_begin_batch) and keeps the buffers alive until pba_solve / pba_end_batch returns.
static int upload_done(pba_handle* h) {
  if (!h->defer_sync) CUDA_TRY(cudaStreamSynchronize(h->stream));
  return PBA_OK;
}

extern "C" {

const char* pba_last_error(void) { return g_err; }
const char* pba_version(void) { return "photobundle_b200 0.1 (sm_100a)"; }

void pba_default_solver_options(pba_solver_options* o) {
  o->max_num_iterations = 500;          // src/photobundle.cc:751
  o->function_tolerance = 1e-6;         // :756
  o->gradient_tolerance = 1e-6;         // :757
  o->parameter_tolerance = 1e-6;        // :758
  o->initial_trust_region_radius = 1e4; // Ceres defaults below
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->max_num_consecutive_invalid_steps = 5;
  o->jacobi_scaling = 1;
}

int pba_create(const pba_config* cfg, pba_handle** out) {
  if (!cfg || !out) return fail(PBA_ERR_ARGUMENT, "pba_create: null argument");
  if (cfg->rows < 8 || cfg->cols < 8) return fail(PBA_ERR_ARGUMENT, "pba_create: image %dx%d too small", cfg->rows, cfg->cols);
  if (cfg->patch_radius < 1 || cfg->patch_radius > PBA_MAX_RADIUS)
    return fail(PBA_ERR_ARGUMENT, "pba_create: patch_radius %d outside [1,%d]", cfg->patch_radius, PBA_MAX_RADIUS);
  if (cfg->n_channels < 1 || cfg->n_channels > PBA_MAX_CHANNELS)
    return fail(PBA_ERR_ARGUMENT, "pba_create: n_channels %d outside [1,%d]", cfg->n_channels, PBA_MAX_CHANNELS);
  if (cfg->max_frames < 1 || cfg->max_frames > PBA_MAX_FRAMES)
    return fail(PBA_ERR_ARGUMENT, "pba_create: max_frames %d outside [1,%d]", cfg->max_frames, PBA_MAX_FRAMES);
  if (cfg->max_points < 1 || cfg->max_observations < 1)
    return fail(PBA_ERR_ARGUMENT, "pba_create: capacities must be positive");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(PBA_ERR_CUDA, "pba_create: no CUDA device (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  pba_handle* h = new pba_handle();
  h->cfg = *cfg;
  if (cfg->device >= 0) {
    if (cfg->device >= ndev) { delete h; return fail(PBA_ERR_ARGUMENT, "pba_create: device %d of %d", cfg->device, ndev); }
    h->device = cfg->device;
  } else {
    cudaGetDevice(&h->device);
  }
#define CREATE_TRY(expr)                                                                           \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      free_all(h); delete h;                                                                       \
      return fail(PBA_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));                  \
    }                                                                                              \
  } while (0)
  CREATE_TRY(cudaSetDevice(h->device));
  cudaDeviceProp prop;
  CREATE_TRY(cudaGetDeviceProperties(&prop, h->device));
  if (prop.major != 10)
    fprintf(stderr, "[pba_b200] warning: device '%s' is sm_%d%d; this library is built for sm_100a only\n",
            prop.name, prop.major, prop.minor);
  h->sm_count = prop.multiProcessorCount;
  CREATE_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CREATE_TRY(cudaEventCreate(&h->ev0));
  CREATE_TRY(cudaEventCreate(&h->ev1));
  const int side = 2 * cfg->patch_radius + 1;
  h->P = side * side;
  h->CP = h->P * cfg->n_channels;
  h->pitch = (cfg->cols + 15) / 16 * 16;
  h->plane = (size_t)cfg->rows * h->pitch;
  const size_t F = cfg->max_frames, n = cfg->max_points, nnz = cfg->max_observations;
  const size_t D = 6 * F;
  CREATE_TRY(cudaMalloc(&h->d_cams, sizeof(double) * 2 * F * 6));
  CREATE_TRY(cudaMalloc(&h->d_pts, sizeof(double) * 2 * n * 3));
  CREATE_TRY(cudaMalloc(&h->d_weights, sizeof(double) * h->P));
  CREATE_TRY(cudaMalloc(&h->d_desc, sizeof(float) * n * h->CP));
  CREATE_TRY(cudaMalloc(&h->d_obs_off, sizeof(int) * (n + 1)));
  CREATE_TRY(cudaMalloc(&h->d_obs_frame, sizeof(int) * nnz));
  CREATE_TRY(cudaMalloc(&h->d_V, sizeof(double) * 2 * n * 6));
  CREATE_TRY(cudaMalloc(&h->d_gp, sizeof(double) * 2 * n * 3));
  CREATE_TRY(cudaMalloc(&h->d_W, sizeof(double) * 2 * nnz * 18));
  CREATE_TRY(cudaMalloc(&h->d_Xacc, sizeof(double) * (F * kUStride + kEacc + kMaxRanks)));
  CREATE_TRY(cudaMalloc(&h->d_Ucur, sizeof(double) * F * kUStride));
  CREATE_TRY(cudaMalloc(&h->d_scale_p, sizeof(double) * n * 3));
  CREATE_TRY(cudaMalloc(&h->d_Vinv, sizeof(double) * n * 6));
  CREATE_TRY(cudaMalloc(&h->d_S, sizeof(double) * kSCopies * reduced_capacity((int)F)));   // multi-GPU: two hypotheses + once more; one GPU: kSReplicas copies
  CREATE_TRY(cudaMalloc(&h->d_ticket, sizeof(unsigned int)));
  CREATE_TRY(cudaMalloc(&h->d_state, 2 * sizeof(LmState)));
  CREATE_TRY(cudaMallocHost(&h->h_state, sizeof(LmState)));
#undef CREATE_TRY
  *out = h;
  return PBA_OK;
}

void pba_destroy(pba_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  free_all(h);
  delete h;
}

// One dense host->device copy per frame (a pitched 2-D copy straight from host memory moves odd-width
// rows one by one: 8 GB/s instead of the link rate), then the rows are spread to the 16-byte pitch on
// the device.
static int upload_plane_u8(pba_handle* h, int slot, const uint8_t* img) {
  const size_t dense = (size_t)h->cfg.rows * h->cfg.cols;
  if (h->pitch == h->cfg.cols) {
    CUDA_TRY(cudaMemcpyAsync(h->d_u8 + (size_t)slot * h->plane, img, dense, cudaMemcpyHostToDevice, h->stream));
    return PBA_OK;
  }
  if (!h->d_stage_u8) CUDA_TRY(cudaMalloc(&h->d_stage_u8, (size_t)h->cfg.max_frames * dense));
  uint8_t* st = h->d_stage_u8 + (size_t)slot * dense;
  CUDA_TRY(cudaMemcpyAsync(st, img, dense, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpy2DAsync(h->d_u8 + (size_t)slot * h->plane, h->pitch, st, h->cfg.cols, h->cfg.cols, h->cfg.rows,
                             cudaMemcpyDeviceToDevice, h->stream));
  return PBA_OK;
}

int pba_set_frames_u8(pba_handle* h, int32_t n_frames, const uint8_t* const* images) {
  if (!h || !images) return fail(PBA_ERR_ARGUMENT, "pba_set_frames_u8: null argument");
  if (h->cfg.n_channels != 1) return fail(PBA_ERR_ARGUMENT, "pba_set_frames_u8: handle has %d channels; uint8 frames are the 1-channel Intensity descriptor", h->cfg.n_channels);
  if (n_frames < 1 || n_frames > h->cfg.max_frames) return fail(PBA_ERR_CAPACITY, "pba_set_frames_u8: %d frames, capacity %d", n_frames, h->cfg.max_frames);
  CUDA_TRY(cudaSetDevice(h->device));
  if (!h->d_u8) {
    CUDA_TRY(cudaMalloc(&h->d_u8, (size_t)h->cfg.max_frames * h->plane + 64));
    CUDA_TRY(cudaMemsetAsync(h->d_u8, 0, (size_t)h->cfg.max_frames * h->plane, h->stream));
  }
  for (int f = 0; f < n_frames; ++f) {
    if (!images[f]) return fail(PBA_ERR_ARGUMENT, "pba_set_frames_u8: images[%d] is null", f);
    int rc = upload_plane_u8(h, f, images[f]);
    if (rc) return rc;
  }
  { int rc_ = upload_done(h); if (rc_) return rc_; }
  h->frames_are_u8 = true; h->have_frames = true; h->n_frames = n_frames;
  return PBA_OK;
}

int pba_set_frames_u8_pyr(pba_handle* h, int32_t n_frames, const uint8_t* const* images, int32_t src_rows,
                          int32_t src_cols, int32_t levels_down) {
  if (!h || !images) return fail(PBA_ERR_ARGUMENT, "pba_set_frames_u8_pyr: null argument");
  if (levels_down == 0 && src_rows == h->cfg.rows && src_cols == h->cfg.cols) return pba_set_frames_u8(h, n_frames, images);
  if (h->cfg.n_channels != 1) return fail(PBA_ERR_ARGUMENT, "pba_set_frames_u8_pyr: uint8 frames are the 1-channel Intensity descriptor");
  if (n_frames < 1 || n_frames > h->cfg.max_frames) return fail(PBA_ERR_CAPACITY, "pba_set_frames_u8_pyr: %d frames, capacity %d", n_frames, h->cfg.max_frames);
  int r = src_rows, c = src_cols;
  for (int l = 0; l < levels_down; ++l) { r = (r + 1) / 2; c = (c + 1) / 2; }
  if (levels_down < 0 || r != h->cfg.rows || c != h->cfg.cols)
    return fail(PBA_ERR_ARGUMENT, "pba_set_frames_u8_pyr: %dx%d reduced %d times is %dx%d, handle is %dx%d", src_rows, src_cols,
                levels_down, r, c, h->cfg.rows, h->cfg.cols);
  CUDA_TRY(cudaSetDevice(h->device));
  if (!h->d_u8) {
    CUDA_TRY(cudaMalloc(&h->d_u8, (size_t)h->cfg.max_frames * h->plane + 64));
    CUDA_TRY(cudaMemsetAsync(h->d_u8, 0, (size_t)h->cfg.max_frames * h->plane, h->stream));
  }
  // two ping-pong scratch planes at level-0 size
  const int p0 = (src_cols + 15) / 16 * 16;
  if (h->pyr_scratch_bytes < 2 * (size_t)src_rows * p0) {
    cudaFree(h->d_pyr_scratch); h->d_pyr_scratch = nullptr; h->pyr_scratch_bytes = 0;
    CUDA_TRY(cudaMalloc(&h->d_pyr_scratch, 2 * (size_t)src_rows * p0));
    h->pyr_scratch_bytes = 2 * (size_t)src_rows * p0;
  }
  uint8_t* scratch = h->d_pyr_scratch;
  for (int f = 0; f < n_frames; ++f) {
    if (!images[f]) return fail(PBA_ERR_ARGUMENT, "pba_set_frames_u8_pyr: images[%d] is null", f);
    uint8_t* a = scratch;
    uint8_t* b = scratch + (size_t)src_rows * p0;
    CUDA_TRY(cudaMemcpyAsync(a, images[f], (size_t)src_rows * src_cols, cudaMemcpyHostToDevice, h->stream));   // dense: pitch = cols
    int rr = src_rows, cc = src_cols, pa = src_cols;
    for (int l = 0; l < levels_down; ++l) {
      const bool last = (l == levels_down - 1);
      uint8_t* dst = last ? h->d_u8 + (size_t)f * h->plane : b;
      const int pd = last ? h->pitch : p0;
      CUDA_TRY(launch_pyrdown_u8(a, rr, cc, pa, dst, pd, h->stream));
      rr = (rr + 1) / 2; cc = (cc + 1) / 2; pa = pd;
      std::swap(a, b);
      if (!last) { /* a now holds the reduced image */ }
    }
  }
  { int rc_ = upload_done(h); if (rc_) return rc_; }
  h->frames_are_u8 = true; h->have_frames = true; h->n_frames = n_frames;
  return PBA_OK;
}

int pba_pyrdown_u8(const uint8_t* src, int32_t rows, int32_t cols, uint8_t* dst, int32_t device) {
  if (!src || !dst || rows < 1 || cols < 1) return fail(PBA_ERR_ARGUMENT, "pba_pyrdown_u8: bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(PBA_ERR_CUDA, "pba_pyrdown_u8: no CUDA device; no CPU fallback");
  if (device >= 0) CUDA_TRY(cudaSetDevice(device));
  const int drows = (rows + 1) / 2, dcols = (cols + 1) / 2;
  uint8_t *ds = nullptr, *dd = nullptr;
  CUDA_TRY(cudaMalloc(&ds, (size_t)rows * cols));
  CUDA_TRY(cudaMalloc(&dd, (size_t)drows * dcols));
  CUDA_TRY(cudaMemcpy(ds, src, (size_t)rows * cols, cudaMemcpyHostToDevice));
  CUDA_TRY(launch_pyrdown_u8(ds, rows, cols, cols, dd, dcols, 0));
  CUDA_TRY(cudaMemcpy(dst, dd, (size_t)drows * dcols, cudaMemcpyDeviceToHost));
  cudaFree(ds); cudaFree(dd);
  return PBA_OK;
}

int pba_set_frame_u8(pba_handle* h, int32_t slot, const uint8_t* image) {
  if (!h || !image) return fail(PBA_ERR_ARGUMENT, "pba_set_frame_u8: null argument");
  if (!h->d_u8 || !h->frames_are_u8) return fail(PBA_ERR_STATE, "pba_set_frame_u8: call pba_set_frames_u8 first");
  if (slot < 0 || slot >= h->cfg.max_frames) return fail(PBA_ERR_ARGUMENT, "pba_set_frame_u8: slot %d", slot);
  CUDA_TRY(cudaSetDevice(h->device));
  int rc = upload_plane_u8(h, slot, image);
  if (rc) return rc;
  { int rc_ = upload_done(h); if (rc_) return rc_; }
  return PBA_OK;
}

int pba_set_frames_f32(pba_handle* h, int32_t n_frames, const float* const* planes) {
  if (!h || !planes) return fail(PBA_ERR_ARGUMENT, "pba_set_frames_f32: null argument");
  if (n_frames < 1 || n_frames > h->cfg.max_frames) return fail(PBA_ERR_CAPACITY, "pba_set_frames_f32: %d frames, capacity %d", n_frames, h->cfg.max_frames);
  CUDA_TRY(cudaSetDevice(h->device));
  const int C = h->cfg.n_channels;
  if (!h->d_f32) {
    CUDA_TRY(cudaMalloc(&h->d_f32, sizeof(float) * (size_t)h->cfg.max_frames * C * h->plane + 64));
    CUDA_TRY(cudaMemsetAsync(h->d_f32, 0, sizeof(float) * (size_t)h->cfg.max_frames * C * h->plane, h->stream));
  }
  for (int i = 0; i < n_frames * C; ++i) {
    if (!planes[i]) return fail(PBA_ERR_ARGUMENT, "pba_set_frames_f32: planes[%d] is null", i);
    CUDA_TRY(cudaMemcpy2DAsync(h->d_f32 + (size_t)i * h->plane, sizeof(float) * h->pitch, planes[i],
                               sizeof(float) * h->cfg.cols, sizeof(float) * h->cfg.cols, h->cfg.rows,
                               cudaMemcpyHostToDevice, h->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->frames_are_u8 = false; h->have_frames = true; h->n_frames = n_frames;
  return PBA_OK;
}

int pba_descriptor_channels(int32_t t) {
  return t == PBA_DESC_INTENSITY ? 1 : t == PBA_DESC_INTENSITY_AND_GRADIENT ? 3 : t == PBA_DESC_BITPLANES ? 8 : -1;
}

static int ensure_descriptor_scratch(pba_handle* h) {
  if (!h->d_scr_a) {
    CUDA_TRY(cudaMalloc(&h->d_scr_a, h->plane));
    CUDA_TRY(cudaMalloc(&h->d_scr_b, h->plane));
  }
  return PBA_OK;
}

int pba_set_frames_u8_descriptor(pba_handle* h, int32_t n_frames, const uint8_t* const* images, int32_t type) {
  if (!h || !images) return fail(PBA_ERR_ARGUMENT, "pba_set_frames_u8_descriptor: null argument");
  if (type == PBA_DESC_INTENSITY) return pba_set_frames_u8(h, n_frames, images);
  const int C = pba_descriptor_channels(type);
  if (C < 0) return fail(PBA_ERR_ARGUMENT, "pba_set_frames_u8_descriptor: unknown descriptor type %d", type);
  if (h->cfg.n_channels != C) return fail(PBA_ERR_ARGUMENT, "pba_set_frames_u8_descriptor: descriptor type %d has %d channels, the handle %d", type, C, h->cfg.n_channels);
  if (n_frames < 1 || n_frames > h->cfg.max_frames) return fail(PBA_ERR_CAPACITY, "pba_set_frames_u8_descriptor: %d frames, capacity %d", n_frames, h->cfg.max_frames);
  CUDA_TRY(cudaSetDevice(h->device));
  if (!h->d_u8) {
    CUDA_TRY(cudaMalloc(&h->d_u8, (size_t)h->cfg.max_frames * h->plane + 64));
    CUDA_TRY(cudaMemsetAsync(h->d_u8, 0, (size_t)h->cfg.max_frames * h->plane, h->stream));
  }
  if (!h->d_f32) {
    CUDA_TRY(cudaMalloc(&h->d_f32, sizeof(float) * (size_t)h->cfg.max_frames * C * h->plane + 64));
    CUDA_TRY(cudaMemsetAsync(h->d_f32, 0, sizeof(float) * (size_t)h->cfg.max_frames * C * h->plane, h->stream));
  }
  int rc = ensure_descriptor_scratch(h);
  if (rc) return rc;
  for (int f = 0; f < n_frames; ++f) {
    if (!images[f]) return fail(PBA_ERR_ARGUMENT, "pba_set_frames_u8_descriptor: images[%d] is null", f);
    rc = upload_plane_u8(h, f, images[f]);
    if (rc) return rc;
    CUDA_TRY(launch_channels(type, h->d_u8 + (size_t)f * h->plane, h->cfg.rows, h->cfg.cols, h->pitch, h->d_scr_a, h->d_scr_b,
                             h->d_f32 + (size_t)f * C * h->plane, h->pitch, h->plane, h->stream));
  }
  { int rc_ = upload_done(h); if (rc_) return rc_; }
  h->frames_are_u8 = false; h->have_frames = true; h->n_frames = n_frames;
  return PBA_OK;
}

int pba_get_channel_plane(pba_handle* h, int32_t frame, int32_t channel, float* out) {
  if (!h || !out) return fail(PBA_ERR_ARGUMENT, "pba_get_channel_plane: null argument");
  if (!h->have_frames || h->frames_are_u8 || !h->d_f32) return fail(PBA_ERR_STATE, "pba_get_channel_plane: no fp32 channel planes on the device");
  if (frame < 0 || frame >= h->n_frames || channel < 0 || channel >= h->cfg.n_channels)
    return fail(PBA_ERR_ARGUMENT, "pba_get_channel_plane: frame %d / channel %d", frame, channel);
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpy2D(out, sizeof(float) * h->cfg.cols, h->d_f32 + ((size_t)frame * h->cfg.n_channels + channel) * h->plane,
                        sizeof(float) * h->pitch, sizeof(float) * h->cfg.cols, h->cfg.rows, cudaMemcpyDeviceToHost));
  return PBA_OK;
}

int pba_prepare_frame_u8(pba_handle* h, const uint8_t* image, int32_t type) {
  if (!h || !image) return fail(PBA_ERR_ARGUMENT, "pba_prepare_frame_u8: null argument");
  const int C = pba_descriptor_channels(type);
  if (C < 0) return fail(PBA_ERR_ARGUMENT, "pba_prepare_frame_u8: unknown descriptor type %d", type);
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t dense = (size_t)h->cfg.rows * h->cfg.cols;
  const int c_alloc = C < 3 ? 3 : C;   // Intensity is built as {I, Ix, Iy} and uses plane 0 only
  if (!h->d_new_u8) {
    CUDA_TRY(cudaMalloc(&h->d_new_u8, h->plane));
    CUDA_TRY(cudaMemsetAsync(h->d_new_u8, 0, h->plane, h->stream));
    CUDA_TRY(cudaMalloc(&h->d_sal, sizeof(float) * dense));
  }
  if (h->new_channels < c_alloc) {
    cudaFree(h->d_new_planes); h->d_new_planes = nullptr; h->new_channels = 0;
    CUDA_TRY(cudaMalloc(&h->d_new_planes, sizeof(float) * (size_t)c_alloc * h->plane));
    CUDA_TRY(cudaMemsetAsync(h->d_new_planes, 0, sizeof(float) * (size_t)c_alloc * h->plane, h->stream));
    h->new_channels = c_alloc;
  }
  int rc = ensure_descriptor_scratch(h);
  if (rc) return rc;
  if (!h->d_stage_u8) CUDA_TRY(cudaMalloc(&h->d_stage_u8, (size_t)h->cfg.max_frames * dense));
  CUDA_TRY(cudaMemcpyAsync(h->d_stage_u8, image, dense, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpy2DAsync(h->d_new_u8, h->pitch, h->d_stage_u8, h->cfg.cols, h->cfg.cols, h->cfg.rows, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(launch_channels(type == PBA_DESC_BITPLANES ? PBA_DESC_BITPLANES : PBA_DESC_INTENSITY_AND_GRADIENT, h->d_new_u8, h->cfg.rows,
                           h->cfg.cols, h->pitch, h->d_scr_a, h->d_scr_b, h->d_new_planes, h->pitch, h->plane, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));   // `image` is only borrowed for the duration of the call
  h->new_active = C;
  return PBA_OK;
}

int pba_saliency_map(pba_handle* h, float* out) {
  if (!h || !out) return fail(PBA_ERR_ARGUMENT, "pba_saliency_map: null argument");
  if (h->new_active <= 0) return fail(PBA_ERR_STATE, "pba_saliency_map: call pba_prepare_frame_u8 first");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(launch_saliency(h->d_new_planes, h->new_active, h->cfg.rows, h->cfg.cols, h->pitch, h->plane, h->d_sal, h->stream));
  CUDA_TRY(cudaMemcpyAsync(out, h->d_sal, sizeof(float) * (size_t)h->cfg.rows * h->cfg.cols, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return PBA_OK;
}

int pba_extract_descriptors(pba_handle* h, int32_t n, const int32_t* xy, double* desc) {
  if (!h || (n > 0 && (!xy || !desc))) return fail(PBA_ERR_ARGUMENT, "pba_extract_descriptors: null argument");
  if (h->new_active <= 0) return fail(PBA_ERR_STATE, "pba_extract_descriptors: call pba_prepare_frame_u8 first");
  if (n < 0) return fail(PBA_ERR_ARGUMENT, "pba_extract_descriptors: n < 0");
  if (n == 0) return PBA_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  const int CP = h->new_active * h->P;
  if (h->xy_cap < n) {
    cudaFree(h->d_xy); cudaFree(h->d_desc_out); h->d_xy = nullptr; h->d_desc_out = nullptr; h->xy_cap = 0;
    CUDA_TRY(cudaMalloc(&h->d_xy, sizeof(int) * 2 * (size_t)n));
    CUDA_TRY(cudaMalloc(&h->d_desc_out, sizeof(double) * (size_t)n * PBA_MAX_CHANNELS * h->P));
    h->xy_cap = n;
  }
  CUDA_TRY(cudaMemcpyAsync(h->d_xy, xy, sizeof(int) * 2 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(launch_extract_patches(h->d_new_planes, h->new_active, h->cfg.rows, h->cfg.cols, h->pitch, h->plane, h->cfg.patch_radius, n,
                                  h->d_xy, h->d_desc_out, h->stream));
  CUDA_TRY(cudaMemcpyAsync(desc, h->d_desc_out, sizeof(double) * (size_t)n * CP, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return PBA_OK;
}

int pba_associate(pba_handle* h, int32_t n, const double* xyz, const float* ref_patch, const float* ref_norm,
                  const double* T_c, const double* K, int32_t border, float* score, int32_t* row_col) {
  if (!h || !T_c || !K || (n > 0 && (!xyz || !ref_patch || !ref_norm || !score || !row_col)))
    return fail(PBA_ERR_ARGUMENT, "pba_associate: null argument");
  if (h->new_active <= 0) return fail(PBA_ERR_STATE, "pba_associate: call pba_prepare_frame_u8 first");
  if (n < 0 || border < 2) return fail(PBA_ERR_ARGUMENT, "pba_associate: n < 0 or border < 2 (the 5x5 ZNCC patch)");
  if (n == 0) return PBA_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->as_cap < n) {
    cudaFree(h->d_as_xyz); cudaFree(h->d_as_patch); cudaFree(h->d_as_norm); cudaFree(h->d_as_score); cudaFree(h->d_as_rc);
    h->d_as_xyz = nullptr; h->d_as_patch = h->d_as_norm = h->d_as_score = nullptr; h->d_as_rc = nullptr; h->as_cap = 0;
    const size_t cap = (size_t)n + n / 2 + 64;
    CUDA_TRY(cudaMalloc(&h->d_as_xyz, sizeof(double) * 3 * cap));
    CUDA_TRY(cudaMalloc(&h->d_as_patch, sizeof(float) * 25 * cap));
    CUDA_TRY(cudaMalloc(&h->d_as_norm, sizeof(float) * cap));
    CUDA_TRY(cudaMalloc(&h->d_as_score, sizeof(float) * cap));
    CUDA_TRY(cudaMalloc(&h->d_as_rc, sizeof(int) * 2 * cap));
    h->as_cap = (int)cap;
  }
  if (!h->d_as_mats) CUDA_TRY(cudaMalloc(&h->d_as_mats, sizeof(double) * 25));
  CUDA_TRY(cudaMemcpyAsync(h->d_as_xyz, xyz, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_as_patch, ref_patch, sizeof(float) * 25 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_as_norm, ref_norm, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_as_mats, T_c, sizeof(double) * 16, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_as_mats + 16, K, sizeof(double) * 9, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(launch_associate(h->d_new_u8, h->cfg.rows, h->cfg.cols, h->pitch, n, h->d_as_xyz, h->d_as_patch, h->d_as_norm, h->d_as_mats,
                            h->d_as_mats + 16, border, h->d_as_score, h->d_as_rc, h->stream));
  CUDA_TRY(cudaMemcpyAsync(score, h->d_as_score, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaMemcpyAsync(row_col, h->d_as_rc, sizeof(int) * 2 * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return PBA_OK;
}

int pba_select_candidates(pba_handle* h, const float* depth, int32_t n_masked, const int32_t* masked_row_col,
                          int32_t mask_radius, int32_t nms_radius, int32_t border, double min_depth, double max_depth,
                          int32_t capacity, int32_t* cand_row_col, float* cand_saliency, int32_t* n_out) {
  if (!h || !n_out || (n_masked > 0 && !masked_row_col) || (capacity > 0 && (!cand_row_col || !cand_saliency)))
    return fail(PBA_ERR_ARGUMENT, "pba_select_candidates: null argument");
  if (h->new_active <= 0) return fail(PBA_ERR_STATE, "pba_select_candidates: call pba_prepare_frame_u8 first");
  if (n_masked < 0 || capacity < 0 || mask_radius < 0 || nms_radius < 0 || border < nms_radius || border < mask_radius)
    return fail(PBA_ERR_ARGUMENT, "pba_select_candidates: negative size, or border smaller than a radius");
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t dense = (size_t)h->cfg.rows * h->cfg.cols;
  if (!h->d_mask) {
    CUDA_TRY(cudaMalloc(&h->d_mask, dense));
    CUDA_TRY(cudaMalloc(&h->d_depth, sizeof(float) * dense));
    CUDA_TRY(cudaMalloc(&h->d_count, sizeof(int)));
  }
  if (h->cand_cap < capacity) {
    cudaFree(h->d_cand_rc); cudaFree(h->d_cand_sal); h->d_cand_rc = nullptr; h->d_cand_sal = nullptr; h->cand_cap = 0;
    CUDA_TRY(cudaMalloc(&h->d_cand_rc, sizeof(int) * 2 * (size_t)capacity));
    CUDA_TRY(cudaMalloc(&h->d_cand_sal, sizeof(float) * (size_t)capacity));
    h->cand_cap = capacity;
  }
  if (h->masked_cap < n_masked) {
    cudaFree(h->d_masked_rc); h->d_masked_rc = nullptr; h->masked_cap = 0;
    CUDA_TRY(cudaMalloc(&h->d_masked_rc, sizeof(int) * 2 * ((size_t)n_masked + n_masked / 2 + 64)));
    h->masked_cap = n_masked + n_masked / 2 + 64;
  }
  if (depth) CUDA_TRY(cudaMemcpyAsync(h->d_depth, depth, sizeof(float) * dense, cudaMemcpyHostToDevice, h->stream));
  if (n_masked > 0) CUDA_TRY(cudaMemcpyAsync(h->d_masked_rc, masked_row_col, sizeof(int) * 2 * (size_t)n_masked, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(launch_saliency(h->d_new_planes, h->new_active, h->cfg.rows, h->cfg.cols, h->pitch, h->plane, h->d_sal, h->stream));
  CUDA_TRY(launch_candidates(h->d_sal, h->d_mask, depth ? h->d_depth : nullptr, h->cfg.rows, h->cfg.cols, border, nms_radius, n_masked, h->d_masked_rc,
                             mask_radius, min_depth, max_depth, capacity, h->d_count, h->d_cand_rc, h->d_cand_sal, h->stream));
  int count = 0;
  CUDA_TRY(cudaMemcpyAsync(&count, h->d_count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  *n_out = count;
  const int m = count < capacity ? count : capacity;
  if (m > 0) {
    std::vector<int32_t> rc(2 * (size_t)m);
    std::vector<float> sal(m);
    CUDA_TRY(cudaMemcpy(rc.data(), h->d_cand_rc, sizeof(int) * 2 * (size_t)m, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(sal.data(), h->d_cand_sal, sizeof(float) * (size_t)m, cudaMemcpyDeviceToHost));
    // the compaction order is arbitrary: back into scan order (row-major), which is what addFrame produces - a counting
    // sort by row, then the few candidates of each row by column
    const int rows = h->cfg.rows;
    std::vector<int> start((size_t)rows + 1, 0), idx(m);
    for (int i = 0; i < m; ++i) ++start[(size_t)rc[2 * i] + 1];
    for (int r = 0; r < rows; ++r) start[(size_t)r + 1] += start[r];
    {
      std::vector<int> fill(start.begin(), start.end() - 1);
      for (int i = 0; i < m; ++i) idx[fill[rc[2 * i]]++] = i;
    }
    for (int r = 0; r < rows; ++r)
      std::sort(idx.begin() + start[r], idx.begin() + start[(size_t)r + 1], [&](int a, int b) { return rc[2 * a + 1] < rc[2 * b + 1]; });
    for (int i = 0; i < m; ++i) { cand_row_col[2 * i] = rc[2 * idx[i]]; cand_row_col[2 * i + 1] = rc[2 * idx[i] + 1]; cand_saliency[i] = sal[idx[i]]; }
  }
  return PBA_OK;
}

int pba_set_poses(pba_handle* h, int32_t n_frames, const double* cam6, int32_t fixed_frame) {
  if (h) h->results_cached = false;
  if (!h || !cam6) return fail(PBA_ERR_ARGUMENT, "pba_set_poses: null argument");
  if (n_frames < 1 || n_frames > h->cfg.max_frames) return fail(PBA_ERR_CAPACITY, "pba_set_poses: %d frames, capacity %d", n_frames, h->cfg.max_frames);
  if (fixed_frame < -1 || fixed_frame >= n_frames) return fail(PBA_ERR_ARGUMENT, "pba_set_poses: fixed_frame %d", fixed_frame);
  if (h->have_frames && n_frames != h->n_frames) return fail(PBA_ERR_ARGUMENT, "pba_set_poses: %d poses for %d frames", n_frames, h->n_frames);
  CUDA_TRY(cudaSetDevice(h->device));
  // buffer 0 holds x; buffer 1 starts as a copy (fixed / unused cameras are never rewritten)
  const size_t stride = (size_t)n_frames * 6;
  CUDA_TRY(cudaMemcpyAsync(h->d_cams, cam6, sizeof(double) * stride, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_cams + stride, cam6, sizeof(double) * stride, cudaMemcpyHostToDevice, h->stream));
  { int rc_ = upload_done(h); if (rc_) return rc_; }
  h->n_frames = n_frames; h->fixed_frame = fixed_frame; h->have_poses = true;
  return PBA_OK;
}

// Contiguous point blocks balanced by observation count: rank r gets points [begin[r], begin[r+1]).
static void shard_split(int n_points, const int32_t* obs_offsets, int n_ranks, std::vector<int>& begin) {
  begin.assign(n_ranks + 1, n_points);
  begin[0] = 0;
  const long long nnz = obs_offsets[n_points];
  int p = 0;
  for (int r = 1; r < n_ranks; ++r) {
    const long long target = nnz * r / n_ranks;
    while (p < n_points && obs_offsets[p] < target) ++p;
    begin[r] = p;
  }
}

int pba_shard_range(int32_t n_points, const int32_t* obs_offsets, int32_t rank, int32_t n_ranks, int32_t* first, int32_t* last) {
  if (!obs_offsets || !first || !last || n_points < 0 || n_ranks < 1 || rank < 0 || rank >= n_ranks)
    return fail(PBA_ERR_ARGUMENT, "pba_shard_range: bad argument");
  std::vector<int> b;
  shard_split(n_points, obs_offsets, n_ranks, b);
  *first = b[rank]; *last = b[rank + 1];
  return PBA_OK;
}

// K_A keeps one point per warp, 14 warps per CTA, two CTAs per SM: up to that many points are ONE wave on one GPU.
// PBA_MGPU_REPLICATE=0 / 1 forces sharding / replication.
static bool replicate_window(const pba_handle* h, int n_points) {
  if (const char* e = getenv("PBA_MGPU_REPLICATE")) return atoi(e) != 0;
  return n_points <= 2 * 14 * h->sm_count;
}

int pba_set_points(pba_handle* h, int32_t n_points, const double* xyz, const double* desc,
                   const int32_t* obs_offsets, const int32_t* obs_frame, const double* weights) {
  if (h) h->results_cached = false;
  if (!h || !xyz || !desc || !obs_offsets || !obs_frame || !weights) return fail(PBA_ERR_ARGUMENT, "pba_set_points: null argument");
  if (n_points < 0 || n_points > h->cfg.max_points) return fail(PBA_ERR_CAPACITY, "pba_set_points: %d points, capacity %d", n_points, h->cfg.max_points);
  if (obs_offsets[0] != 0) return fail(PBA_ERR_ARGUMENT, "pba_set_points: obs_offsets[0] must be 0");
  const int nnz = obs_offsets[n_points];
  if (nnz < 0 || nnz > h->cfg.max_observations) return fail(PBA_ERR_CAPACITY, "pba_set_points: %d observations, capacity %d", nnz, h->cfg.max_observations);
  const int F = h->have_poses || h->have_frames ? h->n_frames : h->cfg.max_frames;
  h->frame_used.assign(h->cfg.max_frames, 0);
  h->max_obs_frame = -1;
  for (int p = 0; p < n_points; ++p) {
    if (obs_offsets[p + 1] < obs_offsets[p]) return fail(PBA_ERR_ARGUMENT, "pba_set_points: obs_offsets not monotone at %d", p);
    // one residual block per (point, frame): at most one observation per frame (the kernels hold a point's
    // observations in per-frame slots)
    if (obs_offsets[p + 1] - obs_offsets[p] > F)
      return fail(PBA_ERR_ARGUMENT, "pba_set_points: point %d has %d observations in a window of %d frames", p, obs_offsets[p + 1] - obs_offsets[p], F);
    unsigned seen = 0;
    for (int o = obs_offsets[p]; o < obs_offsets[p + 1]; ++o) {
      if (obs_frame[o] < 0 || obs_frame[o] >= F) return fail(PBA_ERR_ARGUMENT, "pba_set_points: obs_frame[%d]=%d outside [0,%d)", o, obs_frame[o], F);
      if (seen & (1u << obs_frame[o])) return fail(PBA_ERR_ARGUMENT, "pba_set_points: point %d observes frame %d more than once", p, obs_frame[o]);
      seen |= 1u << obs_frame[o];
      h->frame_used[obs_frame[o]] = 1;
      if (obs_frame[o] > h->max_obs_frame) h->max_obs_frame = obs_frame[o];
    }
  }
  CUDA_TRY(cudaSetDevice(h->device));
  // this rank's shard (the whole window on one GPU, or when the window is too small to gain from sharding)
  h->replicated = h->n_ranks > 1 && replicate_window(h, n_points);
  h->eff_ranks = h->replicated ? 1 : h->n_ranks;
  h->eff_xchg = h->use_xchg && !h->replicated;
  const int my_rank = h->replicated ? 0 : h->rank;
  shard_split(n_points, obs_offsets, h->eff_ranks, h->shard_begin);
  const int p0 = h->shard_begin[my_rank], p1 = h->shard_begin[my_rank + 1];
  const int n_loc = p1 - p0, o_base = obs_offsets[p0], nnz_loc = obs_offsets[p1] - o_base;
  if (!h->p_desc) {   // pinned staging, sized for the handle's capacity
    CUDA_TRY(cudaMallocHost(&h->p_desc, sizeof(float) * (size_t)h->cfg.max_points * h->CP));
    CUDA_TRY(cudaMallocHost(&h->p_off, sizeof(int32_t) * ((size_t)h->cfg.max_points + 1)));
  }
  float* descf = h->p_desc;
  const size_t n_desc = (size_t)n_loc * h->CP;
  const double* dsrc = desc + (size_t)p0 * h->CP;
  for (size_t i = 0; i < n_desc; ++i) descf[i] = (float)dsrc[i];
  int32_t* off_loc = h->p_off;
  for (int i = 0; i <= n_loc; ++i) off_loc[i] = obs_offsets[p0 + i] - o_base;
  const size_t pstride = (size_t)n_loc * 3;
  CUDA_TRY(cudaMemcpyAsync(h->d_pts, xyz + (size_t)p0 * 3, sizeof(double) * pstride, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_pts + pstride, h->d_pts, sizeof(double) * pstride, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_desc, descf, sizeof(float) * n_desc, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_obs_off, off_loc, sizeof(int) * (n_loc + 1), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_obs_frame, obs_frame + o_base, sizeof(int) * nnz_loc, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_weights, weights, sizeof(double) * h->P, cudaMemcpyHostToDevice, h->stream));
  { int rc_ = upload_done(h); if (rc_) return rc_; }
  h->n_points = n_loc; h->nnz = nnz_loc; h->n_points_total = n_points; h->nnz_total = nnz; h->have_points = true;
  return PBA_OK;
}

static int check_ready(pba_handle* h, const char* who) {
  if (!h) return fail(PBA_ERR_ARGUMENT, "%s: null handle", who);
  if (!h->have_frames) return fail(PBA_ERR_STATE, "%s: frames not set", who);
  if (!h->have_poses) return fail(PBA_ERR_STATE, "%s: poses not set", who);
  if (!h->have_points) return fail(PBA_ERR_STATE, "%s: points not set", who);
  if (h->max_obs_frame >= h->n_frames)
    return fail(PBA_ERR_STATE, "%s: an observation refers to frame %d but the window has %d frames (points were set for a wider window)", who, h->max_obs_frame, h->n_frames);
  return PBA_OK;
}

static StepParams make_step_params(pba_handle* h, const LmState* st) {
  StepParams p;
  memset(&p, 0, sizeof(p));
  p.fr.u8 = h->frames_are_u8 ? h->d_u8 : nullptr;
  p.fr.f32 = h->frames_are_u8 ? nullptr : h->d_f32;
  p.fr.rows = h->cfg.rows; p.fr.cols = h->cfg.cols; p.fr.pitch = h->pitch;
  p.fr.n_channels = h->cfg.n_channels; p.fr.plane = h->plane;
  p.n_frames = h->n_frames; p.fixed_frame = h->fixed_frame; p.n_points = h->n_points; p.nnz = h->nnz;
  p.fx = h->cfg.fx; p.fy = h->cfg.fy; p.cx = h->cfg.cx; p.cy = h->cfg.cy; p.huber = h->cfg.huber;
  p.st = st;
  p.cams = h->d_cams; p.pts = h->d_pts; p.desc = h->d_desc; p.obs_off = h->d_obs_off;
  p.obs_frame = h->d_obs_frame; p.weights = h->d_weights;
  p.V = h->d_V; p.gp = h->d_gp; p.W = h->d_W; p.Xacc = h->d_Xacc; p.rank = h->rank;
  p.scale_p = h->d_scale_p; p.Vinv = h->d_Vinv;
  return p;
}

// Multi-GPU K_B: eliminating under both outcomes of the pending decision costs nothing while each half of the SMs
// still has at most one batch of points per CTA (32 points); beyond that it doubles the elimination time, and deciding
// first (one more exchange of ~6 us) is the cheaper iteration.  Wide windows (more than 8 frames) never speculate: the
// message would carry two 90 x 91 systems to every peer (measured at 8 GPUs, 16 frames: K_B 122 us against 88 us).
// PBA_MGPU_SPECULATE=0/1 overrides.
static bool speculate_decision(const pba_handle* h) {
  if (!h->eff_xchg) return false;
  if (const char* e = getenv("PBA_MGPU_SPECULATE")) return atoi(e) != 0;
  return h->n_points <= 32 * (h->sm_count / 2) && h->n_frames <= 8;
}

static LmParams make_lm_params(pba_handle* h) {
  LmParams lp;
  memset(&lp, 0, sizeof(lp));
  lp.trace = h->d_trace; lp.ticket = h->d_ticket; lp.stamps = h->d_stamps;
  lp.n_frames = h->n_frames; lp.n_points = h->n_points; lp.nnz = h->nnz;
  lp.obs_off = h->d_obs_off; lp.obs_frame = h->d_obs_frame;
  lp.cams = h->d_cams; lp.V = h->d_V; lp.gp = h->d_gp; lp.W = h->d_W;
  lp.Xacc = h->d_Xacc; lp.Ucur = h->d_Ucur; lp.split = (h->eff_ranks > 1 && !h->eff_xchg) ? 1 : 0;
  lp.speculate = speculate_decision(h) ? 1 : 0;
  if (h->eff_xchg) lp.xc = h->xc;
  lp.scale_p = h->d_scale_p; lp.Vinv = h->d_Vinv; lp.S = h->d_S;
  lp.s_cap = reduced_capacity(h->cfg.max_frames); lp.Vinv2 = h->d_Vinv2;
  return lp;
}

static int zero_accumulators(pba_handle* h) {
  CUDA_TRY(cudaMemsetAsync(h->d_Xacc, 0, sizeof(double) * ((size_t)h->cfg.max_frames * kUStride + kEacc + kMaxRanks), h->stream));
  return PBA_OK;
}

static void unpack_sym6(const double* u21, double* out36) {
  int k = 0;
  for (int a = 0; a < 6; ++a)
    for (int b = a; b < 6; ++b, ++k) { out36[a * 6 + b] = u21[k]; out36[b * 6 + a] = u21[k]; }
}

int pba_eval(pba_handle* h, pba_eval_out* out) {
  int rc = check_ready(h, "pba_eval");
  if (rc) return rc;
  if (!out) return fail(PBA_ERR_ARGUMENT, "pba_eval: null output");
  CUDA_TRY(cudaSetDevice(h->device));
  const int F = h->n_frames, n = h->n_points, nnz = h->nnz;
  StepParams p = make_step_params(h, nullptr);
  if (out->obs_sqnorm) {
    if (!h->d_obs_sqnorm) CUDA_TRY(cudaMalloc(&h->d_obs_sqnorm, sizeof(double) * h->cfg.max_observations));
    p.obs_sqnorm = h->d_obs_sqnorm;
  }
  if (out->residuals) {
    if (!h->d_residuals) CUDA_TRY(cudaMalloc(&h->d_residuals, sizeof(double) * (size_t)h->cfg.max_observations * h->CP));
    p.residuals = h->d_residuals;
  }
  rc = zero_accumulators(h);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
  CUDA_TRY(launch_k_step(p, h->cfg.patch_radius, h->stream));
  CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  out->device_ms = ms;
  std::vector<double> U((size_t)F * kUStride), E(kEacc);
  CUDA_TRY(cudaMemcpy(U.data(), h->d_Xacc, sizeof(double) * U.size(), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(E.data(), h->d_Xacc + (size_t)F * kUStride, sizeof(double) * kEacc, cudaMemcpyDeviceToHost));
  const double cost = E[0];
  out->cost = cost;
  if (out->U) for (int f = 0; f < F; ++f) unpack_sym6(&U[f * kUStride], out->U + f * 36);
  if (out->gc) for (int f = 0; f < F; ++f) for (int a = 0; a < 6; ++a) out->gc[f * 6 + a] = U[f * kUStride + 21 + a];
  if (out->V) {
    std::vector<double> v((size_t)n * 6);
    CUDA_TRY(cudaMemcpy(v.data(), h->d_V, sizeof(double) * v.size(), cudaMemcpyDeviceToHost));
    for (int q = 0; q < n; ++q) {
      const double* s = &v[(size_t)q * 6];
      double* o = out->V + (size_t)q * 9;
      o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[1]; o[4] = s[3]; o[5] = s[4]; o[6] = s[2]; o[7] = s[4]; o[8] = s[5];
    }
  }
  if (out->gp) CUDA_TRY(cudaMemcpy(out->gp, h->d_gp, sizeof(double) * (size_t)n * 3, cudaMemcpyDeviceToHost));
  if (out->W) CUDA_TRY(cudaMemcpy(out->W, h->d_W, sizeof(double) * (size_t)nnz * 18, cudaMemcpyDeviceToHost));
  if (out->obs_sqnorm) CUDA_TRY(cudaMemcpy(out->obs_sqnorm, h->d_obs_sqnorm, sizeof(double) * nnz, cudaMemcpyDeviceToHost));
  if (out->residuals) CUDA_TRY(cudaMemcpy(out->residuals, h->d_residuals, sizeof(double) * (size_t)nnz * h->CP, cudaMemcpyDeviceToHost));
  return PBA_OK;
}

int pba_eval_timed(pba_handle* h, int32_t iters, double* ms_total) {
  int rc = check_ready(h, "pba_eval_timed");
  if (rc) return rc;
  if (iters < 1 || !ms_total) return fail(PBA_ERR_ARGUMENT, "pba_eval_timed: bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  StepParams p = make_step_params(h, nullptr);
  rc = zero_accumulators(h);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
  for (int i = 0; i < iters; ++i) CUDA_TRY(launch_k_step(p, h->cfg.patch_radius, h->stream));
  CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  *ms_total = ms;
  return PBA_OK;
}

static void format_message(const LmState& s, char* out, size_t cap) {
  switch (s.msg_code) {
    case kMsgGradTol: snprintf(out, cap, "Gradient tolerance reached. Gradient max norm: %e <= %e", s.msg_a, s.msg_b); break;
    case kMsgParamTol: snprintf(out, cap, "Parameter tolerance reached. Relative step_norm: %e <= %e.", s.msg_a, s.msg_b); break;
    case kMsgFuncTol: snprintf(out, cap, "Function tolerance reached. |cost_change|/cost: %e <= %e", s.msg_a, s.msg_b); break;
    case kMsgMaxIter: snprintf(out, cap, "Maximum number of iterations reached. Number of iterations: %d.", (int)s.msg_a); break;
    case kMsgMinRadius: snprintf(out, cap, "Minimum trust region radius reached. Trust region radius: %e <= %e", s.msg_a, s.msg_b); break;
    case kMsgInvalidSteps: snprintf(out, cap, "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps: %d", (int)s.msg_a); break;
    case kMsgXchgTimeout: snprintf(out, cap, "Multi-GPU exchange timed out waiting for a peer (epoch %.0f, %s).", s.msg_a, s.msg_b ? "reduced system" : "pose blocks"); break;
    default: snprintf(out, cap, "no termination message"); break;
  }
}

// (Re)build the device-side LM loop when the kernel parameters changed since the last solve.
// The kernel nodes of a captured linear chain, in launch order.
static bool chain_in_order(cudaGraph_t g, std::vector<cudaGraphNode_t>& out) {
  out.clear();
  size_t n = 0;
  if (cudaGraphGetRootNodes(g, nullptr, &n) != cudaSuccess || n != 1) return false;
  cudaGraphNode_t cur;
  if (cudaGraphGetRootNodes(g, &cur, &n) != cudaSuccess) return false;
  for (;;) {
    cudaGraphNodeType ty;
    if (cudaGraphNodeGetType(cur, &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) return false;
    out.push_back(cur);
    // (_v2: the body's kernel-to-kernel edges are programmatic, which the plain query refuses to report)
    size_t nd = 0;
    if (cudaGraphNodeGetDependentNodes_v2(cur, nullptr, nullptr, &nd) != cudaSuccess) return false;
    if (nd == 0) return true;
    if (nd != 1) return false;
    cudaGraphEdgeData ed;
    if (cudaGraphNodeGetDependentNodes_v2(cur, &cur, &ed, &nd) != cudaSuccess) return false;
  }
}

// Enqueue the body of the LM loop (two iterations: K_B, K_A, K_B, K_A) on the capturing stream.
static cudaError_t enqueue_lm_body(pba_handle* h, const LmParams& lp0, StepParams* sp, int sgrid, int n_free, int pdl) {
  cudaError_t e = cudaSuccess;
  // kernel-to-kernel edges inside the body are programmatic (PDL): the dependent kernel's launch and the part of
  // its prologue that does not read its predecessor's output overlap with the predecessor's tail (PBA_NO_PDL=1: off)
  for (int half = 0; half < 2 && e == cudaSuccess; ++half) {
    LmParams lp = lp0;
    lp.st_in = h->d_state + half; lp.st_out = h->d_state + (1 - half);
    lp.dbg = nullptr; lp.cond = h->lm_cond;
    lp.pdl = half > 0 ? pdl : 0;          // the first kernel of the body follows the loop condition, not a kernel
    e = launch_schur_solve(lp, sgrid, n_free, h->stream);
    sp[half].pdl = pdl;
    if (e == cudaSuccess) e = launch_k_step(sp[half], h->cfg.patch_radius, h->stream);
  }
  return e;
}

static int ensure_lm_graph(pba_handle* h, const LmParams& lp0, int sgrid, int n_free) {
  StepParams sp[2] = {make_step_params(h, h->d_state + 1), make_step_params(h, h->d_state)};
  const int pdl = getenv("PBA_NO_PDL") == nullptr ? 1 : 0;
  std::vector<unsigned char> key(sizeof(sp) + sizeof(LmParams) + 3 * sizeof(int));
  {
    LmParams k = lp0;
    k.cond = 0; k.st_in = nullptr; k.st_out = nullptr; k.dbg = nullptr;
    const int extra[3] = {sgrid, n_free, h->cfg.patch_radius + 16 * pdl};
    memcpy(key.data(), sp, sizeof(sp));
    memcpy(key.data() + sizeof(sp), &k, sizeof(k));
    memcpy(key.data() + sizeof(sp) + sizeof(k), extra, sizeof(extra));
  }
  if (h->lm_exec && key == h->lm_key) return PBA_OK;
  // Same loop, other sizes (a sliding window changes its point count with every frame): capture the four launches
  // again into a scratch graph and move their parameters (arguments, grid, kernel instantiation) into the instantiated
  // graph's nodes in place; anything the driver refuses falls through to a rebuild.
  if (h->lm_exec && h->lm_body_nodes.size() == 4 && getenv("PBA_NO_GRAPH_UPDATE") == nullptr) {
    bool ok = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    cudaGraph_t scratch = nullptr;
    if (ok) {
      const cudaError_t e = enqueue_lm_body(h, lp0, sp, sgrid, n_free, pdl);
      const cudaError_t e2 = cudaStreamEndCapture(h->stream, &scratch);
      ok = e == cudaSuccess && e2 == cudaSuccess && scratch;
    }
    std::vector<cudaGraphNode_t> fresh;
    ok = ok && chain_in_order(scratch, fresh) && fresh.size() == 4;
    const bool dbg = getenv("PBA_DEBUG_GRAPH") != nullptr;
    if (dbg && !ok) fprintf(stderr, "[pba graph] update: scratch capture / chain walk failed (%zu nodes)\n", fresh.size());
    for (int i = 0; i < 4 && ok; ++i) {
      cudaKernelNodeParams kp;
      cudaError_t eg = cudaGraphKernelNodeGetParams(fresh[i], &kp);
      cudaError_t es = eg == cudaSuccess ? cudaGraphExecKernelNodeSetParams(h->lm_exec, h->lm_body_nodes[i], &kp) : eg;
      ok = es == cudaSuccess;
      if (dbg && !ok) fprintf(stderr, "[pba graph] update: node %d: get %s / set %s\n", i, cudaGetErrorString(eg), cudaGetErrorString(es));
    }
    if (scratch) cudaGraphDestroy(scratch);
    cudaGetLastError();
    if (ok) { h->lm_key.swap(key); ++h->lm_updates; return PBA_OK; }
  }
  drop_lm_graph(h);
  CUDA_TRY(cudaGraphCreate(&h->lm_graph, 0));
  cudaGraphConditionalHandle cond;
  CUDA_TRY(cudaGraphConditionalHandleCreate(&cond, h->lm_graph, 1, cudaGraphCondAssignDefault));
  h->lm_cond = (unsigned long long)cond;
  cudaGraphNodeParams np = {};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = cond;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  cudaGraphNode_t node;
  CUDA_TRY(cudaGraphAddNode(&node, h->lm_graph, nullptr, 0, &np));
  cudaGraph_t body = np.conditional.phGraph_out[0];
  CUDA_TRY(cudaStreamBeginCaptureToGraph(h->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  const cudaError_t e = enqueue_lm_body(h, lp0, sp, sgrid, n_free, pdl);
  cudaGraph_t captured = nullptr;
  cudaError_t e2 = cudaStreamEndCapture(h->stream, &captured);
  if (e != cudaSuccess || e2 != cudaSuccess) {
    drop_lm_graph(h);
    return fail(PBA_ERR_CUDA, "pba_solve: capturing the LM loop failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
  }
  if (!chain_in_order(body, h->lm_body_nodes)) {   // no in-place updates then
    if (getenv("PBA_DEBUG_GRAPH")) fprintf(stderr, "[pba graph] body chain walk failed (%zu nodes)\n", h->lm_body_nodes.size());
    h->lm_body_nodes.clear();
  }
  cudaGetLastError();
  CUDA_TRY(cudaGraphInstantiate(&h->lm_exec, h->lm_graph, 0));
  h->lm_key.swap(key);
  ++h->lm_builds;
  return PBA_OK;
}

int pba_solve(pba_handle* h, const pba_solver_options* opt_in, pba_summary* summary) {
  int rc = check_ready(h, "pba_solve");
  if (rc) return rc;
  if (!summary) return fail(PBA_ERR_ARGUMENT, "pba_solve: null summary");
  const auto t_start = std::chrono::steady_clock::now();
  pba_solver_options opt;
  if (opt_in) opt = *opt_in; else pba_default_solver_options(&opt);
  if (opt.max_num_iterations < 0) return fail(PBA_ERR_ARGUMENT, "pba_solve: max_num_iterations < 0");
  CUDA_TRY(cudaSetDevice(h->device));
  const int F = h->n_frames;
  if (h->trace_cap < opt.max_num_iterations + 2) {
    cudaFree(h->d_trace);
    h->d_trace = nullptr;
    h->trace_cap = opt.max_num_iterations + 2;
    CUDA_TRY(cudaMalloc(&h->d_trace, sizeof(IterSummary) * h->trace_cap));
    cudaFree(h->d_stamps);
    h->d_stamps = nullptr;
    CUDA_TRY(cudaMalloc(&h->d_stamps, sizeof(unsigned long long) * 2 * (h->trace_cap + 2)));
    if (h->h_post) { cudaFreeHost(h->h_post); h->h_post = nullptr; }
    auto up = [](size_t x) { return (x + 63) & ~(size_t)63; };
    h->post_trace = 0;
    h->post_stamps = up(sizeof(IterSummary) * h->trace_cap);
    h->post_err = h->post_stamps + up(sizeof(unsigned long long) * 2 * (h->trace_cap + 2));
    h->post_cams = h->post_err + 64;
    h->post_pts = h->post_cams + up(sizeof(double) * 6 * (size_t)h->cfg.max_frames);
    h->post_bytes = h->post_pts + up(sizeof(double) * 3 * (size_t)h->cfg.max_points);
    CUDA_TRY(cudaMallocHost(&h->h_post, h->post_bytes));
  }
  h->results_cached = false;
  CUDA_TRY(cudaMemsetAsync(h->d_stamps, 0, sizeof(unsigned long long) * 2 * (h->trace_cap + 2), h->stream));
  LmState* s = h->h_state;
  memset(s, 0, sizeof(*s));
  s->max_num_iterations = opt.max_num_iterations; s->max_invalid = opt.max_num_consecutive_invalid_steps;
  s->jacobi_scaling = opt.jacobi_scaling;
  s->function_tolerance = opt.function_tolerance; s->gradient_tolerance = opt.gradient_tolerance;
  s->parameter_tolerance = opt.parameter_tolerance; s->initial_radius = opt.initial_trust_region_radius;
  s->max_radius = opt.max_trust_region_radius; s->min_radius = opt.min_trust_region_radius;
  s->min_relative_decrease = opt.min_relative_decrease; s->min_diag = opt.min_lm_diagonal; s->max_diag = opt.max_lm_diagonal;
  s->n_frames = F; s->fixed_frame = h->fixed_frame; s->n_points = h->n_points; s->nnz = h->nnz;
  s->n_free = 0;
  for (int f = 0; f < kMaxFrames; ++f) s->free_index[f] = -1;
  // a camera is a parameter block only if some residual block uses it (src/photobundle.cc:802)
  for (int f = 0; f < F; ++f)
    if (h->frame_used[f] && f != h->fixed_frame) s->free_index[f] = s->n_free++;
  s->cur = 0; s->eval_buf = 0; s->decrease_factor = 2.0; s->radius = opt.initial_trust_region_radius;
  if (h->eff_xchg) {   // every rank advances by the same amount per solve: flags never need a reset
    s->xepoch = h->epoch_next;
    h->epoch_next += (unsigned long long)opt.max_num_iterations + 16ull;
  }
  // x lives in buffer 0 of cams/points; refresh buffer 1 so fixed cameras carry over
  CUDA_TRY(cudaMemcpyAsync(h->d_cams + (size_t)F * 6, h->d_cams, sizeof(double) * F * 6, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_state, s, sizeof(LmState), cudaMemcpyHostToDevice, h->stream));
  rc = zero_accumulators(h);
  if (rc) return rc;
  const size_t D = 6 * (size_t)F;
  CUDA_TRY(cudaMemsetAsync(h->d_S, 0, sizeof(double) * kSCopies * reduced_capacity(h->cfg.max_frames), h->stream));
  CUDA_TRY(cudaMemsetAsync(h->d_ticket, 0, sizeof(unsigned int), h->stream));

  LmParams lp = make_lm_params(h);
  const bool timeline = getenv("PBA_DEBUG_TIMELINE") != nullptr;
  if (timeline && !h->d_dbg) CUDA_TRY(cudaMalloc(&h->d_dbg, sizeof(unsigned long long) * 16 * 1024));
  const int sgrid = (h->eff_xchg && lp.speculate) ? schur_grid_x(h->n_points, h->sm_count) : schur_grid(h->n_points, h->sm_count);
  int launches = 0;
  const bool multi = h->eff_ranks > 1 && !h->eff_xchg;   // NCCL all-reduce path; the peer-memory exchange needs no host-side calls
  // one GPU: the whole loop runs on the device (WHILE node); PBA_NO_GRAPH=1 keeps the stream loop
  const bool use_graph = !multi && !timeline && getenv("PBA_NO_GRAPH") == nullptr;
  if (use_graph) {
    rc = ensure_lm_graph(h, lp, sgrid, s->n_free);
    if (rc) return rc;
  }
  if (h->group && h->eff_xchg) {
    std::lock_guard<std::mutex> lk(h->group->m);
    if ((int)h->group->members.size() != h->n_ranks)
      return fail(PBA_ERR_STATE, "pba_solve: a member of the local communicator has been destroyed (its exchange buffer is gone)");
  }
  if (h->group && h->eff_xchg) {   // ranks of one process: every allocation above is done before any rank launches a kernel that waits on a peer
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (!h->group->arrive_and_wait(60))
      return fail(PBA_ERR_STATE, "pba_solve: the other members of the local communicator did not call pba_solve (one thread per handle, concurrently)");
  }
  if (h->eff_xchg) {   // align the ranks before the clock starts (device-side barrier over peer memory)
    CUDA_TRY(launch_rendezvous(h->xc, s->xepoch, h->stream));
    launches += 1;
  }
  CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
  // iteration 0: evaluate x.  Then per LM iteration: K_B (decide + Schur + solve), K_A
  // (back-substitute + evaluate the candidate).  The state ping-pongs between two structs;
  // kernels turn into no-ops once `done` is set, so iterations are enqueued in groups.
  const size_t xacc_n = (size_t)F * kUStride + kEacc + kMaxRanks;
  int collectives = 0;
  CUDA_TRY(launch_k_step(make_step_params(h, h->d_state), h->cfg.patch_radius, h->stream));
  launches += 1;
  if (multi) {   // candidate's pose blocks + cost scalars, summed over the point shards
    NCCL_TRY(g_nccl.AllReduce(h->d_Xacc, h->d_Xacc, xacc_n, ncclDouble, ncclSum, h->comm, h->stream));
    ++collectives;
  }
  const int group = 4;
  bool done = false;
  int k = 0;
  const size_t stamps_bytes = sizeof(unsigned long long) * 2 * (h->trace_cap + 2);
  bool posted = false;
  if (use_graph) {
    // everything the host needs from this solve is enqueued behind the loop and waited for ONCE: final state, iteration
    // trace, K_B stamps, the exchange's error word, and the results themselves (k_publish moves the accepted x into
    // buffer 0 on the device; sharded points are gathered by pba_get_points instead)
    CUDA_TRY(cudaGraphLaunch(h->lm_exec, h->stream));
    CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
    CUDA_TRY(launch_publish(h->d_state, h->d_cams, h->d_pts, F, h->n_points, h->stream));
    CUDA_TRY(cudaMemcpyAsync(s, h->d_state, sizeof(LmState), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->h_post + h->post_trace, h->d_trace, sizeof(IterSummary) * h->trace_cap, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->h_post + h->post_stamps, h->d_stamps, stamps_bytes, cudaMemcpyDeviceToHost, h->stream));
    if (h->eff_xchg) CUDA_TRY(cudaMemcpyAsync(h->h_post + h->post_err, h->xc.error, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->h_post + h->post_cams, h->d_cams, sizeof(double) * 6 * (size_t)F, cudaMemcpyDeviceToHost, h->stream));
    if (h->eff_ranks == 1)
      CUDA_TRY(cudaMemcpyAsync(h->h_post + h->post_pts, h->d_pts, sizeof(double) * 3 * (size_t)h->n_points, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (!s->done) return fail(PBA_ERR_CUDA, "pba_solve: the LM graph returned before the minimizer terminated");
    launches += 4 * ((s->num_evals + 1) / 2) + 1;   // the body (two iterations, four kernels) ran ceil(decisions / 2) times; k_publish
    done = true; posted = true;
  }
  while (!done) {
    for (int g = 0; g < group; ++g, ++k) {
      lp.st_in = h->d_state + (k & 1);
      lp.st_out = h->d_state + ((k + 1) & 1);
      lp.dbg = (timeline && k < 1024) ? h->d_dbg + 16 * k : nullptr;
      CUDA_TRY(launch_schur_solve(lp, sgrid, s->n_free, h->stream));
      launches += 1;
      if (multi) {   // reduced camera system, summed over the point shards, then the (replicated) solve
        NCCL_TRY(g_nccl.AllReduce(h->d_S, h->d_S, (size_t)(6 * s->n_free) * reduced_ld(6 * s->n_free), ncclDouble, ncclSum, h->comm, h->stream));
        CUDA_TRY(launch_solve_only(lp, s->n_free, h->stream));
        ++collectives; launches += 1;
      }
      CUDA_TRY(launch_k_step(make_step_params(h, lp.st_out), h->cfg.patch_radius, h->stream));
      launches += 1;
      if (multi) {
        NCCL_TRY(g_nccl.AllReduce(h->d_Xacc, h->d_Xacc, xacc_n, ncclDouble, ncclSum, h->comm, h->stream));
        ++collectives;
      }
    }
    CUDA_TRY(cudaMemcpyAsync(s, h->d_state + (k & 1), sizeof(LmState), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    done = s->done != 0 || k > opt.max_num_iterations + 2;
  }
  if (!use_graph) CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  // the accepted x must end up in buffer 0 for pba_get_* and for the next solve
  if (!posted && s->cur != 0) {
    CUDA_TRY(cudaMemcpyAsync(h->d_cams, h->d_cams + (size_t)F * 6, sizeof(double) * F * 6, cudaMemcpyDeviceToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_pts, h->d_pts + (size_t)h->n_points * 3, sizeof(double) * (size_t)h->n_points * 3, cudaMemcpyDeviceToDevice, h->stream));
  }
  if (timeline) {
    std::vector<unsigned long long> t(16 * (size_t)std::min(k, 1024));
    CUDA_TRY(cudaMemcpy(t.data(), h->d_dbg, sizeof(unsigned long long) * t.size(), cudaMemcpyDeviceToHost));
    for (int i = 0; i < std::min(k, 16); ++i)
      fprintf(stderr, "[pba timeline] K_B %2d: decide %6.2f us  schur %6.2f us  solve %6.2f us (assemble %5.2f factor %5.2f subst %5.2f final %5.2f) tail %5.2f us | gap to next K_B start %7.2f us\n", i,
              (t[16 * i + 1] - t[16 * i]) * 1e-3, (t[16 * i + 2] - t[16 * i + 1]) * 1e-3, (t[16 * i + 3] - t[16 * i + 2]) * 1e-3,
              (t[16 * i + 5] - t[16 * i + 2]) * 1e-3, (t[16 * i + 6] - t[16 * i + 5]) * 1e-3, (t[16 * i + 7] - t[16 * i + 6]) * 1e-3,
              (t[16 * i + 3] - t[16 * i + 7]) * 1e-3,
              (t[16 * i + 4] - t[16 * i + 3]) * 1e-3, i + 1 < k ? (t[16 * (i + 1)] - t[16 * i + 4]) * 1e-3 : 0.0);
    for (int i = 1; i < std::min(k, 4); ++i)
      if (t[16 * i + 12])
        fprintf(stderr, "[pba timeline] K_B %2d eliminate (last CTA): setup %5.2f us  points %5.2f us  warp merge %5.2f us  atomics+ticket %5.2f us\n", i,
                (t[16 * i + 12] - t[16 * i + 1]) * 1e-3, (t[16 * i + 13] - t[16 * i + 12]) * 1e-3, (t[16 * i + 14] - t[16 * i + 13]) * 1e-3,
                (t[16 * i + 2] - t[16 * i + 14]) * 1e-3);
    if (h->eff_xchg)
      for (int i = 0; i < std::min(k, 14); ++i)
        fprintf(stderr, "[pba timeline] rank %d K_B %2d exchange: eliminate (both hypotheses) %5.2f us  push %5.2f us  sum+decide %5.2f us  sum P %5.2f us  -> %s\n", h->rank, i,
                (t[16 * i + 1] - t[16 * i]) * 1e-3, (t[16 * i + 8] - t[16 * i + 1]) * 1e-3, (t[16 * i + 9] - t[16 * i + 8]) * 1e-3,
                (t[16 * i + 2] - t[16 * i + 9]) * 1e-3, t[16 * i + 10] == 1 ? "accepted, radius tripled" : t[16 * i + 10] == 2 ? "rejected" : "second elimination + exchange");
  }
  h->trace.resize(s->n_trace);
  if (posted) {
    if (s->n_trace > 0) memcpy(h->trace.data(), h->h_post + h->post_trace, sizeof(IterSummary) * s->n_trace);
  } else {
    if (s->n_trace > 0)
      CUDA_TRY(cudaMemcpyAsync(h->trace.data(), h->d_trace, sizeof(IterSummary) * s->n_trace, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->h_post + h->post_stamps, h->d_stamps, stamps_bytes, cudaMemcpyDeviceToHost, h->stream));
    if (h->eff_xchg) CUDA_TRY(cudaMemcpyAsync(h->h_post + h->post_err, h->xc.error, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));

  memset(summary, 0, sizeof(*summary));
  summary->initial_cost = s->initial_cost; summary->final_cost = s->x_cost; summary->fixed_cost = 0.0;
  summary->num_successful_steps = s->num_successful; summary->num_unsuccessful_steps = s->num_unsuccessful;
  summary->num_residual_blocks = h->nnz_total; summary->num_residuals = h->nnz_total * h->CP;
  summary->num_iterations = s->n_trace; summary->termination_type = s->termination_type;
  summary->num_evaluations = s->num_evals;
  summary->kernel_launches = launches; summary->num_collectives = h->eff_xchg ? s->n_xchg : collectives;
  summary->device_time_in_seconds = ms * 1e-3;
  {
    const unsigned long long* st2 = reinterpret_cast<const unsigned long long*>(h->h_post + h->post_stamps);
    double kb = 0.0;
    for (int i = 0; i < s->num_evals; ++i)
      if (st2[2 * i] && st2[2 * i + 1] > st2[2 * i]) kb += (double)(st2[2 * i + 1] - st2[2 * i]) * 1e-9;
    summary->kb_device_time_in_seconds = kb;
  }
  if (h->eff_xchg) {
    const int xerr = *reinterpret_cast<const int*>(h->h_post + h->post_err);
    if (xerr) {
      CUDA_TRY(cudaMemset(h->xc.error, 0, sizeof(int)));
      return fail(PBA_ERR_NCCL, "pba_solve: the peer-memory exchange timed out (a rank did not reach the same LM iteration)");
    }
  }
  if (!s->done) { s->msg_code = kMsgMaxIter; s->msg_a = opt.max_num_iterations; summary->termination_type = 1; }
  h->defer_sync = false;   // every upload enqueued before this solve has been consumed
  h->results_cached = posted;
  format_message(*s, summary->message, sizeof(summary->message));
  summary->total_time_in_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
  return PBA_OK;
}

int pba_save_state(pba_handle* h) {
  if (!h) return fail(PBA_ERR_ARGUMENT, "pba_save_state: null handle");
  if (!h->have_poses || !h->have_points) return fail(PBA_ERR_STATE, "pba_save_state: poses/points not set");
  CUDA_TRY(cudaSetDevice(h->device));
  if (!h->d_save_cams) {
    CUDA_TRY(cudaMalloc(&h->d_save_cams, sizeof(double) * (size_t)h->cfg.max_frames * 6));
    CUDA_TRY(cudaMalloc(&h->d_save_pts, sizeof(double) * (size_t)h->cfg.max_points * 3));
  }
  CUDA_TRY(cudaMemcpyAsync(h->d_save_cams, h->d_cams, sizeof(double) * (size_t)h->n_frames * 6, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_save_pts, h->d_pts, sizeof(double) * (size_t)h->n_points * 3, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->have_saved = true;
  return PBA_OK;
}

int pba_restore_state(pba_handle* h) {
  if (h) h->results_cached = false;
  if (!h) return fail(PBA_ERR_ARGUMENT, "pba_restore_state: null handle");
  if (!h->have_saved) return fail(PBA_ERR_STATE, "pba_restore_state: nothing saved");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpyAsync(h->d_cams, h->d_save_cams, sizeof(double) * (size_t)h->n_frames * 6, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_pts, h->d_save_pts, sizeof(double) * (size_t)h->n_points * 3, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return PBA_OK;
}

int pba_copy_state(pba_handle* dst, pba_handle* src) {
  if (dst) dst->results_cached = false;
  if (!dst || !src) return fail(PBA_ERR_ARGUMENT, "pba_copy_state: null handle");
  if (!src->have_poses || !src->have_points || !dst->have_poses || !dst->have_points)
    return fail(PBA_ERR_STATE, "pba_copy_state: both handles need poses and points (the copy replaces their values, not their layout)");
  if (src->n_frames != dst->n_frames || src->n_points != dst->n_points || src->n_points_total != dst->n_points_total)
    return fail(PBA_ERR_ARGUMENT, "pba_copy_state: %d frames / %d points into %d frames / %d points", src->n_frames, src->n_points, dst->n_frames, dst->n_points);
  CUDA_TRY(cudaSetDevice(src->device));
  CUDA_TRY(cudaStreamSynchronize(src->stream));    // the source solve has finished (pba_solve returns synchronised; cheap)
  CUDA_TRY(cudaSetDevice(dst->device));
  const size_t cb = sizeof(double) * (size_t)src->n_frames * 6, pb = sizeof(double) * (size_t)src->n_points * 3;
  if (src->device == dst->device) {
    CUDA_TRY(cudaMemcpyAsync(dst->d_cams, src->d_cams, cb, cudaMemcpyDeviceToDevice, dst->stream));
    CUDA_TRY(cudaMemcpyAsync(dst->d_pts, src->d_pts, pb, cudaMemcpyDeviceToDevice, dst->stream));
  } else {
    CUDA_TRY(cudaMemcpyPeerAsync(dst->d_cams, dst->device, src->d_cams, src->device, cb, dst->stream));
    CUDA_TRY(cudaMemcpyPeerAsync(dst->d_pts, dst->device, src->d_pts, src->device, pb, dst->stream));
  }
  // buffer 1 of the points mirrors buffer 0 until the first accepted step (as after pba_set_points)
  CUDA_TRY(cudaMemcpyAsync(dst->d_pts + (size_t)dst->n_points * 3, dst->d_pts, pb, cudaMemcpyDeviceToDevice, dst->stream));
  return PBA_OK;   // ordered on dst's stream: the next pba_solve(dst) sees it
}

int pba_begin_batch(pba_handle* h) {
  if (!h) return fail(PBA_ERR_ARGUMENT, "pba_begin_batch: null handle");
  h->defer_sync = true;
  return PBA_OK;
}

int pba_end_batch(pba_handle* h) {
  if (!h) return fail(PBA_ERR_ARGUMENT, "pba_end_batch: null handle");
  h->defer_sync = false;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return PBA_OK;
}

int pba_set_frame_u8_ex(pba_handle* h, int32_t slot, const uint8_t* image, int32_t src_rows, int32_t src_cols, int32_t levels_down,
                        int32_t descriptor_type) {
  if (!h || !image) return fail(PBA_ERR_ARGUMENT, "pba_set_frame_u8_ex: null argument");
  const int n_slots = h->have_frames ? h->n_frames : h->cfg.max_frames;   // a cold handle: the ring has max_frames slots
  if (slot < 0 || slot >= n_slots) return fail(PBA_ERR_ARGUMENT, "pba_set_frame_u8_ex: slot %d of %d frames", slot, n_slots);
  const int C = pba_descriptor_channels(descriptor_type);
  if (C < 0 || C != h->cfg.n_channels) return fail(PBA_ERR_ARGUMENT, "pba_set_frame_u8_ex: descriptor type %d does not match the handle's %d channels", descriptor_type, h->cfg.n_channels);
  if (levels_down < 0 || (levels_down > 0 && C != 1)) return fail(PBA_ERR_ARGUMENT, "pba_set_frame_u8_ex: pyramid levels are built for the Intensity descriptor only");
  int r = src_rows, cc = src_cols;
  for (int l = 0; l < levels_down; ++l) { r = (r + 1) / 2; cc = (cc + 1) / 2; }
  if (r != h->cfg.rows || cc != h->cfg.cols)
    return fail(PBA_ERR_ARGUMENT, "pba_set_frame_u8_ex: %dx%d reduced %d times is %dx%d, handle is %dx%d", src_rows, src_cols, levels_down, r, cc, h->cfg.rows, h->cfg.cols);
  CUDA_TRY(cudaSetDevice(h->device));
  if (!h->have_frames) {   // no window yet: the slots not written so far read as black frames
    if (!h->d_u8) {
      CUDA_TRY(cudaMalloc(&h->d_u8, (size_t)h->cfg.max_frames * h->plane + 64));
      CUDA_TRY(cudaMemsetAsync(h->d_u8, 0, (size_t)h->cfg.max_frames * h->plane, h->stream));
    }
    if (C > 1 && !h->d_f32) {
      CUDA_TRY(cudaMalloc(&h->d_f32, sizeof(float) * (size_t)h->cfg.max_frames * C * h->plane + 64));
      CUDA_TRY(cudaMemsetAsync(h->d_f32, 0, sizeof(float) * (size_t)h->cfg.max_frames * C * h->plane, h->stream));
    }
    h->frames_are_u8 = (C == 1); h->have_frames = true; h->n_frames = h->cfg.max_frames;
  }
  if (levels_down == 0) {
    int rc = upload_plane_u8(h, slot, image);
    if (rc) return rc;
    if (C > 1) {
      rc = ensure_descriptor_scratch(h);
      if (rc) return rc;
      CUDA_TRY(launch_channels(descriptor_type, h->d_u8 + (size_t)slot * h->plane, h->cfg.rows, h->cfg.cols, h->pitch, h->d_scr_a, h->d_scr_b,
                               h->d_f32 + (size_t)slot * C * h->plane, h->pitch, h->plane, h->stream));
    }
  } else {
    const int p0 = (src_cols + 15) / 16 * 16;
    if (h->pyr_scratch_bytes < 2 * (size_t)src_rows * p0) {
      cudaFree(h->d_pyr_scratch); h->d_pyr_scratch = nullptr; h->pyr_scratch_bytes = 0;
      CUDA_TRY(cudaMalloc(&h->d_pyr_scratch, 2 * (size_t)src_rows * p0));
      h->pyr_scratch_bytes = 2 * (size_t)src_rows * p0;
    }
    uint8_t* a = h->d_pyr_scratch;
    uint8_t* b = h->d_pyr_scratch + (size_t)src_rows * p0;
    CUDA_TRY(cudaMemcpyAsync(a, image, (size_t)src_rows * src_cols, cudaMemcpyHostToDevice, h->stream));
    int rr = src_rows, wc = src_cols, pa = src_cols;
    for (int l = 0; l < levels_down; ++l) {
      const bool last = (l == levels_down - 1);
      uint8_t* dst = last ? h->d_u8 + (size_t)slot * h->plane : b;
      const int pd = last ? h->pitch : p0;
      CUDA_TRY(launch_pyrdown_u8(a, rr, wc, pa, dst, pd, h->stream));
      rr = (rr + 1) / 2; wc = (wc + 1) / 2; pa = pd;
      std::swap(a, b);
    }
  }
  return upload_done(h);
}

int pba_get_results(pba_handle* h, double* cam6, double* xyz) {
  if (!h || !cam6 || !xyz) return fail(PBA_ERR_ARGUMENT, "pba_get_results: null argument");
  if (!h->have_poses || !h->have_points) return fail(PBA_ERR_STATE, "pba_get_results: poses/points not set");
  if (h->eff_ranks > 1) {   // sharded points: gather (pba_get_points), poses are replicated
    int rc = pba_get_poses(h, cam6);
    return rc ? rc : pba_get_points(h, xyz);
  }
  if (h->results_cached) {   // the last pba_solve already brought them to the host together with its summary
    memcpy(cam6, h->h_post + h->post_cams, sizeof(double) * 6 * (size_t)h->n_frames);
    memcpy(xyz, h->h_post + h->post_pts, sizeof(double) * 3 * (size_t)h->n_points);
    return PBA_OK;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpyAsync(cam6, h->d_cams, sizeof(double) * (size_t)h->n_frames * 6, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaMemcpyAsync(xyz, h->d_pts, sizeof(double) * (size_t)h->n_points * 3, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return PBA_OK;
}

int pba_get_poses(pba_handle* h, double* cam6) {
  if (!h || !cam6) return fail(PBA_ERR_ARGUMENT, "pba_get_poses: null argument");
  if (!h->have_poses) return fail(PBA_ERR_STATE, "pba_get_poses: poses not set");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpy(cam6, h->d_cams, sizeof(double) * (size_t)h->n_frames * 6, cudaMemcpyDeviceToHost));
  return PBA_OK;
}

int pba_get_points(pba_handle* h, double* xyz) {
  if (!h || !xyz) return fail(PBA_ERR_ARGUMENT, "pba_get_points: null argument");
  if (!h->have_points) return fail(PBA_ERR_STATE, "pba_get_points: points not set");
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->eff_ranks == 1) {
    CUDA_TRY(cudaMemcpy(xyz, h->d_pts, sizeof(double) * (size_t)h->n_points * 3, cudaMemcpyDeviceToHost));
    return PBA_OK;
  }
  // gather the shards: every rank broadcasts its block into a full-size device array
  if (!h->d_pts_full) CUDA_TRY(cudaMalloc(&h->d_pts_full, sizeof(double) * (size_t)h->cfg.max_points * 3));
  double* full = h->d_pts_full;
  CUDA_TRY(cudaMemcpyAsync(full + (size_t)h->shard_begin[h->rank] * 3, h->d_pts, sizeof(double) * (size_t)h->n_points * 3,
                           cudaMemcpyDeviceToDevice, h->stream));
  if (h->group) {   // ranks of this process: read the peers' shards directly (their solves have returned)
    std::vector<pba_handle*> mem;
    { std::lock_guard<std::mutex> lk(h->group->m); mem = h->group->members; }
    if ((int)mem.size() != h->eff_ranks) return fail(PBA_ERR_STATE, "pba_get_points: a member of the local communicator has been destroyed");
    for (int r = 0; r < h->eff_ranks; ++r) {
      const size_t cnt = (size_t)(h->shard_begin[r + 1] - h->shard_begin[r]) * 3;
      if (!cnt || r == h->rank) continue;
      if (mem[r]->n_points * 3 != (int)cnt) return fail(PBA_ERR_STATE, "pba_get_points: rank %d holds %d points, %zu expected (pba_set_points on every member first)", r, mem[r]->n_points, cnt / 3);
      CUDA_TRY(cudaMemcpyPeerAsync(full + (size_t)h->shard_begin[r] * 3, h->device, mem[r]->d_pts, mem[r]->device, sizeof(double) * cnt, h->stream));
    }
  } else {
    NCCL_TRY(g_nccl.GroupStart());
    for (int r = 0; r < h->eff_ranks; ++r) {
      const size_t cnt = (size_t)(h->shard_begin[r + 1] - h->shard_begin[r]) * 3;
      if (cnt) NCCL_TRY(g_nccl.Broadcast(full + (size_t)h->shard_begin[r] * 3, full + (size_t)h->shard_begin[r] * 3, cnt, ncclDouble, r, h->comm, h->stream));
    }
    NCCL_TRY(g_nccl.GroupEnd());
  }
  CUDA_TRY(cudaMemcpyAsync(xyz, full, sizeof(double) * (size_t)h->n_points_total * 3, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return PBA_OK;
}

int pba_get_iterations(pba_handle* h, pba_iteration_summary* out, int32_t capacity, int32_t* n) {
  if (!h || !n) return fail(PBA_ERR_ARGUMENT, "pba_get_iterations: null argument");
  static_assert(sizeof(pba_iteration_summary) == sizeof(IterSummary), "trace layout mismatch");
  *n = (int32_t)h->trace.size();
  if (out) {
    const int m = std::min<int>(capacity, *n);
    if (m > 0) memcpy(out, h->trace.data(), sizeof(IterSummary) * m);
  }
  return PBA_OK;
}

int pba_comm_unique_id(void* id128) {
  if (!id128) return fail(PBA_ERR_ARGUMENT, "pba_comm_unique_id: null argument");
  if (const char* e = load_nccl()) return fail(PBA_ERR_NCCL, "pba_comm_unique_id: %s", e);
  static_assert(sizeof(ncclUniqueId) == PBA_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  NCCL_TRY(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return PBA_OK;
}

int pba_comm_exchange_kind(const pba_handle* h) {
  if (!h || h->n_ranks <= 1) return PBA_EXCHANGE_NONE;
  return h->use_xchg ? PBA_EXCHANGE_PEER : PBA_EXCHANGE_NCCL;
}

int pba_comm_sharded(const pba_handle* h) { return (h && h->have_points && h->eff_ranks > 1) ? 1 : 0; }

int pba_comm_speculates(const pba_handle* h) { return (h && speculate_decision(h)) ? 1 : 0; }

int pba_comm_init(pba_handle* h, const void* id128, int32_t rank, int32_t n_ranks) {
  if (!h || !id128) return fail(PBA_ERR_ARGUMENT, "pba_comm_init: null argument");
  if (n_ranks < 1 || n_ranks > kMaxRanks || rank < 0 || rank >= n_ranks)
    return fail(PBA_ERR_ARGUMENT, "pba_comm_init: rank %d of %d (at most %d ranks)", rank, n_ranks, kMaxRanks);
  if (h->have_points) return fail(PBA_ERR_STATE, "pba_comm_init: call before pba_set_points (points are sharded at upload)");
  if (const char* e = load_nccl()) return fail(PBA_ERR_NCCL, "pba_comm_init: %s", e);
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->comm) { g_nccl.CommDestroy(h->comm); h->comm = nullptr; }
  if (h->group) return fail(PBA_ERR_STATE, "pba_comm_init: the handle belongs to a local communicator (pba_comm_init_local)");
  h->rank = 0; h->n_ranks = 1;
  if (n_ranks > 1) {
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NCCL_TRY(g_nccl.CommInitRank(&h->comm, n_ranks, id, rank));
    h->rank = rank; h->n_ranks = n_ranks;
    int rc = setup_xchg(h);
    if (rc) return rc;
  }
  h->replicated = false; h->eff_ranks = h->n_ranks; h->eff_xchg = h->use_xchg;   // until pba_set_points has seen the window
  return PBA_OK;
}

void* pba_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (bytes == 0 || cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void pba_host_free(void* p) { if (p) cudaFreeHost(p); }

int pba_graph_counters(const pba_handle* h, int32_t* builds, int32_t* updates) {
  if (!h || !builds || !updates) return fail(PBA_ERR_ARGUMENT, "pba_graph_counters: null argument");
  *builds = h->lm_builds; *updates = h->lm_updates;
  return PBA_OK;
}

int pba_comm_init_local(pba_handle* const* handles, int32_t n) {
  if (!handles || n < 1 || n > kMaxRanks) return fail(PBA_ERR_ARGUMENT, "pba_comm_init_local: %d handles (1..%d)", n, kMaxRanks);
  for (int i = 0; i < n; ++i) {
    pba_handle* h = handles[i];
    if (!h) return fail(PBA_ERR_ARGUMENT, "pba_comm_init_local: handles[%d] is null", i);
    if (h->have_points) return fail(PBA_ERR_STATE, "pba_comm_init_local: call before pba_set_points (points are sharded at upload)");
    if (h->comm || h->group || h->n_ranks != 1) return fail(PBA_ERR_STATE, "pba_comm_init_local: handles[%d] already belongs to a communicator", i);
    if (h->cfg.max_frames != handles[0]->cfg.max_frames) return fail(PBA_ERR_ARGUMENT, "pba_comm_init_local: handles differ in max_frames");
    for (int j = 0; j < i; ++j)
      if (handles[j] == h || handles[j]->device == h->device)
        return fail(PBA_ERR_ARGUMENT, "pba_comm_init_local: handles[%d] and handles[%d] are on the same device %d", j, i, h->device);
  }
  if (n == 1) return PBA_OK;
  for (int i = 0; i < n; ++i) {
    CUDA_TRY(cudaSetDevice(handles[i]->device));
    for (int j = 0; j < n; ++j) {
      if (i == j) continue;
      int can = 0;
      CUDA_TRY(cudaDeviceCanAccessPeer(&can, handles[i]->device, handles[j]->device));
      if (!can) return fail(PBA_ERR_STATE, "pba_comm_init_local: device %d cannot access device %d (use one process per GPU and pba_comm_init)", handles[i]->device, handles[j]->device);
      cudaError_t e = cudaDeviceEnablePeerAccess(handles[j]->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) return fail(PBA_ERR_CUDA, "pba_comm_init_local: cudaDeviceEnablePeerAccess(%d -> %d): %s", handles[i]->device, handles[j]->device, cudaGetErrorString(e));
    }
  }
  const XchgLayout lay = xchg_layout(handles[0]->cfg.max_frames, n);
  for (int i = 0; i < n; ++i) {
    pba_handle* h = handles[i];
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->d_xchg) { cudaFree(h->d_xchg); h->d_xchg = nullptr; }
    CUDA_TRY(cudaMalloc(&h->d_xchg, lay.bytes));
    CUDA_TRY(cudaMemset(h->d_xchg, 0, lay.bytes));
  }
  auto group = std::make_shared<LocalGroup>();
  group->members.assign(handles, handles + n);
  for (int i = 0; i < n; ++i) {
    pba_handle* h = handles[i];
    CUDA_TRY(cudaSetDevice(h->device));
    h->rank = i; h->n_ranks = n; h->group = group;
    for (int q = 0; q < kMaxRanks; ++q) h->peer_xchg[q] = q < n ? handles[q]->d_xchg : nullptr;
    int rc = adopt_xchg(h, lay);
    if (rc) return rc;
    h->replicated = false; h->eff_ranks = n; h->eff_xchg = true;
  }
  return PBA_OK;
}

}  // extern "C"
