// The handle behind the C ABI (include/rsba_cuda.h): host-side problem builder that mirrors
// what CeresHandler::Add feeds to ceres::Problem (CeresHandler.h:208-302, 335-382), plus
// the device-resident state of the evaluator and the LM solver.
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "api_guard.h"
#include "common.cuh"
#include "host_threads.h"
#include "../../include/rsba_cuda.h"

namespace rsba {

void set_last_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define RSBA_CUDA_TRY(expr)                                                        \
  do {                                                                             \
    cudaError_t e__ = (expr);                                                      \
    if (e__ != cudaSuccess) return ::rsba::cuda_fail(e__, #expr, __FILE__, __LINE__); \
  } while (0)

// Device memory comes from the device's default stream-ordered pool with the release threshold lifted: the
// reference builds a new CeresHandler for every BA call (VideoSfMHandler.cc:585), i.e. a fresh problem handle per
// call in one long-lived process, and cudaMalloc / cudaFree of the handle's ~3 GB were 40-380 ms of a 240 ms cold
// call at C3 (profiles/r02_notes.md).  The pool keeps what a handle frees for the next one.  Allocation and release
// are ordered on the legacy default stream behind a device-wide synchronise, which is what cudaFree implied.
// RSBA_CUDA_NO_POOL=1 goes back to cudaMalloc / cudaFree.
bool device_pool_enabled();   // problem.cu: also lifts the release threshold of the current device's pool, once

// Simple owning device buffer.
template <typename T>
struct DeviceBuffer {
  T* ptr = nullptr;
  size_t count = 0;
  bool pooled = false;
  ~DeviceBuffer() { release(); }
  DeviceBuffer() = default;
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  void release() {
    if (ptr) {
      if (pooled) {
        cudaDeviceSynchronize();          // nothing in flight on any stream may still use the buffer
        cudaFreeAsync(ptr, 0);
      } else {
        cudaFree(ptr);
      }
    }
    ptr = nullptr;
    count = 0;
  }
  cudaError_t resize(size_t n) {
    if (n == count) return cudaSuccess;
    release();
    if (n == 0) return cudaSuccess;
    cudaError_t e;
    pooled = device_pool_enabled();
    if (pooled) {
      e = cudaMallocAsync(reinterpret_cast<void**>(&ptr), n * sizeof(T), 0);
      if (e == cudaSuccess) e = cudaStreamSynchronize(0);   // usable from every stream from here on
    } else {
      e = cudaMalloc(&ptr, n * sizeof(T));
    }
    if (e == cudaSuccess) count = n;
    else ptr = nullptr;
    return e;
  }
  size_t bytes() const { return count * sizeof(T); }
};

struct StageTimer {
  cudaEvent_t beg = nullptr, end = nullptr;
  bool pending = false;
  double last_ms = 0.0, total_ms = 0.0;
};

enum Stage { kStageJacobian = 0, kStageResidual, kStageSchur, kStageCholesky, kStageUpdate,
             kStageAllreduce,
             // single kernels inside the stages above (nested events), for the per-kernel rooflines
             kStagePointBlocks, kStageFrameBlocks, kStagePhiBuild, kStageSchurSyrk, kStageSchurReduce,
             kStageFactor, kStageTriSolve, kStagePointStep, kStageFinalize, kStagePnp, kNumStages };

struct LmState;  // solver-side device state (lm_solver.cu)

}  // namespace rsba

struct rsba_problem {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  rsba::CameraModel cm{};
  bool camera_set = false;

  // ---- pointer-identity builder (AddResidualBlock / SetParameterBlockConstant)
  struct PtrObs { double x, y; int frame, point; };
  std::vector<PtrObs> ptr_obs;
  std::unordered_map<const double*, int> pose0_to_frame, pose1_to_frame, point_to_id;
  std::vector<double*> frame_pose0, frame_pose1, point_ptr;
  std::vector<unsigned short> ptr_pose_mask;
  std::vector<unsigned char> ptr_point_const;
  bool ptr_mode = false;
  bool ptr_dirty = false;

  // ---- finalised scene (sorted by frame).  g_* = the whole scene; h_* / n_obs / the device
  // arrays = this rank's share (identical to g_* on one GPU): all observations of the points
  // the rank owns (point-owner rule)
  long n_obs = 0, n_obs_global = 0;
  int n_frames = 0, n_points = 0;
  rsba::HostVec<long> order;               // sorted global position -> caller's observation index
  rsba::HostVec<double2> g_obs_xy;
  rsba::HostVec<int> g_obs_frame, g_obs_point;
  rsba::HostVec<long> local_ids;           // local observation -> sorted global position
  rsba::HostVec<int> h_obs_frame, h_obs_point;  // local share, host copy (structure analysis)
  std::vector<unsigned char> point_owned;  // [points] 1 if this rank eliminates the point
  std::vector<unsigned short> pose_mask;   // [frames] constant-scalar bits
  std::vector<unsigned char> point_const;  // [points]
  bool scene_set = false, params_set = false;

  rsba::DeviceBuffer<double2> d_obs_xy;
  rsba::DeviceBuffer<int> d_obs_frame, d_obs_point;
  rsba::DeviceBuffer<double> d_poses, d_points;
  rsba::DeviceBuffer<double> d_res, d_jac;
  // the solver's own linearisation: compact Jacobian records [N][12] and tau [N] (rsba_reproj_math.h);
  // d_jac ([N][30], the Ceres layout) is only allocated for rsba_cuda_evaluate
  rsba::DeviceBuffer<double> d_jacc, d_tau;
  // uncalibrated variant: the shared intrinsics are the 9 leading parameters of a pseudo-frame stored
  // behind the real frames in d_poses (which always has room for it); d_jac_cam = [N][2][9]
  bool free_cam = false;
  double* ptr_cam = nullptr;               // pointer API: the caller's intrinsics block
  rsba::DeviceBuffer<double> d_jac_cam;
  // free interFrameRatio of the motion priors (CeresHandler.h:156-180): parameter 9 of the same pseudo-frame
  bool free_ratio = false;
  double ratio_value = 1.0;                // host copy: initial value / value after the last solve
  double* ptr_ratio = nullptr;             // pointer API: the caller's scalar block (&opt.ceres.interFrameRatio)
  bool has_pseudo_frame() const { return free_cam || free_ratio; }
  int n_cam_frames() const { return n_frames + (has_pseudo_frame() ? 1 : 0); }
  // lower bound of the ratio block: SetParameterLowerBound(.., 0, _EPS) with acceleration priors, 0.0 with
  // velocity priors (CeresHandler.h:161, 172)
  double ratio_lower_bound() const {
    for (const auto& p : priors) if (p.kind == 2) return 2.220446049250313e-16;
    return 0.0;
  }
  rsba::DeviceBuffer<unsigned char> d_valid;
  rsba::DeviceBuffer<double> d_cost_partials;
  rsba::DeviceBuffer<double> d_scalars;    // [0] cost, misc
  rsba::DeviceBuffer<int> d_invalid;

  // ---- camera-only motion priors (k2_priors.cu)
  struct PriorHost { int kind; double scale, ratio; int frame, prev; };
  std::vector<PriorHost> priors;           // frame indices of the finalised scene
  bool priors_dirty = false;
  rsba::DeviceBuffer<int> d_prior_frame, d_prior_prev, d_prior_cur_of, d_prior_prev_of, d_prior_kind;
  rsba::DeviceBuffer<double> d_prior_coef, d_prior_scale, d_prior_r, d_prior_w2, d_prior_Bx, d_prior_jr;
  rsba::PriorView prior_view() const {
    rsba::PriorView v{};
    v.n = priors_dirty ? 0 : (int)priors.size();
    v.frame = d_prior_frame.ptr; v.prev = d_prior_prev.ptr; v.coef = d_prior_coef.ptr; v.scale = d_prior_scale.ptr;
    v.cur_of = d_prior_cur_of.ptr; v.prev_of = d_prior_prev_of.ptr; v.r = d_prior_r.ptr; v.w2 = d_prior_w2.ptr;
    v.Bx = d_prior_Bx.ptr;
    v.ratio_off = free_ratio ? (long)rsba::kFrameParams * n_frames + 9 : -1;
    v.kind = d_prior_kind.ptr; v.coef_dev = d_prior_coef.ptr; v.jr = d_prior_jr.ptr;
    return v;
  }

  // ---- GoodPosePrior blocks (k2_pose_priors.cu)
  struct PosePriorHost {
    int slot;                // control pose = 2 * frame + (0 | 1); pointer API: resolved when the problem is finalised
    double rot, pos;         // opt.ceres.trustPriorCamRotation / trustPriorCamPosition
    double val[6];           // prior block values (bulk API; pointer API: read from `prior` at solve entry)
    unsigned char constant;  // prior block fixed by the caller
    double* prior;           // pointer API: the caller's prior block (f.priorPoses[i].data())
    double* pose;            // pointer API: the control-pose block (f.poses[i].data())
  };
  std::vector<PosePriorHost> pose_priors;
  std::unordered_map<const double*, int> pose_prior_of_block;   // pointer API: prior block -> index
  bool pose_priors_dirty = false;
  rsba::DeviceBuffer<int> d_pp_slot;
  rsba::DeviceBuffer<unsigned char> d_pp_const;
  rsba::DeviceBuffer<double> d_pp_w, d_pp_val, d_pp_trial, d_pp_r, d_pp_cinv, d_pp_d2;
  rsba::PosePriorView pose_prior_view() const {
    rsba::PosePriorView v{};
    v.n = pose_priors_dirty ? 0 : (int)pose_priors.size();
    v.slot = d_pp_slot.ptr; v.w = d_pp_w.ptr; v.constant = d_pp_const.ptr; v.val = d_pp_val.ptr;
    v.trial = d_pp_trial.ptr; v.r = d_pp_r.ptr; v.cinv = d_pp_cinv.ptr; v.d2 = d_pp_d2.ptr;
    return v;
  }
  long free_pose_prior_params() const {
    long n = 0;
    for (const auto& p : pose_priors) n += p.constant ? 0 : 6;
    return n;
  }

  rsba::StageTimer timers[rsba::kNumStages];
  // events around single kernels (kStagePointBlocks ...) are only recorded when this is set: rsba_cuda_solve
  // leaves them out (every record is a stream operation between two kernels), rsba_cuda_linearize_and_step --
  // what bench.py's per-kernel pass calls -- and RSBA_CUDA_TRACE switch them on
  bool fine_timers = false;
  long launches = 0;

  rsba::LmState* lm = nullptr;
  bool reorder_tiles = true;   // nested-dissection ordering of the reduced system (rsba_solve_options)

  // multi-GPU
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
  bool scatter_owner = true;   // pointer API under rsba_multi: only one rank writes the caller's blocks back

  rsba::ObsView obs_view() const {
    return rsba::ObsView{d_obs_xy.ptr, d_obs_frame.ptr, d_obs_point.ptr, n_obs};
  }
};

namespace rsba {
// stage timing on the handle's stream
void stage_begin(rsba_problem* h, Stage s);
void stage_end(rsba_problem* h, Stage s);
double stage_collect(rsba_problem* h, Stage s);  // syncs on the end event; returns last ms

int materialize_local_share(rsba_problem* h);     // g_* -> this rank's observations on the device
// owner rank of every point: the rank whose contiguous range of frame tiles holds the point's
// median observation (host logic, no device)
void compute_point_owners(int n_frames, int n_points, long n_obs, const int* obs_frame_sorted,
                          const int* obs_point, int world, std::vector<int>* owner);
int allreduce_sum(rsba_problem* h, double* buf, size_t count);   // in place, on the handle's stream
int upload_priors(rsba_problem* h);               // host prior list -> device (after the scene is final)
int upload_pose_priors(rsba_problem* h);          // GoodPosePrior list (+ the prior blocks' current values) -> device
int finalize_pointer_problem(rsba_problem* h);   // pointer API -> sorted SoA on device
int gather_pointer_parameters(rsba_problem* h);  // caller blocks -> device
int scatter_pointer_parameters(rsba_problem* h); // device -> caller blocks
int ensure_eval_buffers(rsba_problem* h, bool jac, bool compact = false);
// residual(+Jacobian) evaluation at (poses, points) on the device; cost lands in d_scalars[0]
// compact (with jac): the Jacobian goes to d_jacc / d_tau as 12-double records instead of d_jac
int run_evaluate(rsba_problem* h, bool jac, const double* poses, const double* points,
                 double* cost_out_host, long* invalid_out_host, bool compact = false,
                 bool store_residuals = false /* without jac: also write d_res / d_valid */);

// priors' cost into the slots behind `np` observation partials, then the fixed-order sum -> d_scalars[0]
int eval_tail(rsba_problem* h, bool store, const double* poses, int np);

void lm_state_free(LmState* s);
}  // namespace rsba
