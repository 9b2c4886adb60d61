// The handle behind the C ABI (include/rsba_cuda.h): host-side problem builder that mirrors
// what CeresHandler::Add feeds to ceres::Problem (CeresHandler.h:208-302, 335-382), plus
// the device-resident state of the evaluator and the LM solver.
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "../../include/rsba_cuda.h"

namespace rsba {

void set_last_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define RSBA_CUDA_TRY(expr)                                                        \
  do {                                                                             \
    cudaError_t e__ = (expr);                                                      \
    if (e__ != cudaSuccess) return ::rsba::cuda_fail(e__, #expr, __FILE__, __LINE__); \
  } while (0)

// Simple owning device buffer.
template <typename T>
struct DeviceBuffer {
  T* ptr = nullptr;
  size_t count = 0;
  ~DeviceBuffer() { release(); }
  DeviceBuffer() = default;
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    count = 0;
  }
  cudaError_t resize(size_t n) {
    if (n == count) return cudaSuccess;
    release();
    if (n == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&ptr, n * sizeof(T));
    if (e == cudaSuccess) count = n;
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
             kStageAllreduce, kNumStages };

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

  // ---- finalised scene (sorted by frame)
  long n_obs = 0;
  int n_frames = 0, n_points = 0;
  std::vector<long> order;                 // sorted position -> caller's observation index
  std::vector<int> h_obs_frame, h_obs_point;  // sorted, host copy (structure analysis)
  std::vector<unsigned short> pose_mask;   // [frames] constant-scalar bits
  std::vector<unsigned char> point_const;  // [points]
  bool scene_set = false, params_set = false;

  rsba::DeviceBuffer<double2> d_obs_xy;
  rsba::DeviceBuffer<int> d_obs_frame, d_obs_point;
  rsba::DeviceBuffer<double> d_poses, d_points;
  rsba::DeviceBuffer<double> d_res, d_jac;
  rsba::DeviceBuffer<unsigned char> d_valid;
  rsba::DeviceBuffer<double> d_cost_partials;
  rsba::DeviceBuffer<double> d_scalars;    // [0] cost, misc
  rsba::DeviceBuffer<int> d_invalid;

  rsba::StageTimer timers[rsba::kNumStages];
  long launches = 0;

  rsba::LmState* lm = nullptr;
  bool reorder_tiles = true;   // nested-dissection ordering of the reduced system (rsba_solve_options)

  // multi-GPU
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;

  rsba::ObsView obs_view() const {
    return rsba::ObsView{d_obs_xy.ptr, d_obs_frame.ptr, d_obs_point.ptr, n_obs};
  }
};

namespace rsba {
// stage timing on the handle's stream
void stage_begin(rsba_problem* h, Stage s);
void stage_end(rsba_problem* h, Stage s);
double stage_collect(rsba_problem* h, Stage s);  // syncs on the end event; returns last ms

int finalize_pointer_problem(rsba_problem* h);   // pointer API -> sorted SoA on device
int gather_pointer_parameters(rsba_problem* h);  // caller blocks -> device
int scatter_pointer_parameters(rsba_problem* h); // device -> caller blocks
int ensure_eval_buffers(rsba_problem* h, bool jac);
// residual(+Jacobian) evaluation at (poses, points) on the device; cost lands in d_scalars[0]
int run_evaluate(rsba_problem* h, bool jac, const double* poses, const double* points,
                 double* cost_out_host, long* invalid_out_host);

void lm_state_free(LmState* s);
}  // namespace rsba
