// Internal declarations shared by the kernels and the host driver (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace rsba {

constexpr int kPoseParams = 6;    // NUM_POSE_PARAMS  (mat/cam.h:20)
constexpr int kPointParams = 3;   // NUM_POINT_PARAMS (mat/cam.h:19)
constexpr int kFrameParams = 12;  // pose0 | pose1
constexpr int kJacDoubles = 30;   // 2x6 | 2x6 | 2x3

// Per-session constants captured by every cost functor (VideoSfmBaRs.h:16-22).
struct CameraModel {
  double cam[9];       // fx fy k1 k2 p1 p2 k3 cx cy
  double scan0;        // scanlines[0]
  double scan_span;    // scanlines[1] - scanlines[0]
  int shutter;         // 0 GLOBAL, 1 HORIZONTAL, 2 VERTICAL
  int interp_rot;      // opt.model.interpolateRotation
  double huber;        // ceres::HuberLoss(a) on every residual block (CeresHandler.h:85-90); 0 = no loss
};

// Observation SoA, sorted by frame.
struct ObsView {
  const double2* xy;
  const int* frame;
  const int* point;
  long n;
};

// ---- launchers (definitions in the .cu files); all asynchronous on `stream` ------------
// K1: residual + Jacobian (+ per-CTA cost partials, invalid count).
void launch_k1(const CameraModel& cm, const ObsView& obs, const double* poses, const double* points,
               double* residuals, double* jac, unsigned char* valid, double* cost_partials,
               int* invalid_count, cudaStream_t stream);
// K1r: cost only at trial parameters.
void launch_k1r(const CameraModel& cm, const ObsView& obs, const double* poses, const double* points,
                double* cost_partials, int* invalid_count, cudaStream_t stream);
// track validation sweep (struct/VideoSfM.cc:159-169): per-observation predicate + squared error
void launch_validate(const CameraModel& cm, const ObsView& obs, const double* poses, const double* points,
                     double sqrd_threshold, double min_distance, unsigned char* ok, double* sqrd_error,
                     cudaStream_t stream);
int k1_num_partials(long n);
// deterministic fixed-order sum of `n` partials into out[0]
void launch_reduce_partials(const double* partials, int n, double* out, cudaStream_t stream);

}  // namespace rsba
