// Internal declarations shared by the kernels and the host driver (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include "../../include/rsba_reproj_math.h"
#include <cstdint>
#include <cstdio>

namespace rsba {


// CameraModel (the per-session constants every cost functor captures) lives in the public header
// include/rsba_reproj_math.h together with the per-observation arithmetic.

// Observation SoA, sorted by frame.
struct ObsView {
  const double2* xy;
  const int* frame;
  const int* point;
  long n;
};

// Camera-only motion priors between a frame and its predecessor (k2_priors.cu), device view.
struct PriorView {
  int n;                 // number of priors
  const int* frame;      // [n] current frame k
  const int* prev;       // [n] previous frame (k-1)
  const double* coef;    // [n][8] rows (half A, half B) x (pose0, end0, pose1, end1)
  const double* scale;   // [n]
  const int* cur_of;     // [F] prior whose current frame is f, or -1
  const int* prev_of;    // [F] prior whose previous frame is f, or -1
  double* r;             // [n][12] loss-corrected residuals at the linearisation point
  double* w2;            // [n] squared loss-correction weight
  double* Bx;            // [F][4][6] coupling diagonals of cur_of[f] (zero without a prior)
  // Free interFrameRatio (the reference's default, interFrameRatio == 1: CeresHandler.h:156-180 leaves the
  // scalar block `&opt.ceres.interFrameRatio` variable, with a lower bound).  The ratio is parameter 9 of
  // the pseudo-frame behind the real frames: poses[ratio_off]; -1 = constant (coef fixed at upload).
  long ratio_off;
  const int* kind;       // [n] 1 velocity, 2 acceleration
  double* coef_dev;      // == coef, writable: refreshed from the current ratio at every linearisation
  double* jr;            // [n][12] loss-corrected d residual / d ratio at the linearisation point
};

// "Good initial guess" priors (GoodPosePrior, CeresHandler.h:55-73, wired at :188-204): residual
// diag(rotation x3, position x3) (prior - pose) between a 6-wide PRIOR block and one control pose.  The
// reference never fixes the prior block, so by default it is a free parameter block that occurs in this one
// residual only: it is eliminated in closed form like a 3-D point seen once (k2_pose_priors.cu).
struct PosePriorView {
  int n;
  const int* slot;               // [n] control pose = 2 * frame + (0 | 1): parameters 6*slot .. 6*slot+5
  const double* w;               // [n][2] rotation, position weights
  const unsigned char* constant; // [n] prior block held constant (set_block_constant)
  double* val;                   // [n][6] current prior block values
  double* trial;                 // [n][6] trial values
  double* r;                     // [n][6] residuals at the linearisation point
  double* cinv;                  // [n][6] s^2 / (s^2 w^2 + D^2): inverse of the damped block (unscaled space)
  double* d2;                    // [n][6] LM diagonal of the block over s^2 (so that D^2 (delta/s)^2 = d2 delta^2)
};

// Kernel attributes (opt-in shared memory) are per device: true the first time `slot` (one per call
// site) is seen on the current device.
inline bool first_use_on_device(bool (&seen)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (seen[dev]) return false;
  seen[dev] = true;
  return true;
}

// ---- launchers (definitions in the .cu files); all asynchronous on `stream` ------------
// K1: residual + Jacobian (+ per-CTA cost partials, invalid count).
// jac_cam: [N][2][9] Jacobian w.r.t. the intrinsics (only with cm.cam_offset >= 0; may be NULL)
void launch_k1(const CameraModel& cm, const ObsView& obs, const double* poses, const double* points,
               double* residuals, double* jac, double* jac_cam, unsigned char* valid, double* cost_partials,
               int* invalid_count, cudaStream_t stream);
// K1 for the solver: compact 12-double Jacobian records + tau per observation (rsba_reproj_math.h)
void launch_k1_compact(const CameraModel& cm, const ObsView& obs, const double* poses, const double* points,
                       double* residuals, double* jac_compact, double* tau, double* jac_cam, double* cost_partials,
                       int* invalid_count, cudaStream_t stream);
// K1r: cost only at trial parameters.
// residuals / valid: optional per-observation outputs (rsba_cuda_evaluate without a Jacobian)
void launch_k1r(const CameraModel& cm, const ObsView& obs, const double* poses, const double* points,
                double* cost_partials, int* invalid_count, cudaStream_t stream, double* residuals = nullptr,
                unsigned char* valid = nullptr);
// track validation sweep (struct/VideoSfM.cc:159-169): per-observation predicate + squared error
void launch_validate(const CameraModel& cm, const ObsView& obs, const double* poses, const double* points,
                     double sqrd_threshold, double min_distance, unsigned char* ok, double* sqrd_error,
                     cudaStream_t stream);
// iterative rolling-shutter re-projection of (frame, point) pairs (struct/VideoSfM.cc:139-155); device arrays
void launch_reproject(const CameraModel& cm, long n, const int* frame, const int* point, const double* poses,
                      const double* points, double sqrd_threshold, double* proj_xy, unsigned char* ok,
                      cudaStream_t stream);
// priors: cost_out[0] = sum rho(|r|^2); store: also residuals and weights for the linearisation
// invalid_count += priors whose functor returns false (ratio below the functor's bound; free ratio only)
void launch_prior_eval(const PriorView& pv, const double* poses, double huber, double* cost_out, bool store,
                       int* invalid_count, cudaStream_t stream);
// pose priors: cost_out[0] = sum |r|^2 at (poses, vals); store: residuals for the linearisation;
// invalid_count += blocks whose functor returns false (rotation residual >= 1, CeresHandler.h:66)
void launch_pose_prior_eval(const PosePriorView& pv, const double* poses, const double* vals, double* cost_out,
                            bool store, int* invalid_count, cudaStream_t stream);
int k1_num_partials(long n);
// deterministic fixed-order sum of `n` partials into out[0]
void launch_reduce_partials(const double* partials, int n, double* out, cudaStream_t stream);

}  // namespace rsba
