// Host cost-functor structs with the reference's signatures -- the spot-check path of the drop-in
// (SURVEY 8b "Signatures kept"): what a maintainer of the reference calls where they used to call the
// functor directly, next to a GPU solve.  T = double only (the reference instantiates its templates
// with double and with ceres::Jet; the derivative that Jet produced is available through Evaluate()).
//
//   vision::ReprojectionError                 video_bundler_free.h:17-101
//     operator()(pose, point, residuals)                       :33-41
//     operator()(camera, pose, point, residuals)               :45-65
//   vision::sfm::RsBundleAdjustment           VideoSfmBaRs.h:15-84
//     operator()(pose0, pose1, point, residuals)               :25-35
//     operator()(camera, pose0, pose1, point, residuals)       :38-49
//   ceres::CostFunction::Evaluate(parameters, residuals, jacobians) of the AutoDiffCostFunction the
//     reference wraps them in (VideoSfmBaRs.h:58-63, 73-79; video_bundler_free.h:73-91)
//
// Same block sizes and order (9 | 6, 6 | 3), two residuals, `false` iff the point is not in front of the
// camera (w2i: z < 1e-8, mat/cam.h:410-412).  The arithmetic is include/rsba_reproj_math.h -- the very code
// the sm_100a kernels run -- compiled for the host; no CUDA headers, any C++11 compiler.
// Results agree with the reference's functor to rounding (<= 1e-12 relative on the reference's own
// mat_test.cc grid, tests/test_functors_host.py), the Jacobian with its Jet<15> / Jet<24> derivative.
#ifndef RSBA_CUDA_FUNCTORS_HPP_
#define RSBA_CUDA_FUNCTORS_HPP_

#include <cstring>

#include "rsba_reproj_math.h"

namespace rsba_cuda {

enum { NUM_CAM_PARAMS = 9, NUM_POSE_PARAMS = 6, NUM_POINT_PARAMS = 3 };   // mat/cam.h:18-20

struct ReprojectionError {
  static const unsigned short NUM_RESIDUALS = 2;

  ReprojectionError() : observed_x(0), observed_y(0) { std::memset(camera_params, 0, sizeof(camera_params)); }
  explicit ReprojectionError(const double observed[2]) : observed_x(observed[0]), observed_y(observed[1]) {
    std::memset(camera_params, 0, sizeof(camera_params));
  }
  ReprojectionError(const double camera[NUM_CAM_PARAMS], const double observed[2])
      : observed_x(observed[0]), observed_y(observed[1]) {
    std::memcpy(camera_params, camera, sizeof(camera_params));
  }

  // video_bundler_free.h:33-41 -- the functor's own copy of the intrinsics
  bool operator()(const double* const pose, const double* const point, double* residuals) const {
    return (*this)(camera_params, pose, point, residuals);
  }
  // video_bundler_free.h:45-65 -- intrinsics as a parameter block
  bool operator()(const double* const camera, const double* const pose, const double* const point,
                  double* residuals) const {
    return project(camera, pose, pose, /*shutter=*/0, 0.0, 1.0, true, point, residuals, nullptr, nullptr);
  }
  // ceres::CostFunction::Evaluate of AutoDiffCostFunction<ReprojectionError, 2, 6, 3> (video_bundler_free.h:84-91):
  // parameters = {pose[6], point[3]}; jacobians (may be NULL, entries may be NULL) = {2x6, 2x3} row-major
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
    double J[rsba::kJacDoubles];
    const bool ok = project(camera_params, parameters[0], parameters[0], 0, 0.0, 1.0, true, parameters[1], residuals,
                            jacobians ? J : nullptr, nullptr);
    if (ok && jacobians) {
      if (jacobians[0]) std::memcpy(jacobians[0], J, 12 * sizeof(double));
      if (jacobians[1]) std::memcpy(jacobians[1], J + 24, 6 * sizeof(double));
    }
    return ok;
  }

  double observed_x;
  double observed_y;
  double camera_params[NUM_CAM_PARAMS];

 protected:
  // one call of the shared per-observation arithmetic; J = [2x6 | 2x6 | 2x3], Jcam = [2x9], both row-major
  bool project(const double* camera, const double* pose0, const double* pose1, int shutter, double scan0,
               double scan_span, bool interpolate_rotation, const double* point, double* residuals, double* J,
               double* Jcam) const {
    rsba::CameraModel cm;
    std::memcpy(cm.cam, camera, sizeof(cm.cam));
    cm.scan0 = scan0;
    cm.scan_span = scan_span;
    cm.shutter = shutter;
    cm.interp_rot = interpolate_rotation ? 1 : 0;
    cm.huber = 0.0;
    cm.cam_offset = -1;
    double frame[12];
    std::memcpy(frame, pose0, 6 * sizeof(double));
    std::memcpy(frame + 6, pose1, 6 * sizeof(double));
    const rsba::Proj p = J ? rsba::reproject<true>(cm, observed_x, observed_y, frame, point[0], point[1], point[2], J, Jcam)
                           : rsba::reproject<false>(cm, observed_x, observed_y, frame, point[0], point[1], point[2],
                                                    nullptr, nullptr);
    if (!p.ok) return false;        // like the reference, residuals are left untouched on failure
    residuals[0] = p.r0;
    residuals[1] = p.r1;
    return true;
  }
};

// Session: anything with  cam (.data() -> 9 doubles), rs (shutter: 0 GLOBAL, 1 HORIZONTAL, 2 VERTICAL),
// scanlines (.data() -> 2 ints);  Options: anything with  model.interpolateRotation  -- the reference's
// gen::Session (gen-cpp/sfm_types.h) and SfmOptions (SfmOptions.h) as they are.
template <class Session, class Options>
struct RsBundleAdjustmentT : public ReprojectionError {
  RsBundleAdjustmentT(const Session& sess, const Options& opt, const double* const observed)
      : ReprojectionError(sess.cam.data(), observed), sess(sess), opt(opt) {}

  // VideoSfmBaRs.h:25-35 (the scan line comes from observed_x for either shutter direction: `obs = {x, x}`)
  bool operator()(const double* const pose0, const double* const pose1, const double* const point,
                  double* residuals) const {
    return rs(camera_params, pose0, pose1, point, residuals, nullptr, nullptr);
  }
  // VideoSfmBaRs.h:38-49
  bool operator()(const double* const camera, const double* const pose0, const double* const pose1,
                  const double* const point, double* residuals) const {
    return rs(camera, pose0, pose1, point, residuals, nullptr, nullptr);
  }
  // AutoDiffCostFunction<RsBundleAdjustment, 2, 6, 6, 3> (VideoSfmBaRs.h:58-63): parameters = {pose0, pose1, point}
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
    double J[rsba::kJacDoubles];
    const bool ok = rs(camera_params, parameters[0], parameters[1], parameters[2], residuals, jacobians ? J : nullptr,
                       nullptr);
    if (ok && jacobians) {
      if (jacobians[0]) std::memcpy(jacobians[0], J, 12 * sizeof(double));
      if (jacobians[1]) std::memcpy(jacobians[1], J + 12, 12 * sizeof(double));
      if (jacobians[2]) std::memcpy(jacobians[2], J + 24, 6 * sizeof(double));
    }
    return ok;
  }
  // AutoDiffCostFunction<RsBundleAdjustment, 2, 9, 6, 6, 3> (CreateWithCam, VideoSfmBaRs.h:68-80):
  // parameters = {camera, pose0, pose1, point}
  bool EvaluateWithCam(double const* const* parameters, double* residuals, double** jacobians) const {
    double J[rsba::kJacDoubles], Jcam[18];
    const bool ok = rs(parameters[0], parameters[1], parameters[2], parameters[3], residuals, jacobians ? J : nullptr,
                       jacobians ? Jcam : nullptr);
    if (ok && jacobians) {
      if (jacobians[0]) std::memcpy(jacobians[0], Jcam, 18 * sizeof(double));
      if (jacobians[1]) std::memcpy(jacobians[1], J, 12 * sizeof(double));
      if (jacobians[2]) std::memcpy(jacobians[2], J + 12, 12 * sizeof(double));
      if (jacobians[3]) std::memcpy(jacobians[3], J + 24, 6 * sizeof(double));
    }
    return ok;
  }

  const Session& sess;
  const Options& opt;

 private:
  bool rs(const double* camera, const double* pose0, const double* pose1, const double* point, double* residuals,
          double* J, double* Jcam) const {
    const int* scan = sess.scanlines.data();
    return project(camera, pose0, pose1, (int)sess.rs, (double)scan[0], (double)(scan[1] - scan[0]),
                   opt.model.interpolateRotation, point, residuals, J, Jcam);
  }
};

}  // namespace rsba_cuda

#endif  // RSBA_CUDA_FUNCTORS_HPP_
