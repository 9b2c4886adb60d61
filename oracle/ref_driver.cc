// TEST INFRASTRUCTURE (oracle/_ref) -- never linked into the product library.
//
// Compiles the reference's OWN hot-path headers, verbatim, from where they lie under
// /root/reference/src (nothing is copied into this repo):
//   rsba/video_bundler_rs_inter.h -> rsba/mat/cam.h, rsba/mat/core.h,
//   rsba/video_bundler_free.h, rsba/tracking.h
// against the small Ceres/Eigen stand-in in oracle/shim/.  The exported C functions
// evaluate exactly what ceres::AutoDiffCostFunction<RsReprojectionError,2,6,6,3>
// would evaluate inside ceres::Solve (CeresHandler.h:250-255 / video_bundler_rs_inter.h:36-47).
//
// Built by oracle/Makefile into oracle/_ref/librsba_ref.so (git-ignored, travels to the
// GPU box as a prebuilt file).  Used by tests/ as the checker and by bench.py as the
// CPU baseline ("kind": "reference").
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "rsba/video_bundler_rs_inter.h"

namespace {

using vision::SHUTTER;

// Glue for the interpolateRotation=false mode.  RsBundleAdjustment (VideoSfmBaRs.h:25-35)
// cannot be included (it drags in Thrift/OpenCV through struct/VideoSfM.h); it differs from
// its compiled twin RsReprojectionError only in forwarding opt.model.interpolateRotation to
// vision::interpolate_rs.  This functor forwards that flag and otherwise calls the same two
// reference functions.
struct RsFlagged : public vision::ReprojectionError {
  SHUTTER shutter;
  int scan[2];
  bool interp_rot;
  RsFlagged(const double* cam, const double* observed, SHUTTER s, const int* sc, bool ir)
      : vision::ReprojectionError(cam, observed), shutter(s), interp_rot(ir) {
    scan[0] = sc[0];
    scan[1] = sc[1];
  }
  template <typename T>
  bool operator()(const T* const p0, const T* const p1, const T* const X, T* r) const {
    T mid[6];
    T xx[2] = {T(observed_x), T(observed_x)};  // F2: both entries are observed_x
    vision::interpolate_rs(p0, p1, shutter, scan, xx, mid, interp_rot);
    return vision::ReprojectionError::operator()(mid, X, r);
  }
};

// The uncalibrated variant: RsBundleAdjustment's 4-block operator() (VideoSfmBaRs.h:38-49) = interpolate_rs
// followed by the reference's ReprojectionError::operator()(camera, pose, point, residuals)
// (video_bundler_free.h:45-65), the intrinsics being a parameter block <2; 9, 6, 6, 3>.
struct RsFlaggedCam : public vision::ReprojectionError {
  SHUTTER shutter;
  int scan[2];
  bool interp_rot;
  RsFlaggedCam(const double* cam, const double* observed, SHUTTER s, const int* sc, bool ir)
      : vision::ReprojectionError(cam, observed), shutter(s), interp_rot(ir) {
    scan[0] = sc[0];
    scan[1] = sc[1];
  }
  template <typename T>
  bool operator()(const T* const camera, const T* const p0, const T* const p1, const T* const X, T* r) const {
    T mid[6];
    T xx[2] = {T(observed_x), T(observed_x)};
    vision::interpolate_rs(p0, p1, shutter, scan, xx, mid, interp_rot);
    return vision::ReprojectionError::operator()(camera, mid, X, r);
  }
};

struct RefProblem {
  std::vector<ceres::CostFunction*> cost;
  vision::framePtr frame;  // carries shutter + scanlines for RsReprojectionError
  double cam[9];
  ~RefProblem() {
    for (auto* c : cost) delete c;
  }
};

}  // namespace

extern "C" {

// Build one cost function per observation exactly as CeresHandler::Add does
// (CeresHandler.h:250), through the reference's own factory when interpolate_rotation != 0.
void* rsba_ref_problem_create(long n, const double* obs_xy, const double* cam9, int shutter,
                              const int* scanlines, int interpolate_rotation) {
  RefProblem* p = new RefProblem;
  memcpy(p->cam, cam9, sizeof(p->cam));
  p->frame = std::make_shared<vision::frame>(0u, p->cam);
  p->frame->shutter = (SHUTTER)shutter;
  p->frame->scanlines[0] = scanlines[0];
  p->frame->scanlines[1] = scanlines[1];
  p->cost.resize(n);
  for (long i = 0; i < n; ++i) {
    if (interpolate_rotation) {
      p->cost[i] = vision::RsReprojectionError::Create(p->cam, p->frame, obs_xy + 2 * i);
    } else {
      p->cost[i] = new ceres::AutoDiffCostFunction<RsFlagged, 2, 6, 6, 3>(
          new RsFlagged(p->cam, obs_xy + 2 * i, (SHUTTER)shutter, scanlines, false));
    }
  }
  return p;
}

void rsba_ref_problem_destroy(void* h) { delete (RefProblem*)h; }

// Evaluate residuals (and Jacobians if jac != NULL) of observations [0, n).
// Layout: residuals[n][2]; jac[n][30] = J_pose0[2][6] | J_pose1[2][6] | J_point[2][3]
// (per-block row-major, Ceres' contract); valid[n] = the functor's bool.
// Returns the number of invalid observations.
long rsba_ref_problem_eval(void* h, long n, const int* frame_idx, const int* point_idx,
                           const double* poses, const double* points, double* residuals,
                           double* jac, unsigned char* valid, int nthreads) {
  RefProblem* p = (RefProblem*)h;
  long bad = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (long i = 0; i < n; ++i) {
    const double* params[3] = {poses + 12 * (long)frame_idx[i], poses + 12 * (long)frame_idx[i] + 6,
                               points + 3 * (long)point_idx[i]};
    double r[2] = {0, 0};
    bool ok;
    if (jac) {
      double* J[3] = {jac + 30 * i, jac + 30 * i + 12, jac + 30 * i + 24};
      ok = p->cost[i]->Evaluate(params, r, J);
      if (!ok) memset(jac + 30 * i, 0, 30 * sizeof(double));
    } else {
      ok = p->cost[i]->Evaluate(params, r, nullptr);
    }
    if (!ok) {
      r[0] = r[1] = 0;
      ++bad;
    }
    if (residuals) {
      residuals[2 * i] = r[0];
      residuals[2 * i + 1] = r[1];
    }
    if (valid) valid[i] = ok ? 1 : 0;
  }
  return bad;
}

// Uncalibrated evaluation: residuals[n][2], jac[n][30] (pose0 | pose1 | point), jac_cam[n][18] = [2][9]
// w.r.t. fx fy k1 k2 p1 p2 k3 cx cy, valid[n].  Returns the number of invalid observations.
long rsba_ref_eval_cam(long n, const double* obs_xy, const int* frame_idx, const int* point_idx, const double* cam9,
                       int shutter, const int* scanlines, int interpolate_rotation, const double* poses,
                       const double* points, double* residuals, double* jac, double* jac_cam, unsigned char* valid,
                       int nthreads) {
  long bad = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (long i = 0; i < n; ++i) {
    ceres::AutoDiffCostFunction<RsFlaggedCam, 2, 9, 6, 6, 3> cf(
        new RsFlaggedCam(cam9, obs_xy + 2 * i, (SHUTTER)shutter, scanlines, interpolate_rotation != 0));
    const double* params[4] = {cam9, poses + 12 * (long)frame_idx[i], poses + 12 * (long)frame_idx[i] + 6,
                               points + 3 * (long)point_idx[i]};
    double r[2] = {0, 0};
    double jc[18];
    double* J[4] = {jc, jac + 30 * i, jac + 30 * i + 12, jac + 30 * i + 24};
    bool ok = cf.Evaluate(params, r, J);
    if (!ok) {
      memset(jac + 30 * i, 0, 30 * sizeof(double));
      memset(jc, 0, sizeof(jc));
      r[0] = r[1] = 0;
      ++bad;
    }
    memcpy(jac_cam + 18 * i, jc, sizeof(jc));
    residuals[2 * i] = r[0];
    residuals[2 * i + 1] = r[1];
    if (valid) valid[i] = ok ? 1 : 0;
  }
  return bad;
}

int rsba_ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// ---- thin wrappers over reference primitives, for the known-answer tests that restate
// ---- src/rsba/test/mat_test.cc
void rsba_ref_rotate(const double* aa, const double* pt, double* out) {
  ceres::AngleAxisRotatePoint(aa, pt, out);  // shim (Ceres is not in the reference tree)
}
void rsba_ref_slerp(const double* r0, const double* r1, double tau, double* out) {
  vision::slerp(r0, r1, tau, out);  // cam.h:251-288
}
void rsba_ref_interpolate_rs(const double* p0, const double* p1, int shutter, const int* scan,
                             const double* obs, double* out, int use_slerp) {
  vision::interpolate_rs(p0, p1, (SHUTTER)shutter, scan, obs, out, use_slerp != 0);  // cam.h:316-349
}
void rsba_ref_w2c(const double* pose, const double* X, double* out) { vision::w2c(pose, X, out); }
void rsba_ref_c2w(const double* pose, const double* pt, double* out) { vision::c2w(pose, pt, out); }
int rsba_ref_w2i(const double* cam, const double* pose, const double* X, double* proj, int validate) {
  return vision::w2i(cam, pose, X, proj, validate != 0) ? 1 : 0;  // cam.h:401-419
}
void rsba_ref_distort(const double* cam, const double* img, double* out) {
  vision::distort(cam, img, out);  // cam.h:49-72
}
double rsba_ref_norm3(const double* v) { return vision::norm3(v); }

// ---- camera-only motion priors: video_bundler_rs_inter.h:55-173
// blocks <1,6,6,6,6>; jac[12][25] row-major over the concatenated 25 parameters.
static int eval_prior(ceres::CostFunction* cf, const double* ifr, const double* pose0,
                      const double* end0, const double* pose1, const double* end1, double* res,
                      double* jac) {
  const double* params[5] = {ifr, pose0, end0, pose1, end1};
  double j0[12], j1[72], j2[72], j3[72], j4[72];
  double* J[5] = {j0, j1, j2, j3, j4};
  bool ok = cf->Evaluate(params, res, jac ? J : nullptr);
  if (jac) {
    const int off[5] = {0, 1, 7, 13, 19};
    const int sz[5] = {1, 6, 6, 6, 6};
    for (int b = 0; b < 5; ++b)
      for (int r = 0; r < 12; ++r)
        for (int c = 0; c < sz[b]; ++c) jac[r * 25 + off[b] + c] = J[b][r * sz[b] + c];
  }
  delete cf;
  return ok ? 1 : 0;
}
int rsba_ref_velo_prior(double scale, const double* ifr, const double* pose0, const double* end0,
                        const double* pose1, const double* end1, double* res, double* jac) {
  return eval_prior(vision::RsConstVeloPrior::Create(scale), ifr, pose0, end0, pose1, end1, res, jac);
}
int rsba_ref_accel_prior(double scale, const double* ifr, const double* pose0, const double* end0,
                         const double* pose1, const double* end1, double* res, double* jac) {
  return eval_prior(vision::RsConstAccelerationPrior::Create(scale), ifr, pose0, end0, pose1, end1,
                    res, jac);
}

}  // extern "C"
