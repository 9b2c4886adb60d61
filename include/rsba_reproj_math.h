/* Per-observation arithmetic of the rolling-shutter reprojection residual and its 2 x (6 + 6 + 3) Jacobian
 * -- what RsBundleAdjustment::operator() (VideoSfmBaRs.h:25-35) computes through interpolate_rs
 * (mat/cam.h:316-349), w2i / w2c (mat/cam.h:401-419, 355-366), ceres::AngleAxisRotatePoint, c2i + distort
 * (mat/cam.h:372-395, 49-72) and ceres::AutoDiffCostFunction<..., 2, 6, 6, 3> (VideoSfmBaRs.h:58-63), with the
 * derivative written out by hand.
 *
 * ONE source for two callers: the sm_100a kernels of the product (rsba_b200/csrc/k1_reproj.cu, k5_pnp.cu:
 * compiled by nvcc, __device__) and the host functor structs of include/rsba_cuda_functors.hpp (compiled by
 * any C++11 compiler, no CUDA headers), the spot-check path a maintainer can call next to the reference's
 * own functors.  The product library itself only ever runs it on the device. */
#ifndef RSBA_REPROJ_MATH_H_
#define RSBA_REPROJ_MATH_H_

#include "rsba_ceres_constants.h"
#include <cfloat>
#include <cmath>

#if defined(__CUDACC__)
#define RSBA_HD __host__ __device__ __forceinline__
#define RSBA_UNROLL _Pragma("unroll")
#else
#define RSBA_HD inline
#define RSBA_UNROLL
#endif

namespace rsba {

constexpr int kPoseParams = 6;    // NUM_POSE_PARAMS  (mat/cam.h:20)
constexpr int kPointParams = 3;   // NUM_POINT_PARAMS (mat/cam.h:19)
constexpr int kFrameParams = 12;  // pose0 | pose1
constexpr int kJacDoubles = 30;   // 2x6 | 2x6 | 2x3
// Compact record of the same Jacobian (what the solver's kernels read; rsba_cuda_evaluate hands out the full
// one): tau is a constant of the observation, so J_pose0 = [wr0 jr | -(1-tau) jx], J_pose1 = [wr1 jr | -tau jx],
// J_point = jx with jr = d res / d rotation and jx = d res / d X of the INTERPOLATED pose (2x3 each), and
// (wr0, wr1) = (1-tau, tau) when the rotation is interpolated, (1, 0) otherwise.  12 doubles = three 32-byte
// sectors instead of 240 bytes:  jx row 0 | jx row 1 | jr row 0 | jr row 1.
constexpr int kJacCompact = 12;

// Per-session constants captured by every cost functor (VideoSfmBaRs.h:16-22).
struct CameraModel {
  double cam[9];       // fx fy k1 k2 p1 p2 k3 cx cy
  double scan0;        // scanlines[0]
  double scan_span;    // scanlines[1] - scanlines[0]
  int shutter;         // 0 GLOBAL, 1 HORIZONTAL, 2 VERTICAL
  int interp_rot;      // opt.model.interpolateRotation
  double huber;        // ceres::HuberLoss(a) on every residual block (CeresHandler.h:85-90); 0 = no loss
  // Uncalibrated variant (RsBundleAdjustment::CreateWithCam <2; 9, 6, 6, 3>, VideoSfmBaRs.h:38-49,68-80): the
  // shared intrinsics are a PARAMETER block.  It lives behind the frames in the pose array, as a
  // pseudo-frame: poses[cam_offset .. cam_offset+8] = fx fy k1 k2 p1 p2 k3 cx cy; -1 = calibrated.
  long cam_offset;
};


struct Proj {
  double r0, r1;   // residual
  bool ok;
  double tau;      // interpolation parameter of the observation, clamped to [0, 1] (0 for a global shutter)
};

// Everything that depends on one observation.  JAC=false skips all derivative work.
// TAU_FROM_Y: the scan line of a VERTICAL shutter is read from observed_y, as interpolate_rs does
// when it is handed the real observation (getPose / validate, struct/VideoSfM.cc:108-111,159-169).
// The BA functor hands it {observed_x, observed_x} (VideoSfmBaRs.h:31), hence false there.
// VALIDATE = false: w2i(..., validate = false) as the RS-PnP functor and its inlier scoring call it
// (solveRSpnp.cpp:65, 233; mat/cam.h:409-416): a point behind the camera is NOT rejected, and a depth
// inside (-eps, eps) is replaced by eps (a constant: its derivative rows vanish).
// COMPACT: J receives the kJacCompact-double record above instead of the 30-double Ceres layout.
template <bool JAC, bool TAU_FROM_Y = false, bool VALIDATE = true, bool COMPACT = false>
RSBA_HD Proj reproject(const CameraModel& cm, double ox, double oy,
                                          const double* __restrict__ p0,  // frame: pose0|pose1
                                          double X0, double X1, double X2,
                                          double* __restrict__ J /* [30], smem, JAC only */,
                                          double* __restrict__ Jcam = nullptr /* [18] = [2][9], JAC only */) {
  Proj out;
  // ---- interpolate_rs (mat/cam.h:316-349): tau is a constant of the observation
  double tau = 0.0, wr0 = 1.0, wr1 = 0.0;
  if (cm.shutter != 0) {
    tau = (((TAU_FROM_Y && cm.shutter == 2) ? oy : ox) - cm.scan0) / cm.scan_span;
    tau = tau < 0.0 ? 0.0 : tau;
    tau = tau > 1.0 ? 1.0 : tau;
    if (cm.interp_rot) { wr0 = 1.0 - tau; wr1 = tau; }
  }
  out.tau = tau;
  const double* p1 = p0 + 6;
  double r[3], c[3];
  if (cm.shutter != 0 && cm.interp_rot) {
RSBA_UNROLL
    for (int k = 0; k < 3; ++k) r[k] = p0[k] + (p1[k] - p0[k]) * tau;
  } else {
RSBA_UNROLL
    for (int k = 0; k < 3; ++k) r[k] = p0[k];
  }
RSBA_UNROLL
  for (int k = 0; k < 3; ++k) c[k] = p0[3 + k] + (p1[3 + k] - p0[3 + k]) * tau;

  // ---- w2c (mat/cam.h:355-366)
  const double q0 = X0 - c[0], q1 = X1 - c[1], q2 = X2 - c[2];
  const double theta2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  double P0, P1, P2;
  double R[9];       // rotation matrix, row-major            (JAC)
  double dPr[9];     // dP/dr, row-major [component][k]       (JAC)
  if (theta2 > RSBA_ANGLE_AXIS_EPS) {   // (recalled Ceres constant: include/rsba_ceres_constants.h)
    const double theta = sqrt(theta2);
    double s, co;
    sincos(theta, &s, &co);
    const double ti = 1.0 / theta;
    const double w0 = r[0] * ti, w1 = r[1] * ti, w2 = r[2] * ti;
    const double x0 = w1 * q2 - w2 * q1;
    const double x1 = w2 * q0 - w0 * q2;
    const double x2 = w0 * q1 - w1 * q0;
    const double omc = 1.0 - co;
    const double wq = w0 * q0 + w1 * q1 + w2 * q2;
    const double tmp = wq * omc;
    P0 = q0 * co + x0 * s + w0 * tmp;
    P1 = q1 * co + x1 * s + w1 * tmp;
    P2 = q2 * co + x2 * s + w2 * tmp;
    if (JAC) {
      // R = co*I + s*[w]x + omc*w w^T
      R[0] = co + omc * w0 * w0;      R[1] = omc * w0 * w1 - s * w2;  R[2] = omc * w0 * w2 + s * w1;
      R[3] = omc * w1 * w0 + s * w2;  R[4] = co + omc * w1 * w1;      R[5] = omc * w1 * w2 - s * w0;
      R[6] = omc * w2 * w0 - s * w1;  R[7] = omc * w2 * w1 + s * w0;  R[8] = co + omc * w2 * w2;
      // Chain rule of the Rodrigues formula with dtheta/dr = w^T, dw/dr = (I - w w^T)/theta:
      //   dP/dr = a w^T + ti * ( -s [q]x + omc (wq I + w q^T) ),
      //   a = -s q + (co - s ti)(w x q) + (s - 2 ti omc) wq w
      const double ka = co - s * ti;
      const double kb = (s - 2.0 * ti * omc) * wq;
      const double a0 = -s * q0 + ka * x0 + kb * w0;
      const double a1 = -s * q1 + ka * x1 + kb * w1;
      const double a2 = -s * q2 + ka * x2 + kb * w2;
      const double st = s * ti, ot = omc * ti, d = ot * wq;
      dPr[0] = a0 * w0 + d + ot * w0 * q0;
      dPr[1] = a0 * w1 + st * q2 + ot * w0 * q1;
      dPr[2] = a0 * w2 - st * q1 + ot * w0 * q2;
      dPr[3] = a1 * w0 - st * q2 + ot * w1 * q0;
      dPr[4] = a1 * w1 + d + ot * w1 * q1;
      dPr[5] = a1 * w2 + st * q0 + ot * w1 * q2;
      dPr[6] = a2 * w0 + st * q1 + ot * w2 * q0;
      dPr[7] = a2 * w1 - st * q0 + ot * w2 * q1;
      dPr[8] = a2 * w2 + d + ot * w2 * q2;
    }
  } else {
    // first-order branch of AngleAxisRotatePoint: P = q + r x q
    P0 = q0 + (r[1] * q2 - r[2] * q1);
    P1 = q1 + (r[2] * q0 - r[0] * q2);
    P2 = q2 + (r[0] * q1 - r[1] * q0);
    if (JAC) {
      R[0] = 1.0;   R[1] = -r[2]; R[2] = r[1];
      R[3] = r[2];  R[4] = 1.0;   R[5] = -r[0];
      R[6] = -r[1]; R[7] = r[0];  R[8] = 1.0;
      dPr[0] = 0.0; dPr[1] = q2;  dPr[2] = -q1;   // d(r x q)/dr = -[q]x
      dPr[3] = -q2; dPr[4] = 0.0; dPr[5] = q0;
      dPr[6] = q1;  dPr[7] = -q0; dPr[8] = 0.0;
    }
  }

  // ---- w2i validity (mat/cam.h:410-412); c2i's own |z| < eps test cannot fire after it
  out.ok = VALIDATE ? !(P2 < 1e-8) : true;
  if (!VALIDATE && P2 < DBL_EPSILON && P2 > -DBL_EPSILON) {
    P2 = DBL_EPSILON;
    if (JAC) {
RSBA_UNROLL
      for (int k = 0; k < 3; ++k) { dPr[6 + k] = 0.0; R[6 + k] = 0.0; }
    }
  }
  if (!out.ok) {
    out.r0 = 0.0;
    out.r1 = 0.0;
    if (JAC) {
RSBA_UNROLL
      for (int k = 0; k < (COMPACT ? kJacCompact : kJacDoubles); ++k) J[k] = 0.0;
      if (Jcam) {
RSBA_UNROLL
        for (int k = 0; k < 18; ++k) Jcam[k] = 0.0;
      }
    }
    return out;
  }

  // ---- c2i + distort (mat/cam.h:372-395, 49-72)
  const double fx = cm.cam[0], fy = cm.cam[1], k1 = cm.cam[2], k2 = cm.cam[3];
  const double t1 = cm.cam[4], t2 = cm.cam[5], k3 = cm.cam[6];
  const double iz = 1.0 / P2;
  const double xp = P0 * iz, yp = P1 * iz;
  const double r2 = xp * xp + yp * yp;
  const double dist = 1.0 + r2 * (k1 + r2 * (k2 + r2 * k3));
  const double xy = xp * yp;
  const double px = dist * xp + (2.0 * t1 * xy + t2 * (r2 + 2.0 * xp * xp));
  const double py = dist * yp + (t1 * (r2 + 2.0 * yp * yp) + 2.0 * t2 * xy);
  out.r0 = (px * fx + cm.cam[7]) - ox;
  out.r1 = (py * fy + cm.cam[8]) - oy;

  if (JAC && Jcam) {
    // d residual / d (fx fy k1 k2 p1 p2 k3 cx cy): c2i + distort (mat/cam.h:372-395, 49-72) differentiated
    const double r4 = r2 * r2, r6 = r4 * r2;
    Jcam[0] = px;            Jcam[9] = 0.0;
    Jcam[1] = 0.0;           Jcam[10] = py;
    Jcam[2] = fx * xp * r2;  Jcam[11] = fy * yp * r2;
    Jcam[3] = fx * xp * r4;  Jcam[12] = fy * yp * r4;
    Jcam[4] = fx * 2.0 * xy; Jcam[13] = fy * (r2 + 2.0 * yp * yp);
    Jcam[5] = fx * (r2 + 2.0 * xp * xp); Jcam[14] = fy * 2.0 * xy;
    Jcam[6] = fx * xp * r6;  Jcam[15] = fy * yp * r6;
    Jcam[7] = 1.0;           Jcam[16] = 0.0;
    Jcam[8] = 0.0;           Jcam[17] = 1.0;
  }
  if (JAC) {
    // d(px,py)/d(xp,yp)
    const double dd = k1 + r2 * (2.0 * k2 + 3.0 * k3 * r2);
    const double gxx = dist + 2.0 * xp * xp * dd + 2.0 * t1 * yp + 6.0 * t2 * xp;
    const double gxy = 2.0 * xy * dd + 2.0 * t1 * xp + 2.0 * t2 * yp;
    const double gyy = dist + 2.0 * yp * yp * dd + 6.0 * t1 * yp + 2.0 * t2 * xp;
    // A = diag(fx,fy) * G * d(xp,yp)/dP,  d(xp,yp)/dP = iz * [1 0 -xp; 0 1 -yp]
    const double A00 = fx * gxx * iz, A01 = fx * gxy * iz;
    const double A02 = -(A00 * xp + A01 * yp);
    const double A10 = fy * gxy * iz, A11 = fy * gyy * iz;
    const double A12 = -(A10 * xp + A11 * yp);
    const double w0c = 1.0 - tau, w1c = tau;
RSBA_UNROLL
    for (int k = 0; k < 3; ++k) {
      const double jr0 = A00 * dPr[k] + A01 * dPr[3 + k] + A02 * dPr[6 + k];  // d res / d rot
      const double jr1 = A10 * dPr[k] + A11 * dPr[3 + k] + A12 * dPr[6 + k];
      const double jx0 = A00 * R[k] + A01 * R[3 + k] + A02 * R[6 + k];        // d res / d X
      const double jx1 = A10 * R[k] + A11 * R[3 + k] + A12 * R[6 + k];
      if (COMPACT) {
        J[k] = jx0;  J[3 + k] = jx1;  J[6 + k] = jr0;  J[9 + k] = jr1;
      } else {
        J[k] = wr0 * jr0;        J[6 + k] = wr0 * jr1;         // J_pose0 rows 0,1: rotation
        J[3 + k] = -w0c * jx0;   J[9 + k] = -w0c * jx1;        //                   centre
        J[12 + k] = wr1 * jr0;   J[18 + k] = wr1 * jr1;        // J_pose1
        J[15 + k] = -w1c * jx0;  J[21 + k] = -w1c * jx1;
        J[24 + k] = jx0;         J[27 + k] = jx1;              // J_point
      }
    }
  }
  return out;
}

// The 30-double Ceres layout back from a compact record (the kernels that are not on the hot path index
// the full record; the hot ones use the structure directly).  rot_interp = shutter != GLOBAL && interpolateRotation.
RSBA_HD void expand_jacobian(const double* __restrict__ rec, double tau, bool rot_interp, double* __restrict__ J) {
  const double wr0 = rot_interp ? 1.0 - tau : 1.0, wr1 = rot_interp ? tau : 0.0;
  const double w0c = 1.0 - tau, w1c = tau;
RSBA_UNROLL
  for (int k = 0; k < 3; ++k) {
    const double jx0 = rec[k], jx1 = rec[3 + k], jr0 = rec[6 + k], jr1 = rec[9 + k];
    J[k] = wr0 * jr0;        J[6 + k] = wr0 * jr1;
    J[3 + k] = -w0c * jx0;   J[9 + k] = -w0c * jx1;
    J[12 + k] = wr1 * jr0;   J[18 + k] = wr1 * jr1;
    J[15 + k] = -w1c * jx0;  J[21 + k] = -w1c * jx1;
    J[24 + k] = jx0;         J[27 + k] = jx1;
  }
}

}  // namespace rsba

#endif  /* RSBA_REPROJ_MATH_H_ */
