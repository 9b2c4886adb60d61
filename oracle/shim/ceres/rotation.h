// TEST INFRASTRUCTURE (oracle) -- never linked into the product library.
//
// Stand-in for <ceres/rotation.h> of Ceres Solver 1.9.0 (third-party, not in the
// reference tree; pinned by /root/reference/.travis.yml:33).  Written from the
// PUBLISHED algorithm (Rodrigues' rotation formula with a first-order small-angle
// branch), not copied.  Call sites in the reference: src/rsba/mat/cam.h:121,135,365
// (the last one in place: pt == result), src/rsba/mat/cam.h:475 (CrossProduct),
// :499-501 (DotProduct), src/rsba/mat/core.h:23,35-36 (MatrixAdapter).
#ifndef RSBA_ORACLE_SHIM_ROTATION_H_
#define RSBA_ORACLE_SHIM_ROTATION_H_

#include <cmath>
#include <limits>
#include "ceres/jet.h"
#include "../../../include/rsba_ceres_constants.h"

namespace ceres {

template <typename T>
inline T DotProduct(const T x[3], const T y[3]) {
  return (x[0] * y[0] + x[1] * y[1] + x[2] * y[2]);
}

template <typename T>
inline void CrossProduct(const T x[3], const T y[3], T x_cross_y[3]) {
  x_cross_y[0] = x[1] * y[2] - x[2] * y[1];
  x_cross_y[1] = x[2] * y[0] - x[0] * y[2];
  x_cross_y[2] = x[0] * y[1] - x[1] * y[0];
}

// Strided 2-D view of a flat array: element (r, c) lives at r*row_stride + c*col_stride.
template <typename T, int row_stride, int col_stride>
struct MatrixAdapter {
  T* pointer_;
  explicit MatrixAdapter(T* pointer) : pointer_(pointer) {}
  T& operator()(const int r, const int c) const { return pointer_[r * row_stride + c * col_stride]; }
};

// result = R(angle_axis) * pt.   Safe when result aliases pt: every input
// component is consumed into temporaries before the first store.
//
// Threshold of the small-angle branch: RSBA_ANGLE_AXIS_EPS in include/rsba_ceres_constants.h, the one
// header of recalled Ceres constants.  Ceres releases differ between `> 0.0` and
// `> epsilon`; for theta2 in (0, eps] the two branches agree to < 1e-16 in value.
template <typename T>
inline void AngleAxisRotatePoint(const T angle_axis[3], const T pt[3], T result[3]) {
  using std::sqrt; using std::cos; using std::sin;
  const T theta2 = DotProduct(angle_axis, angle_axis);
  if (theta2 > T(RSBA_ANGLE_AXIS_EPS)) {
    const T theta = sqrt(theta2);
    const T costheta = cos(theta);
    const T sintheta = sin(theta);
    const T theta_inverse = T(1.0) / theta;
    const T w[3] = { angle_axis[0] * theta_inverse,
                     angle_axis[1] * theta_inverse,
                     angle_axis[2] * theta_inverse };
    const T w_cross_pt[3] = { w[1] * pt[2] - w[2] * pt[1],
                              w[2] * pt[0] - w[0] * pt[2],
                              w[0] * pt[1] - w[1] * pt[0] };
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
    const T r0 = pt[0] * costheta + w_cross_pt[0] * sintheta + w[0] * tmp;
    const T r1 = pt[1] * costheta + w_cross_pt[1] * sintheta + w[1] * tmp;
    const T r2 = pt[2] * costheta + w_cross_pt[2] * sintheta + w[2] * tmp;
    result[0] = r0; result[1] = r1; result[2] = r2;
  } else {
    // R ~ I + [angle_axis]_x near zero; keeps the derivative finite at angle_axis == 0.
    const T w_cross_pt[3] = { angle_axis[1] * pt[2] - angle_axis[2] * pt[1],
                              angle_axis[2] * pt[0] - angle_axis[0] * pt[2],
                              angle_axis[0] * pt[1] - angle_axis[1] * pt[0] };
    const T r0 = pt[0] + w_cross_pt[0];
    const T r1 = pt[1] + w_cross_pt[1];
    const T r2 = pt[2] + w_cross_pt[2];
    result[0] = r0; result[1] = r1; result[2] = r2;
  }
}

}  // namespace ceres
#endif  // RSBA_ORACLE_SHIM_ROTATION_H_
