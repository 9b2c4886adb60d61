/* TEST INFRASTRUCTURE (oracle "port") -- never linked into the product library.
 *
 * Plain-C restatement of the reference's rolling-shutter reprojection residual and of
 * the forward-mode (dual number) Jacobian that ceres::AutoDiffCostFunction<.,2,6,6,3>
 * produces for it.  Each function cites the reference lines it follows
 * (paths relative to /root/reference/src/rsba/).  Parity of this port is PINNED against
 * oracle/_ref (the reference's own headers compiled verbatim) by
 * tests/test_oracle_cpu.py and by the committed fixtures in tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this.
 */
#include <math.h>
#include <float.h>
#include "../include/rsba_ceres_constants.h"
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NV 15 /* 6 (pose0) + 6 (pose1) + 3 (point): Jet<double,15> */

typedef struct {
  double a;
  double v[NV];
} dual;

/* ---- dual-number arithmetic: exact chain rule, same operation order as a Jet ---- */
static inline dual d_const(double a) { dual r; r.a = a; memset(r.v, 0, sizeof(r.v)); return r; }
static inline dual d_var(double a, int k) { dual r = d_const(a); r.v[k] = 1.0; return r; }
static inline dual d_add(dual f, dual g) { dual h; h.a = f.a + g.a; for (int i = 0; i < NV; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
static inline dual d_sub(dual f, dual g) { dual h; h.a = f.a - g.a; for (int i = 0; i < NV; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
static inline dual d_mul(dual f, dual g) { dual h; h.a = f.a * g.a; for (int i = 0; i < NV; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
static inline dual d_div(dual f, dual g) {
  dual h; const double gi = 1.0 / g.a; h.a = f.a * gi;
  for (int i = 0; i < NV; ++i) h.v[i] = (f.v[i] - h.a * g.v[i]) * gi;
  return h;
}
static inline dual d_sqrt(dual f) { dual h; h.a = sqrt(f.a); const double t = 1.0 / (2.0 * h.a); for (int i = 0; i < NV; ++i) h.v[i] = f.v[i] * t; return h; }
static inline dual d_cos(dual f) { dual h; h.a = cos(f.a); const double ms = -sin(f.a); for (int i = 0; i < NV; ++i) h.v[i] = ms * f.v[i]; return h; }
static inline dual d_sin(dual f) { dual h; h.a = sin(f.a); const double c = cos(f.a); for (int i = 0; i < NV; ++i) h.v[i] = c * f.v[i]; return h; }

/* The arithmetic below is written once over the macro type T with operator macros and
 * instantiated twice: T = double (cost-only evaluation) and T = dual (Jacobian). */

/* ================= instantiation 1: double ================= */
#define T double
#define K(x) (x)
#define ADD(a, b) ((a) + (b))
#define SUB(a, b) ((a) - (b))
#define MUL(a, b) ((a) * (b))
#define DIV(a, b) ((a) / (b))
#define SQRT(a) sqrt(a)
#define COS(a) cos(a)
#define SIN(a) sin(a)
#define VAL(a) (a)
#define FN(name) name##_d
#include "rsba_oracle_body.inc"
#undef T
#undef K
#undef ADD
#undef SUB
#undef MUL
#undef DIV
#undef SQRT
#undef COS
#undef SIN
#undef VAL
#undef FN

/* ================= instantiation 2: dual ================= */
#define T dual
#define K(x) d_const(x)
#define ADD(a, b) d_add(a, b)
#define SUB(a, b) d_sub(a, b)
#define MUL(a, b) d_mul(a, b)
#define DIV(a, b) d_div(a, b)
#define SQRT(a) d_sqrt(a)
#define COS(a) d_cos(a)
#define SIN(a) d_sin(a)
#define VAL(x_) ((x_).a)
#define FN(name) name##_j
#include "rsba_oracle_body.inc"

/* ---- exported entry points -------------------------------------------------------- */

/* primitives, for the known-answer tests restating test/mat_test.cc */
void rsba_oracle_rotate(const double* aa, const double* pt, double* out) { rotate_d(aa, pt, out); }
void rsba_oracle_slerp(const double* r0, const double* r1, double tau, double* out) { lerp3_d(r0, r1, tau, out); }
void rsba_oracle_interpolate_rs(const double* p0, const double* p1, int shutter, const int* scan,
                                const double* obs, double* out, int use_slerp) {
  interpolate_rs_d(p0, p1, shutter, scan, obs, out, use_slerp);
}
void rsba_oracle_w2c(const double* pose, const double* X, double* out) { w2c_d(pose, X, out); }
int rsba_oracle_w2i(const double* cam, const double* pose, const double* X, double* proj, int validate) {
  return w2i_d(cam, pose, X, proj, validate);
}
void rsba_oracle_distort(const double* cam, const double* img, double* out) { distort_d(cam, img, out); }

/* Same contract as rsba_ref_problem_eval (oracle/ref_driver.cc):
 * residuals[n][2]; jac[n][30] = J_pose0[2][6] | J_pose1[2][6] | J_point[2][3]; valid[n].
 * Returns the number of invalid observations (functor returned false: cam.h:410-412). */
long rsba_oracle_eval(long n, const double* obs_xy, const int* frame_idx, const int* point_idx,
                      const double* poses, const double* points, const double* cam9, int shutter,
                      const int* scanlines, int interpolate_rotation, double* residuals,
                      double* jac, unsigned char* valid, int nthreads) {
  long bad = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (long i = 0; i < n; ++i) {
    const double* p0 = poses + 12 * (long)frame_idx[i];
    const double* p1 = p0 + 6;
    const double* X = points + 3 * (long)point_idx[i];
    double r[2] = {0, 0};
    int ok;
    if (jac) {
      /* AutoDiffCostFunction: one dual per parameter scalar, seeded with a unit partial */
      dual jp0[6], jp1[6], jX[3], jr[2];
      for (int k = 0; k < 6; ++k) { jp0[k] = d_var(p0[k], k); jp1[k] = d_var(p1[k], 6 + k); }
      for (int k = 0; k < 3; ++k) jX[k] = d_var(X[k], 12 + k);
      ok = rs_residual_j(cam9, obs_xy + 2 * i, shutter, scanlines, interpolate_rotation, jp0, jp1, jX, jr);
      double* J = jac + 30 * i;
      if (ok) {
        r[0] = jr[0].a; r[1] = jr[1].a;
        for (int row = 0; row < 2; ++row) {
          for (int k = 0; k < 6; ++k) J[row * 6 + k] = jr[row].v[k];
          for (int k = 0; k < 6; ++k) J[12 + row * 6 + k] = jr[row].v[6 + k];
          for (int k = 0; k < 3; ++k) J[24 + row * 3 + k] = jr[row].v[12 + k];
        }
      } else {
        memset(J, 0, 30 * sizeof(double));
      }
    } else {
      ok = rs_residual_d(cam9, obs_xy + 2 * i, shutter, scanlines, interpolate_rotation, p0, p1, X, r);
      if (!ok) r[0] = r[1] = 0;
    }
    if (!ok) ++bad;
    if (residuals) { residuals[2 * i] = r[0]; residuals[2 * i + 1] = r[1]; }
    if (valid) valid[i] = (unsigned char)(ok ? 1 : 0);
  }
  return bad;
}

int rsba_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
