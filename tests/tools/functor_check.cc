// Test tool: drives the host functor structs of include/rsba_cuda_functors.hpp (the reference's
// ReprojectionError / RsBundleAdjustment signatures, T = double) over a flat scene so that
// tests/test_functors_host.py can compare them with the reference's own functors (oracle/_ref, golden files).
// Built with plain g++ -std=c++11 -Wall -Werror: the header must not need CUDA.
#include <cstring>
#include <vector>

#include "rsba_cuda_functors.hpp"

namespace {
struct Session {            // the members of gen::Session (gen-cpp/sfm_types.h) that the functor reads
  std::vector<double> cam;
  int rs;
  std::vector<int> scanlines;
};
struct Options {            // SfmOptions.h:20-28
  struct { bool interpolateRotation; } model;
};
typedef rsba_cuda::RsBundleAdjustmentT<Session, Options> RsBundleAdjustment;
}  // namespace

extern "C" {

// returns the number of observations whose operator() and Evaluate() disagree (must be 0)
long functor_eval_rs(long n, const double* xy, const int* frame, const int* point, const double* poses,
                     const double* points, const double* cam, int shutter, const int* scan, int interp, double* res,
                     double* J, unsigned char* valid, double* Jcam) {
  Session sess;
  sess.cam.assign(cam, cam + 9);
  sess.rs = shutter;
  sess.scanlines.assign(scan, scan + 2);
  Options opt;
  opt.model.interpolateRotation = interp != 0;
  long mismatches = 0;
  for (long i = 0; i < n; ++i) {
    const RsBundleAdjustment f(sess, opt, xy + 2 * i);
    const double* p0 = poses + 12L * frame[i];
    const double* p1 = p0 + 6;
    const double* X = points + 3L * point[i];
    double r[2] = {0, 0}, r2[2] = {0, 0}, r3[2] = {0, 0};
    const bool ok = f(p0, p1, X, r);                          // VideoSfmBaRs.h:25-35
    const bool ok2 = f(cam, p0, p1, X, r2);                   // VideoSfmBaRs.h:38-49
    double j0[12], j1[12], jx[6], jc[18];
    const double* params[3] = {p0, p1, X};
    double* jac[3] = {j0, j1, jx};
    const bool ok3 = f.Evaluate(params, r3, jac);
    if (ok != ok2 || ok != ok3 || std::memcmp(r, r2, sizeof(r)) || std::memcmp(r, r3, sizeof(r))) ++mismatches;
    valid[i] = ok ? 1 : 0;
    res[2 * i] = ok ? r[0] : 0.0;
    res[2 * i + 1] = ok ? r[1] : 0.0;
    std::memset(J + 30 * i, 0, 30 * sizeof(double));
    if (ok) {
      std::memcpy(J + 30 * i, j0, sizeof(j0));
      std::memcpy(J + 30 * i + 12, j1, sizeof(j1));
      std::memcpy(J + 30 * i + 24, jx, sizeof(jx));
    }
    if (Jcam) {
      const double* params4[4] = {cam, p0, p1, X};
      double k0[12], k1[12], kx[6];
      double* jac4[4] = {jc, k0, k1, kx};
      double r4[2] = {0, 0};
      const bool ok4 = f.EvaluateWithCam(params4, r4, jac4);
      if (ok4 != ok || (ok && (std::memcmp(r, r4, sizeof(r)) || std::memcmp(k0, j0, sizeof(j0)) ||
                               std::memcmp(k1, j1, sizeof(j1)) || std::memcmp(kx, jx, sizeof(jx)))))
        ++mismatches;
      std::memset(Jcam + 18 * i, 0, 18 * sizeof(double));
      if (ok4) std::memcpy(Jcam + 18 * i, jc, sizeof(jc));
    }
  }
  return mismatches;
}

// global-shutter single-pose functor (video_bundler_free.h:33-41): pose = the frame's first control pose
long functor_eval_single(long n, const double* xy, const int* frame, const int* point, const double* poses,
                         const double* points, const double* cam, double* res, double* Jpose, double* Jpoint,
                         unsigned char* valid) {
  long mismatches = 0;
  for (long i = 0; i < n; ++i) {
    const rsba_cuda::ReprojectionError f(cam, xy + 2 * i);
    const double* p0 = poses + 12L * frame[i];
    const double* X = points + 3L * point[i];
    double r[2] = {0, 0}, r2[2] = {0, 0};
    const bool ok = f(p0, X, r);
    const double* params[2] = {p0, X};
    double* jac[2] = {Jpose + 12 * i, Jpoint + 6 * i};
    std::memset(jac[0], 0, 12 * sizeof(double));
    std::memset(jac[1], 0, 6 * sizeof(double));
    const bool ok2 = f.Evaluate(params, r2, jac);
    if (ok != ok2 || std::memcmp(r, r2, sizeof(r))) ++mismatches;
    valid[i] = ok ? 1 : 0;
    res[2 * i] = ok ? r[0] : 0.0;
    res[2 * i + 1] = ok ? r[1] : 0.0;
  }
  return mismatches;
}

}  // extern "C"
