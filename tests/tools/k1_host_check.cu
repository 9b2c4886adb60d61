// TEST TOOL (not shipped): runs the product's per-observation K1 arithmetic
// (rsba_b200/csrc/reproj_math.cuh, __host__ __device__) on the CPU so that the math can be
// checked against the oracle in the build container, which has no GPU.
#include "../../rsba_b200/csrc/reproj_math.cuh"
extern "C" __attribute__((visibility("default")))
long k1_host_eval(long n, const double* xy, const int* fr, const int* pt, const double* poses,
                  const double* points, const double* cam9, int shutter, const int* scan, int interp_rot,
                  double* res, double* jac, unsigned char* valid) {
  rsba::CameraModel cm;
  for (int i = 0; i < 9; ++i) cm.cam[i] = cam9[i];
  cm.scan0 = scan[0];
  cm.scan_span = scan[1] - scan[0];
  cm.shutter = shutter;
  cm.interp_rot = interp_rot;
  long bad = 0;
  for (long i = 0; i < n; ++i) {
    const double* X = points + 3L * pt[i];
    rsba::Proj p = jac ? rsba::reproject<true>(cm, xy[2 * i], xy[2 * i + 1], poses + 12L * fr[i], X[0], X[1], X[2], jac + 30 * i)
                       : rsba::reproject<false>(cm, xy[2 * i], xy[2 * i + 1], poses + 12L * fr[i], X[0], X[1], X[2], nullptr);
    res[2 * i] = p.r0;
    res[2 * i + 1] = p.r1;
    valid[i] = p.ok;
    bad += !p.ok;
  }
  return bad;
}
