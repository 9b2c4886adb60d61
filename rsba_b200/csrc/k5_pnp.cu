// K5 -- batched rolling-shutter PnP.
//
// Supersedes the inner solve of vision::solveRsPnP (solveRSpnp.cpp:100-192): a ceres::Solve over the
// two control poses of ONE frame (12 parameters) with RsBA<float> residual blocks <2; 6, 6>
// (solveRSpnp.cpp:25-97: the BA functor with the 3-D point fixed and w2i(..., validate = false)),
// max_num_iterations = 10 -- which RANSAC runs once per hypothesis on a minimal sample
// (pnpTask, solveRSpnp.cpp:265-335) -- and the inlier count each hypothesis is scored with
// (project3dPoints + the reprojectionError test, solveRSpnp.cpp:226-263, 318-323; there the pose is
// interpolated at the observation's OWN scan line, x or y by shutter direction).
//
// One warp per hypothesis, the whole trust-region loop on the device (no host round trips): lanes
// evaluate the sample's residuals and 2x12 Jacobians (the K1 arithmetic, reproj_math.cuh) into shared
// memory, the 12x12 normal equations are reduced cooperatively, lane 0 runs the damped Cholesky
// solve; the loop restates the same Ceres 1.9.0 Levenberg-Marquardt rules as lm_solver.cu (Jacobi
// scaling fixed at the first iteration, D^2 = clamp(diag)/radius, rho test, radius update,
// function / parameter / gradient tolerances).  Then the warp sweeps all points for the inlier count.
#include "common.cuh"
#include "reproj_math.cuh"
#include "problem.cuh"

namespace rsba {
namespace {

constexpr int kPnpWarps = 4;
constexpr int kPnpMaxSample = 32;   // points of one hypothesis (RANSAC minimal samples are ~6)

struct PnpOptions {
  int max_iterations;
  double radius0, max_radius, min_radius, min_rel_decrease, min_diag, max_diag, f_tol, g_tol, p_tol;
  double inlier_threshold;   // pixels (norm, not squared: solveRSpnp.cpp:320)
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// lane 0: solve (diag(s) H diag(s) + D) y = s g in place by Cholesky; returns false if not positive definite.
// H (12x12, full, row-major) and g are left untouched; y [12] out; d2 [12] out.
__device__ bool damped_solve12(const double* H, const double* g, const double* s, double radius, double min_diag,
                               double max_diag, double* L /*[144] scratch*/, double* y, double* d2) {
  for (int i = 0; i < 12; ++i)
    for (int j = 0; j <= i; ++j) L[i * 12 + j] = s[i] * H[i * 12 + j] * s[j];
  for (int i = 0; i < 12; ++i) {
    d2[i] = fmin(fmax(L[i * 13], min_diag), max_diag) / radius;
    L[i * 13] += d2[i];
  }
  for (int j = 0; j < 12; ++j) {
    double d = L[j * 13];
    for (int m = 0; m < j; ++m) d -= L[j * 12 + m] * L[j * 12 + m];
    if (!(d > 0.0)) return false;
    d = sqrt(d);
    L[j * 13] = d;
    for (int i = j + 1; i < 12; ++i) {
      double v = L[i * 12 + j];
      for (int m = 0; m < j; ++m) v -= L[i * 12 + m] * L[j * 12 + m];
      L[i * 12 + j] = v / d;
    }
  }
  for (int i = 0; i < 12; ++i) {          // forward
    double v = s[i] * g[i];
    for (int m = 0; m < i; ++m) v -= L[i * 12 + m] * y[m];
    y[i] = v / L[i * 13];
  }
  for (int i = 11; i >= 0; --i) {         // backward
    double v = y[i];
    for (int m = i + 1; m < 12; ++m) v -= L[m * 12 + i] * y[m];
    y[i] = v / L[i * 13];
  }
  return true;
}

struct WarpSmem {
  double J[2 * kPnpMaxSample * 12];   // rows of the sample's camera Jacobian
  double r[2 * kPnpMaxSample];
  double H[144], L[144];
  double g[12], s[12], y[12], d2[12], x[12], xt[12];
  double sc[4];                       // cost, ok flag, ...
};

__global__ void __launch_bounds__(kPnpWarps * 32)
pnp_batch_kernel(const CameraModel cm, const double* __restrict__ points, const double* __restrict__ obs_xy,
                 int n_points, int n_hyp, int sample_size, const int* __restrict__ sample_idx,
                 double* __restrict__ poses, PnpOptions opt, double* __restrict__ final_cost,
                 int* __restrict__ usable, int* __restrict__ iterations, int* __restrict__ inliers) {
  extern __shared__ __align__(16) unsigned char pnp_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hyp = blockIdx.x * kPnpWarps + warp;
  if (hyp >= n_hyp) return;
  WarpSmem& W = reinterpret_cast<WarpSmem*>(pnp_smem)[warp];
  const int* idx = sample_idx + (long)hyp * sample_size;
  if (lane < 12) W.x[lane] = poses[12L * hyp + lane];
  __syncwarp();

  // residuals (+ Jacobian rows) of the sample at pose `x`; returns cost = 1/2 sum r^2 (all lanes)
  auto evaluate = [&](const double* x, bool jac) -> double {
    double c = 0.0;
    for (int q = lane; q < sample_size; q += 32) {
      const int i = idx[q];
      const double X0 = points[3L * i], X1 = points[3L * i + 1], X2 = points[3L * i + 2];
      const double ox = obs_xy[2L * i], oy = obs_xy[2L * i + 1];
      double pose[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) pose[k] = x[k];
      double Jl[kJacDoubles];
      Proj pr;
      if (jac) pr = reproject<true, false, false>(cm, ox, oy, pose, X0, X1, X2, Jl);
      else     pr = reproject<false, false, false>(cm, ox, oy, pose, X0, X1, X2, nullptr);
      c += pr.r0 * pr.r0 + pr.r1 * pr.r1;
      if (jac) {
        W.r[2 * q] = pr.r0;
        W.r[2 * q + 1] = pr.r1;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          W.J[(2 * q) * 12 + k] = Jl[k];            W.J[(2 * q) * 12 + 6 + k] = Jl[12 + k];
          W.J[(2 * q + 1) * 12 + k] = Jl[6 + k];    W.J[(2 * q + 1) * 12 + 6 + k] = Jl[18 + k];
        }
      }
    }
    return 0.5 * warp_sum_d(c);
  };
  // H = J^T J, g = J^T r from the rows in shared memory (fixed summation order)
  auto normal_equations = [&]() {
    __syncwarp();
    for (int e = lane; e < 144 + 12; e += 32) {
      double sacc = 0.0;
      if (e < 144) {
        const int a = e / 12, b = e % 12;
        for (int row = 0; row < 2 * sample_size; ++row) sacc += W.J[row * 12 + a] * W.J[row * 12 + b];
        W.H[e] = sacc;
      } else {
        const int a = e - 144;
        for (int row = 0; row < 2 * sample_size; ++row) sacc += W.J[row * 12 + a] * W.r[row];
        W.g[a] = sacc;
      }
    }
    __syncwarp();
  };

  double cost = evaluate(W.x, true);
  normal_equations();
  if (lane < 12) W.s[lane] = 1.0 / (1.0 + sqrt(W.H[lane * 13]));   // Jacobi scaling, kept
  __syncwarp();
  double radius = opt.radius0, decrease = 2.0;
  int it = 0, ok = 1;
  double gmax = 0.0, xnorm = 0.0;
  {
    double gl = lane < 12 ? fabs(W.g[lane]) : 0.0, xl = lane < 12 ? W.x[lane] * W.x[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gl = fmax(gl, __shfl_xor_sync(0xffffffffu, gl, o));
    gmax = gl;
    xnorm = sqrt(warp_sum_d(xl));
  }
  while (ok && gmax > opt.g_tol && it < opt.max_iterations) {
    ++it;
    if (lane == 0) W.sc[1] = damped_solve12(W.H, W.g, W.s, radius, opt.min_diag, opt.max_diag, W.L, W.y, W.d2) ? 1.0 : 0.0;
    __syncwarp();
    if (W.sc[1] == 0.0) { ok = 0; break; }
    // step, model cost change  -1/2 g.delta + 1/2 sum D^2 y^2   (y = scaled step, as lm_solver.cu)
    double dl = 0.0, gd = 0.0, dd = 0.0;
    if (lane < 12) {
      dl = -W.s[lane] * W.y[lane];
      W.xt[lane] = W.x[lane] + dl;
      gd = W.g[lane] * dl;
      dd = W.d2[lane] * W.y[lane] * W.y[lane];
    }
    __syncwarp();
    const double mcc = -0.5 * warp_sum_d(gd) + 0.5 * warp_sum_d(dd);
    const double step_norm = sqrt(warp_sum_d(dl * dl));
    const double new_cost = evaluate(W.xt, false);
    bool accepted = false;
    double rho = 0.0;
    if (mcc > 0.0) {
      if (step_norm <= opt.p_tol * (xnorm + opt.p_tol)) break;
      if (fabs(cost - new_cost) < opt.f_tol * cost) break;
      rho = (cost - new_cost) / mcc;
      accepted = rho > opt.min_rel_decrease;
    }
    if (accepted) {
      if (lane < 12) W.x[lane] = W.xt[lane];
      __syncwarp();
      radius = fmin(opt.max_radius, radius / fmax(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) * (2.0 * rho - 1.0) * (2.0 * rho - 1.0)));
      decrease = 2.0;
      cost = evaluate(W.x, true);
      normal_equations();
      double gl = lane < 12 ? fabs(W.g[lane]) : 0.0, xl = lane < 12 ? W.x[lane] * W.x[lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) gl = fmax(gl, __shfl_xor_sync(0xffffffffu, gl, o));
      gmax = gl;
      xnorm = sqrt(warp_sum_d(xl));
    } else {
      radius /= decrease;
      decrease *= 2.0;
      if (radius < opt.min_radius) break;
    }
  }
  if (lane < 12) poses[12L * hyp + lane] = W.x[lane];
  if (lane == 0) {
    final_cost[hyp] = cost;
    usable[hyp] = ok;
    iterations[hyp] = it;
  }
  // ---- inlier count of the refined hypothesis over ALL points (solveRSpnp.cpp:318-323)
  if (inliers) {
    __syncwarp();
    double pose[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) pose[k] = W.x[k];
    int cnt = 0;
    for (int i = lane; i < n_points; i += 32) {
      const Proj pr = reproject<false, true, false>(cm, obs_xy[2L * i], obs_xy[2L * i + 1], pose, points[3L * i],
                                                    points[3L * i + 1], points[3L * i + 2], nullptr);
      cnt += sqrt(pr.r0 * pr.r0 + pr.r1 * pr.r1) < opt.inlier_threshold ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) inliers[hyp] = cnt;
  }
}

}  // namespace
}  // namespace rsba

using namespace rsba;

extern "C" int rsba_cuda_pnp_batch(rsba_problem* h, const double cam9[9], int shutter, const int scanlines[2],
                                   int n_points, const double* points3d, const double* obs_xy, int n_hyp,
                                   int sample_size, const int* sample_idx, double* poses,
                                   const rsba_solve_options* options, double inlier_threshold, double* final_cost,
                                   int* usable, int* iterations, int* inlier_count) {
  return rsba::api_guard([&]() -> int {
  auto bad = [](const char* m) { set_last_error(m); return (int)RSBA_ERR_INVALID_ARGUMENT; };
  if (!h || !cam9 || !scanlines || !points3d || !obs_xy || !sample_idx || !poses || !options) return bad("NULL argument");
  if (n_points <= 0 || n_hyp < 0 || sample_size <= 0 || sample_size > kPnpMaxSample)
    return bad("sample_size must be in 1..32 (larger point sets: rsba_cuda_solve with constant points)");
  if (shutter < 0 || shutter > 2) return bad("shutter must be 0, 1 or 2");
  for (long k = 0; k < (long)n_hyp * sample_size; ++k)
    if (sample_idx[k] < 0 || sample_idx[k] >= n_points) return bad("sample index out of range");
  if (n_hyp == 0) return RSBA_OK;
  RSBA_CUDA_TRY(cudaSetDevice(h->device));
  CameraModel cm{};
  memcpy(cm.cam, cam9, sizeof(cm.cam));
  cm.shutter = shutter;
  cm.scan0 = (double)scanlines[0];
  cm.scan_span = (double)(scanlines[1] - scanlines[0]);
  cm.interp_rot = 1;   // RsBA calls interpolate_rs with its default useSlerp = true (solveRSpnp.cpp:57)
  cm.huber = 0.0;
  DeviceBuffer<double> d_pts, d_obs, d_poses, d_cost;
  DeviceBuffer<int> d_idx, d_flags;
  RSBA_CUDA_TRY(d_pts.resize(3 * (size_t)n_points));
  RSBA_CUDA_TRY(d_obs.resize(2 * (size_t)n_points));
  RSBA_CUDA_TRY(d_poses.resize(12 * (size_t)n_hyp));
  RSBA_CUDA_TRY(d_cost.resize(n_hyp));
  RSBA_CUDA_TRY(d_idx.resize((size_t)n_hyp * sample_size));
  RSBA_CUDA_TRY(d_flags.resize(3 * (size_t)n_hyp));
  cudaStream_t s = h->stream;
  RSBA_CUDA_TRY(cudaMemcpyAsync(d_pts.ptr, points3d, d_pts.bytes(), cudaMemcpyHostToDevice, s));
  RSBA_CUDA_TRY(cudaMemcpyAsync(d_obs.ptr, obs_xy, d_obs.bytes(), cudaMemcpyHostToDevice, s));
  RSBA_CUDA_TRY(cudaMemcpyAsync(d_poses.ptr, poses, d_poses.bytes(), cudaMemcpyHostToDevice, s));
  RSBA_CUDA_TRY(cudaMemcpyAsync(d_idx.ptr, sample_idx, d_idx.bytes(), cudaMemcpyHostToDevice, s));
  PnpOptions o{options->max_num_iterations, options->initial_trust_region_radius, options->max_trust_region_radius,
               options->min_trust_region_radius, options->min_relative_decrease, options->min_lm_diagonal,
               options->max_lm_diagonal, options->function_tolerance, options->gradient_tolerance,
               options->parameter_tolerance, inlier_threshold};
  static bool seen[64] = {};
  const size_t smem = kPnpWarps * sizeof(WarpSmem);
  if (first_use_on_device(seen))
    cudaFuncSetAttribute(pnp_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  stage_begin(h, kStagePnp);
  pnp_batch_kernel<<<(n_hyp + kPnpWarps - 1) / kPnpWarps, kPnpWarps * 32, smem, s>>>(
      cm, d_pts.ptr, d_obs.ptr, n_points, n_hyp, sample_size, d_idx.ptr, d_poses.ptr, o, d_cost.ptr, d_flags.ptr,
      d_flags.ptr + n_hyp, inlier_count ? d_flags.ptr + 2 * (size_t)n_hyp : nullptr);
  stage_end(h, kStagePnp);
  h->launches += 1;
  RSBA_CUDA_TRY(cudaGetLastError());
  RSBA_CUDA_TRY(cudaMemcpyAsync(poses, d_poses.ptr, d_poses.bytes(), cudaMemcpyDeviceToHost, s));
  if (final_cost) RSBA_CUDA_TRY(cudaMemcpyAsync(final_cost, d_cost.ptr, d_cost.bytes(), cudaMemcpyDeviceToHost, s));
  if (usable) RSBA_CUDA_TRY(cudaMemcpyAsync(usable, d_flags.ptr, n_hyp * sizeof(int), cudaMemcpyDeviceToHost, s));
  if (iterations) RSBA_CUDA_TRY(cudaMemcpyAsync(iterations, d_flags.ptr + n_hyp, n_hyp * sizeof(int), cudaMemcpyDeviceToHost, s));
  if (inlier_count) RSBA_CUDA_TRY(cudaMemcpyAsync(inlier_count, d_flags.ptr + 2 * (size_t)n_hyp, n_hyp * sizeof(int), cudaMemcpyDeviceToHost, s));
  RSBA_CUDA_TRY(cudaStreamSynchronize(s));
  stage_collect(h, kStagePnp);
  return RSBA_OK;
  });
}
