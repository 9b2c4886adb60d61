// K1 / K1r -- rolling-shutter reprojection residual and its 2x(6+6+3) Jacobian, FP64.
//
// Supersedes, per observation, RsBundleAdjustment::operator() evaluated under
// ceres::Jet<double,15> by AutoDiffCostFunction<.,2,6,6,3> (VideoSfmBaRs.h:25-35,58-63):
//   interpolate_rs   mat/cam.h:316-349   tau from observed_x for BOTH shutter directions
//                                        (the functor passes obs = {x, x}: VideoSfmBaRs.h:31)
//   interpolate/slerp mat/cam.h:294-311, 251-288   component-wise LINEAR interpolation of the
//                                        angle-axis vector and of the camera centre
//   w2c              mat/cam.h:355-366   X - c, then ceres::AngleAxisRotatePoint
//   w2i              mat/cam.h:401-419   fails (functor returns false) when z < 1e-8
//   c2i + distort    mat/cam.h:372-395, 49-72
//   residual         video_bundler_free.h:45-65
//   loss             optional ceres::HuberLoss, CeresHandler.h:85-90 (SfmOptions.h:64, default off)
// The Jacobian is the exact chain rule of those formulas (what forward-mode autodiff yields),
// written out by hand: because tau does not depend on the parameters,
//   J_pose0 = (1-tau) * J_pose,  J_pose1 = tau * J_pose  (rotation columns: 1 and 0 when
//   interpolateRotation is off), with J_pose the 2x6 Jacobian w.r.t. the interpolated pose.
//
// Data movement (B200): one warp = 32 consecutive observations (sorted by frame).  The
// frame's control poses are staged once per CTA in shared memory; xy / indices are read
// coalesced; the 32x30 Jacobian tile is assembled in shared memory and leaves the SM as ONE
// 7680-byte TMA bulk store (cp.async.bulk.global.shared::cta), so the 240 B/observation
// write stream -- 85 % of this kernel's HBM traffic -- is fully coalesced.
#include "common.cuh"
#include "reproj_math.cuh"

namespace rsba {

namespace {

constexpr int kK1Threads = 128;
constexpr int kK1Warps = kK1Threads / 32;
constexpr int kStageFrames = 8;  // control poses staged per CTA (frames spanned by 128 obs)
constexpr bool kStagePoses = false;       // experiment: stage the CTA's control poses in shared memory
constexpr int kPrefetchTiles = 148 * 5;   // CTAs resident at once on a B200 (148 SMs x 5 per SM by registers)

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Stage the control poses of the frames this CTA's observations span (sorted by frame, so
// [frame[first], frame[last]]).  Returns the first staged frame, or -1 if the span is too
// wide and poses must be read through L1 instead.
__device__ __forceinline__ int stage_poses(const ObsView& obs, const double* __restrict__ poses,
                                           long base, int cnt, double* s_pose) {
  const int f_lo = obs.frame[base];
  const int f_hi = obs.frame[base + cnt - 1];
  const int span = f_hi - f_lo + 1;
  if (span > kStageFrames || span <= 0) return -1;
  for (int i = threadIdx.x; i < span * kFrameParams; i += blockDim.x)
    s_pose[i] = poses[(long)f_lo * kFrameParams + i];
  return f_lo;
}

// COMPACT (the solver's own linearisation): `jac` receives the 12-double record of rsba_reproj_math.h and
// `tau_out` the observation's interpolation parameter -- 120 instead of 256 bytes written per observation.
template <bool JAC, bool CAM, bool COMPACT = false>
__global__ void __launch_bounds__(kK1Threads)
k1_kernel(const CameraModel cm_in, const ObsView obs, const double* __restrict__ poses,
          const double* __restrict__ points, double* __restrict__ residuals,
          double* __restrict__ jac, double* __restrict__ jac_cam, unsigned char* __restrict__ valid,
          double* __restrict__ cost_partials, int* __restrict__ invalid_count, double* __restrict__ tau_out = nullptr) {
  constexpr int kRec = COMPACT ? kJacCompact : kJacDoubles;
  CameraModel cm_local;
  if (CAM) {   // uncalibrated: the intrinsics are parameters, read at the point of evaluation
    cm_local = cm_in;
#pragma unroll
    for (int k = 0; k < 9; ++k) cm_local.cam[k] = __ldg(poses + cm_in.cam_offset + k);
  }
  const CameraModel& cm = CAM ? cm_local : cm_in;
  __shared__ double s_pose[kStageFrames * kFrameParams];
  __shared__ double s_cost[kK1Warps];
  extern __shared__ __align__(128) double s_jac[];  // [warps][32][kRec], JAC only

  const long base = (long)blockIdx.x * kK1Threads;
  const int cnt = (int)min((long)kK1Threads, obs.n - base);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long i = base + threadIdx.x;
  // The observation and its point are requested BEFORE the pose staging and its barrier: the two dependent
  // round trips (index -> point, frame -> poses) then overlap instead of queueing behind each other.
  double2 o = make_double2(0.0, 0.0);
  int f = 0;
  double X0 = 0.0, X1 = 0.0, X2 = 0.0;
  if (i < obs.n) {
    o = obs.xy[i];
    f = obs.frame[i];
    const double* pp = points + 3L * obs.point[i];
    X0 = pp[0]; X1 = pp[1]; X2 = pp[2];
  }
  // Half of this kernel's stall samples sat on those input round trips (profiles/r01_notes.md).  CTAs run in
  // blockIdx order, about one resident wave (kPrefetchTiles) at a time, so each CTA pulls the inputs of the
  // CTAs two and one wave ahead into L2: the observation lines of tile + 2 waves, and -- through the point
  // indices of tile + 1 wave, themselves prefetched a wave ago -- the 3-D points of tile + 1 wave.
  int p_ahead = -1;
  {
    const long j2 = i + 2L * kPrefetchTiles * kK1Threads, j1 = i + 1L * kPrefetchTiles * kK1Threads;
    if (j2 < obs.n) {
      if ((lane & 7) == 0) prefetch_l2(obs.xy + j2);        // 8 x 16 B = one 128-byte line
      if (lane == 0) { prefetch_l2(obs.frame + j2); prefetch_l2(obs.point + j2); }
    }
    if (j1 < obs.n) p_ahead = obs.point[j1];                // consumed after the arithmetic below
  }
  double pose[kFrameParams];
  if (kStagePoses) {
    const int staged = stage_poses(obs, poses, base, cnt, s_pose);
    __syncthreads();
    if (i < obs.n) {
      if (staged >= 0) {
        const double* sp = s_pose + (f - staged) * kFrameParams;
#pragma unroll
        for (int k = 0; k < kFrameParams; ++k) pose[k] = sp[k];
      } else {
        const double* gp = poses + (long)f * kFrameParams;
#pragma unroll
        for (int k = 0; k < kFrameParams; ++k) pose[k] = __ldg(gp + k);
      }
    }
  } else if (i < obs.n) {
    // no staging, no barrier: the frame's 96 bytes come through L1 (the threads of a warp share a frame, so
    // these are broadcast hits) as six 128-bit loads that queue right behind the point gather
    const double2* gp = reinterpret_cast<const double2*>(poses + (long)f * kFrameParams);
#pragma unroll
    for (int k = 0; k < kFrameParams / 2; ++k) {
      const double2 v = __ldg(gp + k);
      pose[2 * k] = v.x;
      pose[2 * k + 1] = v.y;
    }
  }

  double cost = 0.0;
  bool bad = false;
  double* Jrow = JAC ? s_jac + (warp * 32 + lane) * kRec : nullptr;
  if (i < obs.n) {
    double Jc[CAM ? 18 : 1];
    constexpr bool want_cam = JAC && CAM;
    Proj pr = reproject<JAC, false, true, COMPACT>(cm, o.x, o.y, pose, X0, X1, X2, Jrow, want_cam ? Jc : nullptr);
    if (COMPACT && tau_out) tau_out[i] = pr.tau;
    cost = pr.r0 * pr.r0 + pr.r1 * pr.r1;
    bad = !pr.ok;
    // ceres::HuberLoss(a) through Ceres' Corrector (third-party; CeresHandler.h:85-90 passes the loss to
    // every AddResidualBlock): rho(s) = s for s <= a^2, else 2 a sqrt(s) - a^2; rho'' <= 0, so the
    // correction is a plain rescaling of residual and Jacobian by sqrt(rho') and cost = 1/2 rho(s).
    if (cm.huber > 0.0 && cost > cm.huber * cm.huber) {
      const double sr = sqrt(cost);
      const double w = sqrt(cm.huber / sr);
      cost = 2.0 * cm.huber * sr - cm.huber * cm.huber;
      pr.r0 *= w;
      pr.r1 *= w;
      if (JAC) {
#pragma unroll
        for (int k = 0; k < kRec; ++k) Jrow[k] *= w;
        if (want_cam) {
#pragma unroll
          for (int k = 0; k < 18; ++k) Jc[k] *= w;
        }
      }
    }
    if (want_cam) {
      double2* dst = reinterpret_cast<double2*>(jac_cam + 18 * i);
#pragma unroll
      for (int k = 0; k < 9; ++k) dst[k] = make_double2(Jc[2 * k], Jc[2 * k + 1]);
    }
    if (residuals) reinterpret_cast<double2*>(residuals)[i] = make_double2(pr.r0, pr.r1);
    if (valid) valid[i] = pr.ok ? 1 : 0;
  }

  if (JAC) {
    // Hand the warp's tile to the async proxy and push it out with one bulk store.
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const long wbase = base + warp * 32;
    if (lane == 0 && wbase < obs.n) {
      const int wcnt = (int)min(32L, obs.n - wbase);
      const unsigned bytes = (unsigned)(wcnt * kRec * sizeof(double));  // multiple of 16
      const unsigned src = (unsigned)__cvta_generic_to_shared(s_jac + warp * 32 * kRec);
      double* dst = jac + wbase * kRec;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src),
                   "r"(bytes)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }

  if (p_ahead >= 0) prefetch_l2(points + 3L * p_ahead);

  // cost partial of this CTA (fixed reduction order -> deterministic cost)
  cost = warp_sum(cost);
  const unsigned badmask = __ballot_sync(0xffffffffu, bad);
  if (lane == 0) {
    s_cost[warp] = cost;
    if (badmask) atomicAdd(invalid_count, __popc(badmask));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kK1Warps; ++w) t += s_cost[w];
    cost_partials[blockIdx.x] = t;
  }
  if (JAC) {
    // shared memory must outlive the bulk store's read
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

// Track validation sweep: the predicate of validate(sess, f, opt, pt, obs) (struct/VideoSfM.cc:159-169)
// for every observation, as evalTracks runs it after each BA (VideoSfMHandler.cc:377-410, 599-600):
//   |c(tau) - X| >= minDistanceToCamera  and  w2i succeeds  and  |proj - obs|^2 < sqrdThreshold
// with the pose interpolated at the observation's own scan line (x or y by shutter direction).
__global__ void __launch_bounds__(kK1Threads)
validate_kernel(const CameraModel cm, const ObsView obs, const double* __restrict__ poses,
                const double* __restrict__ points, double sqrd_threshold, double min_distance,
                unsigned char* __restrict__ ok, double* __restrict__ sqrd_error) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= obs.n) return;
  const double2 o = obs.xy[i];
  const double* gp = poses + (long)obs.frame[i] * kFrameParams;
  const double* pp = points + 3L * obs.point[i];
  double pose[kFrameParams];
#pragma unroll
  for (int k = 0; k < kFrameParams; ++k) pose[k] = __ldg(gp + k);
  const double X0 = pp[0], X1 = pp[1], X2 = pp[2];
  CameraModel plain = cm;
  plain.huber = 0.0;
  if (cm.cam_offset >= 0) {
#pragma unroll
    for (int k = 0; k < 9; ++k) plain.cam[k] = __ldg(poses + cm.cam_offset + k);
  }
  const Proj pr = reproject<false, true>(plain, o.x, o.y, pose, X0, X1, X2, nullptr);
  // camera centre at the observation's scan line (interpolate, mat/cam.h:294-311)
  double tau = 0.0;
  if (cm.shutter != 0) {
    tau = ((cm.shutter == 2 ? o.y : o.x) - cm.scan0) / cm.scan_span;
    tau = tau < 0.0 ? 0.0 : (tau > 1.0 ? 1.0 : tau);
  }
  const double d0 = pose[3] + (pose[9] - pose[3]) * tau - X0;
  const double d1 = pose[4] + (pose[10] - pose[4]) * tau - X1;
  const double d2 = pose[5] + (pose[11] - pose[5]) * tau - X2;
  const double err = pr.r0 * pr.r0 + pr.r1 * pr.r1;
  const bool good = pr.ok && sqrt(d0 * d0 + d1 * d1 + d2 * d2) >= min_distance && err < sqrd_threshold;
  if (ok) ok[i] = good ? 1 : 0;
  if (sqrd_error) sqrd_error[i] = pr.ok ? err : -1.0;
}

// Iterative re-projection: reproject(sess, f, opt, pt, obs) (struct/VideoSfM.cc:139-155).  The scan line --
// hence the interpolated pose -- of a rolling-shutter projection is unknown, so the reference starts at the
// principal point and repeats  pose = getPose(proj); proj = w2i(cam, pose, pt)  until the projection moves by
// less than 1e-3 px (squared 1e-6), at most 49 times (limit = 50, pre-decremented); fails when w2i fails
// (z < 1e-8) or the limit is hit; the final ::vision::validate compares the projection with itself, i.e.
// passes iff sqrdThreshold > 0.  One thread per (frame, point) pair.
__global__ void __launch_bounds__(kK1Threads)
reproject_kernel(const CameraModel cm, long n, const int* __restrict__ frame, const int* __restrict__ point,
                 const double* __restrict__ poses, const double* __restrict__ points, double sqrd_threshold,
                 double2* __restrict__ proj_xy, unsigned char* __restrict__ ok) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* gp = poses + (long)frame[i] * kFrameParams;
  const double* pp = points + 3L * point[i];
  double pose[kFrameParams];
#pragma unroll
  for (int k = 0; k < kFrameParams; ++k) pose[k] = __ldg(gp + k);
  const double X0 = pp[0], X1 = pp[1], X2 = pp[2];
  CameraModel plain = cm;
  plain.huber = 0.0;
  if (cm.cam_offset >= 0) {
#pragma unroll
    for (int k = 0; k < 9; ++k) plain.cam[k] = __ldg(poses + cm.cam_offset + k);
  }
  double px = plain.cam[7], py = plain.cam[8];
  bool good = false;
  for (int limit = 49; limit >= 1; --limit) {
    // residual of the current guess against itself as "observation": r = w2i(pose(guess)) - guess
    const Proj pr = reproject<false, true>(plain, px, py, pose, X0, X1, X2, nullptr);
    if (!pr.ok) break;
    px += pr.r0;
    py += pr.r1;
    if (!(pr.r0 * pr.r0 + pr.r1 * pr.r1 > 1e-6)) { good = true; break; }
  }
  proj_xy[i] = make_double2(px, py);
  ok[i] = (good && 0.0 < sqrd_threshold) ? 1 : 0;
}

__global__ void __launch_bounds__(1024)
reduce_partials_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
  __shared__ double s[32];
  double t = 0.0;
  // eight loads in flight per thread (a rolled loop paid one L2 round trip per partial: ~20 us per call)
  int i = threadIdx.x;
  for (; i + 7 * (int)blockDim.x < n; i += 8 * blockDim.x) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = partials[i + u * blockDim.x];
#pragma unroll
    for (int u = 0; u < 8; ++u) t += v[u];
  }
  for (; i < n; i += blockDim.x) t += partials[i];
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x < 32) {
    t = s[threadIdx.x];
    t = warp_sum(t);
    if (threadIdx.x == 0) out[0] = 0.5 * t;  // cost = 1/2 sum r^2
  }
}

}  // namespace

int k1_num_partials(long n) { return (int)((n + kK1Threads - 1) / kK1Threads); }

void launch_k1(const CameraModel& cm, const ObsView& obs, const double* poses, const double* points,
               double* residuals, double* jac, double* jac_cam, unsigned char* valid, double* cost_partials,
               int* invalid_count, cudaStream_t stream) {
  if (obs.n <= 0) return;
  const int grid = k1_num_partials(obs.n);
  const size_t smem = (size_t)kK1Threads * kJacDoubles * sizeof(double);
  if (cm.cam_offset >= 0)
    k1_kernel<true, true><<<grid, kK1Threads, smem, stream>>>(cm, obs, poses, points, residuals, jac, jac_cam, valid,
                                                              cost_partials, invalid_count);
  else
    k1_kernel<true, false><<<grid, kK1Threads, smem, stream>>>(cm, obs, poses, points, residuals, jac, nullptr, valid,
                                                               cost_partials, invalid_count);
}

void launch_k1_compact(const CameraModel& cm, const ObsView& obs, const double* poses, const double* points,
                       double* residuals, double* jac_compact, double* tau, double* jac_cam, double* cost_partials,
                       int* invalid_count, cudaStream_t stream) {
  if (obs.n <= 0) return;
  const int grid = k1_num_partials(obs.n);
  const size_t smem = (size_t)kK1Threads * kJacCompact * sizeof(double);
  if (cm.cam_offset >= 0)
    k1_kernel<true, true, true><<<grid, kK1Threads, smem, stream>>>(cm, obs, poses, points, residuals, jac_compact, jac_cam,
                                                                    nullptr, cost_partials, invalid_count, tau);
  else
    k1_kernel<true, false, true><<<grid, kK1Threads, smem, stream>>>(cm, obs, poses, points, residuals, jac_compact, nullptr,
                                                                     nullptr, cost_partials, invalid_count, tau);
}

void launch_k1r(const CameraModel& cm, const ObsView& obs, const double* poses, const double* points,
                double* cost_partials, int* invalid_count, cudaStream_t stream, double* residuals,
                unsigned char* valid) {
  if (obs.n <= 0) return;
  const int grid = k1_num_partials(obs.n);
  if (cm.cam_offset >= 0)
    k1_kernel<false, true><<<grid, kK1Threads, 0, stream>>>(cm, obs, poses, points, residuals, nullptr, nullptr, valid,
                                                            cost_partials, invalid_count);
  else
    k1_kernel<false, false><<<grid, kK1Threads, 0, stream>>>(cm, obs, poses, points, residuals, nullptr, nullptr, valid,
                                                             cost_partials, invalid_count);
}

void launch_validate(const CameraModel& cm, const ObsView& obs, const double* poses, const double* points,
                     double sqrd_threshold, double min_distance, unsigned char* ok, double* sqrd_error,
                     cudaStream_t stream) {
  if (obs.n <= 0) return;
  validate_kernel<<<k1_num_partials(obs.n), kK1Threads, 0, stream>>>(cm, obs, poses, points, sqrd_threshold,
                                                                      min_distance, ok, sqrd_error);
}

void launch_reproject(const CameraModel& cm, long n, const int* frame, const int* point, const double* poses,
                      const double* points, double sqrd_threshold, double* proj_xy, unsigned char* ok,
                      cudaStream_t stream) {
  if (n <= 0) return;
  reproject_kernel<<<k1_num_partials(n), kK1Threads, 0, stream>>>(cm, n, frame, point, poses, points, sqrd_threshold,
                                                                 reinterpret_cast<double2*>(proj_xy), ok);
}

void launch_reduce_partials(const double* partials, int n, double* out, cudaStream_t stream) {
  reduce_partials_kernel<<<1, 1024, 0, stream>>>(partials, n, out);
}

}  // namespace rsba
