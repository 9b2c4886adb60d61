// K4 -- back-substitution of the point blocks, trial parameters and the scalars the
// trust-region logic needs.  Supersedes SchurEliminator::BackSubstitute and the bookkeeping
// inside TrustRegionMinimizer (Ceres 1.9.0, third-party; reached through ceres::Solve,
// CeresHandler.h:419).
//
//   delta_c = -s_c * y_c
//   delta_p = -Cinv_p ( g_p + sum_{i in p} Jx_i^T (Jc_i delta_c[frame_i]) )      (unscaled)
//   model_cost_change = -1/2 g^T delta + 1/2 sum D^2 (delta / s)^2
// (the last line equals Ceres' -m.(r + m/2), m = J delta, because the linear solve is direct:
//  (J'^T J' + D^2) delta' = -g').  All reductions are two-stage with a fixed order.
#include "lm.cuh"

#include <cstdlib>

namespace rsba {
namespace {

constexpr int kRedThreads = 256;
constexpr int kWideThreads = 1024;   // single-CTA reductions over 12 F / n_blocks values: latency-bound

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) t += sh[w];
  __syncthreads();
  return t;  // valid on thread 0
}

__device__ __forceinline__ double block_max(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) t = fmax(t, sh[w]);
  __syncthreads();
  return t;
}

// frames: delta_c, trial poses, partial scalars -> scratch[0..2] (single CTA)
__global__ void __launch_bounds__(kWideThreads)
frame_step_kernel(NormalEq ne, const int* __restrict__ tile_pos, const double* __restrict__ y, int n,
                  const double* __restrict__ poses, double* __restrict__ delta_c, double* __restrict__ trial,
                  double* __restrict__ scratch, int bounded, double lower_bound) {
  __shared__ double sh[32];
  double gd = 0.0, dd = 0.0, nn = 0.0;
#pragma unroll 4   // single CTA, latency-bound: keep several iterations' loads in flight
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const double sc = ne.scale_c[t];
    // y lives in the (tile-permuted) order of the reduced system
    const double ys = y[(long)tile_pos[t / kTile] * kTile + t % kTile];
    const double d = -sc * ys;
    delta_c[t] = d;
    // the one bounded parameter (free interFrameRatio, SetParameterLowerBound at CeresHandler.h:161,172):
    // Ceres' Plus projects the trial point onto the feasible set
    trial[t] = (t == bounded) ? fmax(lower_bound, poses[t] + d) : poses[t] + d;
    gd += ne.gc[t] * d;
    dd += ne.d2_c[t] * ys * ys;
    nn += d * d;
  }
  gd = block_sum(gd, sh);
  dd = block_sum(dd, sh);
  nn = block_sum(nn, sh);
  if (threadIdx.x == 0) { scratch[0] = gd; scratch[1] = dd; scratch[2] = nn; }
}

// points: back-substitution, trial points, per-CTA partial scalars -> scratch[3 + 3*block ...]
// One WARP per point: the lanes gather the point's observations in parallel (a serial loop of
// dependent 240-byte gathers per thread ran at 1 TB/s), then a fixed-order butterfly sums them.
constexpr int kPointStepWarps = 8;

__global__ void __launch_bounds__(kPointStepWarps * 32)
point_step_kernel(SchurStructure st, JacView jv,
                  const double* __restrict__ jac_cam, int cam_frame, NormalEq ne,
                  const double* __restrict__ delta_c, int n_points, const double* __restrict__ points,
                  double* __restrict__ delta_p, double* __restrict__ trial, double* __restrict__ scratch, int block_offset) {
  __shared__ double sh[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * kPointStepWarps + warp;
  double gd = 0.0, dd = 0.0, nn = 0.0;
  if (k < ne.n_owned) {
    const int p = owned_point(ne, k);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    const int beg = st.pt_ptr[p], end = st.pt_ptr[p + 1];
    for (int e = beg + lane; e < end; e += 32) {
      const long i = st.pt_obs[e];
      // compact record: Jc delta_c = jr . (wr0 d_rot0 + wr1 d_rot1) - jx . ((1-tau) d_c0 + tau d_c1), i.e. the
      // Jacobian of the INTERPOLATED pose applied to the interpolated camera step -- 96 bytes per observation
      const double2* rp = reinterpret_cast<const double2*>(jv.rec + (jv.point_major ? (long)e : i) * kJacCompact);
      const double2 a01 = rp[0], a2b0 = rp[1], b12 = rp[2], c01 = rp[3], c2d0 = rp[4], d12 = rp[5];
      const double tau = st.pt_tau[e];
      const double2* dc = reinterpret_cast<const double2*>(delta_c + 12L * st.pt_frame[e]);
      const double2 u0 = dc[0], u1 = dc[1], u2 = dc[2], u3 = dc[3], u4 = dc[4], u5 = dc[5];
      const double th0 = 1.0 - tau, wr0 = jv.rot_interp ? th0 : 1.0, wr1 = jv.rot_interp ? tau : 0.0;
      const double dr0 = wr0 * u0.x + wr1 * u3.x, dr1 = wr0 * u0.y + wr1 * u3.y, dr2 = wr0 * u1.x + wr1 * u4.x;
      const double dq0 = th0 * u1.y + tau * u4.y, dq1 = th0 * u2.x + tau * u5.x, dq2 = th0 * u2.y + tau * u5.y;
      double m0 = c01.x * dr0 + c01.y * dr1 + c2d0.x * dr2 - (a01.x * dq0 + a01.y * dq1 + a2b0.x * dq2);
      double m1 = c2d0.y * dr0 + d12.x * dr1 + d12.y * dr2 - (a2b0.y * dq0 + b12.x * dq1 + b12.y * dq2);
      if (jac_cam) {   // uncalibrated variant: + Jcam * delta_intrinsics
        const double* jc = jac_cam + i * 18;
        const double* di = delta_c + 12L * cam_frame;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          m0 += jc[k] * di[k];
          m1 += jc[9 + k] * di[k];
        }
      }
      a0 += a01.x * m0 + a2b0.y * m1;
      a1 += a01.y * m0 + b12.x * m1;
      a2 += a2b0.x * m0 + b12.y * m1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) {
      const double* g = ne.gp + 3L * p;
      a0 += g[0]; a1 += g[1]; a2 += g[2];
      const double* Ci = ne.Cinv + 6L * p;
      const double d0 = -(Ci[0] * a0 + Ci[1] * a1 + Ci[2] * a2);
      const double d1 = -(Ci[1] * a0 + Ci[3] * a1 + Ci[4] * a2);
      const double d2 = -(Ci[2] * a0 + Ci[4] * a1 + Ci[5] * a2);
      delta_p[3L * p] = d0; delta_p[3L * p + 1] = d1; delta_p[3L * p + 2] = d2;
      trial[3L * p] = points[3L * p] + d0;
      trial[3L * p + 1] = points[3L * p + 1] + d1;
      trial[3L * p + 2] = points[3L * p + 2] + d2;
      gd = g[0] * d0 + g[1] * d1 + g[2] * d2;
      nn = d0 * d0 + d1 * d1 + d2 * d2;
      const double* sp = ne.scale_p + 3L * p;
      const double* e2 = ne.d2_p + 3L * p;
      const double u0 = d0 / sp[0], u1 = d1 / sp[1], u2 = d2 / sp[2];
      dd = e2[0] * u0 * u0 + e2[1] * u1 * u1 + e2[2] * u2 * u2;
    }
  }
  gd = block_sum(gd, sh);
  dd = block_sum(dd, sh);
  nn = block_sum(nn, sh);
  if (threadIdx.x == 0) {
    double* o = scratch + 3 + 3L * (blockIdx.x + block_offset);
    o[0] = gd; o[1] = dd; o[2] = nn;
  }
}

// The same back-substitution with one thread per observation over GROUPS of whole points (host: greedy over the
// point-major order, <= 256 observations and <= 128 points per CTA, ~10 points at C3; PointGroups in lm.cuh): the observation's term Jx^T (Jc delta_c) goes to shared memory, (component, point) pairs are
// summed serially in observation order, one thread per point finishes.  Reads the point-major records coalesced.
constexpr int kStepGroupThreads = 256;
constexpr int kStepGroupMaxPoints = 128;

__global__ void __launch_bounds__(kStepGroupThreads)
point_step_group_kernel(SchurStructure st, const int2* __restrict__ groups, JacView jv, NormalEq ne,
                        const double* __restrict__ delta_c, const double* __restrict__ points, double* __restrict__ delta_p,
                        double* __restrict__ trial, double* __restrict__ scratch) {
  __shared__ double prod[3][kStepGroupThreads];
  __shared__ double sums[3][kStepGroupMaxPoints];
  __shared__ int s_ptr[kStepGroupMaxPoints + 1];
  __shared__ double sh[32];
  const int tid = threadIdx.x;
  const int2 g = groups[blockIdx.x];
  const int p_lo = g.x, npts = g.y - g.x;
  const int e_lo = st.pt_ptr[p_lo], e_hi = st.pt_ptr[g.y];
  if (tid <= npts) s_ptr[tid] = st.pt_ptr[p_lo + tid] - e_lo;
  const int e = e_lo + tid;
  double c0 = 0.0, c1 = 0.0, c2 = 0.0;
  if (e < e_hi) {
    const double2* rp = reinterpret_cast<const double2*>(jv.rec + (long)e * kJacCompact);
    const double2 a01 = rp[0], a2b0 = rp[1], b12 = rp[2], c01 = rp[3], c2d0 = rp[4], d12 = rp[5];
    const double tau = st.pt_tau[e];
    const double2* dc = reinterpret_cast<const double2*>(delta_c + 12L * st.pt_frame[e]);
    const double2 u0 = dc[0], u1 = dc[1], u2 = dc[2], u3 = dc[3], u4 = dc[4], u5 = dc[5];
    const double th0 = 1.0 - tau, wr0 = jv.rot_interp ? th0 : 1.0, wr1 = jv.rot_interp ? tau : 0.0;
    const double dr0 = wr0 * u0.x + wr1 * u3.x, dr1 = wr0 * u0.y + wr1 * u3.y, dr2 = wr0 * u1.x + wr1 * u4.x;
    const double dq0 = th0 * u1.y + tau * u4.y, dq1 = th0 * u2.x + tau * u5.x, dq2 = th0 * u2.y + tau * u5.y;
    const double m0 = c01.x * dr0 + c01.y * dr1 + c2d0.x * dr2 - (a01.x * dq0 + a01.y * dq1 + a2b0.x * dq2);
    const double m1 = c2d0.y * dr0 + d12.x * dr1 + d12.y * dr2 - (a2b0.y * dq0 + b12.x * dq1 + b12.y * dq2);
    c0 = a01.x * m0 + a2b0.y * m1;
    c1 = a01.y * m0 + b12.x * m1;
    c2 = a2b0.x * m0 + b12.y * m1;
  }
  prod[0][tid] = c0; prod[1][tid] = c1; prod[2][tid] = c2;
  __syncthreads();
  for (int u = tid; u < 3 * npts; u += kStepGroupThreads) {
    const int comp = u / npts, pq = u - comp * npts;
    double sum = 0.0;
    for (int j = s_ptr[pq]; j < s_ptr[pq + 1]; ++j) sum += prod[comp][j];
    sums[comp][pq] = sum;
  }
  __syncthreads();
  double gd = 0.0, dd = 0.0, nn = 0.0;
  if (tid < npts) {
    const int p = p_lo + tid;
    if (ne.point_owned[p]) {
      const double* gp = ne.gp + 3L * p;
      const double a0 = sums[0][tid] + gp[0], a1 = sums[1][tid] + gp[1], a2 = sums[2][tid] + gp[2];
      const double* Ci = ne.Cinv + 6L * p;
      const double d0 = -(Ci[0] * a0 + Ci[1] * a1 + Ci[2] * a2);
      const double d1 = -(Ci[1] * a0 + Ci[3] * a1 + Ci[4] * a2);
      const double d2 = -(Ci[2] * a0 + Ci[4] * a1 + Ci[5] * a2);
      delta_p[3L * p] = d0; delta_p[3L * p + 1] = d1; delta_p[3L * p + 2] = d2;
      trial[3L * p] = points[3L * p] + d0;
      trial[3L * p + 1] = points[3L * p + 1] + d1;
      trial[3L * p + 2] = points[3L * p + 2] + d2;
      gd = gp[0] * d0 + gp[1] * d1 + gp[2] * d2;
      nn = d0 * d0 + d1 * d1 + d2 * d2;
      const double* sp = ne.scale_p + 3L * p;
      const double* e2 = ne.d2_p + 3L * p;
      const double u0 = d0 / sp[0], u1 = d1 / sp[1], u2 = d2 / sp[2];
      dd = e2[0] * u0 * u0 + e2[1] * u1 * u1 + e2[2] * u2 * u2;
    }
  }
  gd = block_sum(gd, sh);
  dd = block_sum(dd, sh);
  nn = block_sum(nn, sh);
  if (threadIdx.x == 0) {
    double* o = scratch + 3 + 3L * blockIdx.x;
    o[0] = gd; o[1] = dd; o[2] = nn;
  }
}

// final: scalars[0..2] = frame part, scalars[8..10] = sum of point partials (fixed order)
__global__ void __launch_bounds__(kWideThreads)
step_final_kernel(const double* __restrict__ scratch, int n_blocks, double* __restrict__ scalars) {
  __shared__ double sh[32];
  double a = 0.0, b = 0.0, c = 0.0;
#pragma unroll 8
  for (int k = threadIdx.x; k < n_blocks; k += blockDim.x) {
    a += scratch[3 + 3L * k];
    b += scratch[4 + 3L * k];
    c += scratch[5 + 3L * k];
  }
  a = block_sum(a, sh);
  b = block_sum(b, sh);
  c = block_sum(c, sh);
  if (threadIdx.x == 0) {
    // camera part (identical on every rank) | point part (this rank's points; summed over ranks)
    scalars[0] = scratch[0]; scalars[8] = a;    // g . delta
    scalars[1] = scratch[1]; scalars[9] = b;    // sum D^2 delta'^2
    scalars[2] = scratch[2]; scalars[10] = c;   // |delta|^2
  }
}

// |x|^2 over parameter blocks that are not entirely constant, max |g| over free parameters:
// point part (owned points only) ...
__global__ void __launch_bounds__(kRedThreads)
point_norms_kernel(NormalEq ne, int n_points, const double* __restrict__ points, double* __restrict__ scratch) {
  __shared__ double sh[32];
  double xx = 0.0, gm = 0.0;
  const long np = 3L * ne.n_owned;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < np; t += (long)gridDim.x * blockDim.x) {
    const long p = owned_point(ne, (int)(t / 3));
    const long u = 3 * p + t % 3;
    if (!ne.point_const[p]) {
      xx += points[u] * points[u];
      gm = fmax(gm, fabs(ne.gp[u]));
    }
  }
  xx = block_sum(xx, sh);
  gm = block_max(gm, sh);
  if (threadIdx.x == 0) { scratch[2L * blockIdx.x] = xx; scratch[2L * blockIdx.x + 1] = gm; }
}

__global__ void __launch_bounds__(kRedThreads)
point_norms_final_kernel(const double* __restrict__ scratch, int n_blocks, double* __restrict__ out_xx,
                         double* __restrict__ out_gmax) {
  __shared__ double sh[32];
  double xx = 0.0, gm = 0.0;
  for (int k = threadIdx.x; k < n_blocks; k += blockDim.x) {
    xx += scratch[2L * k];
    gm = fmax(gm, scratch[2L * k + 1]);
  }
  xx = block_sum(xx, sh);
  gm = block_max(gm, sh);
  if (threadIdx.x == 0) { *out_xx = xx; *out_gmax = gm; }
}

// ... and camera part (single CTA; 12 F values)
__global__ void __launch_bounds__(kWideThreads)
camera_norms_kernel(NormalEq ne, int n_frames, const double* __restrict__ poses, double* __restrict__ scalars) {
  __shared__ double sh[32];
  double xx = 0.0, gm = 0.0;
#pragma unroll 4
  for (int t = threadIdx.x; t < 12 * n_frames; t += blockDim.x) {
    const int f = t / 12, k = t % 12;
    const unsigned m = ne.pose_mask[f];
    const unsigned blockbits = (k < 6) ? (m & 0x3F) : ((m >> 6) & 0x3F);
    if (blockbits != 0x3F) xx += poses[t] * poses[t];
    if (!((m >> k) & 1)) gm = fmax(gm, fabs(ne.gc[t]));
  }
  xx = block_sum(xx, sh);
  gm = block_max(gm, sh);
  if (threadIdx.x == 0) { scalars[3] = xx; scalars[4] = gm; }
}

constexpr int kStateBlocks = 296;

}  // namespace

void launch_step_update(const SchurStructure& st, const ObsView& obs, const JacView& jv, const double* jac_cam,
                        int cam_frame, NormalEq ne,
                        const double* y_c, int n_frames, int n_points, const double* poses,
                        const double* points, double* delta_c, double* delta_p, double* trial_poses,
                        double* trial_points, double* scalars, double* scratch, int bounded_param,
                        double lower_bound, const PointGroups& pg, cudaStream_t s) {
  frame_step_kernel<<<1, kWideThreads, 0, s>>>(ne, st.tile_pos, y_c, 12 * n_frames, poses, delta_c, trial_poses, scratch,
                                               bounded_param, lower_bound);
  // (RSBA_CUDA_STEP=warp: the warp-per-point kernel, 0.277 against 0.248 ms at C3)
  static const bool use_groups = [] { const char* e = getenv("RSBA_CUDA_STEP"); return !(e && e[0] == 'w'); }();
  if (use_groups && pg.groups && jv.point_major && !jac_cam) {
    // one thread per observation over groups of whole points; tracks longer than a CTA through the warp-per-point kernel
    if (pg.n_groups > 0)
      point_step_group_kernel<<<pg.n_groups, kStepGroupThreads, 0, s>>>(st, pg.groups, jv, ne, delta_c, points, delta_p,
                                                                       trial_points, scratch);
    int nb = pg.n_groups;
    if (pg.n_big > 0) {
      NormalEq nbig = ne;
      nbig.owned_ids = pg.big_ids;
      nbig.n_owned = pg.n_big;
      const int blocks = (pg.n_big + kPointStepWarps - 1) / kPointStepWarps;
      point_step_kernel<<<blocks, kPointStepWarps * 32, 0, s>>>(st, jv, jac_cam, cam_frame, nbig, delta_c, n_points, points,
                                                                delta_p, trial_points, scratch, nb);
      nb += blocks;
    }
    step_final_kernel<<<1, kWideThreads, 0, s>>>(scratch, nb, scalars);
    return;
  }
  const int nb = (ne.n_owned + kPointStepWarps - 1) / kPointStepWarps;
  if (nb > 0)
    point_step_kernel<<<nb, kPointStepWarps * 32, 0, s>>>(st, jv, jac_cam, cam_frame, ne, delta_c, n_points, points, delta_p,
                                                          trial_points, scratch, 0);
  step_final_kernel<<<1, kWideThreads, 0, s>>>(scratch, nb, scalars);
}

void launch_point_norms(NormalEq ne, int n_points, const double* points, double* out_xx, double* out_gmax,
                        double* scratch, cudaStream_t s) {
  point_norms_kernel<<<kStateBlocks, kRedThreads, 0, s>>>(ne, n_points, points, scratch);
  point_norms_final_kernel<<<1, kRedThreads, 0, s>>>(scratch, kStateBlocks, out_xx, out_gmax);
}

void launch_camera_norms(NormalEq ne, int n_frames, const double* poses, double* scalars, cudaStream_t s) {
  camera_norms_kernel<<<1, kWideThreads, 0, s>>>(ne, n_frames, poses, scalars);
}

}  // namespace rsba
