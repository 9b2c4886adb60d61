// "Good initial guess" priors: GoodPosePrior <6; 6, 6> (CeresHandler.h:55-73) as
// CeresHandler::Add wires it for every control pose of a frame that carries priorPoses when
// opt.ceres.trustPriorCamRotation / trustPriorCamPosition are set (CeresHandler.h:188-204, loss = nullptr):
//     r = diag(rot, rot, rot, pos, pos, pos) (prior - pose),      functor false iff r[0] >= 1.
// Both blocks are parameter blocks and the reference never calls SetParameterBlockConstant on the prior
// block, so it is a FREE 6-wide block that occurs in this one residual only.  Everything is diagonal:
// per component  J_prior = w, J_pose = -w,  and the prior block is eliminated in closed form exactly
// like a 3-D point seen once (Ceres' Schur ordering may do the same; the LM step does not depend on
// the elimination order):
//     s  = 1 / (1 + w)                      Jacobi scaling of the prior column (fixed: w is constant)
//     D2 = clamp(s^2 w^2, min, max) / radius
//     cinv = s^2 / (s^2 w^2 + D2)           inverse of the damped block in the unscaled space
//     B_qq += w^2 - w^4 cinv,  g_q += -w r,  wf_q += (-w^2) cinv (w r),  diag(B)_q += w^2
//     delta_prior = -cinv (w r - w^2 delta_q)
// A prior block the caller fixed (rsba_cuda_set_block_constant) keeps cinv = 0.
#include "lm.cuh"

namespace rsba {
namespace {

__device__ __forceinline__ void sum2_256(double& a, double& b, double (*sh)[256]) {
  sh[0][threadIdx.x] = a;
  sh[1][threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {       // fixed-order tree: deterministic
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  a = sh[0][0];
  b = sh[1][0];
  __syncthreads();
}

__global__ void __launch_bounds__(256)
pose_prior_eval_kernel(PosePriorView pv, const double* __restrict__ poses, const double* __restrict__ vals,
                       double* __restrict__ cost_out, double* __restrict__ r_out, int* __restrict__ invalid) {
  __shared__ double sh[2][256];
  double cost = 0.0, unused = 0.0;
  for (int t = threadIdx.x; t < 6 * pv.n; t += blockDim.x) {
    const int i = t / 6, j = t % 6;
    const double w = pv.w[2 * i + (j < 3 ? 0 : 1)];
    const double r = w * (vals[t] - poses[6L * pv.slot[i] + j]);   // minus6(pose0 = prior, pose) then the weights
    if (j == 0 && !(r < 1.0)) atomicAdd(invalid, 1);              // "rotation limit" (CeresHandler.h:66)
    cost += r * r;
    if (r_out) r_out[t] = r;
  }
  sum2_256(cost, unused, sh);
  if (threadIdx.x == 0) cost_out[0] = cost;
}

__global__ void __launch_bounds__(192)
pose_prior_blocks_kernel(PosePriorView pv, NormalEq ne, LmOptionsDev o, int jacobi, int add) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 6 * pv.n) return;
  const int i = t / 6, j = t % 6;
  const double w = pv.w[2 * i + (j < 3 ? 0 : 1)], r = pv.r[t];
  double cinv = 0.0, d2s = 1.0;
  if (!pv.constant[i]) {
    const double s = jacobi ? 1.0 / (1.0 + fabs(w)) : 1.0;
    const double d2 = fmin(fmax(s * s * w * w, o.min_diag), o.max_diag) / o.radius;
    cinv = s * s / (s * s * w * w + d2);
    d2s = d2 / (s * s);
  }
  pv.cinv[t] = cinv;
  pv.d2[t] = d2s;
  if (!add) return;
  const long q = 6L * pv.slot[i] + j;            // the pose parameter, index into the [12 F] arrays
  const long f = q / kFrameParams;
  const int k = (int)(q % kFrameParams);
  const double w2 = w * w;
  ne.B[f * 144 + k * 12 + k] += w2 - w2 * w2 * cinv;
  ne.diagB[q] += w2;
  ne.gc[q] += -w * r;
  ne.wf[q] += -w2 * cinv * (w * r);
}

__global__ void __launch_bounds__(256)
pose_prior_step_kernel(PosePriorView pv, const double* __restrict__ delta_c, double* __restrict__ scalars) {
  __shared__ double sh[2][256];
  double gd = 0.0, dd = 0.0, nn = 0.0, unused = 0.0;
  for (int t = threadIdx.x; t < 6 * pv.n; t += blockDim.x) {
    const int i = t / 6, j = t % 6;
    const double w = pv.w[2 * i + (j < 3 ? 0 : 1)];
    const double gp = w * pv.r[t];
    const double d = -pv.cinv[t] * (gp - w * w * delta_c[6L * pv.slot[i] + j]);
    pv.trial[t] = pv.val[t] + d;
    if (!pv.constant[i]) {
      gd += gp * d;
      nn += d * d;
      dd += pv.d2[t] * d * d;     // D^2 (delta / s)^2
    }
  }
  sum2_256(gd, dd, sh);
  sum2_256(nn, unused, sh);
  if (threadIdx.x == 0) { scalars[0] += gd; scalars[1] += dd; scalars[2] += nn; }
}

__global__ void __launch_bounds__(256)
pose_prior_norms_kernel(PosePriorView pv, double* __restrict__ scalars) {
  __shared__ double sh[2][256];
  double xx = 0.0, gm = 0.0;
  for (int t = threadIdx.x; t < 6 * pv.n; t += blockDim.x) {
    const int i = t / 6, j = t % 6;
    if (pv.constant[i]) continue;
    xx += pv.val[t] * pv.val[t];
    gm = fmax(gm, fabs(pv.w[2 * i + (j < 3 ? 0 : 1)] * pv.r[t]));
  }
  sh[0][threadIdx.x] = xx;
  sh[1][threadIdx.x] = gm;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] = fmax(sh[1][threadIdx.x], sh[1][threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { scalars[3] += sh[0][0]; scalars[4] = fmax(scalars[4], sh[1][0]); }
}

}  // namespace

void launch_pose_prior_eval(const PosePriorView& pv, const double* poses, const double* vals, double* cost_out,
                            bool store, int* invalid_count, cudaStream_t s) {
  pose_prior_eval_kernel<<<1, 256, 0, s>>>(pv, poses, vals, cost_out, store ? pv.r : nullptr, invalid_count);
}

void launch_pose_prior_blocks(const PosePriorView& pv, NormalEq ne, LmOptionsDev o, bool jacobi, bool add,
                              cudaStream_t s) {
  if (pv.n > 0) pose_prior_blocks_kernel<<<(6 * pv.n + 191) / 192, 192, 0, s>>>(pv, ne, o, jacobi ? 1 : 0, add ? 1 : 0);
}

void launch_pose_prior_step(const PosePriorView& pv, const double* delta_c, double* scalars, cudaStream_t s) {
  if (pv.n > 0) pose_prior_step_kernel<<<1, 256, 0, s>>>(pv, delta_c, scalars);
}

void launch_pose_prior_norms(const PosePriorView& pv, double* scalars, cudaStream_t s) {
  if (pv.n > 0) pose_prior_norms_kernel<<<1, 256, 0, s>>>(pv, scalars);
}

}  // namespace rsba
