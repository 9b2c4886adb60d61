// Camera-only motion priors: RsConstVeloPrior / RsConstAccelerationPrior
// (video_bundler_rs_inter.h:55-108, 113-173) as CeresHandler::Add wires them between frame k and
// frame k-1 (CeresHandler.h:148-186), with the interFrameRatio block held constant (the reference
// fixes it whenever it differs from 1, CeresHandler.h:178-180).  With a constant ratio both
// functors are LINEAR in the four pose blocks: each 6-residual half is
//     r = sigma * (c0 pose0 + c1 end0 + c2 pose1 + c3 end1),  sigma = scale * (.01,.01,.01,1,1,1)
// (pose0/end0 = first/last control pose of frame k, pose1/end1 of frame k-1), so residuals, the
// gradient and J^T J have closed forms with 6x6 DIAGONAL blocks.  tests/test_priors_cpu.py pins the
// coefficients against the reference functors compiled verbatim under Jet autodiff.
// The optional Huber loss applies to the 12-residual block exactly as to the reprojection blocks.
#include "lm.cuh"

namespace rsba {
namespace {

__device__ __forceinline__ double sigma_of(double scale, int j) { return j < 3 ? 0.01 * scale : scale; }

// residuals of prior i at `poses`; returns |r|^2
__device__ __forceinline__ double prior_residuals(const PriorView& pv, int i, const double* __restrict__ poses,
                                                  double r[12]) {
  const double* fk = poses + 12L * pv.frame[i];
  const double* fp = poses + 12L * pv.prev[i];
  const double* c = pv.coef + 8L * i;
  double s = 0.0;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const double v = sigma_of(pv.scale[i], j) *
                       (c[4 * h] * fk[j] + c[4 * h + 1] * fk[6 + j] + c[4 * h + 2] * fp[j] + c[4 * h + 3] * fp[6 + j]);
      r[6 * h + j] = v;
      s += v * v;
    }
  return s;
}

// One CTA.  cost_out[0] = sum_i rho(|r_i|^2)  (the caller's reduction applies the 1/2); with
// r_out != NULL also stores the (loss-corrected) residuals and the squared correction weight.
__global__ void __launch_bounds__(256)
prior_eval_kernel(PriorView pv, const double* __restrict__ poses, double huber, double* __restrict__ cost_out,
                  double* __restrict__ r_out, double* __restrict__ w2_out) {
  __shared__ double sh[256];
  double cost = 0.0;
  for (int i = threadIdx.x; i < pv.n; i += blockDim.x) {
    double r[12];
    double s = prior_residuals(pv, i, poses, r);
    double w = 1.0;
    if (huber > 0.0 && s > huber * huber) {   // Ceres Corrector for HuberLoss, as in K1
      const double sr = sqrt(s);
      w = sqrt(huber / sr);
      s = 2.0 * huber * sr - huber * huber;
    }
    cost += s;
    if (r_out) {
#pragma unroll
      for (int k = 0; k < 12; ++k) r_out[12L * i + k] = w * r[k];
      w2_out[i] = w * w;
    }
  }
  sh[threadIdx.x] = cost;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {       // fixed-order tree: deterministic
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) cost_out[0] = sh[0];
}

// Thread (frame f, component j): adds the priors' J^T J and J^T r to the frame's diagonal block and
// gradient -- as the current frame of prior a = cur_of[f] and as the previous frame of prior
// b = prev_of[f] -- and writes the coupling diagonals of prior a:  Bx[f][(X,Y)][j], X in {pose0,end0}
// of frame f, Y in {pose1,end1} of its previous frame.
__global__ void __launch_bounds__(192)
prior_blocks_kernel(PriorView pv, NormalEq ne, int n_frames) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_frames * 6) return;
  const int f = t / 6, j = t % 6;
  double* B = ne.B + 144L * f;
  double* Bx = pv.Bx + 24L * f;
  const int a = pv.cur_of[f], b = pv.prev_of[f];
  double bxx[4] = {0.0, 0.0, 0.0, 0.0};
  if (a >= 0) {
    const double* c = pv.coef + 8L * a;
    const double sg = sigma_of(pv.scale[a], j), w2 = pv.w2[a];
    const double q = sg * sg * w2;
    const double* r = pv.r + 12L * a;
    // blocks X = pose0 (0), end0 (1) of this frame
    const double d00 = q * (c[0] * c[0] + c[4] * c[4]), d01 = q * (c[0] * c[1] + c[4] * c[5]);
    const double d11 = q * (c[1] * c[1] + c[5] * c[5]);
    B[j * 12 + j] += d00;
    B[j * 12 + 6 + j] += d01;
    B[(6 + j) * 12 + j] += d01;
    B[(6 + j) * 12 + 6 + j] += d11;
    ne.diagB[12L * f + j] += d00;
    ne.diagB[12L * f + 6 + j] += d11;
    // gradient: the stored residuals already carry w, the Jacobian carries another w
    const double sw = sg * sqrt(w2);
    ne.gc[12L * f + j] += sw * (c[0] * r[j] + c[4] * r[6 + j]);
    ne.gc[12L * f + 6 + j] += sw * (c[1] * r[j] + c[5] * r[6 + j]);
    bxx[0] = q * (c[0] * c[2] + c[4] * c[6]);   // pose0 x pose1
    bxx[1] = q * (c[0] * c[3] + c[4] * c[7]);   // pose0 x end1
    bxx[2] = q * (c[1] * c[2] + c[5] * c[6]);   // end0  x pose1
    bxx[3] = q * (c[1] * c[3] + c[5] * c[7]);   // end0  x end1
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) Bx[6 * k + j] = bxx[k];
  if (b >= 0) {
    const double* c = pv.coef + 8L * b;
    const double sg = sigma_of(pv.scale[b], j), w2 = pv.w2[b];
    const double q = sg * sg * w2;
    const double* r = pv.r + 12L * b;
    // blocks Y = pose1 (2), end1 (3): this frame is the prior's previous frame
    const double d00 = q * (c[2] * c[2] + c[6] * c[6]), d01 = q * (c[2] * c[3] + c[6] * c[7]);
    const double d11 = q * (c[3] * c[3] + c[7] * c[7]);
    B[j * 12 + j] += d00;
    B[j * 12 + 6 + j] += d01;
    B[(6 + j) * 12 + j] += d01;
    B[(6 + j) * 12 + 6 + j] += d11;
    ne.diagB[12L * f + j] += d00;
    ne.diagB[12L * f + 6 + j] += d11;
    const double sw = sg * sqrt(w2);
    ne.gc[12L * f + j] += sw * (c[2] * r[j] + c[6] * r[6 + j]);
    ne.gc[12L * f + 6 + j] += sw * (c[3] * r[j] + c[7] * r[6 + j]);
  }
}

}  // namespace

void launch_prior_eval(const PriorView& pv, const double* poses, double huber, double* cost_out, bool store,
                       cudaStream_t s) {
  prior_eval_kernel<<<1, 256, 0, s>>>(pv, poses, huber, cost_out, store ? pv.r : nullptr, store ? pv.w2 : nullptr);
}

void launch_prior_blocks(const PriorView& pv, NormalEq ne, int n_frames, cudaStream_t s) {
  if (n_frames > 0) prior_blocks_kernel<<<(n_frames * 6 + 191) / 192, 192, 0, s>>>(pv, ne, n_frames);
}

}  // namespace rsba
