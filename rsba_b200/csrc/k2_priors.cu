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
//
// FREE ratio (the reference's default interFrameRatio == 1 leaves the scalar block variable, with a lower
// bound: CeresHandler.h:156-180): the ratio is parameter 9 of the pseudo-frame behind the real frames.  The
// coefficients c(ratio) are then refreshed on the device at every linearisation, each prior also carries the
// column d r / d ratio = sigma * (c'(ratio) . blocks), and the priors add  B[9][9], g[9]  of the pseudo-frame
// and the dense couplings  B_f,ratio  (column 9 of ne.Bcam) -- the same border the intrinsics block uses.
#include "lm.cuh"

namespace rsba {
namespace {

__device__ __forceinline__ double sigma_of(double scale, int j) { return j < 3 ? 0.01 * scale : scale; }

constexpr double kEps = 2.220446049250313e-16;   // _EPS of the reference (mat/core.h)

// c(ratio) and dc/dratio in the block order pose0, end0, pose1, end1 (video_bundler_rs_inter.h:69-84
// velocity, :127-148 acceleration); returns the functor's bool (:92, :157)
__device__ __forceinline__ bool prior_coef(int kind, double r, double c[8], double dc[8]) {
  const double h = kind == 1 ? 1.0 : 0.5;
  c[0] = h; c[1] = 0.0; c[2] = h * r; c[3] = -h * (1.0 + r);
  dc[0] = 0.0; dc[1] = 0.0; dc[2] = h; dc[3] = -h;
  if (kind != 1 || r > kEps) {
    const double ir = 1.0 / r;
    c[4] = -h * (1.0 + ir); c[5] = h; c[6] = 0.0; c[7] = h * ir;
    dc[4] = h * ir * ir; dc[5] = 0.0; dc[6] = 0.0; dc[7] = -h * ir * ir;
  } else {
    c[4] = -1.0; c[5] = 1.0; c[6] = 1.0; c[7] = -1.0;
    dc[4] = dc[5] = dc[6] = dc[7] = 0.0;
  }
  return kind == 1 ? (r >= 0.0) : (r >= kEps);
}

// residuals of prior i at `poses`; returns |r|^2
__device__ __forceinline__ double prior_residuals(const PriorView& pv, int i, const double* __restrict__ poses,
                                                  const double* __restrict__ c, double r[12]) {
  const double* fk = poses + 12L * pv.frame[i];
  const double* fp = poses + 12L * pv.prev[i];
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
                  double* __restrict__ r_out, double* __restrict__ w2_out, int* __restrict__ invalid) {
  __shared__ double sh[256];
  double cost = 0.0;
  const bool free_ratio = pv.ratio_off >= 0;
  for (int i = threadIdx.x; i < pv.n; i += blockDim.x) {
    double r[12], c[8], dc[8];
    if (free_ratio) {
      if (!prior_coef(pv.kind[i], poses[pv.ratio_off], c, dc)) atomicAdd(invalid, 1);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) c[k] = pv.coef[8L * i + k];
    }
    double s = prior_residuals(pv, i, poses, c, r);
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
      if (free_ratio) {
        double jr[12];
        prior_residuals(pv, i, poses, dc, jr);   // linear in the coefficients: same form with c' for c
#pragma unroll
        for (int k = 0; k < 12; ++k) pv.jr[12L * i + k] = w * jr[k];
#pragma unroll
        for (int k = 0; k < 8; ++k) pv.coef_dev[8L * i + k] = c[k];
      }
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
  double bq0 = 0.0, bq1 = 0.0;   // free ratio: coupling of the frame's (pose0[j], end0[j]) with the ratio
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
    if (pv.ratio_off >= 0) {
      const double* jq = pv.jr + 12L * a;
      bq0 += sw * (c[0] * jq[j] + c[4] * jq[6 + j]);
      bq1 += sw * (c[1] * jq[j] + c[5] * jq[6 + j]);
    }
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
    if (pv.ratio_off >= 0) {
      const double* jq = pv.jr + 12L * b;
      bq0 += sw * (c[2] * jq[j] + c[6] * jq[6 + j]);
      bq1 += sw * (c[3] * jq[j] + c[7] * jq[6 + j]);
    }
  }
  if (pv.ratio_off >= 0) {
    ne.Bcam[144L * f + j * 12 + 9] = bq0;
    ne.Bcam[144L * f + (6 + j) * 12 + 9] = bq1;
  }
}

// One CTA: the ratio's own block of the pseudo-frame (index n_frames):  B[9][9] = sum jr.jr,  g[9] = sum jr.r
// (fixed-order tree).  frame_reduce_kernel has zeroed the pseudo-frame's row 9 before.
__global__ void __launch_bounds__(256)
prior_ratio_kernel(PriorView pv, NormalEq ne, int n_frames) {
  __shared__ double sh[2][256];
  double qq = 0.0, qr = 0.0;
  for (int i = threadIdx.x; i < pv.n; i += blockDim.x)
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const double q = pv.jr[12L * i + k];
      qq += q * q;
      qr += q * pv.r[12L * i + k];
    }
  sh[0][threadIdx.x] = qq;
  sh[1][threadIdx.x] = qr;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const long F = n_frames;
    ne.B[F * 144 + 9 * 12 + 9] = sh[0][0];
    ne.diagB[F * 12 + 9] = sh[0][0];
    ne.gc[F * 12 + 9] = sh[1][0];
  }
}

}  // namespace

void launch_prior_eval(const PriorView& pv, const double* poses, double huber, double* cost_out, bool store,
                       int* invalid_count, cudaStream_t s) {
  prior_eval_kernel<<<1, 256, 0, s>>>(pv, poses, huber, cost_out, store ? pv.r : nullptr, store ? pv.w2 : nullptr,
                                      invalid_count);
}

void launch_prior_blocks(const PriorView& pv, NormalEq ne, int n_frames, cudaStream_t s) {
  if (n_frames > 0) prior_blocks_kernel<<<(n_frames * 6 + 191) / 192, 192, 0, s>>>(pv, ne, n_frames);
  if (pv.ratio_off >= 0) prior_ratio_kernel<<<1, 256, 0, s>>>(pv, ne, n_frames);
}

}  // namespace rsba
