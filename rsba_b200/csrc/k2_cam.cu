// Uncalibrated variant: the session's intrinsics are a shared 9-wide parameter block of
// every residual block -- RsBundleAdjustment::CreateWithCam <2; 9, 6, 6, 3> (VideoSfmBaRs.h:38-49,68-80),
// wired at CeresHandler.h:256-264 when opt.model.calibrated is false.  Here that block is a PSEUDO-FRAME
// behind the real frames (parameters 0..8 of "frame" F, 9..11 constant), so the reduced camera system,
// its scaling / damping, the factorisation and the step logic treat it like any other frame; what is
// specific to it lives in this file:
//   cam_blocks / cam_reduce : B_II = sum Jcam^T Jcam, g_I = sum Jcam^T r, w_I = sum Jcam^T Jx t_p (one block
//                             for the whole problem) and the frame couplings B_fI = sum_{i in f} Jc^T Jcam
//   phi_cam                 : the pseudo-frame rows of every point's Schur panel,
//                             sum_{i in p} Jcam_i^T (Jx_i s_p) L^-T  (9 x 3)
// The Schur border (S_II, S_fI) then falls out of the ordinary sub-tile-pair SYRK.
#include "lm.cuh"

namespace rsba {
namespace {

constexpr int kCamChunk = 128;
constexpr int kCamPartial = 108 + 81 + 9 + 9;   // B_fI | B_II | g_I | w_I

// One CTA per frame chunk.  The chunk's rows are staged in shared memory, then thread k < 207 owns one
// output and runs over the observations in a fixed order.
__global__ void __launch_bounds__(256)
cam_blocks_kernel(SchurStructure st, ObsView obs, JacView jv, const double* __restrict__ jac_cam,
                  const double* __restrict__ res, NormalEq ne, double* __restrict__ partials) {
  __shared__ double sC[kCamChunk * 2 * 12];   // camera rows of J
  __shared__ double sI[kCamChunk * 2 * 9];    // intrinsics rows
  __shared__ double sR[kCamChunk * 2], sQ[kCamChunk * 2];
  const int c = blockIdx.x;
  const long beg = st.chunk_beg[c];
  const int cnt = st.chunk_cnt[c];
  for (int o = threadIdx.x; o < cnt; o += blockDim.x) {
    double Jf[kJacDoubles];
    load_full_jacobian(jv, beg + o, Jf);
#pragma unroll
    for (int k = 0; k < 24; ++k) {
      const int row = k / 12, col = k % 12;
      sC[o * 24 + k] = Jf[col < 6 ? row * 6 + col : 12 + row * 6 + (col - 6)];
    }
  }
  for (int t = threadIdx.x; t < cnt * 18; t += blockDim.x) sI[t] = jac_cam[beg * 18 + t];
  for (int t = threadIdx.x; t < cnt * 2; t += blockDim.x) {
    const int o = t >> 1, row = t & 1;
    sR[t] = res[beg * 2 + t];
    const int p = obs.point[beg + o];
    const double* jx = jv.rec + (beg + o) * kJacCompact + 3 * row;   // point part: first six doubles of the record
    sQ[t] = jx[0] * ne.tp[3L * p] + jx[1] * ne.tp[3L * p + 1] + jx[2] * ne.tp[3L * p + 2];
  }
  __syncthreads();
  const int k = threadIdx.x;
  if (k >= kCamPartial) return;
  double s = 0.0;
  if (k < 108) {                      // B_fI[a][b]
    const int a = k / 9, b = k % 9;
    for (int r = 0; r < 2 * cnt; ++r) s += sC[r * 12 + a] * sI[r * 9 + b];
  } else if (k < 189) {               // B_II[a][b]
    const int a = (k - 108) / 9, b = (k - 108) % 9;
    for (int r = 0; r < 2 * cnt; ++r) s += sI[r * 9 + a] * sI[r * 9 + b];
  } else if (k < 198) {               // g_I
    const int a = k - 189;
    for (int r = 0; r < 2 * cnt; ++r) s += sI[r * 9 + a] * sR[r];
  } else {                            // w_I
    const int a = k - 198;
    for (int r = 0; r < 2 * cnt; ++r) s += sI[r * 9 + a] * sQ[r];
  }
  partials[(long)c * kCamPartial + k] = s;
}

// grid = real frames: B_fI of the frame, and the frame's share of the problem-wide sums -> scratch[f][99]
__global__ void __launch_bounds__(256)
cam_reduce_frames_kernel(SchurStructure st, NormalEq ne, const double* __restrict__ partials,
                         double* __restrict__ scratch) {
  const int f = blockIdx.x, k = threadIdx.x;
  if (k >= kCamPartial) return;
  double s = 0.0;
  for (int c = st.frame_chunk_ptr[f]; c < st.frame_chunk_ptr[f + 1]; ++c) s += partials[(long)c * kCamPartial + k];
  if (k < 108) ne.Bcam[(long)f * 144 + (k / 9) * 12 + k % 9] = s;
  else scratch[(long)f * 99 + (k - 108)] = s;
}

// one CTA: sums the frames' shares in frame order and fills the pseudo-frame's blocks
__global__ void __launch_bounds__(128)
cam_reduce_final_kernel(NormalEq ne, const double* __restrict__ scratch, int n_frames) {
  const int k = threadIdx.x;
  if (k >= 99) return;
  double s = 0.0;
  for (int f = 0; f < n_frames; ++f) s += scratch[(long)f * 99 + k];
  const long F = n_frames;
  if (k < 81) {
    const int a = k / 9, b = k % 9;
    ne.B[F * 144 + a * 12 + b] = s;
    if (a == b) ne.diagB[F * 12 + a] = s;
  } else if (k < 90) {
    ne.gc[F * 12 + (k - 81)] = s;
  } else {
    ne.wf[F * 12 + (k - 90)] = s;
  }
}

// One warp per point: pseudo-frame rows of the point's Schur panel.
__global__ void __launch_bounds__(256)
phi_cam_kernel(SchurStructure st, JacView jv, const double* __restrict__ jac_cam, NormalEq ne,
               int n_points, int slot) {
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (p >= n_points) return;
  const int inc = st.cam_inc[p];
  if (inc < 0) return;               // constant point: not eliminated
  const double* Mi = ne.Minv + 6L * p;
  const double m00 = Mi[0], m10 = Mi[1], m11 = Mi[2], m20 = Mi[3], m21 = Mi[4], m22 = Mi[5];
  const double sp0 = ne.scale_p[3L * p], sp1 = ne.scale_p[3L * p + 1], sp2 = ne.scale_p[3L * p + 2];
  double F[27];
#pragma unroll
  for (int k = 0; k < 27; ++k) F[k] = 0.0;
  for (int e = st.pt_ptr[p] + lane; e < st.pt_ptr[p + 1]; e += 32) {
    const long i = st.pt_obs[e];
    const double* jx = jv.rec + i * kJacCompact;
    const double a0 = jx[0] * sp0, a1 = jx[1] * sp1, a2 = jx[2] * sp2;
    const double b0 = jx[3] * sp0, b1 = jx[4] * sp1, b2 = jx[5] * sp2;
    const double xa[3] = {a0 * m00, a0 * m10 + a1 * m11, a0 * m20 + a1 * m21 + a2 * m22};
    const double xb[3] = {b0 * m00, b0 * m10 + b1 * m11, b0 * m20 + b1 * m21 + b2 * m22};
    const double* jc = jac_cam + i * 18;
#pragma unroll
    for (int a = 0; a < 9; ++a)
#pragma unroll
      for (int k = 0; k < 3; ++k) F[3 * a + k] += jc[a] * xa[k] + jc[9 + a] * xb[k];
  }
#pragma unroll
  for (int k = 0; k < 27; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) F[k] += __shfl_xor_sync(0xffffffffu, F[k], o);
  // rows 12*slot .. 12*slot+8 of the panel (k-major); rows 9..11 of the pseudo-frame stay zero
  double* dst = ne.Phi + (long)inc * kPanelDoubles + slot * kFrameParams;
  if (lane == 0) {   // (one store per lane compiled to a jump table of 27 divergent paths)
#pragma unroll
    for (int k = 0; k < 27; ++k) dst[(k % 3) * kPanelLd + k / 3] = F[k];
  }
}

}  // namespace

void launch_cam_blocks(const SchurStructure& st, const ObsView& obs, const JacView& jv, const double* jac_cam,
                       const double* res, NormalEq ne, int n_frames, double* partials, double* scratch,
                       cudaStream_t s) {
  if (st.n_chunks > 0) cam_blocks_kernel<<<st.n_chunks, 256, 0, s>>>(st, obs, jv, jac_cam, res, ne, partials);
  if (n_frames > 0) cam_reduce_frames_kernel<<<n_frames, 256, 0, s>>>(st, ne, partials, scratch);
  cam_reduce_final_kernel<<<1, 128, 0, s>>>(ne, scratch, n_frames);
}

void launch_phi_cam(const SchurStructure& st, const JacView& jv, const double* jac_cam, NormalEq ne, int n_points,
                    int n_frames, cudaStream_t s) {
  if (n_points > 0)
    phi_cam_kernel<<<(n_points + 7) / 8, 256, 0, s>>>(st, jv, jac_cam, ne, n_points, n_frames % kSubFrames);
}

}  // namespace rsba
