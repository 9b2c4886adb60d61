// K2 -- block normal equations and the Schur complement on the point blocks, FP64.
//
// Supersedes Ceres' SchurEliminator::Eliminate (third-party, reached through ceres::Solve with
// linear_solver_type = SPARSE_SCHUR, CeresHandler.h:403,419).  Input is the per-observation
// Jacobian written by K1 ([N][30] = J_pose0[2][6] | J_pose1[2][6] | J_point[2][3]) and the
// residuals.  A "frame" is the 12-wide camera block pose0|pose1; every observation couples
// one frame with one point, so
//   B_f  = sum_{i in f} Jc_i^T Jc_i        12x12   (frame_blocks)
//   C_p  = sum_{i in p} Jx_i^T Jx_i         3x3    (point_blocks)
//   S_ab = [a==b](s B s + D^2) - s_a ( sum_{p seen by a and b} Jc_i^T Jx_i Cinv_p Jx_j^T Jc_j ) s_b
// with Cinv_p = s_p (s_p C_p s_p + D_p^2)^-1 s_p inverted in registers (3x3 Cholesky).  This file
// holds the diagonal blocks and the point inverses; the Schur product itself is k2_schur.cu.
// Jacobi scaling s and the LM diagonal D^2 = clamp(diag)/radius follow
// LevenbergMarquardtStrategy::ComputeStep / TrustRegionMinimizer (Ceres 1.9.0).
//
// Every output element is produced by exactly one thread in a fixed summation order: the
// normal equations are bit-reproducible from run to run (no floating-point atomics).
#include "lm.cuh"

#include <algorithm>

namespace rsba {
namespace {

constexpr int kChunk = 128;        // observations per frame chunk (host builds the chunk table)
constexpr int kPartial = 168;      // 144 (B) + 12 (gc) + 12 (wf)

// ---------------------------------------------------------------- points: C_p, g_p
// One warp per point; lanes gather the observations in parallel, fixed-order butterfly sum.  The point part of
// the Jacobian is the first 48 bytes (two sectors) of the observation's compact record.
constexpr int kPointBlockWarps = 8;

__global__ void __launch_bounds__(kPointBlockWarps * 32)
point_blocks_kernel(const int* __restrict__ pt_ptr, const int* __restrict__ pt_obs,
                    const double* __restrict__ rec, const double* __restrict__ res, NormalEq ne,
                    double* __restrict__ C, double* __restrict__ gp) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * kPointBlockWarps + (threadIdx.x >> 5);
  if (k >= ne.n_owned) return;
  const int p = owned_point(ne, k);
  double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // c0..c5, g0..g2
  const int beg = pt_ptr[p], end = pt_ptr[p + 1];
  for (int e = beg + lane; e < end; e += 32) {
    const long i = pt_obs[e];
    const double2* jx = reinterpret_cast<const double2*>(rec + i * kJacCompact);
    const double2 q0 = jx[0], q1 = jx[1], q2 = jx[2];   // row0: q0.x q0.y q1.x ; row1: q1.y q2.x q2.y
    const double2 r = reinterpret_cast<const double2*>(res)[i];
    const double a0 = q0.x, a1 = q0.y, a2 = q1.x, b0 = q1.y, b1 = q2.x, b2 = q2.y;
    v[0] += a0 * a0 + b0 * b0;
    v[1] += a0 * a1 + b0 * b1;
    v[2] += a0 * a2 + b0 * b2;
    v[3] += a1 * a1 + b1 * b1;
    v[4] += a1 * a2 + b1 * b2;
    v[5] += a2 * a2 + b2 * b2;
    v[6] += a0 * r.x + b0 * r.y;
    v[7] += a1 * r.x + b1 * r.y;
    v[8] += a2 * r.x + b2 * r.y;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  if (lane == 0) {   // every lane holds the full sums ("lane k stores element k" compiles to a jump table of divergent paths)
#pragma unroll
    for (int k = 0; k < 3; ++k) reinterpret_cast<double2*>(C + 6L * p)[k] = make_double2(v[2 * k], v[2 * k + 1]);
#pragma unroll
    for (int k = 0; k < 3; ++k) gp[3L * p + k] = v[6 + k];
  }
}

// tau and frame of every observation in point-major order (coalesced for the warp-per-point kernels)
__global__ void point_major_obs_kernel(const int* __restrict__ pt_obs, const int* __restrict__ frame,
                                       const double* __restrict__ tau, long n, double* __restrict__ pt_tau,
                                       int* __restrict__ pt_frame) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const long i = pt_obs[e];
  if (tau) pt_tau[e] = tau[i];
  if (pt_frame) pt_frame[e] = frame[i];
}

__global__ void point_scale_kernel(NormalEq ne, const double* __restrict__ C,
                                   const unsigned char* __restrict__ point_const, int enabled,
                                   double* __restrict__ scale_p) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ne.n_owned) return;
  const int p = owned_point(ne, t);
  const bool cst = point_const[p] != 0;
  const double* Cp = C + 6L * p;
  const double d[3] = {Cp[0], Cp[3], Cp[5]};
#pragma unroll
  for (int k = 0; k < 3; ++k) scale_p[3L * p + k] = (cst || !enabled) ? 1.0 : 1.0 / (1.0 + sqrt(d[k]));
}

__global__ void frame_scale_kernel(int n_frames, const double* __restrict__ diagB,
                                   const unsigned short* __restrict__ pose_mask, int enabled,
                                   double* __restrict__ scale_c) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_frames * kFrameParams) return;
  const int f = t / kFrameParams, k = t % kFrameParams;
  const bool cst = (pose_mask[f] >> k) & 1;
  scale_c[t] = (cst || !enabled) ? 1.0 : 1.0 / (1.0 + sqrt(diagB[t]));
}

// ---------------------------------------------------------------- points: damped inverse
__global__ void __launch_bounds__(128)
point_invert_kernel(NormalEq ne, LmOptionsDev o) {
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (t0 >= ne.n_owned) return;
  const int p = owned_point(ne, t0);
  double* Ci = ne.Cinv + 6L * p;
  double* t = ne.tp + 3L * p;
  double* d2 = ne.d2_p + 3L * p;
  if (ne.point_const[p]) {
#pragma unroll
    for (int k = 0; k < 6; ++k) Ci[k] = 0.0;
    t[0] = t[1] = t[2] = 0.0;
    d2[0] = d2[1] = d2[2] = 1.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) ne.Minv[6L * p + k] = 0.0;
#pragma unroll
    for (int k = 0; k < kPointRec; ++k) ne.prec[(long)kPointRec * p + k] = 0.0;
    return;
  }
  const double* Cp = ne.C + 6L * p;
  const double s0 = ne.scale_p[3L * p], s1 = ne.scale_p[3L * p + 1], s2 = ne.scale_p[3L * p + 2];
  double c00 = s0 * Cp[0] * s0, c10 = s1 * Cp[1] * s0, c20 = s2 * Cp[2] * s0;
  double c11 = s1 * Cp[3] * s1, c21 = s2 * Cp[4] * s1, c22 = s2 * Cp[5] * s2;
  const double e0 = fmin(fmax(c00, o.min_diag), o.max_diag) / o.radius;
  const double e1 = fmin(fmax(c11, o.min_diag), o.max_diag) / o.radius;
  const double e2 = fmin(fmax(c22, o.min_diag), o.max_diag) / o.radius;
  d2[0] = e0; d2[1] = e1; d2[2] = e2;
  c00 += e0; c11 += e1; c22 += e2;
  // 3x3 Cholesky C' = L L^T, M = L^-1, C'^-1 = M^T M  (all in registers)
  const double l00 = sqrt(c00);
  const double m00 = 1.0 / l00;
  const double l10 = c10 * m00, l20 = c20 * m00;
  const double l11 = sqrt(c11 - l10 * l10);
  const double m11 = 1.0 / l11;
  const double l21 = (c21 - l20 * l10) * m11;
  const double l22 = sqrt(c22 - l20 * l20 - l21 * l21);
  const double m22 = 1.0 / l22;
  const double m10 = -l10 * m00 * m11;
  const double m21 = -l21 * m11 * m22;
  const double m20 = -(l20 * m00 + l21 * m10) * m22;
  {
    double* Mi = ne.Minv + 6L * p;
    Mi[0] = m00; Mi[1] = m10; Mi[2] = m11; Mi[3] = m20; Mi[4] = m21; Mi[5] = m22;
  }
  // inverse (scaled space), then fold the point scaling back in: Cinv'' = s Cinv' s
  const double i00 = (m00 * m00 + m10 * m10 + m20 * m20) * s0 * s0;
  const double i10 = (m10 * m11 + m20 * m21) * s1 * s0;
  const double i20 = (m20 * m22) * s2 * s0;
  const double i11 = (m11 * m11 + m21 * m21) * s1 * s1;
  const double i21 = (m21 * m22) * s2 * s1;
  const double i22 = (m22 * m22) * s2 * s2;
  Ci[0] = i00; Ci[1] = i10; Ci[2] = i20; Ci[3] = i11; Ci[4] = i21; Ci[5] = i22;
  const double* g = ne.gp + 3L * p;
  t[0] = i00 * g[0] + i10 * g[1] + i20 * g[2];
  t[1] = i10 * g[0] + i11 * g[1] + i21 * g[2];
  t[2] = i20 * g[0] + i21 * g[1] + i22 * g[2];
  // what frame_blocks gathers per observation, as one 96-byte record
  double2* pr = reinterpret_cast<double2*>(ne.prec + (long)kPointRec * p);
  pr[0] = make_double2(s0 * m00, s0 * m10);
  pr[1] = make_double2(s1 * m11, s0 * m20);
  pr[2] = make_double2(s1 * m21, s2 * m22);
  pr[3] = make_double2(t[0], t[1]);
  pr[4] = make_double2(t[2], 0.0);
  pr[5] = make_double2(0.0, 0.0);
}

// ---------------------------------------------------------------- frames: B_f, g_c, w_f (+ Schur panels)
// One CTA per frame chunk (<= 128 observations of one frame), one thread per observation.  The thread reads its
// compact Jacobian record, tau, the residual and the 72 bytes its point contributes (W = s_p L^-T and t_p, one
// gathered 96-byte record), and
//   * writes the observation's Schur panel rows  F = Jc^T (Jx W)  (12 x 3) straight into the point's
//     (sub-tile, point) panel (k2_schur.cu describes the layout; zero rows of unobserved frames were written once,
//     at allocation).  With Jc = [wr0 jr | -(1-tau) jx | wr1 jr | -tau jx] the twelve rows are two 3x3 products
//     G_rot = jr^T Xm, G_c = jx^T Xm scaled by the four weights;
//   * stores its two rows of  Z = [Jc | r | q]  (q = Jx t_p) transposed in shared memory.
// The chunk's sums  B_f = Jc^T Jc, g_c = Jc^T r, w_f = Jc^T q  are then the lower tiles of the Gram matrix
// Z^T Z (16 x 16, K = 256): three DMMA.8x8x4 per four rows, 16 k-steps per warp, summed over the four warps in
// a fixed order -- no per-thread accumulators, no shared-memory operand traffic beyond two fragment loads per
// k-step (the scalar version was bound by its barriers and 34 M bank conflicts per launch).
constexpr int kFrameThreads = kChunk;              // 128
constexpr int kZld = 2 * kChunk + 4;               // row stride of Z^T: == 4 (mod 16) -> conflict-free DMMA fragment loads
constexpr size_t kFrameSmem = (size_t)(16 * kZld + 4 * 3 * 64) * sizeof(double);

__device__ __forceinline__ void dmma_gram(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(kFrameThreads, 4)
frame_blocks_kernel(SchurStructure st, ObsView obs, JacView jv, const double* __restrict__ res, NormalEq ne) {
  extern __shared__ __align__(16) double fsm[];
  double* Zt = fsm;                        // [16][kZld]: column-major Z, row index = (residual row) * 128 + observation
  double* red = fsm + 16 * kZld;           // [4 warps][3 tiles][64]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.x;
  const long beg = st.chunk_beg[c];
  const int cnt = st.chunk_cnt[c];
  double z0[14], z1[14];
#pragma unroll
  for (int k = 0; k < 14; ++k) z0[k] = z1[k] = 0.0;
  if (tid < cnt) {
    const long i = beg + tid;
    const double2* rp = reinterpret_cast<const double2*>(jv.rec + i * kJacCompact);
    const double2 a01 = __ldg(rp), a2b0 = __ldg(rp + 1), b12 = __ldg(rp + 2);      // jx rows
    const double2 c01 = __ldg(rp + 3), c2d0 = __ldg(rp + 4), d12 = __ldg(rp + 5);  // jr rows
    const double tau = jv.tau[i];
    const double2 r = reinterpret_cast<const double2*>(res)[i];
    const int p = obs.point[i];
    const int off = st.obs_phi_off[i];
    const double2* pp = reinterpret_cast<const double2*>(ne.prec + (long)kPointRec * p);
    const double2 w0 = __ldg(pp), w1 = __ldg(pp + 1), w2 = __ldg(pp + 2), t01 = __ldg(pp + 3), t2_ = __ldg(pp + 4);
    const double jx0[3] = {a01.x, a01.y, a2b0.x}, jx1[3] = {a2b0.y, b12.x, b12.y};
    const double jr0[3] = {c01.x, c01.y, c2d0.x}, jr1[3] = {c2d0.y, d12.x, d12.y};
    const double th0 = 1.0 - tau, th1 = tau;
    const double wr0 = jv.rot_interp ? th0 : 1.0, wr1 = jv.rot_interp ? th1 : 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      z0[k] = wr0 * jr0[k];      z1[k] = wr0 * jr1[k];
      z0[3 + k] = -th0 * jx0[k]; z1[3 + k] = -th0 * jx1[k];
      z0[6 + k] = wr1 * jr0[k];  z1[6 + k] = wr1 * jr1[k];
      z0[9 + k] = -th1 * jx0[k]; z1[9 + k] = -th1 * jx1[k];
    }
    z0[12] = r.x; z1[12] = r.y;
    z0[13] = jx0[0] * t01.x + jx0[1] * t01.y + jx0[2] * t2_.x;     // q = Jx t_p
    z1[13] = jx1[0] * t01.x + jx1[1] * t01.y + jx1[2] * t2_.x;
    if (off >= 0) {
      // Xm = Jx W:  xa[k] = sum_{c <= k} jx0[c] W[k][c]
      const double xa[3] = {jx0[0] * w0.x, jx0[0] * w0.y + jx0[1] * w1.x, jx0[0] * w1.y + jx0[1] * w2.x + jx0[2] * w2.y};
      const double xb[3] = {jx1[0] * w0.x, jx1[0] * w0.y + jx1[1] * w1.x, jx1[0] * w1.y + jx1[1] * w2.x + jx1[2] * w2.y};
      const unsigned mask = ne.pose_mask[st.chunk_frame[c]];
      const double wgt[4] = {wr0, -th0, wr1, -th1};
      double* dst = ne.Phi + off;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double gr[3], gc3[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          gr[j] = jr0[j] * xa[k] + jr1[j] * xb[k];
          gc3[j] = jx0[j] * xa[k] + jx1[j] * xb[k];
        }
        double f[12];
#pragma unroll
        for (int a = 0; a < 12; ++a) {
          const double g = ((a / 3) & 1) ? gc3[a % 3] : gr[a % 3];
          f[a] = ((mask >> a) & 1) ? 0.0 : wgt[a / 3] * g;
        }
        // the 96-byte row leaves as three full 32-byte sectors (256-bit stores, sm_100): 16-byte stores would
        // double the number of L2 write transactions
#pragma unroll
        for (int a = 0; a < 12; a += 4)
          asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + k * kPanelLd + a), "d"(f[a]),
                       "d"(f[a + 1]), "d"(f[a + 2]), "d"(f[a + 3])
                       : "memory");
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 14; ++k) {
    Zt[k * kZld + tid] = z0[k];
    Zt[k * kZld + kChunk + tid] = z1[k];
  }
  Zt[14 * kZld + tid] = 0.0; Zt[14 * kZld + kChunk + tid] = 0.0;
  Zt[15 * kZld + tid] = 0.0; Zt[15 * kZld + kChunk + tid] = 0.0;
  __syncthreads();
  // Gram tiles: D00 = rows/cols 0..7, D10 = rows 8..15 x cols 0..7, D11 = rows/cols 8..15
  {
    const int fr = lane >> 2, fc = lane & 3;
    double d00[2] = {0.0, 0.0}, d10[2] = {0.0, 0.0}, d11[2] = {0.0, 0.0};
    const double* zlo = Zt + fr * kZld + warp * 64 + fc;
    const double* zhi = zlo + 8 * kZld;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
      const double lo = zlo[4 * ks], hi = zhi[4 * ks];
      dmma_gram(d00[0], d00[1], lo, lo);
      dmma_gram(d10[0], d10[1], hi, lo);
      dmma_gram(d11[0], d11[1], hi, hi);
    }
    double2* rw = reinterpret_cast<double2*>(red + warp * 192);
    rw[lane] = make_double2(d00[0], d00[1]);            // element (fr, 2 fc + {0, 1}) of the tile
    rw[32 + lane] = make_double2(d10[0], d10[1]);
    rw[64 + lane] = make_double2(d11[0], d11[1]);
  }
  __syncthreads();
  // fixed-order sum over the warps -> chunk partial  B (144, row-major, both triangles) | g_c (12) | w_f (12)
  double* out = ne.partials + (long)c * kPartial;
  for (int k = tid; k < kPartial; k += kFrameThreads) {
    int r, cc;
    if (k < 144) { r = k / 12; cc = k % 12; if (cc > r) { const int t = r; r = cc; cc = t; } }
    else if (k < 156) { r = 12; cc = k - 144; }
    else { r = 13; cc = k - 156; }
    const int tile = (r < 8) ? 0 : ((cc < 8) ? 1 : 2);
    const int e = tile * 64 + (r & 7) * 8 + (cc & 7);
    double sum = 0.0;
#pragma unroll
    for (int w = 0; w < 4; ++w) sum += red[w * 192 + e];
    out[k] = sum;
  }
}

__global__ void __launch_bounds__(192)
frame_reduce_kernel(SchurStructure st, NormalEq ne, int n_frames) {
  const int f = blockIdx.x;
  const int k = threadIdx.x;
  if (f >= n_frames || k >= kPartial) return;
  double s = 0.0;
  for (int c = st.frame_chunk_ptr[f]; c < st.frame_chunk_ptr[f + 1]; ++c) s += ne.partials[(long)c * kPartial + k];
  if (k < 144) {
    ne.B[(long)f * 144 + k] = s;
    if (k % 13 == 0) ne.diagB[(long)f * 12 + k / 13] = s;
  } else if (k < 156) ne.gc[(long)f * 12 + (k - 144)] = s;
  else ne.wf[(long)f * 12 + (k - 156)] = s;
}

}  // namespace

void launch_point_blocks(const SchurStructure& st, const JacView& jv, const double* res, NormalEq ne, cudaStream_t s) {
  if (ne.n_owned <= 0) return;
  point_blocks_kernel<<<(ne.n_owned + kPointBlockWarps - 1) / kPointBlockWarps, kPointBlockWarps * 32, 0, s>>>(
      st.pt_ptr, st.pt_obs, jv.rec, res, ne, ne.C, ne.gp);
}

void launch_point_major_obs(const SchurStructure& st, const ObsView& obs, const double* tau, long n, double* pt_tau,
                            int* pt_frame, cudaStream_t s) {
  if (n > 0) point_major_obs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(st.pt_obs, obs.frame, tau, n, pt_tau, pt_frame);
}

void launch_frame_blocks(const SchurStructure& st, const ObsView& obs, const JacView& jv, const double* res,
                         int n_frames, NormalEq ne, cudaStream_t s) {
  static bool seen[64] = {};
  if (first_use_on_device(seen))
    cudaFuncSetAttribute(frame_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFrameSmem);
  if (st.n_chunks > 0) frame_blocks_kernel<<<st.n_chunks, kFrameThreads, kFrameSmem, s>>>(st, obs, jv, res, ne);
  if (n_frames > 0) frame_reduce_kernel<<<n_frames, 192, 0, s>>>(st, ne, n_frames);
}

void launch_frame_reduce(const SchurStructure& st, NormalEq ne, int n_frames, cudaStream_t s) {
  if (n_frames > 0) frame_reduce_kernel<<<n_frames, 192, 0, s>>>(st, ne, n_frames);
}

void launch_jacobi_scale(int n_frames, bool points, NormalEq ne, bool enabled, cudaStream_t s) {
  // split in two by the caller's ordering needs: points first (n_frames == 0), frames later
  if (points && ne.n_owned > 0)
    point_scale_kernel<<<(ne.n_owned + 255) / 256, 256, 0, s>>>(ne, ne.C, ne.point_const, enabled ? 1 : 0, ne.scale_p);
  if (n_frames > 0)
    frame_scale_kernel<<<(n_frames * kFrameParams + 255) / 256, 256, 0, s>>>(n_frames, ne.diagB, ne.pose_mask,
                                                                             enabled ? 1 : 0, ne.scale_c);
}

void launch_point_invert(NormalEq ne, LmOptionsDev o, cudaStream_t s) {
  if (ne.n_owned <= 0) return;
  point_invert_kernel<<<(ne.n_owned + 127) / 128, 128, 0, s>>>(ne, o);
}

}  // namespace rsba
