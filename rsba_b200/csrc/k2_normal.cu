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

// offset of columns 3*t .. 3*t+2 (t = 0..3) of camera row `row` inside a 30-double record
__device__ __forceinline__ int jc_off(int t, int row) { return ((t & 2) ? 12 : 0) + row * 6 + (t & 1) * 3; }

// ---------------------------------------------------------------- points: C_p, g_p
// One warp per point; lanes gather the observations in parallel, fixed-order butterfly sum.
constexpr int kPointBlockWarps = 8;

__global__ void __launch_bounds__(kPointBlockWarps * 32)
point_blocks_kernel(const int* __restrict__ pt_ptr, const int* __restrict__ pt_obs,
                    const double* __restrict__ jac, const double* __restrict__ res, NormalEq ne,
                    double* __restrict__ C, double* __restrict__ gp) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * kPointBlockWarps + (threadIdx.x >> 5);
  if (k >= ne.n_owned) return;
  const int p = owned_point(ne, k);
  double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // c0..c5, g0..g2
  const int beg = pt_ptr[p], end = pt_ptr[p + 1];
  for (int e = beg + lane; e < end; e += 32) {
    const long i = pt_obs[e];
    const double2* jx = reinterpret_cast<const double2*>(jac + i * kJacDoubles + 24);
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
#pragma unroll
  for (int k = 0; k < 9; ++k)                        // every lane holds the full sums
    if (lane == k) {
      if (k < 6) C[6L * p + k] = v[k];
      else gp[3L * p + (k - 6)] = v[k];
    }
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
}

// ---------------------------------------------------------------- frames: B_f, g_c, w_f (+ Schur panels)
// Persistent CTAs (two per SM) walk the chunk list (<= 128 observations of one frame each).  A
// chunk's Jacobian records and residuals are contiguous, so each is ONE TMA bulk copy
// (cp.async.bulk, 30 KB + 2 KB) into a double-buffered shared-memory stage: the next chunk streams
// in while the current one is reduced.  Thread (grp, tr, tc) owns the 3x3 tile (tr, tc) of the
// 12x12 block over the observations grp, grp+16, ...; the per-observation pass also writes the
// observation's Schur panel rows.
constexpr int kFrameThreads = 256;
constexpr int kFrameGroups = kFrameThreads / 16;
constexpr size_t kFrameSmem = (size_t)(2 * kChunk * kJacDoubles + 2 * kChunk * 2 + kChunk * 2) * sizeof(double) + 64;

__device__ __forceinline__ unsigned fsmem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kFrameThreads, 2)
frame_blocks_kernel(SchurStructure st, ObsView obs, const double* __restrict__ jac,
                    const double* __restrict__ res, NormalEq ne, int with_wf) {
  extern __shared__ __align__(128) unsigned char fsmem[];
  double* sJbuf = reinterpret_cast<double*>(fsmem);                    // [2][128*30] raw records
  double* sRbuf = sJbuf + 2 * kChunk * kJacDoubles;                    // [2][256]
  double* sQ = sRbuf + 2 * kChunk * 2;                                 // [256] Jx_i t_p
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sQ + kChunk * 2);
  const int tid = threadIdx.x;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fsmem_u32(&bars[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fsmem_u32(&bars[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int c, int buf) {      // one thread
    const long beg = st.chunk_beg[c];
    const unsigned cnt = (unsigned)st.chunk_cnt[c];
    const unsigned bar = fsmem_u32(&bars[buf]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cnt * 256u) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     fsmem_u32(sJbuf + buf * kChunk * kJacDoubles)),
                 "l"(jac + beg * kJacDoubles), "r"(cnt * 240u), "r"(bar)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     fsmem_u32(sRbuf + buf * kChunk * 2)),
                 "l"(res + beg * 2), "r"(cnt * 16u), "r"(bar)
                 : "memory");
  };
  if (tid == 0 && (int)blockIdx.x < st.n_chunks) issue(blockIdx.x, 0);

  const int grp = tid >> 4, tp = tid & 15, tr = tp >> 2, tc = tp & 3;
  int it = 0;
  for (int c = blockIdx.x; c < st.n_chunks; c += gridDim.x, ++it) {
    const int buf = it & 1;
    const long beg = st.chunk_beg[c];
    const int cnt = st.chunk_cnt[c];
    // the other stage was fully consumed in the previous iteration (trailing __syncthreads)
    if (tid == 0 && c + (int)gridDim.x < st.n_chunks) issue(c + gridDim.x, buf ^ 1);
    {
      const unsigned bar = fsmem_u32(&bars[buf]), parity = (unsigned)((it >> 1) & 1);
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "WAIT_%=:\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
          "@p bra DONE_%=;\n"
          "bra WAIT_%=;\n"
          "DONE_%=:\n"
          "}\n" ::"r"(bar), "r"(parity)
          : "memory");
    }
    double* sJ = sJbuf + buf * kChunk * kJacDoubles;
    const double* sR = sRbuf + buf * kChunk * 2;
    // ---- phase 0: issue the per-point gathers of this chunk (point id -> t_p, L^-1, s_p, panel offset);
    // they are consumed after the B / g_c accumulation below, which hides their latency
    int pf_off = -1;
    double pf_t[3] = {0.0, 0.0, 0.0}, pf_m[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, pf_s[3] = {0.0, 0.0, 0.0};
    if (with_wf && tid < cnt) {
      const int p = obs.point[beg + tid];
      pf_off = st.obs_phi_off[beg + tid];
#pragma unroll
      for (int k = 0; k < 3; ++k) { pf_t[k] = ne.tp[3L * p + k]; pf_s[k] = ne.scale_p[3L * p + k]; }
#pragma unroll
      for (int k = 0; k < 6; ++k) pf_m[k] = ne.Minv[6L * p + k];
    }
    // ---- phase 1: B_f and g_c partial sums (independent of the points)
    double acc[9], ga[3], wa[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.0;
    ga[0] = ga[1] = ga[2] = wa[0] = wa[1] = wa[2] = 0.0;
    for (int o = grp; o < cnt; o += kFrameGroups) {
      const double* rec = sJ + o * kJacDoubles;
#pragma unroll
      for (int row = 0; row < 2; ++row) {
        const double* pa = rec + jc_off(tr, row);
        const double* pb = rec + jc_off(tc, row);
        const double a0 = pa[0], a1 = pa[1], a2 = pa[2];
        const double b0 = pb[0], b1 = pb[1], b2 = pb[2];
        acc[0] += a0 * b0; acc[1] += a0 * b1; acc[2] += a0 * b2;
        acc[3] += a1 * b0; acc[4] += a1 * b1; acc[5] += a1 * b2;
        acc[6] += a2 * b0; acc[7] += a2 * b1; acc[8] += a2 * b2;
        if (tc == 0) {
          const double r = sR[2 * o + row];
          ga[0] += a0 * r; ga[1] += a1 * r; ga[2] += a2 * r;
        }
      }
    }
    // ---- phase 2: per observation  q = Jx t_p  and the Schur panel rows  F = Jc^T (Jx s_p) L^-T (12 x 3),
    // written straight into the point's (sub-tile, point) panel while the record is in shared memory
    // (k2_schur.cu describes the layout; zero rows of unobserved frames were written once, at allocation)
    if (tid < cnt) {
      double q0 = 0.0, q1 = 0.0;
      const double* rec = sJ + tid * kJacDoubles;
      if (with_wf) {
        const double* jx = rec + 24;
        q0 = jx[0] * pf_t[0] + jx[1] * pf_t[1] + jx[2] * pf_t[2];
        q1 = jx[3] * pf_t[0] + jx[4] * pf_t[1] + jx[5] * pf_t[2];
        if (pf_off >= 0) {
          const double m00 = pf_m[0], m10 = pf_m[1], m11 = pf_m[2], m20 = pf_m[3], m21 = pf_m[4], m22 = pf_m[5];
          const double a0 = jx[0] * pf_s[0], a1 = jx[1] * pf_s[1], a2 = jx[2] * pf_s[2];
          const double b0 = jx[3] * pf_s[0], b1 = jx[4] * pf_s[1], b2 = jx[5] * pf_s[2];
          const double xa[3] = {a0 * m00, a0 * m10 + a1 * m11, a0 * m20 + a1 * m21 + a2 * m22};
          const double xb[3] = {b0 * m00, b0 * m10 + b1 * m11, b0 * m20 + b1 * m21 + b2 * m22};
          const unsigned mask = ne.pose_mask[st.chunk_frame[c]];
          double* dst = ne.Phi + pf_off;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            double f[12];
#pragma unroll
            for (int a = 0; a < 12; ++a) {
              // camera column a of the record: rows 0/1 at jc offsets
              const int o0 = (a < 6) ? a : 12 + (a - 6);
              f[a] = ((mask >> a) & 1) ? 0.0 : rec[o0] * xa[k] + rec[o0 + 6] * xb[k];
            }
            // the 96-byte row leaves as three full 32-byte sectors (256-bit stores, sm_100): 16-byte
            // stores would double the number of L2 write transactions, which is what bounds this phase
#pragma unroll
            for (int a = 0; a < 12; a += 4)
              asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + k * kPanelLd + a), "d"(f[a]),
                           "d"(f[a + 1]), "d"(f[a + 2]), "d"(f[a + 3])
                           : "memory");
          }
        }
      }
      sQ[2 * tid] = q0;
      sQ[2 * tid + 1] = q1;
    }
    __syncthreads();
    // ---- phase 3: w_f partial sums (need q)
    if (tc == 0) {
      for (int o = grp; o < cnt; o += kFrameGroups) {
        const double* rec = sJ + o * kJacDoubles;
#pragma unroll
        for (int row = 0; row < 2; ++row) {
          const double* pa = rec + jc_off(tr, row);
          const double q = sQ[2 * o + row];
          wa[0] += pa[0] * q; wa[1] += pa[1] * q; wa[2] += pa[2] * q;
        }
      }
    }
    __syncthreads();                 // sJ is dead: reuse it for the cross-group reduction
    double* sAcc = sJ;               // [16][16][15]
    double* my = sAcc + (grp * 16 + tp) * 15;
#pragma unroll
    for (int k = 0; k < 9; ++k) my[k] = acc[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) { my[9 + k] = ga[k]; my[12 + k] = wa[k]; }
    __syncthreads();
    // fixed-order sum over the groups -> chunk partial
    double* out = ne.partials + (long)c * kPartial;
    for (int k = tid; k < kPartial; k += kFrameThreads) {
      int tpos, slot;
      if (k < 144) {
        const int r = k / 12, cc = k % 12;
        tpos = (r / 3) * 4 + (cc / 3);
        slot = (r % 3) * 3 + (cc % 3);
      } else if (k < 156) {
        const int r = k - 144;
        tpos = (r / 3) * 4;
        slot = 9 + r % 3;
      } else {
        const int r = k - 156;
        tpos = (r / 3) * 4;
        slot = 12 + r % 3;
      }
      double sum = 0.0;
#pragma unroll
      for (int g = 0; g < kFrameGroups; ++g) sum += sAcc[(g * 16 + tpos) * 15 + slot];
      out[k] = sum;
    }
    // generic-proxy reads/writes of this stage must be ordered before the async-proxy refill
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
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

void launch_point_blocks(const SchurStructure& st, const ObsView& obs, const double* jac, const double* res,
                         NormalEq ne, cudaStream_t s) {
  if (ne.n_owned <= 0) return;
  point_blocks_kernel<<<(ne.n_owned + kPointBlockWarps - 1) / kPointBlockWarps, kPointBlockWarps * 32, 0, s>>>(
      st.pt_ptr, st.pt_obs, jac, res, ne, ne.C, ne.gp);
}

void launch_frame_blocks(const SchurStructure& st, const ObsView& obs, const double* jac, const double* res,
                         int n_frames, NormalEq ne, bool with_wf, cudaStream_t s) {
  static bool seen[64] = {};
  static int sm_count[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (first_use_on_device(seen)) {
    cudaFuncSetAttribute(frame_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFrameSmem);
    cudaDeviceGetAttribute(&sm_count[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  }
  const int n_sm = sm_count[dev & 63] > 0 ? sm_count[dev & 63] : 148;
  if (st.n_chunks > 0)
    frame_blocks_kernel<<<std::min(st.n_chunks, 2 * n_sm), kFrameThreads, kFrameSmem, s>>>(st, obs, jac, res, ne, with_wf ? 1 : 0);
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
