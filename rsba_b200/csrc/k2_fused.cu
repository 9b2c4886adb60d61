// The solver's linearisation as two passes that never materialise the Jacobian in observation order:
//
//   point_pass  (one warp per point, one lane per observation)   -- "KP"
//       evaluates the rolling-shutter functor + its Jacobian for every observation of the point (the K1 arithmetic
//       of rsba_reproj_math.h, Huber correction included), forms C_p = sum Jx^T Jx and g_p = sum Jx^T r by a
//       fixed-order butterfly, applies the Jacobi scaling and the LM diagonal, inverts the 3x3 block by Cholesky in
//       registers, and -- still holding jr, jx and tau of its observation -- every lane writes
//         * its 12 Schur panel rows  F = Jc^T (Jx s_p) L^-T  into the point's (sub-tile, point) panel; adjacent lanes
//           are adjacent frames, so the four frames of a sub-tile leave as one contiguous 384-byte run per k-row;
//         * its compact Jacobian record and tau in POINT-major order (coalesced), for the back-substitution (K4).
//   frame_pass  (one CTA per frame chunk, one thread per observation)   -- "KF"
//       re-evaluates the functor from the 24-byte observation (coalesced) and the point's (X_p | t_p) record, and
//       reduces  B_f = Jc^T Jc, g_c = Jc^T r, w_f = Jc^T (Jx t_p)  and the cost as the lower tiles of the Gram matrix
//       of Z = [Jc | r | q] on the FP64 tensor path (three DMMA.8x8x4 per four rows).
//
// Against K1 -> point_blocks -> point_invert -> frame_blocks (k1_reproj.cu, k2_normal.cu), which this replaces for
// calibrated scenes, nothing is gathered twice: K1's 0.56 GB of records are neither written in frame order nor
// re-read by two gathering kernels, and the 600 FP64 flops per observation that are computed twice instead cost
// ~0.08 ms of pipe time.  Supersedes, like those kernels, ceres::AutoDiffCostFunction evaluation +
// SchurEliminator::Eliminate's per-chunk work (third-party Ceres 1.9.0, reached through ceres::Solve,
// CeresHandler.h:403,419).  Every sum has a fixed order: results are bit-reproducible.
#include "lm.cuh"

namespace rsba {
namespace {

struct ObsEval {
  double r0, r1;
  double rec[kJacCompact];   // jx row 0 | jx row 1 | jr row 0 | jr row 1   (loss-corrected)
  double tau;
  double cost;               // rho(|r|^2)
  bool ok;
};

// K1's per-observation work (k1_reproj.cu), registers only
__device__ __forceinline__ void eval_observation(const CameraModel& cm, double ox, double oy, const double* pose,
                                                 double X0, double X1, double X2, ObsEval& e) {
  const Proj pr = reproject<true, false, true, true>(cm, ox, oy, pose, X0, X1, X2, e.rec);
  e.r0 = pr.r0; e.r1 = pr.r1; e.tau = pr.tau; e.ok = pr.ok;
  e.cost = pr.r0 * pr.r0 + pr.r1 * pr.r1;
  // ceres::HuberLoss through Ceres' Corrector: see k1_kernel
  if (cm.huber > 0.0 && e.cost > cm.huber * cm.huber) {
    const double sr = sqrt(e.cost);
    const double w = sqrt(cm.huber / sr);
    e.cost = 2.0 * cm.huber * sr - cm.huber * cm.huber;
    e.r0 *= w; e.r1 *= w;
#pragma unroll
    for (int k = 0; k < kJacCompact; ++k) e.rec[k] *= w;
  }
}

__device__ __forceinline__ void load_pose(const double* __restrict__ poses, int f, double* pose) {
  const double2* gp = reinterpret_cast<const double2*>(poses + (long)f * kFrameParams);
#pragma unroll
  for (int k = 0; k < kFrameParams / 2; ++k) {
    const double2 v = __ldg(gp + k);
    pose[2 * k] = v.x;
    pose[2 * k + 1] = v.y;
  }
}

// ---------------------------------------------------------------- point pass
constexpr int kPointPassWarps = 4;
constexpr int kPointPassAhead = 148 * 4 * kPointPassWarps;   // points of one resident wave (4 CTAs per SM)

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// What the point pass needs of an observation, in point-major order and one 32-byte sector: no dependent
// index -> observation gather on the critical path of a warp (it was 13 k of the 17.7 k cycles a point took).
struct __align__(16) PtObsRec {
  double x, y;
  int frame, phi_off, obs, point;
};
static_assert(sizeof(PtObsRec) == 32, "one sector");

__global__ void pack_point_major_kernel(const int* __restrict__ pt_obs, ObsView obs, const int* __restrict__ obs_phi_off,
                                        long n, PtObsRec* __restrict__ out) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int i = pt_obs[e];
  const double2 xy = obs.xy[i];
  PtObsRec r;
  r.x = xy.x; r.y = xy.y; r.frame = obs.frame[i]; r.phi_off = obs_phi_off[i]; r.obs = i; r.point = obs.point[i];
  out[e] = r;
}

__global__ void __launch_bounds__(kPointPassWarps * 32, 4)
point_pass_kernel(const CameraModel cm, SchurStructure st, const PtObsRec* __restrict__ prec_pm, const double* __restrict__ poses,
                  const double* __restrict__ points, NormalEq ne, LmOptionsDev o, int compute_scale, int jacobi,
                  int rot_interp, int write_phi, double* __restrict__ rec_pt, double* __restrict__ tau_pt,
                  double* __restrict__ xt, long n_obs, int n_points) {
  const int lane = threadIdx.x & 31;
  const int kk = blockIdx.x * kPointPassWarps + (threadIdx.x >> 5);
  if (kk >= ne.n_owned) return;
  const int p = owned_point(ne, kk);
  const int beg = st.pt_ptr[p], end = st.pt_ptr[p + 1];
  // The top stall of this kernel is the chain of global round trips at a warp's start (CSR pointers -> records ->
  // pose) at 16 warps per SM.  CTAs run in blockIdx order, about one resident wave at a time, and the records are
  // point-major, so every warp pulls into L2 what the warp one wave behind it will ask for: its records (this
  // track's length as the estimate of the others'), its point, its CSR pointers.
  {
    const long ahead = (long)kPointPassAhead * (end - beg);
    const long e = beg + ahead + 4L * lane;                      // 4 records of 32 bytes per 128-byte line
    if (4 * lane < end - beg && e < n_obs) prefetch_l2(prec_pm + e);
    const int pa = p + kPointPassAhead;
    if (lane == 0 && pa < n_points) {
      prefetch_l2(points + 3L * pa);
      prefetch_l2(st.pt_ptr + pa);
      if (!compute_scale) prefetch_l2(ne.scale_p + 3L * pa);
    }
  }
  const double X0 = points[3L * p], X1 = points[3L * p + 1], X2 = points[3L * p + 2];
  const bool cst = ne.point_const[p] != 0;
  // the fixed Jacobi scaling of later linearisations, requested now (it is consumed behind the butterfly)
  double sp_in[3] = {1.0, 1.0, 1.0};
  if (!compute_scale) { sp_in[0] = ne.scale_p[3L * p]; sp_in[1] = ne.scale_p[3L * p + 1]; sp_in[2] = ne.scale_p[3L * p + 2]; }

  // ---- phase A: C_p, g_p over all observations (one per lane and round; the last round stays in registers)
  ObsEval ev;
  int frame = 0, off = -1;
  unsigned mask = 0;
  double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // c0..c5, g0..g2
  const bool reuse = end - beg <= 32;   // one round: phase B finds the lane's observation still in registers
  for (int e0 = beg; e0 < end; e0 += 32) {
    const int e = e0 + lane;
    if (e < end) {
      const PtObsRec pr = prec_pm[e];
      frame = pr.frame;
      off = pr.phi_off;
      double pose[kFrameParams];
      load_pose(poses, frame, pose);
      mask = ne.pose_mask[frame];
      eval_observation(cm, pr.x, pr.y, pose, X0, X1, X2, ev);
      const double a0 = ev.rec[0], a1 = ev.rec[1], a2 = ev.rec[2], b0 = ev.rec[3], b1 = ev.rec[4], b2 = ev.rec[5];
      v[0] += a0 * a0 + b0 * b0;
      v[1] += a0 * a1 + b0 * b1;
      v[2] += a0 * a2 + b0 * b2;
      v[3] += a1 * a1 + b1 * b1;
      v[4] += a1 * a2 + b1 * b2;
      v[5] += a2 * a2 + b2 * b2;
      v[6] += a0 * ev.r0 + b0 * ev.r1;
      v[7] += a1 * ev.r0 + b1 * ev.r1;
      v[8] += a2 * ev.r0 + b2 * ev.r1;
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], s);

  // ---- the point block: scaling, damping, 3x3 Cholesky inverse (every lane, identically; see point_invert_kernel)
  double s0, s1, s2;
  if (compute_scale) {
    s0 = (cst || !jacobi) ? 1.0 : 1.0 / (1.0 + sqrt(v[0]));
    s1 = (cst || !jacobi) ? 1.0 : 1.0 / (1.0 + sqrt(v[3]));
    s2 = (cst || !jacobi) ? 1.0 : 1.0 / (1.0 + sqrt(v[5]));
  } else {
    s0 = sp_in[0]; s1 = sp_in[1]; s2 = sp_in[2];
  }
  double W[6] = {0, 0, 0, 0, 0, 0}, t[3] = {0, 0, 0};   // W = s_p L^-T entries (frame side), t_p = Cinv g_p
  double ci[6] = {0, 0, 0, 0, 0, 0}, mi[6] = {0, 0, 0, 0, 0, 0}, d2[3] = {1.0, 1.0, 1.0};
  if (!cst) {
    double c00 = s0 * v[0] * s0, c10 = s1 * v[1] * s0, c20 = s2 * v[2] * s0;
    double c11 = s1 * v[3] * s1, c21 = s2 * v[4] * s1, c22 = s2 * v[5] * s2;
    d2[0] = fmin(fmax(c00, o.min_diag), o.max_diag) / o.radius;
    d2[1] = fmin(fmax(c11, o.min_diag), o.max_diag) / o.radius;
    d2[2] = fmin(fmax(c22, o.min_diag), o.max_diag) / o.radius;
    c00 += d2[0]; c11 += d2[1]; c22 += d2[2];
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
    mi[0] = m00; mi[1] = m10; mi[2] = m11; mi[3] = m20; mi[4] = m21; mi[5] = m22;
    ci[0] = (m00 * m00 + m10 * m10 + m20 * m20) * s0 * s0;
    ci[1] = (m10 * m11 + m20 * m21) * s1 * s0;
    ci[2] = (m20 * m22) * s2 * s0;
    ci[3] = (m11 * m11 + m21 * m21) * s1 * s1;
    ci[4] = (m21 * m22) * s2 * s1;
    ci[5] = (m22 * m22) * s2 * s2;
    t[0] = ci[0] * v[6] + ci[1] * v[7] + ci[2] * v[8];
    t[1] = ci[1] * v[6] + ci[3] * v[7] + ci[4] * v[8];
    t[2] = ci[2] * v[6] + ci[4] * v[7] + ci[5] * v[8];
    W[0] = s0 * m00; W[1] = s0 * m10; W[2] = s1 * m11; W[3] = s0 * m20; W[4] = s1 * m21; W[5] = s2 * m22;
  }
  // per-point outputs (what K4, the norms and the dup / intrinsics kernels read): every lane holds all of them and
  // lane 0 stores them.  "lane k stores element k" compiles to jump tables -- ten divergent paths behind indirect
  // branches per point (the same pattern cost K3's panel 800 cycles: profiles/r02_notes.md)
  if (lane == 0) {
    double2* c2 = reinterpret_cast<double2*>(ne.C + 6L * p);
    double2* ci2 = reinterpret_cast<double2*>(ne.Cinv + 6L * p);
    double2* mi2 = reinterpret_cast<double2*>(ne.Minv + 6L * p);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      c2[k] = make_double2(v[2 * k], v[2 * k + 1]);
      ci2[k] = make_double2(ci[2 * k], ci[2 * k + 1]);
      mi2[k] = make_double2(mi[2 * k], mi[2 * k + 1]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      ne.gp[3L * p + k] = v[6 + k];
      ne.tp[3L * p + k] = t[k];
      ne.d2_p[3L * p + k] = d2[k];
    }
    if (compute_scale) { ne.scale_p[3L * p] = s0; ne.scale_p[3L * p + 1] = s1; ne.scale_p[3L * p + 2] = s2; }
    double2* x = reinterpret_cast<double2*>(xt + 6L * p);   // what the frame pass gathers: X_p | t_p
    x[0] = make_double2(X0, X1);
    x[1] = make_double2(X2, t[0]);
    x[2] = make_double2(t[1], t[2]);
  }

  // ---- phase B: panel rows and point-major records
  for (int e0 = beg; e0 < end; e0 += 32) {
    const int e = e0 + lane;
    const bool mine = e < end;
    if (mine && !reuse) {   // a track longer than a warp: evaluate again
      const PtObsRec pr = prec_pm[e];
      frame = pr.frame;
      off = pr.phi_off;
      double pose[kFrameParams];
      load_pose(poses, frame, pose);
      mask = ne.pose_mask[frame];
      eval_observation(cm, pr.x, pr.y, pose, X0, X1, X2, ev);
    }
    if (!mine) continue;
    double2* rp = reinterpret_cast<double2*>(rec_pt + (long)e * kJacCompact);
#pragma unroll
    for (int k = 0; k < kJacCompact / 2; ++k) rp[k] = make_double2(ev.rec[2 * k], ev.rec[2 * k + 1]);
    tau_pt[e] = ev.tau;
    if (off < 0 || !write_phi) continue;
    const double* jx0 = ev.rec, *jx1 = ev.rec + 3, *jr0 = ev.rec + 6, *jr1 = ev.rec + 9;
    const double xa[3] = {jx0[0] * W[0], jx0[0] * W[1] + jx0[1] * W[2], jx0[0] * W[3] + jx0[1] * W[4] + jx0[2] * W[5]};
    const double xb[3] = {jx1[0] * W[0], jx1[0] * W[1] + jx1[1] * W[2], jx1[0] * W[3] + jx1[1] * W[4] + jx1[2] * W[5]};
    const double th0 = 1.0 - ev.tau, th1 = ev.tau;
    const double wgt[4] = {rot_interp ? th0 : 1.0, -th0, rot_interp ? th1 : 0.0, -th1};
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
#pragma unroll
      for (int a = 0; a < 12; a += 4)
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + k * kPanelLd + a), "d"(f[a]),
                     "d"(f[a + 1]), "d"(f[a + 2]), "d"(f[a + 3])
                     : "memory");
    }
  }
}

// ---------------------------------------------------------------- frame pass
constexpr int kChunk = 128;
constexpr int kPartial = 168;                      // 144 (B) + 12 (gc) + 12 (wf): layout of frame_reduce_kernel
constexpr int kZld = 2 * kChunk + 4;               // == 4 (mod 16): conflict-free DMMA fragment loads
constexpr size_t kFramePassSmem = (size_t)(16 * kZld + 4 * 3 * 64) * sizeof(double);

__device__ __forceinline__ void dmma_gram(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(kChunk)
frame_pass_kernel(const CameraModel cm, SchurStructure st, ObsView obs, const double* __restrict__ poses,
                  const double* __restrict__ xt, NormalEq ne, int rot_interp, double* __restrict__ cost_partials,
                  int* __restrict__ invalid_count) {
  extern __shared__ __align__(16) double fsm[];
  double* Zt = fsm;                        // [16][kZld]: Z transposed, row index = (residual row) * 128 + observation
  double* red = fsm + 16 * kZld;           // [4 warps][3 tiles][64]
  __shared__ double s_cost[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.x;
  const long beg = st.chunk_beg[c];
  const int cnt = st.chunk_cnt[c];
  double cost = 0.0;
  bool bad = false;
  double z0[14], z1[14];
#pragma unroll
  for (int k = 0; k < 14; ++k) z0[k] = z1[k] = 0.0;
  if (tid < cnt) {
    const long i = beg + tid;
    const double2 xy = obs.xy[i];
    const int p = obs.point[i];
    double pose[kFrameParams];
    load_pose(poses, st.chunk_frame[c], pose);
    const double2* xp = reinterpret_cast<const double2*>(xt + 6L * p);
    const double2 x01 = __ldg(xp), x2t0 = __ldg(xp + 1), t12 = __ldg(xp + 2);
    ObsEval ev;
    eval_observation(cm, xy.x, xy.y, pose, x01.x, x01.y, x2t0.x, ev);
    cost = ev.cost;
    bad = !ev.ok;
    const double* jx0 = ev.rec, *jx1 = ev.rec + 3, *jr0 = ev.rec + 6, *jr1 = ev.rec + 9;
    const double th0 = 1.0 - ev.tau, th1 = ev.tau;
    const double wr0 = rot_interp ? th0 : 1.0, wr1 = rot_interp ? th1 : 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      z0[k] = wr0 * jr0[k];      z1[k] = wr0 * jr1[k];
      z0[3 + k] = -th0 * jx0[k]; z1[3 + k] = -th0 * jx1[k];
      z0[6 + k] = wr1 * jr0[k];  z1[6 + k] = wr1 * jr1[k];
      z0[9 + k] = -th1 * jx0[k]; z1[9 + k] = -th1 * jx1[k];
    }
    z0[12] = ev.r0; z1[12] = ev.r1;
    z0[13] = jx0[0] * x2t0.y + jx0[1] * t12.x + jx0[2] * t12.y;     // q = Jx t_p
    z1[13] = jx1[0] * x2t0.y + jx1[1] * t12.x + jx1[2] * t12.y;
  }
#pragma unroll
  for (int k = 0; k < 14; ++k) {
    Zt[k * kZld + tid] = z0[k];
    Zt[k * kZld + kChunk + tid] = z1[k];
  }
  Zt[14 * kZld + tid] = 0.0; Zt[14 * kZld + kChunk + tid] = 0.0;
  Zt[15 * kZld + tid] = 0.0; Zt[15 * kZld + kChunk + tid] = 0.0;
  // cost partial of the chunk (fixed order) and the invalid count, as K1 leaves them
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, s);
  const unsigned badmask = __ballot_sync(0xffffffffu, bad);
  if (lane == 0) {
    s_cost[warp] = cost;
    if (badmask) atomicAdd(invalid_count, __popc(badmask));
  }
  __syncthreads();
  if (tid == 0) cost_partials[c] = (s_cost[0] + s_cost[1]) + (s_cost[2] + s_cost[3]);
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
  double* out = ne.partials + (long)c * kPartial;
  for (int k = tid; k < kPartial; k += kChunk) {
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

}  // namespace

size_t point_pass_record_bytes() { return sizeof(PtObsRec); }

void launch_pack_point_major(const SchurStructure& st, const ObsView& obs, long n, void* packed, cudaStream_t s) {
  if (n > 0)
    pack_point_major_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(st.pt_obs, obs, st.obs_phi_off, n,
                                                                        static_cast<PtObsRec*>(packed));
}

void launch_point_pass(const CameraModel& cm, const SchurStructure& st, const void* packed, const double* poses,
                       const double* points, NormalEq ne, LmOptionsDev o, bool compute_scale, bool jacobi,
                       double* rec_pt, double* tau_pt, double* xt, bool write_phi, long n_obs, int n_points,
                       cudaStream_t s) {
  if (ne.n_owned <= 0) return;
  const int rot_interp = (cm.shutter != 0 && cm.interp_rot) ? 1 : 0;
  // Measured beside this form and not kept (profiles/r02_notes.md, code at commit b6103ee): the evaluation in a kernel of
  // its own + this pass reading the records back (0.67-0.70 ms), and one thread per observation over groups of whole
  // points (0.70 ms; that grouping is what the back-substitution in k4_update.cu uses).
  const int grid = (ne.n_owned + kPointPassWarps - 1) / kPointPassWarps;
  point_pass_kernel<<<grid, kPointPassWarps * 32, 0, s>>>(cm, st, static_cast<const PtObsRec*>(packed), poses, points, ne, o,
                                                          compute_scale ? 1 : 0, jacobi ? 1 : 0, rot_interp, write_phi ? 1 : 0,
                                                          rec_pt, tau_pt, xt, n_obs, n_points);
}

void launch_frame_pass(const CameraModel& cm, const SchurStructure& st, const ObsView& obs, const double* poses,
                       const double* xt, NormalEq ne, double* cost_partials, int* invalid_count, cudaStream_t s) {
  static bool seen[64] = {};
  if (first_use_on_device(seen))
    cudaFuncSetAttribute(frame_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFramePassSmem);
  if (st.n_chunks <= 0) return;
  const int rot_interp = (cm.shutter != 0 && cm.interp_rot) ? 1 : 0;
  frame_pass_kernel<<<st.n_chunks, kChunk, kFramePassSmem, s>>>(cm, st, obs, poses, xt, ne, rot_interp, cost_partials,
                                                                  invalid_count);
}

}  // namespace rsba
