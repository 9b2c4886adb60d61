// TEST INFRASTRUCTURE / CPU BASELINE -- never linked into the product library.
//
// Multi-threaded C++ restatement of ONE linear solve of the Levenberg-Marquardt subproblem as
// Ceres Solver 1.9.0 performs it for rsba (ceres::Solve with SPARSE_SCHUR,
// CeresHandler.h:394-419, VideoSfMHandler.cc:579-583).  PARITY UNPINNED: Ceres is a
// third-party dependency that is absent from /root/reference (pinned by .travis.yml:33);
// the algorithm is restated from its published implementation:
//   levenberg_marquardt_strategy.cc  D^2 = clamp(diag(J'^T J'), min, max) / radius
//   trust_region_minimizer.cc        Jacobi scaling 1/(1+|col|), model_cost_change = -m.(r+m/2)
//   schur_eliminator_impl.h          per point:  C^-1, S -= E C^-1 E^T, rhs -= E C^-1 g_p,
//                                    BackSubstitute
// Ceres parallelises Eliminate over chunks of points with a mutex per S cell; here every
// thread owns whole block rows of S (all points seen by its frames), which needs no locks and
// is at least as fast.  The reduced camera system is written in LAPACK lower band storage;
// the caller factorises it with LAPACK dpbsv (scipy.linalg.solveh_banded) -- the stand-in for
// CHOLMOD's sparse Cholesky, which on a video-like (banded) scene does the same work.
//
// Used by bench.py --impl reference (the CPU LM iteration timed beside the GPU one) and by
// tests/ as a second, independent checker of the GPU step next to oracle/lm_oracle.py.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct CpuLm {
  long N = 0;
  int F = 0, P = 0;
  std::vector<int> fr, pt;
  std::vector<long> pt_ptr, fr_ptr;
  std::vector<long> pt_obs, fr_obs;       // CSR lists (observation ids)
  std::vector<unsigned short> pose_mask;
  std::vector<unsigned char> point_const;
  int max_span = 0;                       // max |frame_i - frame_j| over pairs sharing a point
  int kd = 0;                             // sub-diagonals of the band: 12*(max_span+1)-1
  // numeric state of the last build
  std::vector<double> scale_c, scale_p, d2_c, d2_p;  // [12F], [3P]
  std::vector<double> M;                  // [P][9] lower-triangular L^-1 of the damped point block
  std::vector<double> gp_s;               // [3P] scaled point gradient
  std::vector<double> Fm;                 // [N][36]  F_i = Jc'^T Jx' M^T  (12x3 row-major)
  std::vector<double> gc_s;               // [12F]
  bool have_scale = false;
};

inline bool cconst(const CpuLm& s, int f, int k) { return (s.pose_mask[f] >> k) & 1; }

// camera Jacobian entry (row, col 0..11) of a 30-double record
inline double jc(const double* rec, int row, int col) {
  return col < 6 ? rec[row * 6 + col] : rec[12 + row * 6 + (col - 6)];
}

}  // namespace

extern "C" {

void* cpu_lm_create(long N, int F, int P, const int* obs_frame, const int* obs_point,
                    const unsigned short* pose_mask, const unsigned char* point_const) {
  CpuLm* s = new CpuLm;
  s->N = N; s->F = F; s->P = P;
  s->fr.assign(obs_frame, obs_frame + N);
  s->pt.assign(obs_point, obs_point + N);
  s->pose_mask.assign(F, 0);
  s->point_const.assign(P, 0);
  if (pose_mask) for (int f = 0; f < F; ++f) s->pose_mask[f] = pose_mask[f] & 0xFFF;
  if (point_const) for (int p = 0; p < P; ++p) s->point_const[p] = point_const[p] ? 1 : 0;
  s->pt_ptr.assign(P + 1, 0);
  s->fr_ptr.assign(F + 1, 0);
  for (long i = 0; i < N; ++i) { s->pt_ptr[s->pt[i] + 1]++; s->fr_ptr[s->fr[i] + 1]++; }
  for (int p = 0; p < P; ++p) s->pt_ptr[p + 1] += s->pt_ptr[p];
  for (int f = 0; f < F; ++f) s->fr_ptr[f + 1] += s->fr_ptr[f];
  s->pt_obs.resize(N); s->fr_obs.resize(N);
  {
    std::vector<long> cp(s->pt_ptr.begin(), s->pt_ptr.end() - 1), cf(s->fr_ptr.begin(), s->fr_ptr.end() - 1);
    for (long i = 0; i < N; ++i) { s->pt_obs[cp[s->pt[i]]++] = i; s->fr_obs[cf[s->fr[i]]++] = i; }
  }
  int span = 0;
  for (int p = 0; p < P; ++p) {
    if (s->point_const[p]) continue;
    int lo = F, hi = -1;
    for (long e = s->pt_ptr[p]; e < s->pt_ptr[p + 1]; ++e) {
      lo = std::min(lo, s->fr[s->pt_obs[e]]);
      hi = std::max(hi, s->fr[s->pt_obs[e]]);
    }
    if (hi >= lo) span = std::max(span, hi - lo);
  }
  s->max_span = span;
  s->kd = std::min(12 * (span + 1) - 1, std::max(12 * F - 1, 0));
  s->scale_c.assign(12L * F, 1.0); s->scale_p.assign(3L * P, 1.0);
  s->d2_c.assign(12L * F, 1.0); s->d2_p.assign(3L * P, 1.0);
  s->M.assign(9L * P, 0.0); s->gp_s.assign(3L * P, 0.0); s->gc_s.assign(12L * F, 0.0);
  s->Fm.assign(36L * N, 0.0);
  return s;
}

void cpu_lm_destroy(void* h) { delete (CpuLm*)h; }
int cpu_lm_band(void* h) { return ((CpuLm*)h)->kd; }

// Builds the reduced camera system  S y_c = rhs  (scaled space; constant parameters are identity
// rows) from the Jacobian J[N][30] and residuals r[N][2].
//   ab  [(kd+1)][n]  LAPACK lower band storage: ab[(i-j)*n + j] = S[i][j], n = 12F
//   rhs [n]
//   gmax             max |g_k| over free parameters (unscaled gradient)
// compute_scale != 0 recomputes the Jacobi scaling (first iteration only, as Ceres does).
void cpu_lm_build(void* h, const double* J, const double* r, double radius, double min_diag,
                  double max_diag, int jacobi_scaling, int compute_scale, double* ab, double* rhs,
                  double* gmax_out, int nthreads) {
  CpuLm& s = *(CpuLm*)h;
  const int F = s.F, P = s.P;
  const long n = 12L * F;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  std::vector<double> C(6L * P), gp(3L * P), B(144L * F), gc(12L * F);
  // ---- unscaled diagonal blocks and gradient
#pragma omp parallel for schedule(dynamic, 256)
  for (int p = 0; p < P; ++p) {
    double c[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
    for (long e = s.pt_ptr[p]; e < s.pt_ptr[p + 1]; ++e) {
      const long i = s.pt_obs[e];
      const double* x = J + 30 * i + 24;
      const double r0 = r[2 * i], r1 = r[2 * i + 1];
      c[0] += x[0] * x[0] + x[3] * x[3]; c[1] += x[0] * x[1] + x[3] * x[4]; c[2] += x[0] * x[2] + x[3] * x[5];
      c[3] += x[1] * x[1] + x[4] * x[4]; c[4] += x[1] * x[2] + x[4] * x[5]; c[5] += x[2] * x[2] + x[5] * x[5];
      g[0] += x[0] * r0 + x[3] * r1; g[1] += x[1] * r0 + x[4] * r1; g[2] += x[2] * r0 + x[5] * r1;
    }
    memcpy(&C[6L * p], c, sizeof(c));
    memcpy(&gp[3L * p], g, sizeof(g));
  }
#pragma omp parallel for schedule(dynamic, 1)
  for (int f = 0; f < F; ++f) {
    double b[144], g[12];
    memset(b, 0, sizeof(b)); memset(g, 0, sizeof(g));
    for (long e = s.fr_ptr[f]; e < s.fr_ptr[f + 1]; ++e) {
      const long i = s.fr_obs[e];
      const double* rec = J + 30 * i;
      for (int row = 0; row < 2; ++row) {
        double v[12];
        for (int k = 0; k < 12; ++k) v[k] = jc(rec, row, k);
        const double rr = r[2 * i + row];
        for (int a = 0; a < 12; ++a) {
          g[a] += v[a] * rr;
          for (int c2 = 0; c2 < 12; ++c2) b[a * 12 + c2] += v[a] * v[c2];
        }
      }
    }
    memcpy(&B[144L * f], b, sizeof(b));
    memcpy(&gc[12L * f], g, sizeof(g));
  }
  // ---- Jacobi scaling (kept from the first iteration)
  if (compute_scale || !s.have_scale) {
    for (int f = 0; f < F; ++f)
      for (int k = 0; k < 12; ++k)
        s.scale_c[12L * f + k] = (!jacobi_scaling || cconst(s, f, k)) ? 1.0 : 1.0 / (1.0 + std::sqrt(B[144L * f + 13 * k]));
    const int dix[3] = {0, 3, 5};
    for (int p = 0; p < P; ++p)
      for (int k = 0; k < 3; ++k)
        s.scale_p[3L * p + k] = (!jacobi_scaling || s.point_const[p]) ? 1.0 : 1.0 / (1.0 + std::sqrt(C[6L * p + dix[k]]));
    s.have_scale = true;
  }
  double gmax = 0.0;
  for (int f = 0; f < F; ++f)
    for (int k = 0; k < 12; ++k)
      if (!cconst(s, f, k)) gmax = std::max(gmax, std::fabs(gc[12L * f + k]));
  for (int p = 0; p < P; ++p)
    if (!s.point_const[p])
      for (int k = 0; k < 3; ++k) gmax = std::max(gmax, std::fabs(gp[3L * p + k]));
  if (gmax_out) *gmax_out = gmax;

  // ---- damped point blocks: C' = s C s + D^2 = L L^T,  M = L^-1
#pragma omp parallel for schedule(static)
  for (int p = 0; p < P; ++p) {
    double* Mp = &s.M[9L * p];
    if (s.point_const[p]) {
      memset(Mp, 0, 9 * sizeof(double));
      s.d2_p[3L * p] = s.d2_p[3L * p + 1] = s.d2_p[3L * p + 2] = 1.0;
      s.gp_s[3L * p] = s.gp_s[3L * p + 1] = s.gp_s[3L * p + 2] = 0.0;
      continue;
    }
    const double* c = &C[6L * p];
    const double* sp = &s.scale_p[3L * p];
    double c00 = sp[0] * c[0] * sp[0], c10 = sp[1] * c[1] * sp[0], c20 = sp[2] * c[2] * sp[0];
    double c11 = sp[1] * c[3] * sp[1], c21 = sp[2] * c[4] * sp[1], c22 = sp[2] * c[5] * sp[2];
    const double e0 = std::min(std::max(c00, min_diag), max_diag) / radius;
    const double e1 = std::min(std::max(c11, min_diag), max_diag) / radius;
    const double e2 = std::min(std::max(c22, min_diag), max_diag) / radius;
    s.d2_p[3L * p] = e0; s.d2_p[3L * p + 1] = e1; s.d2_p[3L * p + 2] = e2;
    c00 += e0; c11 += e1; c22 += e2;
    const double l00 = std::sqrt(c00), l10 = c10 / l00, l20 = c20 / l00;
    const double l11 = std::sqrt(c11 - l10 * l10), l21 = (c21 - l20 * l10) / l11;
    const double l22 = std::sqrt(c22 - l20 * l20 - l21 * l21);
    const double m00 = 1 / l00, m11 = 1 / l11, m22 = 1 / l22;
    const double m10 = -l10 * m00 * m11, m21 = -l21 * m11 * m22, m20 = -(l20 * m00 + l21 * m10) * m22;
    Mp[0] = m00; Mp[1] = 0; Mp[2] = 0; Mp[3] = m10; Mp[4] = m11; Mp[5] = 0; Mp[6] = m20; Mp[7] = m21; Mp[8] = m22;
    for (int k = 0; k < 3; ++k) s.gp_s[3L * p + k] = sp[k] * gp[3L * p + k];
  }
  // ---- F_i = Jc'^T Jx' M^T
#pragma omp parallel for schedule(static)
  for (long i = 0; i < s.N; ++i) {
    const int f = s.fr[i], p = s.pt[i];
    double* Fi = &s.Fm[36L * i];
    if (s.point_const[p]) { memset(Fi, 0, 36 * sizeof(double)); continue; }
    const double* rec = J + 30 * i;
    const double* sp = &s.scale_p[3L * p];
    const double* Mp = &s.M[9L * p];
    double xm[2][3];   // Jx' M^T
    for (int row = 0; row < 2; ++row) {
      const double x0 = rec[24 + 3 * row] * sp[0], x1 = rec[25 + 3 * row] * sp[1], x2 = rec[26 + 3 * row] * sp[2];
      for (int k = 0; k < 3; ++k) xm[row][k] = x0 * Mp[3 * k] + x1 * Mp[3 * k + 1] + x2 * Mp[3 * k + 2];
    }
    for (int a = 0; a < 12; ++a) {
      const double sa = cconst(s, f, a) ? 0.0 : s.scale_c[12L * f + a];
      const double j0 = jc(rec, 0, a) * sa, j1 = jc(rec, 1, a) * sa;
      for (int k = 0; k < 3; ++k) Fi[3 * a + k] = j0 * xm[0][k] + j1 * xm[1][k];
    }
  }
  // ---- S (band) and rhs: thread owns block rows  (frame b = row block, frames a <= b = column blocks)
  const int kd = s.kd;
  memset(ab, 0, sizeof(double) * (size_t)(kd + 1) * n);
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < F; ++b) {
    // diagonal block: s B s + D^2, constants -> identity
    for (int rI = 0; rI < 12; ++rI) {
      const long gi = 12L * b + rI;
      const bool rc = cconst(s, b, rI);
      for (int cI = 0; cI <= rI; ++cI) {
        const long gj = 12L * b + cI;
        const bool cc = cconst(s, b, cI);
        double v;
        if (rc || cc) v = (rI == cI) ? 1.0 : 0.0;
        else {
          v = s.scale_c[gi] * B[144L * b + 12 * rI + cI] * s.scale_c[gj];
          if (rI == cI) {
            const double d2 = std::min(std::max(v, min_diag), max_diag) / radius;
            s.d2_c[gi] = d2;
            v += d2;
          }
        }
        ab[(gi - gj) * n + gj] = v;
      }
      if (rc) s.d2_c[gi] = 1.0;
      s.gc_s[gi] = rc ? 0.0 : s.scale_c[gi] * gc[gi];
    }
    double rb[12];
    for (int k = 0; k < 12; ++k) rb[k] = s.gc_s[12L * b + k];
    // Schur terms of block row b, accumulated in a dense local strip [a = b-span .. b][12][12]
    const int a_lo = std::max(0, b - s.max_span);
    std::vector<double> strip((size_t)(b - a_lo + 1) * 144, 0.0);
    for (long e = s.fr_ptr[b]; e < s.fr_ptr[b + 1]; ++e) {
      const long j = s.fr_obs[e];
      const int p = s.pt[j];
      if (s.point_const[p]) continue;
      const double* Fj = &s.Fm[36L * j];
      // rhs -= F_j (M g_p')
      const double* Mp = &s.M[9L * p];
      const double* g = &s.gp_s[3L * p];
      const double t0 = Mp[0] * g[0], t1 = Mp[3] * g[0] + Mp[4] * g[1], t2 = Mp[6] * g[0] + Mp[7] * g[1] + Mp[8] * g[2];
      for (int k = 0; k < 12; ++k) rb[k] -= Fj[3 * k] * t0 + Fj[3 * k + 1] * t1 + Fj[3 * k + 2] * t2;
      for (long e2 = s.pt_ptr[p]; e2 < s.pt_ptr[p + 1]; ++e2) {
        const long i = s.pt_obs[e2];
        const int a = s.fr[i];
        if (a > b) continue;
        const double* Fi = &s.Fm[36L * i];
        double* blk = &strip[(size_t)(a - a_lo) * 144];      // block (b, a) += F_j F_i^T
        for (int rI = 0; rI < 12; ++rI) {
          const double f0 = Fj[3 * rI], f1 = Fj[3 * rI + 1], f2 = Fj[3 * rI + 2];
          for (int cI = 0; cI < 12; ++cI)
            blk[12 * rI + cI] += f0 * Fi[3 * cI] + f1 * Fi[3 * cI + 1] + f2 * Fi[3 * cI + 2];
        }
      }
    }
    for (int a = a_lo; a <= b; ++a) {
      const double* blk = &strip[(size_t)(a - a_lo) * 144];
      for (int rI = 0; rI < 12; ++rI) {
        const long gi = 12L * b + rI;
        const int cmax = (a == b) ? rI : 11;        // diagonal block: lower triangle only
        for (int cI = 0; cI <= cmax; ++cI) {
          const long gj = 12L * a + cI;
          ab[(gi - gj) * n + gj] -= blk[12 * rI + cI];
        }
      }
    }
    for (int k = 0; k < 12; ++k) rhs[12L * b + k] = cconst(s, b, k) ? 0.0 : rb[k];
  }
}

// Back-substitution and step.  y_c [12F] solves S y_c = rhs.  Outputs the unscaled step and
//   scalars[0] = model_cost_change = -m.(r + m/2), m = J delta   (trust_region_minimizer.cc)
//   scalars[1] = |delta|
void cpu_lm_backsub(void* h, const double* J, const double* r, const double* y_c, double* delta_c,
                    double* delta_p, double* scalars, int nthreads) {
  CpuLm& s = *(CpuLm*)h;
  const int F = s.F, P = s.P;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  for (long k = 0; k < 12L * F; ++k) delta_c[k] = cconst(s, (int)(k / 12), (int)(k % 12)) ? 0.0 : -s.scale_c[k] * y_c[k];
#pragma omp parallel for schedule(dynamic, 256)
  for (int p = 0; p < P; ++p) {
    double* d = delta_p + 3L * p;
    if (s.point_const[p]) { d[0] = d[1] = d[2] = 0.0; continue; }
    // y_p = Cinv' (g_p' - sum_i Jx'^T Jc' y_c) = M^T ( M g_p' - sum_i F_i^T y_c[frame_i] )
    const double* Mp = &s.M[9L * p];
    const double* g = &s.gp_s[3L * p];
    double t[3] = {Mp[0] * g[0], Mp[3] * g[0] + Mp[4] * g[1], Mp[6] * g[0] + Mp[7] * g[1] + Mp[8] * g[2]};
    for (long e = s.pt_ptr[p]; e < s.pt_ptr[p + 1]; ++e) {
      const long i = s.pt_obs[e];
      const double* Fi = &s.Fm[36L * i];
      const double* y = y_c + 12L * s.fr[i];
      for (int a = 0; a < 12; ++a) { t[0] -= Fi[3 * a] * y[a]; t[1] -= Fi[3 * a + 1] * y[a]; t[2] -= Fi[3 * a + 2] * y[a]; }
    }
    const double y0 = Mp[0] * t[0] + Mp[3] * t[1] + Mp[6] * t[2];
    const double y1 = Mp[4] * t[1] + Mp[7] * t[2];
    const double y2 = Mp[8] * t[2];
    d[0] = -s.scale_p[3L * p] * y0; d[1] = -s.scale_p[3L * p + 1] * y1; d[2] = -s.scale_p[3L * p + 2] * y2;
  }
  double mcc = 0.0, nn = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : mcc)
  for (long i = 0; i < s.N; ++i) {
    const double* rec = J + 30 * i;
    const double* dc = delta_c + 12L * s.fr[i];
    const double* dp = delta_p + 3L * s.pt[i];
    for (int row = 0; row < 2; ++row) {
      double m = rec[24 + 3 * row] * dp[0] + rec[25 + 3 * row] * dp[1] + rec[26 + 3 * row] * dp[2];
      for (int k = 0; k < 12; ++k) m += jc(rec, row, k) * dc[k];
      mcc -= m * (r[2 * i + row] + 0.5 * m);
    }
  }
  for (long k = 0; k < 12L * F; ++k) nn += delta_c[k] * delta_c[k];
  for (long k = 0; k < 3L * P; ++k) nn += delta_p[k] * delta_p[k];
  scalars[0] = mcc;
  scalars[1] = std::sqrt(nn);
}

}  // extern "C"
