// K3 -- Cholesky factorisation of the reduced camera system S and the triangular solves, FP64.
//
// Supersedes Ceres' SparseSchurComplementSolver -> CHOLMOD (third-party; reached through
// ceres::Solve with SPARSE_SCHUR, CeresHandler.h:403,419).
//
// S is cut into 96x96 tiles (8 frames) and stored TILE-PACKED in HBM: only the structurally
// non-zero lower tiles exist, slot = tile_slot[i*T + j], each tile row-major with ld = 96.
// The host runs a nested-dissection ordering and a symbolic factorisation on the tile graph once
// per scene (tile_plan.cu).  The numeric right-looking factorisation below only touches
// structurally non-zero tiles, so a video-like (banded) scene costs O(n b^2) while a fully
// covisible scene degenerates to the classic dense blocked algorithm; panels of one elimination
// level are independent and share one launch of each kernel.  Per panel k:
//   potrf_inv : L_kk = chol(A_kk) in shared memory (8x8-blocked), plus L_kk^-1 (explicit, so
//               that everything below is a GEMM / GEMV and not a substitution chain)
//   trsm      : L_ik = A_ik L_kk^-T              -- tile GEMM on the FP64 tensor cores (DMMA)
//   update    : A_ij -= L_ik L_jk^T              -- tile GEMM on the FP64 tensor cores (DMMA)
// tcgen05 has no FP64 kind; mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) is the FP64 tensor path on
// sm_100a.  This is the only place in the pipeline where tensor cores apply (dense GEMM).
#include "lm.cuh"

namespace rsba {
namespace {

constexpr int kLd = kTile + 1;      // potrf smem leading dimension (conflict-free columns)
constexpr int kGemmLd = kTile + 4;  // GEMM smem leading dimension (conflict-free DMMA fragments)

// ---------------------------------------------------------------- clear structurally non-zero tiles
__global__ void __launch_bounds__(256)
clear_tiles_kernel(double* __restrict__ S) {
  // every diagonal tile is rewritten by schur_reduce (incl. the identity padding rows); fill tiles
  // that no tile pair touches must start from zero
  double2* base = reinterpret_cast<double2*>(S + (long)blockIdx.x * kTile * kTile);
  for (int e = threadIdx.x; e < kTile * kTile / 2; e += blockDim.x) base[e] = make_double2(0.0, 0.0);
}

// ---------------------------------------------------------------- diagonal tile: Cholesky + inverse
__global__ void __launch_bounds__(256)
potrf_inv_kernel(double* __restrict__ S, const int* __restrict__ tile_slot, int T,
                 const int* __restrict__ panels, double* __restrict__ Dinv, int* __restrict__ info) {
  const int k = panels[blockIdx.x];
  constexpr long ld = kTile;
  extern __shared__ double smem[];
  double* A = smem;                 // [96][97]  factor
  double* Li = smem + kTile * kLd;  // [96][97]  inverse
  double* Tm = Li + kTile * kLd;    // [8][96]
  const int tid = threadIdx.x;
  double* g = S + (long)tile_slot[k * T + k] * kTile * kTile;
  for (int e = tid; e < kTile * kTile; e += blockDim.x) {
    const int r = e / kTile, c = e % kTile;
    A[r * kLd + c] = (c <= r) ? g[(long)r * ld + c] : 0.0;
    Li[r * kLd + c] = 0.0;
  }
  __syncthreads();

  for (int I = 0; I < kTile / 8; ++I) {
    const int o = 8 * I;
    // (i) 8x8 diagonal block, unblocked, by the first 8 lanes of warp 0
    if (tid < 32) {
      for (int j = 0; j < 8; ++j) {
        if (tid == j) {
          const double d = A[(o + j) * kLd + o + j];
          if (!(d > 0.0)) atomicExch(info, k * kTile + o + j + 1);
          A[(o + j) * kLd + o + j] = sqrt(d);
        }
        __syncwarp();
        if (tid < 8 && tid > j) A[(o + tid) * kLd + o + j] /= A[(o + j) * kLd + o + j];
        __syncwarp();
        if (tid < 8 && tid > j) {
          const double l = A[(o + tid) * kLd + o + j];
          for (int c = j + 1; c <= tid; ++c) A[(o + tid) * kLd + o + c] -= l * A[(o + c) * kLd + o + j];
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // (ii) rows below the block: forward substitution against the 8x8 factor, one thread per row
    {
      const int r = o + 8 + tid;
      if (r < kTile) {
        double x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          double s = A[r * kLd + o + j];
#pragma unroll
          for (int m = 0; m < j; ++m) s -= x[m] * A[(o + j) * kLd + o + m];
          x[j] = s / A[(o + j) * kLd + o + j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) A[r * kLd + o + j] = x[j];
      }
    }
    __syncthreads();
    // (iii) trailing lower triangle  A[i][c] -= sum_m P[i][m] P[c][m]
    {
      const int t = kTile - o - 8;
      for (int e = tid; e < t * t; e += blockDim.x) {
        const int i = e / t, c = e % t;
        if (c > i) continue;
        const double* pi = A + (o + 8 + i) * kLd + o;
        const double* pc = A + (o + 8 + c) * kLd + o;
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < 8; ++m) s += pi[m] * pc[m];
        A[(o + 8 + i) * kLd + o + 8 + c] -= s;
      }
    }
    __syncthreads();
  }

  // ---- inverse of the lower-triangular factor, 8x8-blocked
  if (tid < kTile) {
    const int o = (tid / 8) * 8, cc = tid % 8;
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double s = (i == cc) ? 1.0 : 0.0;
#pragma unroll
      for (int m = 0; m < i; ++m)
        if (m >= cc) s -= A[(o + i) * kLd + o + m] * x[m];
      x[i] = (i >= cc) ? s / A[(o + i) * kLd + o + i] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) Li[(o + i) * kLd + o + cc] = x[i];
  }
  __syncthreads();
  for (int I = 1; I < kTile / 8; ++I) {
    const int o = 8 * I;
    // T = L[I][0..I) * Li[0..I)[0..I)  (8 x 8I)
    for (int e = tid; e < 8 * o; e += blockDim.x) {
      const int r = e / o, gc = e % o;
      double s = 0.0;
      for (int m = (gc / 8) * 8; m < o; ++m) s += A[(o + r) * kLd + m] * Li[m * kLd + gc];
      Tm[r * kTile + gc] = s;
    }
    __syncthreads();
    for (int e = tid; e < 8 * o; e += blockDim.x) {
      const int r = e / o, gc = e % o;
      double s = 0.0;
      for (int q = 0; q <= r; ++q) s += Li[(o + r) * kLd + o + q] * Tm[q * kTile + gc];
      Li[(o + r) * kLd + gc] = -s;
    }
    __syncthreads();
  }

  double* di = Dinv + (long)k * kTile * kTile;
  for (int e = tid; e < kTile * kTile; e += blockDim.x) {
    const int r = e / kTile, c = e % kTile;
    if (c <= r) g[(long)r * ld + c] = A[r * kLd + c];
    di[e] = Li[r * kLd + c];
  }
}

// ---------------------------------------------------------------- tile GEMM  C (-)= A B^T  on DMMA
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void load_tile(double* dst, const double* __restrict__ src, long ld) {
  // 96x96 doubles, rows of 48 double2
  for (int e = threadIdx.x; e < kTile * (kTile / 2); e += blockDim.x) {
    const int r = e / (kTile / 2), c2 = e % (kTile / 2);
    const double2 v = reinterpret_cast<const double2*>(src + (long)r * ld)[c2];
    dst[r * kGemmLd + 2 * c2] = v.x;
    dst[r * kGemmLd + 2 * c2 + 1] = v.y;
  }
}

// MODE 0 (trsm):   S(i,k) = S(i,k) * Dinv[k]^T           one CTA per (i,k) of the level's trsm list
// MODE 1 (update): S(i,j) -= S(i,k) * S(j,k)^T           one CTA per (i,j,k) of the update group
template <int MODE>
__global__ void __launch_bounds__(256)
tile_gemm_kernel(double* S, const int* __restrict__ tile_slot, int T, const int2* __restrict__ trsm,
                 const int4* __restrict__ upd, const double* __restrict__ Dinv) {
  constexpr long ld = kTile;
  extern __shared__ double smem[];
  double* As = smem;
  double* Bs = smem + kTile * kGemmLd;
  int ti, tj, k;
  const double* Bsrc;
  if (MODE == 0) {
    const int2 p = trsm[blockIdx.x];
    ti = p.x;
    tj = k = p.y;
    Bsrc = Dinv + (long)k * kTile * kTile;
  } else {
    const int4 p = upd[blockIdx.x];
    ti = p.x;
    tj = p.y;
    k = p.z;
    Bsrc = S + (long)tile_slot[tj * T + k] * kTile * kTile;
  }
  const double* Asrc = S + (long)tile_slot[ti * T + k] * kTile * kTile;
  load_tile(As, Asrc, ld);
  load_tile(Bs, Bsrc, ld);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = (warp >> 2) * 48, n0 = (warp & 3) * 24;
  const int fr = lane >> 2, fc = lane & 3;
  double acc[6][3][2];
#pragma unroll
  for (int mi = 0; mi < 6; ++mi)
#pragma unroll
    for (int ni = 0; ni < 3; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
#pragma unroll 4
  for (int k0 = 0; k0 < kTile; k0 += 4) {
    double a[6], b[3];
#pragma unroll
    for (int mi = 0; mi < 6; ++mi) a[mi] = As[(m0 + 8 * mi + fr) * kGemmLd + k0 + fc];
#pragma unroll
    for (int ni = 0; ni < 3; ++ni) b[ni] = Bs[(n0 + 8 * ni + fr) * kGemmLd + k0 + fc];
#pragma unroll
    for (int mi = 0; mi < 6; ++mi)
#pragma unroll
      for (int ni = 0; ni < 3; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
  }
  double* Cg = S + (long)tile_slot[ti * T + tj] * kTile * kTile;
#pragma unroll
  for (int mi = 0; mi < 6; ++mi)
#pragma unroll
    for (int ni = 0; ni < 3; ++ni) {
      double2* dst = reinterpret_cast<double2*>(Cg + (long)(m0 + 8 * mi + fr) * ld + n0 + 8 * ni + 2 * fc);
      if (MODE == 0) {
        *dst = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
      } else {
        double2 v = *dst;
        v.x -= acc[mi][ni][0];
        v.y -= acc[mi][ni][1];
        *dst = v;
      }
    }
}

// ---------------------------------------------------------------- triangular solves, one launch per level
// x holds the right-hand side on entry and the solution of (L L^T) x = b on exit.
// forward:  z_k = Linv_kk (b_k - sum_{j<k} L_kj z_j)      every j sits in a lower level
__global__ void __launch_bounds__(256)
solve_forward_kernel(const double* __restrict__ S, TileSchedule ts, const int* __restrict__ panels,
                     double* __restrict__ x) {
  constexpr long ld = kTile;
  __shared__ double tmp[kTile];
  const int k = panels[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = ts.n_tiles;
  for (int r = warp; r < kTile; r += 8) {
    double s = 0.0;
    for (int q = ts.lrow_ptr[k]; q < ts.lrow_ptr[k + 1]; ++q) {
      const int j = ts.lrow_cols[q];
      const double* lj = S + (long)ts.tile_slot[k * T + j] * kTile * kTile + (long)r * ld;
      const double* xj = x + (long)j * kTile;
#pragma unroll
      for (int c = 0; c < kTile; c += 32) s += lj[c + lane] * xj[c + lane];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) tmp[r] = x[(long)k * kTile + r] - s;
  }
  __syncthreads();
  const double* di = ts.Dinv + (long)k * kTile * kTile;
  for (int r = warp; r < kTile; r += 8) {
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < kTile; c += 32)
      if (c + lane <= r) s += di[r * kTile + c + lane] * tmp[c + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) x[(long)k * kTile + r] = s;
  }
}

// backward: y_k = Linv_kk^T (z_k - sum_{i>k} L_ik^T y_i)   every i sits in a higher level
__global__ void __launch_bounds__(256)
solve_backward_kernel(const double* __restrict__ S, TileSchedule ts, const int* __restrict__ panels,
                      double* __restrict__ x) {
  constexpr long ld = kTile;
  __shared__ double tmp[kTile];
  __shared__ double part[8][kTile + 1];
  const int k = panels[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = ts.n_tiles;
  // thread (g = warp, r = lane + 32 m): partial over tile rows c = g, g+8, ... of every L_ik
  for (int m = 0; m < 3; ++m) {
    const int r = lane + 32 * m;
    double s = 0.0;
    for (int q = ts.row_ptr[k]; q < ts.row_ptr[k + 1]; ++q) {
      const int i = ts.rows[q];
      const double* lik = S + (long)ts.tile_slot[i * T + k] * kTile * kTile;
      const double* yi = x + (long)i * kTile;
      for (int c = warp; c < kTile; c += 8) s += lik[(long)c * ld + r] * yi[c];
    }
    part[warp][r] = s;
  }
  __syncthreads();
  if (threadIdx.x < kTile) {
    double s = 0.0;
#pragma unroll
    for (int g = 0; g < 8; ++g) s += part[g][threadIdx.x];
    tmp[threadIdx.x] = x[(long)k * kTile + threadIdx.x] - s;
  }
  __syncthreads();
  const double* di = ts.Dinv + (long)k * kTile * kTile;
  for (int m = 0; m < 3; ++m) {
    const int r = lane + 32 * m;
    double s = 0.0;
    for (int c = warp; c < kTile; c += 8)
      if (c >= r) s += di[c * kTile + r] * tmp[c];
    part[warp][r] = s;
  }
  __syncthreads();
  if (threadIdx.x < kTile) {
    double s = 0.0;
#pragma unroll
    for (int g = 0; g < 8; ++g) s += part[g][threadIdx.x];
    x[(long)k * kTile + threadIdx.x] = s;
  }
}

}  // namespace

void launch_clear_tiles(double* S, const TileSchedule& ts, cudaStream_t s) {
  if (ts.n_nz > 0) clear_tiles_kernel<<<ts.n_nz, 256, 0, s>>>(S);
}

int launch_tile_cholesky(double* S, const TileSchedule& ts, const TilePlan& plan, int* info, cudaStream_t s) {
  static bool attr_done = false;
  const size_t potrf_smem = (size_t)(2 * kTile * kLd + 8 * kTile) * sizeof(double);
  const size_t gemm_smem = (size_t)(2 * kTile * kGemmLd) * sizeof(double);
  if (!attr_done) {
    cudaFuncSetAttribute(potrf_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)potrf_smem);
    cudaFuncSetAttribute(tile_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem);
    cudaFuncSetAttribute(tile_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem);
    attr_done = true;
  }
  int launches = 0;
  for (int l = 0; l < plan.n_levels; ++l) {
    const int np = plan.panel_ptr[l + 1] - plan.panel_ptr[l];
    potrf_inv_kernel<<<np, 256, potrf_smem, s>>>(S, ts.tile_slot, ts.n_tiles, ts.panels + plan.panel_ptr[l], ts.Dinv, info);
    ++launches;
    const int nt = plan.trsm_ptr[l + 1] - plan.trsm_ptr[l];
    if (nt > 0) {
      tile_gemm_kernel<0><<<nt, 256, gemm_smem, s>>>(S, ts.tile_slot, ts.n_tiles, ts.trsm + plan.trsm_ptr[l], nullptr, ts.Dinv);
      ++launches;
    }
    for (int g = plan.level_group_ptr[l]; g < plan.level_group_ptr[l + 1]; ++g) {
      const long nu = plan.group_ptr[g + 1] - plan.group_ptr[g];
      if (nu <= 0) continue;
      tile_gemm_kernel<1><<<(unsigned)nu, 256, gemm_smem, s>>>(S, ts.tile_slot, ts.n_tiles, nullptr, ts.upd + plan.group_ptr[g], nullptr);
      ++launches;
    }
  }
  return launches;
}

int launch_tile_solve(const double* S, const TileSchedule& ts, const TilePlan& plan, double* x, cudaStream_t s) {
  int launches = 0;
  for (int l = 0; l < plan.n_levels; ++l) {
    const int np = plan.panel_ptr[l + 1] - plan.panel_ptr[l];
    solve_forward_kernel<<<np, 256, 0, s>>>(S, ts, ts.panels + plan.panel_ptr[l], x);
    ++launches;
  }
  for (int l = plan.n_levels - 1; l >= 0; --l) {
    const int np = plan.panel_ptr[l + 1] - plan.panel_ptr[l];
    solve_backward_kernel<<<np, 256, 0, s>>>(S, ts, ts.panels + plan.panel_ptr[l], x);
    ++launches;
  }
  return launches;
}

}  // namespace rsba
