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

#include <algorithm>

namespace rsba {
namespace {

constexpr int kLd = kTile + 4;   // smem leading dimension: conflict-free DMMA fragment loads (ld % 16 == 4)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// 16-byte asynchronous global -> shared copies (LDGSTS): a whole operand tile is put in flight at once.
// These kernels are short and latency-bound; a load -> store loop through registers serialised one L2
// round trip per iteration (18-36 of them per CTA) and was most of their run time.
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------- clear structurally non-zero tiles
__global__ void __launch_bounds__(256)
clear_tiles_kernel(double* __restrict__ S) {
  // every diagonal tile is rewritten by schur_reduce (incl. the identity padding rows); fill tiles
  // that no tile pair touches must start from zero
  double2* base = reinterpret_cast<double2*>(S + (long)blockIdx.x * kTile * kTile);
  for (int e = threadIdx.x; e < kTile * kTile / 2; e += blockDim.x) base[e] = make_double2(0.0, 0.0);
}

// ---------------------------------------------------------------- small GEMMs on shared memory
// C[m x n] = alpha * A[m x k] * op(B) (+ C), row-major smem operands with leading dimension kLd,
// m, n multiples of 8, k a multiple of 4.  op(B) = B (k x n) or B^T (B stored n x k).  The 8x8
// output blocks are dealt round-robin to warps wid, wid + nw, ...
template <bool TRANS_B, bool ACCUMULATE>
__device__ __forceinline__ void smem_gemm(double* C, const double* A, const double* B, int m, int n, int k,
                                          double alpha, int wid, int nw, int lane) {
  const int fr = lane >> 2, fc = lane & 3;
  const int nbn = n >> 3;
  for (int blk = wid; blk < (m >> 3) * nbn; blk += nw) {
    const int bi = blk / nbn, bj = blk % nbn;
    double c0 = 0.0, c1 = 0.0;
    const double* ap = A + (8 * bi + fr) * kLd + fc;
    const double* bp = TRANS_B ? B + (8 * bj + fr) * kLd + fc : B + fc * kLd + 8 * bj + fr;
    for (int kk = 0; kk < k; kk += 4) dmma(c0, c1, ap[kk], TRANS_B ? bp[kk] : bp[kk * kLd]);
    double* cp = C + (8 * bi + fr) * kLd + 8 * bj + 2 * fc;
    if (ACCUMULATE) { cp[0] += alpha * c0; cp[1] += alpha * c1; }
    else            { cp[0] = alpha * c0;  cp[1] = alpha * c1; }
  }
}

// ---------------------------------------------------------------- diagonal tile: Cholesky + inverse
// One CTA (8 warps) per panel of the level.  Right-looking, 8-column blocks:
//   (i)   the 8x8 diagonal block is factorised by warp 0 in registers (one row per lane, shuffles)
//   (ii)  the rows below solve against it (one thread per row, reciprocal pivots)
//   (iii) the trailing lower triangle is updated with DMMA (K = 8); warp 0 takes the next diagonal
//         block first and runs its step (i) while the other warps finish the update (look-ahead)
// then L^-1 by recursive doubling (8 -> 16 -> 32 -> 96; every level is a pair of small DMMA GEMMs),
// so that trsm and the triangular solves are GEMM / GEMV and not substitution chains.
constexpr size_t kPotrfSmem = (size_t)(2 * kTile * kLd + 3 * 32 * kLd + kTile) * sizeof(double);

__global__ void __launch_bounds__(256)
potrf_inv_kernel(double* __restrict__ S, const int* __restrict__ tile_slot, int T,
                 const int* __restrict__ panels, double* __restrict__ Dinv, int* __restrict__ info,
                 double* __restrict__ x, const int* __restrict__ lrow_ptr, const double* __restrict__ fwd_partials) {
  const int k = panels[blockIdx.x];
  extern __shared__ __align__(16) double smem[];
  double* A = smem;                      // [96][kLd]  factor (lower)
  double* X = A + kTile * kLd;           // [96][kLd]  inverse (lower)
  double* W = X + kTile * kLd;           // [3][32][kLd] products of the inverse
  double* rdiag = W + 3 * 32 * kLd;      // [96] reciprocal pivots
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double* g = S + (long)tile_slot[k * T + k] * kTile * kTile;
  for (int e = tid; e < kTile * kTile / 2; e += blockDim.x) {
    const int r = e / (kTile / 2), c = 2 * (e % (kTile / 2));
    cp_async16(A + r * kLd + c, g + 2 * e);
    X[r * kLd + c] = 0.0;
    X[r * kLd + c + 1] = 0.0;
  }
  cp_async_wait_all();
  __syncthreads();
  for (int e = tid; e < kTile * kTile; e += blockDim.x) {   // the factorisation works on the lower triangle
    const int r = e / kTile, c = e % kTile;
    if (c > r) A[r * kLd + c] = 0.0;
  }
  __syncthreads();

  // (i) 8x8 diagonal block at offset o, in registers: lane l < 8 owns row l (warp 0 only)
  auto potrf8 = [&](int o) {
    const int l = lane & 7;
    double a[8];   // lanes >= 8 only take part in the shuffles (sources are always lanes < 8)
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] = lane < 8 ? A[(o + l) * kLd + o + c] : 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const double d = __shfl_sync(0xffffffffu, a[j], j);
      if (lane == 0 && !(d > 0.0)) atomicExch(info, k * kTile + o + j + 1);
      const double rs = rsqrt(d);
      const double lj = (l == j) ? d * rs : a[j] * rs;   // column j of L (rows >= j are meaningful)
      a[j] = lj;
      if (lane == 0) rdiag[o + j] = rs;
#pragma unroll
      for (int c = j + 1; c < 8; ++c) a[c] -= lj * __shfl_sync(0xffffffffu, lj, c);
    }
    if (lane < 8) {
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c <= l) A[(o + l) * kLd + o + c] = a[c];
    }
  };
  if (warp == 0) potrf8(0);
  __syncthreads();

  for (int I = 0; I < kTile / 8; ++I) {
    const int o = 8 * I;
    // ---- (ii) rows below: x L8^T = a, one thread per row
    const int nrem = kTile - o - 8;
    if (tid < nrem) {
      double* row = A + (o + 8 + tid) * kLd + o;
      double x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        double sacc = row[j];
#pragma unroll
        for (int m = 0; m < j; ++m) sacc -= x[m] * A[(o + j) * kLd + o + m];
        x[j] = sacc * rdiag[o + j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) row[j] = x[j];
    }
    __syncthreads();
    // ---- (iii) trailing lower block triangle: A22 -= P P^T  (K = 8), DMMA.  Look-ahead: warp 0
    // updates the next diagonal block and factorises it straight away while warps 1..7 update the
    // rest, so the serial 8x8 factorisation is off the critical path.
    {
      const int nb = nrem >> 3, fr = lane >> 2, fc = lane & 3;
      const double* Pm = A + (o + 8) * kLd + o;
      double* C22 = A + (o + 8) * kLd + o + 8;
      auto update_block = [&](int bi, int bj) {
        double c0 = 0.0, c1 = 0.0;
        const double* ap = Pm + (8 * bi + fr) * kLd + fc;
        const double* bp = Pm + (8 * bj + fr) * kLd + fc;
        dmma(c0, c1, ap[0], bp[0]);
        dmma(c0, c1, ap[4], bp[4]);
        double* cp = C22 + (8 * bi + fr) * kLd + 8 * bj + 2 * fc;
        cp[0] -= c0;
        cp[1] -= c1;
      };
      if (warp == 0) {
        if (nb > 0) {
          update_block(0, 0);
          __syncwarp();
          potrf8(o + 8);
        }
      } else {
        // warp-strided walk over the lower block triangle (bi >= bj) without block (0,0)
        int bi = 1, bj = 0;
        for (int blk = 0, mine = warp - 1; bi < nb; ++blk) {
          if (blk == mine) {
            update_block(bi, bj);
            mine += 7;
          }
          if (++bj > bi) { bj = 0; ++bi; }
        }
      }
    }
    __syncthreads();
  }

  // ---- inverse, level 0: the twelve 8x8 diagonal blocks (thread = (block, column))
  if (tid < kTile) {
    const int o = (tid >> 3) * 8, cc = tid & 7;
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double sacc = (i == cc) ? 1.0 : 0.0;
#pragma unroll
      for (int m = 0; m < i; ++m)
        if (m >= cc) sacc -= A[(o + i) * kLd + o + m] * x[m];
      x[i] = (i >= cc) ? sacc * rdiag[o + i] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) X[(o + i) * kLd + o + cc] = x[i];
  }
  __syncthreads();
  // ---- level 1 (8 -> 16) and level 2 (16 -> 32):  X_ba = -X_b (L_ba X_a)
#pragma unroll 1
  for (int h = 8; h <= 16; h *= 2) {
    const int npair = kTile / (2 * h);              // 6 pairs of 8-blocks, then 3 pairs of 16-blocks
    const int wpp = (h == 8) ? 1 : 2;               // warps per pair
    if (warp < npair * wpp) {
      const int pr = warp / wpp, o = pr * 2 * h;
      smem_gemm<false, false>(W + pr * 16 * kLd, A + (o + h) * kLd + o, X + o * kLd + o, h, h, h, 1.0,
                              warp % wpp, wpp, lane);
    }
    __syncthreads();
    if (warp < npair * wpp) {
      const int pr = warp / wpp, o = pr * 2 * h;
      smem_gemm<false, false>(X + (o + h) * kLd + o, X + (o + h) * kLd + o + h, W + pr * 16 * kLd, h, h, h, -1.0,
                              warp % wpp, wpp, lane);
    }
    __syncthreads();
  }
  // ---- level 3 (32 -> 96): X21 = -X2 (L21 X1), X32 = -X3 (L32 X2), X31 = -X3 (L31 X1 + L32 X21)
  {
    double* T21 = W;
    double* T32 = W + 32 * kLd;
    double* T31 = W + 64 * kLd;
    const int half = warp >> 2, wq = warp & 3;      // warps 0-3 / 4-7 work on different products
    if (half == 0) smem_gemm<false, false>(T21, A + 32 * kLd, X, 32, 32, 32, 1.0, wq, 4, lane);
    else           smem_gemm<false, false>(T32, A + 64 * kLd + 32, X + 32 * kLd + 32, 32, 32, 32, 1.0, wq, 4, lane);
    __syncthreads();
    if (half == 0) smem_gemm<false, false>(X + 32 * kLd, X + 32 * kLd + 32, T21, 32, 32, 32, -1.0, wq, 4, lane);
    else           smem_gemm<false, false>(X + 64 * kLd + 32, X + 64 * kLd + 64, T32, 32, 32, 32, -1.0, wq, 4, lane);
    __syncthreads();
    smem_gemm<false, false>(T31, A + 64 * kLd, X, 32, 32, 32, 1.0, warp, 8, lane);
    __syncthreads();
    smem_gemm<false, true>(T31, A + 64 * kLd + 32, X + 32 * kLd, 32, 32, 32, 1.0, warp, 8, lane);
    __syncthreads();
    smem_gemm<false, false>(X + 64 * kLd, X + 64 * kLd + 64, T31, 32, 32, 32, -1.0, warp, 8, lane);
    __syncthreads();
  }

  double* di = Dinv + (long)k * kTile * kTile;
  for (int e = tid; e < kTile * kTile / 2; e += blockDim.x) {
    const int r = e / (kTile / 2), c = 2 * (e % (kTile / 2));
    reinterpret_cast<double2*>(g)[e] = make_double2(c <= r ? A[r * kLd + c] : 0.0, c + 1 <= r ? A[r * kLd + c + 1] : 0.0);
    reinterpret_cast<double2*>(di)[e] = make_double2(X[r * kLd + c], X[r * kLd + c + 1]);
  }
  // ---- forward substitution of this panel, while its inverse is in shared memory:
  //   z_k = L_kk^-1 (b_k - sum_{j<k} L_kj z_j);  the terms L_kj z_j were left by tile_trsm_kernel at the levels
  //   of the panels j (all lower than this one), one slot per tile, and are summed in list order
  double* tvec = rdiag;   // the reciprocal pivots are dead
  if (tid < kTile) {
    double sum = 0.0;
    for (int q = lrow_ptr[k]; q < lrow_ptr[k + 1]; ++q) sum += fwd_partials[(long)q * kTile + tid];
    tvec[tid] = x[(long)k * kTile + tid] - sum;
  }
  __syncthreads();
  {
    double zs[12];
#pragma unroll
    for (int t = 0; t < 12; ++t) {
      const int r = warp + 8 * t;
      double v = 0.0;
#pragma unroll
      for (int c = 0; c < kTile; c += 32)
        if (c + lane <= r) v += X[r * kLd + c + lane] * tvec[c + lane];
      zs[t] = v;
    }
#pragma unroll
    for (int t = 0; t < 12; ++t) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) zs[t] += __shfl_xor_sync(0xffffffffu, zs[t], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int t = 0; t < 12; ++t) x[(long)k * kTile + warp + 8 * t] = zs[t];
    }
  }
}

// ---------------------------------------------------------------- tile GEMMs  C (-)= A B^T  on DMMA
// Both kernels split one 96x96 output tile over 4 CTAs (blockIdx.y), so that a level with few tiles
// still spreads over many SMs; 4 warps per CTA, each a 24x24 patch (3x3 DMMA tiles), K = 96.
//   trsm   : S(i,k) = S(i,k) * Dinv[k]^T  in place; CTA = 24 rows (it reads and writes only its own rows)
//   update : S(i,j) -= S(i,k) * S(j,k)^T; CTA = 48x48 quadrant
constexpr int kQ = kTile / 2;
constexpr int kTrsmRows = kTile / 4;
constexpr size_t kUpdateSmem = (size_t)(2 * kQ * kLd) * sizeof(double);
constexpr size_t kTrsmSmem = (size_t)((kTrsmRows + kTile) * kLd) * sizeof(double);

__device__ __forceinline__ void load_rows(double* dst, const double* __restrict__ src, int rows) {
  // rows x 96 doubles, rows of 48 double2
  // (asynchronous: the caller runs cp_async_wait_all() + __syncthreads() before reading dst)
  for (int e = threadIdx.x; e < rows * (kTile / 2); e += blockDim.x) {
    const int r = e / (kTile / 2), c2 = e % (kTile / 2);
    cp_async16(dst + r * kLd + 2 * c2, src + (long)r * kTile + 2 * c2);
  }
}

// acc[3][3][2] += As[m0.., :] * Bs[n0.., :]^T over K = 96
__device__ __forceinline__ void warp_gemm_24x24(const double* As, const double* Bs, int m0, int n0, int lane,
                                                double (&acc)[3][3][2]) {
  const int fr = lane >> 2, fc = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 3; ++mi)
#pragma unroll
    for (int ni = 0; ni < 3; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
#pragma unroll 4
  for (int k0 = 0; k0 < kTile; k0 += 4) {
    double a[3], b[3];
#pragma unroll
    for (int mi = 0; mi < 3; ++mi) a[mi] = As[(m0 + 8 * mi + fr) * kLd + k0 + fc];
#pragma unroll
    for (int ni = 0; ni < 3; ++ni) b[ni] = Bs[(n0 + 8 * ni + fr) * kLd + k0 + fc];
#pragma unroll
    for (int mi = 0; mi < 3; ++mi)
#pragma unroll
      for (int ni = 0; ni < 3; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
  }
}

__global__ void __launch_bounds__(128)
tile_trsm_kernel(double* S, const int* __restrict__ tile_slot, int T, const int2* __restrict__ trsm,
                 const double* __restrict__ Dinv, const double* __restrict__ x, const int* __restrict__ fwd_slot,
                 double* __restrict__ fwd_partials) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double zs[kTile];             // z_k of the panel (written by this level's potrf_inv launch)
  __shared__ double red[4][kTrsmRows];
  double* As = smem;                       // [24][kLd]  this CTA's rows of S(i,k)
  double* Bs = smem + kTrsmRows * kLd;     // [96][kLd]  Dinv[k]
  const int2 p = trsm[blockIdx.x];
  if (threadIdx.x < kTile) zs[threadIdx.x] = x[(long)p.y * kTile + threadIdx.x];
  double* tile = S + (long)tile_slot[p.x * T + p.y] * kTile * kTile + (long)blockIdx.y * kTrsmRows * kTile;
  load_rows(As, tile, kTrsmRows);
  load_rows(Bs, Dinv + (long)p.y * kTile * kTile, kTile);
  cp_async_wait_all();
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int fr = lane >> 2, fc = lane & 3, n0 = warp * 24;
  double acc[3][3][2];
  warp_gemm_24x24(As, Bs, 0, n0, lane, acc);
#pragma unroll
  for (int mi = 0; mi < 3; ++mi)
#pragma unroll
    for (int ni = 0; ni < 3; ++ni)
      *reinterpret_cast<double2*>(tile + (long)(8 * mi + fr) * kTile + n0 + 8 * ni + 2 * fc) =
          make_double2(acc[mi][ni][0], acc[mi][ni][1]);
  // forward-substitution term of this tile: (L_ik z_k)[rows of this CTA], fixed summation order
  // (zs was written before the __syncthreads above)
#pragma unroll
  for (int mi = 0; mi < 3; ++mi) {
    double v = 0.0;
#pragma unroll
    for (int ni = 0; ni < 3; ++ni)
      v += acc[mi][ni][0] * zs[n0 + 8 * ni + 2 * fc] + acc[mi][ni][1] * zs[n0 + 8 * ni + 2 * fc + 1];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (fc == 0) red[warp][8 * mi + fr] = v;
  }
  __syncthreads();
  if (threadIdx.x < kTrsmRows)
    fwd_partials[(long)fwd_slot[blockIdx.x] * kTile + blockIdx.y * kTrsmRows + threadIdx.x] =
        (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
}

__global__ void __launch_bounds__(128)
tile_update_kernel(double* S, const int* __restrict__ tile_slot, int T, const int4* __restrict__ upd) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + kQ * kLd;
  const int4 p = upd[blockIdx.x];
  const int ti = p.x, tj = p.y, k = p.z;
  const int qi = blockIdx.y >> 1, qj = blockIdx.y & 1;
  if (ti == tj && qj > qi) return;   // diagonal target: the factorisation reads the lower part only
  load_rows(As, S + (long)tile_slot[ti * T + k] * kTile * kTile + (long)qi * kQ * kTile, kQ);
  load_rows(Bs, S + (long)tile_slot[tj * T + k] * kTile * kTile + (long)qj * kQ * kTile, kQ);
  cp_async_wait_all();
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = (warp >> 1) * 24, n0 = (warp & 1) * 24;
  const int fr = lane >> 2, fc = lane & 3;
  double acc[3][3][2];
  double* Cg = S + (long)tile_slot[ti * T + tj] * kTile * kTile + (long)qi * kQ * kTile + qj * kQ;
  double2 cv[3][3];   // the target's old values: requested before the GEMM, consumed after it
#pragma unroll
  for (int mi = 0; mi < 3; ++mi)
#pragma unroll
    for (int ni = 0; ni < 3; ++ni)
      cv[mi][ni] = *reinterpret_cast<const double2*>(Cg + (long)(m0 + 8 * mi + fr) * kTile + n0 + 8 * ni + 2 * fc);
  warp_gemm_24x24(As, Bs, m0, n0, lane, acc);
#pragma unroll
  for (int mi = 0; mi < 3; ++mi)
#pragma unroll
    for (int ni = 0; ni < 3; ++ni)
      *reinterpret_cast<double2*>(Cg + (long)(m0 + 8 * mi + fr) * kTile + n0 + 8 * ni + 2 * fc) =
          make_double2(cv[mi][ni].x - acc[mi][ni][0], cv[mi][ni].y - acc[mi][ni][1]);
}

// ---------------------------------------------------------------- triangular solves, level by level
// The FORWARD substitution z = L^-1 b rides in the factorisation's own launches (potrf_inv_kernel's epilogue
// forms z_k while the panel's inverse is in shared memory, tile_trsm_kernel's epilogue leaves L_ik z_k in the
// tile's slot): the 19 + 12 extra launches of a separate forward sweep were a quarter of the solve time.
// Below: the backward substitution.  x holds z on entry and the solution of (L L^T) x = b on exit.  The tile row /
// column of a separator panel can hold dozens of tiles, so the matrix-vector products of one
// panel are split over `split` CTAs (blockIdx.y) that write partial sums; a finish kernel adds
// them in a fixed order and applies the inverse of the diagonal factor.  split == 1 fuses both.
constexpr int kMaxSolveSplit = 16;

__device__ __forceinline__ void apply_dinv_backward(const double* __restrict__ di, const double* tmp,
                                                    double (*part)[kTile + 1], double* __restrict__ xk,
                                                    int warp, int lane) {
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    const int r = lane + 32 * m;
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < 12; ++t) {     // fully unrolled: the 36 loads of the inverse go out together
      const int c = warp + 8 * t;
      if (c >= r) s += __ldg(di + c * kTile + r) * tmp[c];
    }
    part[warp][r] = s;
  }
  __syncthreads();
  if (threadIdx.x < kTile) {
    double s = 0.0;
#pragma unroll
    for (int g = 0; g < 8; ++g) s += part[g][threadIdx.x];
    xk[threadIdx.x] = s;
  }
}

// backward: y_k = Linv_kk^T (z_k - sum_{i>k} L_ik^T y_i)   every i sits in a higher level
__global__ void __launch_bounds__(256)
solve_backward_kernel(const double* __restrict__ S, TileSchedule ts, const int* __restrict__ panels,
                      double* __restrict__ x, int split, double* __restrict__ partials) {
  constexpr long ld = kTile;
  __shared__ double tmp[kTile];
  __shared__ double part[8][kTile + 1];
  const int k = panels[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = ts.n_tiles;
  // thread (g = warp, r = lane + 32 m): partial over tile rows c = g, g+8, ... of every L_ik
  double acc[3] = {0.0, 0.0, 0.0};
  for (int q = ts.row_ptr[k] + blockIdx.y; q < ts.row_ptr[k + 1]; q += split) {
    const int i = ts.rows[q];
    const double* lik = S + (long)ts.tile_slot[i * T + k] * kTile * kTile;
    const double* yi = x + (long)i * kTile;
#pragma unroll
    for (int t = 0; t < 12; ++t) {
      const int c = warp + 8 * t;
      const double y = yi[c];
      const double* row = lik + (long)c * ld;
      acc[0] += row[lane] * y;
      acc[1] += row[lane + 32] * y;
      acc[2] += row[lane + 64] * y;
    }
  }
#pragma unroll
  for (int m = 0; m < 3; ++m) part[warp][lane + 32 * m] = acc[m];
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x < kTile) {
#pragma unroll
    for (int g = 0; g < 8; ++g) s += part[g][threadIdx.x];
  }
  if (split > 1) {
    if (threadIdx.x < kTile) partials[((long)k * kMaxSolveSplit + blockIdx.y) * kTile + threadIdx.x] = s;
    return;
  }
  if (threadIdx.x < kTile) tmp[threadIdx.x] = x[(long)k * kTile + threadIdx.x] - s;
  __syncthreads();
  apply_dinv_backward(ts.Dinv + (long)k * kTile * kTile, tmp, part, x + (long)k * kTile, warp, lane);
}

__global__ void __launch_bounds__(256)
solve_backward_finish_kernel(TileSchedule ts, const int* __restrict__ panels, double* __restrict__ x, int split,
                             const double* __restrict__ partials) {
  __shared__ double tmp[kTile];
  __shared__ double part[8][kTile + 1];
  const int k = panels[blockIdx.x];
  if (threadIdx.x < kTile) {
    double s = 0.0;
    for (int g = 0; g < split; ++g) s += partials[((long)k * kMaxSolveSplit + g) * kTile + threadIdx.x];
    tmp[threadIdx.x] = x[(long)k * kTile + threadIdx.x] - s;
  }
  __syncthreads();
  apply_dinv_backward(ts.Dinv + (long)k * kTile * kTile, tmp, part, x + (long)k * kTile, threadIdx.x >> 5,
                      threadIdx.x & 31);
}

}  // namespace

void launch_clear_tiles(double* S, const TileSchedule& ts, cudaStream_t s) {
  if (ts.n_nz > 0) clear_tiles_kernel<<<ts.n_nz, 256, 0, s>>>(S);
}

void k3_prepare() {
  static bool seen[64] = {};
  if (first_use_on_device(seen)) {
    cudaFuncSetAttribute(potrf_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPotrfSmem);
    cudaFuncSetAttribute(tile_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTrsmSmem);
    cudaFuncSetAttribute(tile_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUpdateSmem);
  }
}

int launch_tile_cholesky(double* S, const TileSchedule& ts, const TilePlan& plan, double* x, int* info, cudaStream_t s) {
  k3_prepare();
  int launches = 0;
  for (int l = 0; l < plan.n_levels; ++l) {
    const int np = plan.panel_ptr[l + 1] - plan.panel_ptr[l];
    potrf_inv_kernel<<<np, 256, kPotrfSmem, s>>>(S, ts.tile_slot, ts.n_tiles, ts.panels + plan.panel_ptr[l], ts.Dinv, info,
                                                 x, ts.lrow_ptr, ts.fwd_partials);
    ++launches;
    const int nt = plan.trsm_ptr[l + 1] - plan.trsm_ptr[l];
    if (nt > 0) {
      tile_trsm_kernel<<<dim3(nt, 4), 128, kTrsmSmem, s>>>(S, ts.tile_slot, ts.n_tiles, ts.trsm + plan.trsm_ptr[l], ts.Dinv,
                                                           x, ts.fwd_slot + plan.trsm_ptr[l], ts.fwd_partials);
      ++launches;
    }
    for (int g = plan.level_group_ptr[l]; g < plan.level_group_ptr[l + 1]; ++g) {
      const long nu = plan.group_ptr[g + 1] - plan.group_ptr[g];
      if (nu <= 0) continue;
      tile_update_kernel<<<dim3((unsigned)nu, 4), 128, kUpdateSmem, s>>>(S, ts.tile_slot, ts.n_tiles, ts.upd + plan.group_ptr[g]);
      ++launches;
    }
  }
  return launches;
}

int launch_tile_solve(const double* S, const TileSchedule& ts, const TilePlan& plan, double* x, cudaStream_t s) {
  int launches = 0;
  auto split_of = [](int most) { return most <= 4 ? 1 : std::min(kMaxSolveSplit, (most + 2) / 3); };
  // (the forward substitution z = L^-1 b was done inside launch_tile_cholesky)
  for (int l = plan.n_levels - 1; l >= 0; --l) {
    const int np = plan.panel_ptr[l + 1] - plan.panel_ptr[l];
    int most = 0;
    for (int q = plan.panel_ptr[l]; q < plan.panel_ptr[l + 1]; ++q)
      most = std::max(most, plan.row_ptr[plan.panels[q] + 1] - plan.row_ptr[plan.panels[q]]);
    const int split = split_of(most);
    solve_backward_kernel<<<dim3(np, split), 256, 0, s>>>(S, ts, ts.panels + plan.panel_ptr[l], x, split, ts.solve_partials);
    ++launches;
    if (split > 1) {
      solve_backward_finish_kernel<<<np, 256, 0, s>>>(ts, ts.panels + plan.panel_ptr[l], x, split, ts.solve_partials);
      ++launches;
    }
  }
  return launches;
}

}  // namespace rsba
