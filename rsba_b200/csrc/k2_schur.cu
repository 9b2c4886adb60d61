// K2s -- the Schur complement of the point blocks as a tile-level SYRK on the FP64 tensor path.
//
// Supersedes the reduction loop of Ceres' SchurEliminator::Eliminate (third-party, reached
// through ceres::Solve with SPARSE_SCHUR, CeresHandler.h:403,419):
//     S = diag(s B s + D^2) - sum_p E_p C_p^-1 E_p^T .
// With C_p'^-1 = L^-T L^-1 (3x3 Cholesky inverted in registers, k2_normal.cu) the per-observation
// block  F_i = Jc_i^T (Jx_i s_p) L^-T  (12x3) makes the point term a plain Gram product:
//     sum_p E_p C_p^-1 E_p^T = Phi Phi^T,   Phi = [F_i] block-sparse, 12F x 3P.
// Frames are cut into sub-tiles of 4 (48 rows, a quadrant of the 96-row Cholesky tile).  For every
// (sub-tile, point) incidence phi_build writes one dense k-major panel [3][48(+4 pad)]; rows of
// frames that do not see the point are zero.  A sub-tile pair (a <= b) then is a dense GEMM over
// the points seen from both, K = 3 per point:
//     S(b, a) -= sum_p Phi(b,p) Phi(a,p)^T                      48 x 48, mma.sync.m8n8k4.f64
// (48 rather than 96 rows: a point seen from 25 consecutive frames fills 86 % of its 4-frame
// panels but only 78 % of 8-frame ones, and its 7 x 8 / 2 sub-tile pairs waste less of the
// squared fill -- 0.72x the DMMAs of 96-row tiles.)
// Long pairs are split along K into work items of <= 512 points whose 48x48 partial products
// are summed in a fixed order by schur_reduce (bit-reproducible, no atomics), which also adds
// the camera diagonal blocks and places the tile at its (permuted) position of the tile-packed
// reduced matrix.  schur_finalize then applies the camera-side Jacobi scaling, the LM diagonal
// and the constant-parameter identity rows -- after the NCCL all-reduce when the observations
// are sharded over several GPUs, because those depend on the global diag(B).
//
// Data movement: panels are 1248-byte contiguous records, fetched by TMA bulk copies
// (cp.async.bulk.shared::cluster.global.mbarrier) into a 3-stage shared-memory ring; the kernel
// is bound by the FP64 pipe (tcgen05 has no FP64 kind: DMMA == DFMA rate, 37.1 TFLOP/s measured).
#include "lm.cuh"

#include <type_traits>

namespace rsba {
namespace {

constexpr unsigned kPanelBytes = kPanelDoubles * sizeof(double);
// pipeline shape: points per stage (CHUNK: 8 = 24 K columns), stages, CTAs per SM (the launcher at the bottom picks it)
template <int CHUNK, int STAGES>
constexpr size_t syrk_smem() { return (size_t)STAGES * 2 * CHUNK * kPanelDoubles * sizeof(double) + 64; }

// ---------------------------------------------------------------- panels
// The panel rows of an observation are normally written by frame_blocks_kernel (k2_normal.cu), which
// already holds the Jacobian records in shared memory.  This kernel only rebuilds the incidences in
// which a point was observed twice in one frame (st.dup_inc), where the rows are sums:
// one thread per (incidence, frame slot): F = sum over the slot's observations
// of  Jc^T (Jx s_p) L^-T ; constant camera
// parameters get zero rows, constant points have L^-1 = 0 and never appear in a pair.  The
// camera-side Jacobi scaling is applied to S afterwards (schur_finalize), so that partial sums of
// different GPUs can be added before the scaling -- which depends on the global diag(B) -- is known.
__global__ void __launch_bounds__(256)
phi_build_kernel(SchurStructure st, ObsView obs, JacView jv, NormalEq ne) {
  const long u = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= (long)st.n_dup * kSubFrames) return;
  const int inc = st.dup_inc[u / kSubFrames], fs = (int)(u % kSubFrames);
  const long t = (long)inc * kSubFrames + fs;
  double F[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) F[k] = 0.0;
  const int cnt = st.slot_cnt[t];
  if (cnt > 0) {
    const int p = st.inc_point[inc];
    const int f = st.inc_tile[inc] * kSubFrames + fs;
    const double* Mi = ne.Minv + 6L * p;
    const double m00 = Mi[0], m10 = Mi[1], m11 = Mi[2], m20 = Mi[3], m21 = Mi[4], m22 = Mi[5];
    const double sp0 = ne.scale_p[3L * p], sp1 = ne.scale_p[3L * p + 1], sp2 = ne.scale_p[3L * p + 2];
    const unsigned mask = ne.pose_mask[f];
    double sc[12];
#pragma unroll
    for (int a = 0; a < 12; ++a) sc[a] = ((mask >> a) & 1) ? 0.0 : 1.0;
    const int beg = st.slot_beg[t];
    for (int d = 0; d < cnt; ++d) {
      const long i = st.pt_obs[beg + d];
      double Jf[kJacDoubles];
      load_full_jacobian(jv, jv.point_major ? (long)(beg + d) : i, Jf);
      const double2* J = reinterpret_cast<const double2*>(Jf);
      const double2 x0 = J[12], x1 = J[13], x2 = J[14];   // Jx rows: (x0.x x0.y x1.x) (x1.y x2.x x2.y)
      // xm[row][k] = sum_c Jx[row][c] s_p[c] L^-1[k][c]
      const double a0 = x0.x * sp0, a1 = x0.y * sp1, a2 = x1.x * sp2;
      const double b0 = x1.y * sp0, b1 = x2.x * sp1, b2 = x2.y * sp2;
      const double xa0 = a0 * m00, xa1 = a0 * m10 + a1 * m11, xa2 = a0 * m20 + a1 * m21 + a2 * m22;
      const double xb0 = b0 * m00, xb1 = b0 * m10 + b1 * m11, xb2 = b0 * m20 + b1 * m21 + b2 * m22;
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const double2 r0 = J[h2 * 6 + q], r1 = J[h2 * 6 + 3 + q];   // rows 0/1, columns 2q, 2q+1 of this half
          const int a = h2 * 6 + 2 * q;
          const double j00 = r0.x * sc[a], j10 = r1.x * sc[a];
          const double j01 = r0.y * sc[a + 1], j11 = r1.y * sc[a + 1];
          F[3 * a + 0] += j00 * xa0 + j10 * xb0;
          F[3 * a + 1] += j00 * xa1 + j10 * xb1;
          F[3 * a + 2] += j00 * xa2 + j10 * xb2;
          F[3 * a + 3] += j01 * xa0 + j11 * xb0;
          F[3 * a + 4] += j01 * xa1 + j11 * xb1;
          F[3 * a + 5] += j01 * xa2 + j11 * xb2;
        }
    }
  }
  double* dst = ne.Phi + (long)inc * kPanelDoubles + fs * kFrameParams;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double2* d2 = reinterpret_cast<double2*>(dst + k * kPanelLd);
#pragma unroll
    for (int a = 0; a < 12; a += 2) d2[a >> 1] = make_double2(F[3 * a + k], F[3 * (a + 1) + k]);
  }
}

// ---------------------------------------------------------------- mbarrier / TMA helpers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
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
__device__ __forceinline__ void tma_load_1d(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---------------------------------------------------------------- sub-tile-pair SYRK
// One CTA (4 warps) per work item.  60 KB of shared memory per CTA: three CTAs per SM keep 12 warps on
// the FP64 pipe.
template <int kChunkPts, int kStages, int MIN_CTAS>
__global__ void __launch_bounds__(128, MIN_CTAS)
schur_syrk_kernel(const double* __restrict__ Phi, const int2* __restrict__ entries,
                  const int4* __restrict__ items, double* __restrict__ partial) {
  constexpr int kOperandDoubles = kChunkPts * kPanelDoubles;       // one side of a stage
  constexpr int kStageDoubles = 2 * kOperandDoubles;               // row side | column side
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stage_base = reinterpret_cast<double*>(smem_raw);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)kStages * kStageDoubles * sizeof(double));

  const int4 item = items[blockIdx.x];
  const int beg = item.y, nchunks = item.z / kChunkPts;
  const bool diag = (item.w & 1) != 0;
  // off-diagonal items: which 2-frame halves of the column (A) / row (B) side are populated for EVERY entry
  const unsigned half_a = ((unsigned)item.w >> 4) & 3u, half_b = ((unsigned)item.w >> 8) & 3u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&bars[s]), 1);              // "full": the stage's TMA bytes have landed
      mbar_init(smem_u32(&bars[kStages + s]), 4);    // "empty": all four warps are done reading the stage
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // producer: warp 0; lane q < 8 fetches the row-side panel of point q, lane 8+q the column side.  The entry
  // (panel indices) of the chunk after the one being issued is already in a register, so the refill does not
  // wait on a global load.
  int2 e_next = make_int2(0, 0);
  auto load_entry = [&](int c) {
    if (lane < (diag ? kChunkPts : 2 * kChunkPts) && c < nchunks) e_next = entries[beg + c * kChunkPts + (lane % kChunkPts)];
  };
  auto issue = [&](int c) {
    const int s = c % kStages;
    const unsigned bar = smem_u32(&bars[s]);
    if (lane == 0) mbar_expect_tx(bar, (diag ? 1u : 2u) * kChunkPts * kPanelBytes);
    __syncwarp();
    const int2 e = e_next;
    load_entry(c + 1);
    if (lane < (diag ? kChunkPts : 2 * kChunkPts)) {
      const int q = lane % kChunkPts, side = lane / kChunkPts;
      const int inc = side ? e.y : e.x;
      double* dst = stage_base + (size_t)s * kStageDoubles + side * kOperandDoubles + q * kPanelDoubles;
      tma_load_1d(smem_u32(dst), Phi + (long)inc * kPanelDoubles, kPanelBytes, bar);
    }
  };
  if (warp == 0) {
    load_entry(0);
    for (int c = 0; c < kStages && c < nchunks; ++c) issue(c);
  }
  // The warps are NOT kept in lockstep: a warp that has read stage s arrives on the stage's "empty" barrier
  // and moves on; warp 0 refills the stage of chunk c - 1 when it reaches chunk c, by which time the others
  // have normally passed it.  (A CTA-wide barrier per 8-point chunk was 30 % of this kernel's stall samples.)
  auto refill = [&](int c) {          // warp 0, at the top of chunk c
    const int prev = c - 1;
    if (prev >= 0 && prev + kStages < nchunks) {
      mbar_wait(smem_u32(&bars[kStages + prev % kStages]), (unsigned)((prev / kStages) & 1));
      issue(prev + kStages);
    }
  };
  auto release = [&](int s) {         // every warp, after its last read of stage s
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[kStages + s])) : "memory");
  };

  // The 48 x 48 block is a 6 x 6 grid of 8 x 8 DMMA tiles.  Off-diagonal pairs: warp w owns the 3 x 3
  // tiles at (3 (w>>1), 3 (w&1)).  Diagonal pairs (a == b) are symmetric and only their lower tiles are
  // ever read, so the 21 lower tiles are dealt 6 / 6 / 5 / 4 to the warps instead of 9 each: every
  // variant is "a 3 x 3 window at (r0, c0) with a compile-time tile mask" (bit 3 i + j).
  constexpr unsigned kAll = 0x1FF;
  constexpr unsigned kLower = (1u << 0) | (1u << 3) | (1u << 4) | (1u << 6) | (1u << 7) | (1u << 8);
  constexpr unsigned kRows34 = (1u << 0) | (1u << 1) | (1u << 2) | (1u << 3) | (1u << 4);       // window (3, 0)
  constexpr unsigned kRows45 = (1u << 5) | (1u << 6) | (1u << 7) | (1u << 8);                   // window (3, 0)
  double* out = partial + (long)blockIdx.x * kSub * kSub;
  // ksplit: the warp only multiplies the chunks with (c & 1) == kphase -- the K-split of an item with two live
  // patches, whose other two warps would otherwise idle (13.7 % of the DMMAs but 23.5 % of the CTA time at C3); the
  // two halves of a patch are added in a fixed order through shared memory behind the loop.
  auto run = [&](auto mask_tag, int r0, int c0, bool ksplit, int kphase, int zr0, int zc0) {
    constexpr unsigned MASK = decltype(mask_tag)::value;
    const int m0 = 8 * r0, n0 = 8 * c0;
    const int fr = lane >> 2, fc = lane & 3;
    double acc[3][3][2];
#pragma unroll
    for (int mi = 0; mi < 3; ++mi)
#pragma unroll
      for (int ni = 0; ni < 3; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % kStages;
      if (warp == 0) refill(c);
      mbar_wait(smem_u32(&bars[s]), (unsigned)((c / kStages) & 1));
      const double* R = stage_base + (size_t)s * kStageDoubles;
      const double* Cc = diag ? R : R + kOperandDoubles;
      if (!ksplit || (c & 1) == kphase)
#pragma unroll
      for (int ks = 0; ks < (3 * kChunkPts) / 4; ++ks) {
        const int krow = (4 * ks + fc) * kPanelLd + fr;
        double a[3], b[3];
#pragma unroll
        for (int mi = 0; mi < 3; ++mi)
          if ((MASK >> (3 * mi)) & 7u) a[mi] = R[krow + m0 + 8 * mi];
#pragma unroll
        for (int ni = 0; ni < 3; ++ni)
          if ((MASK >> ni) & 0x49u) b[ni] = Cc[krow + n0 + 8 * ni];
#pragma unroll
        for (int mi = 0; mi < 3; ++mi)
#pragma unroll
          for (int ni = 0; ni < 3; ++ni)
            if ((MASK >> (3 * mi + ni)) & 1u) dmma8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
      }
      release(s);
    }
    if (ksplit) {   // CTA-uniform: every warp of a two-patch item comes through here
      __syncthreads();                       // all stages consumed: the ring is free
      double2* buf = reinterpret_cast<double2*>(stage_base) + ((r0 + c0) & 1 ? 288 : 0);   // one slab per live patch
      if (kphase == 1) {
#pragma unroll
        for (int mi = 0; mi < 3; ++mi)
#pragma unroll
          for (int ni = 0; ni < 3; ++ni) buf[(3 * mi + ni) * 32 + lane] = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
      }
      __syncthreads();
      if (kphase == 1) {                     // this warp's own patch of the 48 x 48 block is structurally zero
#pragma unroll
        for (int mi = 0; mi < 3; ++mi)
#pragma unroll
          for (int ni = 0; ni < 3; ++ni)
            *reinterpret_cast<double2*>(out + (8 * zr0 + 8 * mi + fr) * kSub + 8 * zc0 + 8 * ni + 2 * fc) = make_double2(0.0, 0.0);
        return;
      }
#pragma unroll
      for (int mi = 0; mi < 3; ++mi)
#pragma unroll
        for (int ni = 0; ni < 3; ++ni) {
          const double2 o = buf[(3 * mi + ni) * 32 + lane];
          acc[mi][ni][0] += o.x;
          acc[mi][ni][1] += o.y;
        }
    }
#pragma unroll
    for (int mi = 0; mi < 3; ++mi)
#pragma unroll
      for (int ni = 0; ni < 3; ++ni)
        if ((MASK >> (3 * mi + ni)) & 1u)
          *reinterpret_cast<double2*>(out + (m0 + 8 * mi + fr) * kSub + n0 + 8 * ni + 2 * fc) =
              make_double2(acc[mi][ni][0], acc[mi][ni][1]);
  };
  // A warp whose 24 x 24 patch is structurally zero for this item (a track that starts or ends inside the
  // sub-tile leaves a 2-frame half empty) only keeps the pipeline's barriers -- and, for warp 0, the TMA refills --
  // going and writes zeros: its DMMA slots go to the other CTAs of the SM.
  auto idle = [&](int r0, int c0) {
    for (int c = 0; c < nchunks; ++c) {      // paced by the TMA like everyone else (one arrival per phase)
      const int s = c % kStages;
      if (warp == 0) refill(c);
      mbar_wait(smem_u32(&bars[s]), (unsigned)((c / kStages) & 1));
      release(s);
    }
    const int fr = lane >> 2, fc = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 3; ++mi)
#pragma unroll
      for (int ni = 0; ni < 3; ++ni)
        *reinterpret_cast<double2*>(out + (8 * r0 + 8 * mi + fr) * kSub + 8 * c0 + 8 * ni + 2 * fc) = make_double2(0.0, 0.0);
  };
  if (!diag) {
    const bool one_b = half_b == 1u || half_b == 2u, one_a = half_a == 1u || half_a == 2u;
    if (one_b != one_a) {
      // two live patches (one side has a single populated half): warp = (patch, K half).  The patch a warp owns
      // in the plain layout but not here -- (zr, zc) -- is structurally zero and written as such.
      int pr, pc, kphase, zr, zc;
      if (one_b) { pr = half_b == 2u; pc = warp & 1; kphase = warp >> 1; zr = 1 - pr; zc = pc; }
      else       { pc = half_a == 2u; pr = warp >> 1; kphase = warp & 1; zr = pr; zc = 1 - pc; }
      if (kphase == 0) { zr = pr; zc = pc; }   // (unused: the first half writes the sum into the live patch)
      run(std::integral_constant<unsigned, kAll>{}, 3 * pr, 3 * pc, true, kphase, 3 * zr, 3 * zc);
    } else {
      const bool live = ((half_b >> (warp >> 1)) & 1u) && ((half_a >> (warp & 1)) & 1u);
      if (live) run(std::integral_constant<unsigned, kAll>{}, 3 * (warp >> 1), 3 * (warp & 1), false, 0, 0, 0);
      else idle(3 * (warp >> 1), 3 * (warp & 1));
    }
  }
  else if (warp < 2) run(std::integral_constant<unsigned, kLower>{}, 3 * warp, 3 * warp, false, 0, 0, 0);
  else if (warp == 2) run(std::integral_constant<unsigned, kRows34>{}, 3, 0, false, 0, 0, 0);
  else run(std::integral_constant<unsigned, kRows45>{}, 3, 0, false, 0, 0, 0);
}

// ---------------------------------------------------------------- reduce
// grid (n_pairs, 3): 768 elements of the pair's 48 x 48 block per CTA, 3 per thread.  Writes the
// unscaled block  [a == b] B_f - sum of the pair's partial products  into quadrant (b%2, a%2) of
// Cholesky tile (b/2, a/2) at its (permuted) place.
__global__ void __launch_bounds__(256)
schur_reduce_kernel(SchurStructure st, NormalEq ne, PriorView pv, int cam_frame, double* __restrict__ S,
                    const int* __restrict__ tile_slot, int T) {
  const int pr = blockIdx.x;
  const int a = st.pair_a[pr], b = st.pair_b[pr];
  const int ib = st.pair_item_ptr[pr], ie = st.pair_item_ptr[pr + 1];
  const int A = a >> 1, B = b >> 1;
  const int pa = st.tile_pos[A], pb = st.tile_pos[B];
  const bool transposed = pb < pa;           // rows must be the later position (lower triangle)
  double* tile = S + (long)tile_slot[(transposed ? pa : pb) * T + (transposed ? pb : pa)] * kTile * kTile;
  const int r0 = (b & 1) * kSub, c0 = (a & 1) * kSub;
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int e = blockIdx.y * 768 + u * 256 + threadIdx.x;
    const int r = e / kSub, c = e % kSub;
    double sum = 0.0;
    if (a != b || (c >> 3) <= (r >> 3))     // diagonal pairs: the SYRK leaves the strictly upper 8x8 tiles unwritten
      for (int it = ib; it < ie; ++it) sum += ne.partial[(long)it * kSub * kSub + e];
    double val = -sum;
    const int fr = b * kSubFrames + r / kFrameParams, fc = a * kSubFrames + c / kFrameParams;
    if (fr == fc) {
      if ((long)fr * kFrameParams < st.n_cam_params)
        val += ne.B[(long)fr * 144 + (r % kFrameParams) * 12 + c % kFrameParams];
    } else if (fc == cam_frame || fr == cam_frame) {
      // uncalibrated variant: coupling of a real frame with the intrinsics pseudo-frame
      const int rp = r % kFrameParams, cp = c % kFrameParams;
      // (the pseudo-frame's sub-tile can hold padding slots behind it: no frame, no coupling)
      if (fr == cam_frame) { if (fc < cam_frame) val += ne.Bcam[(long)fc * 144 + cp * 12 + rp]; }   // rows = intrinsics, cols = frame
      else if (fr < cam_frame) val += ne.Bcam[(long)fr * 144 + rp * 12 + cp];
    } else if (pv.n > 0 && (r % 6) == (c % 6) && (long)fr * kFrameParams < st.n_cam_params &&
               (long)fc * kFrameParams < st.n_cam_params) {
      // motion-prior coupling between a frame and its previous frame: 6x6 diagonal blocks
      const int rb = (r % kFrameParams) / 6, cb = (c % kFrameParams) / 6;
      const int pa = pv.cur_of[fr], pc = pv.cur_of[fc];
      if (pa >= 0 && pv.prev[pa] == fc) val += pv.Bx[24L * fr + 6 * (2 * rb + cb) + r % 6];
      else if (pc >= 0 && pv.prev[pc] == fr) val += pv.Bx[24L * fc + 6 * (2 * cb + rb) + r % 6];
    }
    if (transposed) tile[(c0 + c) * kTile + r0 + r] = val;
    else            tile[(r0 + r) * kTile + c0 + c] = val;
  }
}

// ---------------------------------------------------------------- finalize (after the all-reduce)
// grid (n_nz tiles, 9).  S' = s S s + D^2 with D^2 = clamp(s^2 diag(B)) / radius on the diagonal;
// constant parameters and the padding rows of the last frame tile become identity rows.
__global__ void __launch_bounds__(256)
schur_finalize_kernel(SchurStructure st, NormalEq ne, LmOptionsDev o, double* __restrict__ S,
                      const int2* __restrict__ nz_tiles) {
  const int2 t = nz_tiles[blockIdx.x];
  double* tile = S + (long)blockIdx.x * kTile * kTile;
  const long base_r = (long)st.pos_tile[t.x] * kTile, base_c = (long)st.pos_tile[t.y] * kTile;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int e = blockIdx.y * 1024 + u * 256 + threadIdx.x;
    const int r = e / kTile, c = e % kTile;
    const long gi = base_r + r, gj = base_c + c;
    bool fixed = gi >= st.n_cam_params || gj >= st.n_cam_params;
    if (!fixed) {
      const unsigned mi = ne.pose_mask[gi / kFrameParams], mj = ne.pose_mask[gj / kFrameParams];
      fixed = ((mi >> (gi % kFrameParams)) & 1) || ((mj >> (gj % kFrameParams)) & 1);
    }
    double val;
    if (fixed) {
      val = (gi == gj) ? 1.0 : 0.0;
    } else {
      const double si = ne.scale_c[gi], sj = ne.scale_c[gj];
      val = si * tile[e] * sj;
      if (gi == gj) val += fmin(fmax(si * ne.diagB[gi] * si, o.min_diag), o.max_diag) / o.radius;
    }
    tile[e] = val;
  }
}

// LM diagonal of the camera parameters and the right-hand side of  S y = rhs  (y = -scaled step),
// the latter in the permuted order of the reduced system.
__global__ void __launch_bounds__(256)
camera_rhs_kernel(SchurStructure st, NormalEq ne, LmOptionsDev o, int n_frames, double* __restrict__ rhs) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_frames * kFrameParams) return;
  const int f = t / kFrameParams, k = t % kFrameParams;
  const bool cst = (ne.pose_mask[f] >> k) & 1;
  const double s = ne.scale_c[t];
  ne.d2_c[t] = cst ? 1.0 : fmin(fmax(s * ne.diagB[t] * s, o.min_diag), o.max_diag) / o.radius;
  const long pos = (long)st.tile_pos[f / kFramesPerTile] * kTile + (f % kFramesPerTile) * kFrameParams + k;
  rhs[pos] = cst ? 0.0 : s * (ne.gc[t] - ne.wf[t]);
}

}  // namespace

void launch_phi_build(const SchurStructure& st, const ObsView& obs, const JacView& jv, NormalEq ne,
                      cudaStream_t s) {
  if (st.n_dup <= 0) return;   // the common case: frame_blocks_kernel has written every panel
  const long n = (long)st.n_dup * kSubFrames;
  phi_build_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(st, obs, jv, ne);
}

// Pipeline shape: 8-point chunks, a 2-stage ring, 5 CTAs per SM.  Measured at C3 (profiles/r02_notes.md): <8,3,3> 2.16 ms
// (round 1's shape), <8,2,5> 1.85, <4,3,5> 1.99, <4,4,5> 1.97, <4,2,6> 1.91, <4,2,7> 1.92, <4,3,6> 1.93: the number of live warps
// per SM sub-partition, not the ring depth, is what keeps the FP64 pipe fed.
constexpr int kSyrkChunk = 8, kSyrkStages = 2, kSyrkCtasPerSm = 5;

void launch_schur_syrk(const SchurStructure& st, NormalEq ne, cudaStream_t s) {
  if (st.n_items <= 0) return;
  static bool seen[64] = {};
  constexpr size_t smem = syrk_smem<kSyrkChunk, kSyrkStages>();
  auto kernel = schur_syrk_kernel<kSyrkChunk, kSyrkStages, kSyrkCtasPerSm>;
  if (first_use_on_device(seen)) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kernel<<<st.n_items, 128, smem, s>>>(ne.Phi, st.entries, st.items, ne.partial);
}

void launch_schur_reduce(const SchurStructure& st, NormalEq ne, const PriorView& pv, int cam_frame, double* S,
                         const int* tile_slot, int n_tiles, cudaStream_t s) {
  if (st.n_pairs > 0)
    schur_reduce_kernel<<<dim3(st.n_pairs, 3), 256, 0, s>>>(st, ne, pv, cam_frame, S, tile_slot, n_tiles);
}

void launch_schur_finalize(const SchurStructure& st, NormalEq ne, LmOptionsDev o, double* S,
                           const TileSchedule& ts, int n_frames, double* rhs, cudaStream_t s) {
  if (ts.n_nz > 0) schur_finalize_kernel<<<dim3(ts.n_nz, 9), 256, 0, s>>>(st, ne, o, S, ts.nz_tiles);
  if (n_frames > 0)
    camera_rhs_kernel<<<(n_frames * kFrameParams + 255) / 256, 256, 0, s>>>(st, ne, o, n_frames, rhs);
}

}  // namespace rsba
