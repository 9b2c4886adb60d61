// K2s, second form -- the Schur complement SYRK regrouped by CHOLESKY-TILE pairs.
//
// schur_syrk_kernel (k2_schur.cu) gives one CTA one pair of 4-frame sub-tiles: every entry loads two 1 248-byte
// panels for one 48 x 48 product, 5.5 flop per byte of L2 -> shared-memory fill, and the kernel sits on that fill
// (40 % of its stall samples wait for TMA data, profiles/r02f_full.txt) as much as on the FP64 pipe.  Here a CTA
// owns a pair of 8-frame Cholesky tiles (A <= B): a point seen from both contributes up to FOUR panels -- sub-tiles
// 2B, 2B+1 on the row side, 2A, 2A+1 on the column side -- and warp (i, j) multiplies row panel i by column panel j,
// i.e. the four sub-tile pairs of the tile pair are computed from one copy of the panels: 11 flop per loaded byte,
// 36 DMMA per 12 fragment loads instead of 9 per 6.  Executed flops do not grow: the entries of a tile pair are
// sorted by class = which 2-frame halves of the 8 + 8 frames are populated, classes are padded to whole 4-point
// chunks, and per chunk a warp runs the variant of its 48 x 48 product that touches the populated 24 x 24 patches
// only (warp-uniform switch around fully unrolled bodies); absent panels are not loaded at all.
// Diagonal tile pairs (A == B) need sub-tile pairs (0,0), (1,1) -- lower tiles only -- and (0,1): warps 0 and 3 take
// the two symmetric ones (21 DMMA tiles each), warps 2 and 1 the upper and lower half of the off-diagonal one (18
// each).
// Work item = <= 512 entries of one tile pair; its (up to) four 48 x 48 partial products go to
// partial2[slot][quadrant], summed per sub-tile pair in slot order by schur_reduce2_kernel (bit-reproducible).
// Supersedes, like k2_schur.cu, the reduction loop of Ceres' SchurEliminator::Eliminate (third-party, reached
// through ceres::Solve with SPARSE_SCHUR, CeresHandler.h:403,419).
#include "lm.cuh"
#include "structure.cuh"

#include <algorithm>
#include <type_traits>

namespace rsba {
namespace {

constexpr int kC2 = kSyrk2Chunk;                              // 4 points = 12 K rows = 3 k-steps per stage
constexpr int kStages2 = 5;
constexpr int kOperand2 = kC2 * kPanelDoubles;                // 624 doubles: one side's sub-tile, 4 points
constexpr int kStage2Doubles = 4 * kOperand2;                 // B0 | B1 | A0 | A1
constexpr unsigned kPanelBytes2 = kPanelDoubles * sizeof(double);
constexpr size_t kSyrk2Smem = (size_t)kStages2 * kStage2Doubles * sizeof(double) + 128;

__device__ __forceinline__ unsigned s2_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void s2_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void s2_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s2_mbar_wait(unsigned bar, unsigned parity) {
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
__device__ __forceinline__ void s2_tma_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void s2_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- a warp's share of one chunk (3 k-steps).  Eight warps; every role owns <= 18 of the 8 x 8 DMMA tiles of one
// 48 x 48 sub-tile-pair product, held in acc[18][2]; the role -> tile -> accumulator maps are compile-time.
//   Rect<M0, CH>   tile rows M0 .. M0+2 (one 2-frame half of the row panel) x the column halves in CH (bit 0:
//                  tile columns 0-2, bit 1: 3-5); accumulator (mi - M0) * 6 + ni
//   Quarter<M0, N0> tile rows M0 .. M0+2 x tile columns N0 .. N0+2; accumulator (mi - M0) * 3 + (ni - N0)
//   LowerA         rows 0-3 of a symmetric product, columns <= row (10 tiles); accumulator mi (mi + 1) / 2 + ni
//   LowerB         rows 4-5, columns <= row (11 tiles); accumulator (mi - 4) * 5 + ni
template <int M0, int CH>
__device__ __forceinline__ void chunk_rect(double (&acc)[18][2], const double* __restrict__ Rp,
                                           const double* __restrict__ Cp, const int (&koff)[3]) {
#pragma unroll
  for (int ks = 0; ks < 3; ++ks) {
    double a[3], b[6];
#pragma unroll
    for (int mi = 0; mi < 3; ++mi) a[mi] = Rp[koff[ks] + 8 * (M0 + mi)];
#pragma unroll
    for (int ni = 0; ni < 6; ++ni)
      if ((CH >> (ni / 3)) & 1) b[ni] = Cp[koff[ks] + 8 * ni];
#pragma unroll
    for (int mi = 0; mi < 3; ++mi)
#pragma unroll
      for (int ni = 0; ni < 6; ++ni)
        if ((CH >> (ni / 3)) & 1) s2_dmma(acc[mi * 6 + ni][0], acc[mi * 6 + ni][1], a[mi], b[ni]);
  }
}

template <int M0, int N0>
__device__ __forceinline__ void chunk_quarter(double (&acc)[18][2], const double* __restrict__ Rp,
                                              const double* __restrict__ Cp, const int (&koff)[3]) {
#pragma unroll
  for (int ks = 0; ks < 3; ++ks) {
    double a[3], b[3];
#pragma unroll
    for (int mi = 0; mi < 3; ++mi) a[mi] = Rp[koff[ks] + 8 * (M0 + mi)];
#pragma unroll
    for (int ni = 0; ni < 3; ++ni) b[ni] = Cp[koff[ks] + 8 * (N0 + ni)];
#pragma unroll
    for (int mi = 0; mi < 3; ++mi)
#pragma unroll
      for (int ni = 0; ni < 3; ++ni) s2_dmma(acc[mi * 3 + ni][0], acc[mi * 3 + ni][1], a[mi], b[ni]);
  }
}

template <bool B>   // false: rows 0-3, true: rows 4-5
__device__ __forceinline__ void chunk_lower(double (&acc)[18][2], const double* __restrict__ P, const int (&koff)[3]) {
  constexpr int M0 = B ? 4 : 0, M1 = B ? 6 : 4;
#pragma unroll
  for (int ks = 0; ks < 3; ++ks) {
    double v[6];
#pragma unroll
    for (int t = 0; t < M1; ++t) v[t] = P[koff[ks] + 8 * t];     // rows and columns come from the same panel
#pragma unroll
    for (int mi = M0; mi < M1; ++mi)
#pragma unroll
      for (int ni = 0; ni <= mi; ++ni) {
        const int k = B ? (mi - 4) * 5 + ni : mi * (mi + 1) / 2 + ni;
        s2_dmma(acc[k][0], acc[k][1], v[mi], v[ni]);
      }
  }
}

constexpr int kSyrk2Warps = 8;

__global__ void __launch_bounds__(kSyrk2Warps * 32, 2)
schur_syrk2_kernel(const double* __restrict__ Phi, const int4* __restrict__ entries, const unsigned char* __restrict__ chunk_mask,
                   const int4* __restrict__ items, double* __restrict__ partial, int zero_panel) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stage_base = reinterpret_cast<double*>(smem_raw);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)kStages2 * kStage2Doubles * sizeof(double));
  int* s_mask = reinterpret_cast<int*>(bars + 2 * kStages2);     // class of the chunk in each stage

  const int4 item = items[blockIdx.x];
  const int slot = item.x, beg = item.y, nchunks = item.z / kC2;
  const bool diag = item.w != 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages2; ++s) {
      s2_mbar_init(s2_u32(&bars[s]), 1);                         // "full": the stage's TMA bytes have landed
      s2_mbar_init(s2_u32(&bars[kStages2 + s]), kSyrk2Warps);    // "empty": every warp is done reading the stage
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // producer: warp 0.  Lane o*4 + q (o = operand B0 B1 A0 A1, q = point of the chunk) fetches one panel if the
  // chunk's class says the operand's sub-tile is populated (diagonal tile pairs: the column side IS the row side).
  int inc_next = zero_panel, mask_next = 0;
  auto load_entry = [&](int c) {
    if (c < nchunks) {
      mask_next = chunk_mask[beg / kC2 + c];
      if (lane < 16) {
        const int4 e = entries[beg + c * kC2 + (lane & 3)];
        const int o = lane >> 2;
        inc_next = o == 0 ? e.x : (o == 1 ? e.y : (o == 2 ? e.z : e.w));
      }
    }
  };
  auto issue = [&](int c) {
    const int s = c % kStages2;
    const unsigned bar = s2_u32(&bars[s]);
    const int cm = mask_next, inc = inc_next;
    const int o = lane >> 2, q = lane & 3;
    // populated halves of operand o: row side = high nibble, column side = low nibble
    const int h = o < 2 ? (cm >> (4 + 2 * o)) & 3 : (cm >> (2 * (o & 1))) & 3;
    const bool want = lane < (diag ? 8 : 16) && h != 0;
    const unsigned n_loads = __popc(__ballot_sync(0xffffffffu, want));
    if (lane == 0) {
      s_mask[s] = cm;
      s2_mbar_expect_tx(bar, n_loads * kPanelBytes2);
    }
    __syncwarp();
    load_entry(c + 1);
    if (want) {
      double* dst = stage_base + (size_t)s * kStage2Doubles + o * kOperand2 + q * kPanelDoubles;
      s2_tma_load(s2_u32(dst), Phi + (long)inc * kPanelDoubles, kPanelBytes2, bar);
    }
  };
  if (warp == 0) {
    load_entry(0);
    for (int c = 0; c < kStages2 && c < nchunks; ++c) issue(c);
  }
  auto refill = [&](int c) {          // warp 0, at the top of chunk c: the stage of chunk c - 1 is refilled
    const int prev = c - 1;
    if (prev >= 0 && prev + kStages2 < nchunks) {
      s2_mbar_wait(s2_u32(&bars[kStages2 + prev % kStages2]), (unsigned)((prev / kStages2) & 1));
      issue(prev + kStages2);
    }
  };
  auto release = [&](int s) {
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s2_u32(&bars[kStages2 + s])) : "memory");
  };

  // fragment addressing: K row 4 ks + fc of the chunk = point (4 ks + fc) / 3, k = (4 ks + fc) % 3
  const int fr = lane >> 2, fc = lane & 3;
  int koff[3];
#pragma unroll
  for (int ks = 0; ks < 3; ++ks) {
    const int kr = 4 * ks + fc;
    koff[ks] = (kr / 3) * kPanelDoubles + (kr % 3) * kPanelLd + fr;
  }
  double acc[18][2];
#pragma unroll
  for (int k = 0; k < 18; ++k) acc[k][0] = acc[k][1] = 0.0;

  // roles.  Off-diagonal tile pair: warp = sub-tile pair * 2 + row half; sub-tile pair = (row sub-tile bi) * 2 +
  // (column sub-tile aj).  Diagonal: warps 0/1 the symmetric product of sub-tile 0 (rows 0-3 / 4-5 of its lower
  // tiles), 2/3 that of sub-tile 1, 4-7 the four 24 x 24 quarters of (rows sub-tile 1) x (columns sub-tile 0).
  // A track that starts inside tile A and ends inside tile B populates the LAST halves of A and the FIRST of B: the
  // role (B0, A1, row half 0) carries 18 % of all DMMAs, (B1, A0, row half 1) 7 %, and with a fixed warp -> role map
  // one SM sub-partition would get 32 % of the work.  Warps w and w ^ x share nothing but the sub-partition w % 4,
  // so consecutive CTAs rotate the two low role bits.
  const int role = diag ? warp : (warp ^ (int)(blockIdx.x & 3u));
  const int sp = role >> 1, rh = role & 1;
  for (int c = 0; c < nchunks; ++c) {
    const int s = c % kStages2;
    if (warp == 0) refill(c);
    s2_mbar_wait(s2_u32(&bars[s]), (unsigned)((c / kStages2) & 1));
    const int cm = s_mask[s];
    const double* st = stage_base + (size_t)s * kStage2Doubles;
    if (!diag) {
      const int bi = sp >> 1, aj = sp & 1;
      const int hb = (cm >> (4 + 2 * bi)) & 3, ha = (cm >> (2 * aj)) & 3;
      const double* Rp = st + bi * kOperand2;
      const double* Cp = st + (2 + aj) * kOperand2;
      if ((hb >> rh) & 1) {
        switch (rh * 4 + ha) {
          case 1: chunk_rect<0, 1>(acc, Rp, Cp, koff); break;
          case 2: chunk_rect<0, 2>(acc, Rp, Cp, koff); break;
          case 3: chunk_rect<0, 3>(acc, Rp, Cp, koff); break;
          case 5: chunk_rect<3, 1>(acc, Rp, Cp, koff); break;
          case 6: chunk_rect<3, 2>(acc, Rp, Cp, koff); break;
          case 7: chunk_rect<3, 3>(acc, Rp, Cp, koff); break;
          default: break;
        }
      }
    } else {
      const int h0 = (cm >> 4) & 3, h1 = (cm >> 6) & 3;     // halves of sub-tile 0 / 1 (row nibble == column nibble)
      const double* P0 = st;
      const double* P1 = st + kOperand2;
      switch (warp) {
        case 0: if (h0) chunk_lower<false>(acc, P0, koff); break;
        case 1: if (h0) chunk_lower<true>(acc, P0, koff); break;
        case 2: if (h1) chunk_lower<false>(acc, P1, koff); break;
        case 3: if (h1) chunk_lower<true>(acc, P1, koff); break;
        case 4: if ((h1 & 1) && (h0 & 1)) chunk_quarter<0, 0>(acc, P1, P0, koff); break;
        case 5: if ((h1 & 1) && (h0 & 2)) chunk_quarter<0, 3>(acc, P1, P0, koff); break;
        case 6: if ((h1 & 2) && (h0 & 1)) chunk_quarter<3, 0>(acc, P1, P0, koff); break;
        default: if ((h1 & 2) && (h0 & 2)) chunk_quarter<3, 3>(acc, P1, P0, koff); break;
      }
    }
    release(s);
  }

  // quadrant = (row sub-tile) * 2 + (column sub-tile) of the tile pair's four sub-tile pairs
  auto put = [&](double* out, int mi, int ni, int k) {
    *reinterpret_cast<double2*>(out + (8 * mi + fr) * kSub + 8 * ni + 2 * fc) = make_double2(acc[k][0], acc[k][1]);
  };
  if (!diag) {
    double* out = partial + ((long)slot * 4 + sp) * kSub * kSub;
#pragma unroll
    for (int mi = 0; mi < 3; ++mi)
#pragma unroll
      for (int ni = 0; ni < 6; ++ni) put(out, 3 * rh + mi, ni, mi * 6 + ni);
  } else if (warp < 4) {
    double* out = partial + ((long)slot * 4 + (warp < 2 ? 0 : 3)) * kSub * kSub;
    if ((warp & 1) == 0) {
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni <= mi; ++ni) put(out, mi, ni, mi * (mi + 1) / 2 + ni);
    } else {
#pragma unroll
      for (int mi = 4; mi < 6; ++mi)
#pragma unroll
        for (int ni = 0; ni <= mi; ++ni) put(out, mi, ni, (mi - 4) * 5 + ni);
    }
  } else {
    double* out = partial + ((long)slot * 4 + 2) * kSub * kSub;
    const int m0 = 3 * ((warp >> 1) & 1), n0 = 3 * (warp & 1);
#pragma unroll
    for (int mi = 0; mi < 3; ++mi)
#pragma unroll
      for (int ni = 0; ni < 3; ++ni) put(out, m0 + mi, n0 + ni, mi * 3 + ni);
  }
}

// grid (n sub-tile pairs, 3): as schur_reduce_kernel, reading the quadrant partials of the pair's tile pair
__global__ void __launch_bounds__(256)
schur_reduce2_kernel(SchurStructure st, Syrk2View sv, NormalEq ne, PriorView pv, int cam_frame, double* __restrict__ S,
                     const int* __restrict__ tile_slot, int T) {
  const int pr = blockIdx.x;
  const int a = st.pair_a[pr], b = st.pair_b[pr];
  const int tp = sv.pair_tp[pr];
  const int ib = sv.tp_item_ptr[tp], ie = sv.tp_item_ptr[tp + 1];
  const int quad = (b & 1) * 2 + (a & 1);
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
      for (int it = ib; it < ie; ++it) sum += sv.partial[((long)it * 4 + quad) * kSub * kSub + e];
    double val = -sum;
    const int fr = b * kSubFrames + r / kFrameParams, fc = a * kSubFrames + c / kFrameParams;
    if (fr == fc) {
      if ((long)fr * kFrameParams < st.n_cam_params)
        val += ne.B[(long)fr * 144 + (r % kFrameParams) * 12 + c % kFrameParams];
    } else if (fc == cam_frame || fr == cam_frame) {
      const int rp = r % kFrameParams, cp = c % kFrameParams;
      if (fr == cam_frame) { if (fc < cam_frame) val += ne.Bcam[(long)fc * 144 + cp * 12 + rp]; }
      else if (fr < cam_frame) val += ne.Bcam[(long)fr * 144 + rp * 12 + cp];
    } else if (pv.n > 0 && (r % 6) == (c % 6) && (long)fr * kFrameParams < st.n_cam_params &&
               (long)fc * kFrameParams < st.n_cam_params) {
      const int rb = (r % kFrameParams) / 6, cb = (c % kFrameParams) / 6;
      const int pa2 = pv.cur_of[fr], pc = pv.cur_of[fc];
      if (pa2 >= 0 && pv.prev[pa2] == fc) val += pv.Bx[24L * fr + 6 * (2 * rb + cb) + r % 6];
      else if (pc >= 0 && pv.prev[pc] == fr) val += pv.Bx[24L * fc + 6 * (2 * cb + rb) + r % 6];
    }
    if (transposed) tile[(c0 + c) * kTile + r0 + r] = val;
    else            tile[(r0 + r) * kTile + c0 + c] = val;
  }
}

}  // namespace

void launch_schur_syrk2(const SchurStructure& st, const Syrk2View& sv, NormalEq ne, cudaStream_t s) {
  static bool seen[64] = {};
  if (first_use_on_device(seen))
    cudaFuncSetAttribute(schur_syrk2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSyrk2Smem);
  if (sv.n_items <= 0) return;
  schur_syrk2_kernel<<<sv.n_items, kSyrk2Warps * 32, kSyrk2Smem, s>>>(ne.Phi, sv.entries, sv.chunk_mask, sv.items, sv.partial, st.n_inc);
}

void launch_schur_reduce2(const SchurStructure& st, const Syrk2View& sv, NormalEq ne, const PriorView& pv, int cam_frame,
                          double* S, const int* tile_slot, int n_tiles, cudaStream_t s) {
  if (st.n_pairs > 0)
    schur_reduce2_kernel<<<dim3(st.n_pairs, 3), 256, 0, s>>>(st, sv, ne, pv, cam_frame, S, tile_slot, n_tiles);
}

}  // namespace rsba
