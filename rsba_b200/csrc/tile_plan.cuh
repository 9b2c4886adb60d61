// Host-side symbolic analysis of the reduced camera system at tile granularity: fill-reducing,
// parallelism-exposing ordering (nested dissection by BFS level sets), symbolic Cholesky on
// the tile graph, elimination levels and conflict-free update groups.  This is the analogue of
// what CHOLMOD's analyse phase does for Ceres' SparseSchurComplementSolver (third-party, reached
// through ceres::Solve with SPARSE_SCHUR, CeresHandler.h:403,419); it runs once per scene.
#pragma once
#include <utility>
#include <vector>

#include <cuda_runtime.h>

namespace rsba {

struct TilePlan {
  int T = 0;                          // tiles per dimension
  std::vector<int> tile_pos;          // frame tile -> position in the reduced system
  std::vector<int> pos_tile;          // inverse
  // structurally non-zero lower tiles after fill (positions): slot = tile_slot[i*T + j], i >= j
  std::vector<int2> nz_tiles;
  std::vector<int> tile_slot;
  std::vector<int> row_ptr, rows;     // per panel k: tiles (i,k), i > k
  std::vector<int> lrow_ptr, lrow_cols;  // per tile row i: tiles (i,j), j < i
  // elimination levels: panels of one level are independent
  int n_levels = 0;
  std::vector<int> panels, panel_ptr;    // panels sorted by level; panel_ptr[n_levels+1]
  std::vector<int2> trsm;                // (row tile i, panel k), sorted by level of k
  std::vector<int> trsm_ptr;             // [n_levels+1]
  // trailing updates (i, j, k): S(i,j) -= L(i,k) L(j,k)^T, grouped so that no two updates of one
  // group write the same tile (groups of a level run back to back: deterministic summation)
  std::vector<int4> upd;
  std::vector<long> group_ptr;           // [n_groups+1]
  std::vector<int> level_group_ptr;      // [n_levels+1] groups of each level
  double flops = 0.0;                    // numeric factorisation work at tile granularity
};

// tile_pairs: structurally non-zero tile pairs (A <= B) in FRAME-tile numbering (diagonal pairs may
// be omitted).  dense: treat every tile as non-zero.  reorder: nested dissection (else natural order).
// border_tile >= 0: a tile that couples with every other one (the intrinsics pseudo-frame of the uncalibrated
// variant); it is kept out of the dissection and eliminated last.
void build_tile_plan(int T, const std::vector<std::pair<int, int>>& tile_pairs, bool dense, bool reorder,
                     int border_tile, TilePlan* plan);

}  // namespace rsba
