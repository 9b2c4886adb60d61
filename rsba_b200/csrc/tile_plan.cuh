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

namespace rsba {

// ---- the numeric factorisation and the triangular solves as ONE static task graph (k3_dag.cu) ----
// The level-batched launch sequence above costs one launch boundary and one straggler wait per kernel
// and level (~70 launches for the factorisation and ~30 for the solves of a video scene) and keeps
// every update of level l in front of the panels of level l+1, needed or not.  The same work as a task
// list for ONE persistent kernel: CTAs fetch tasks in list order and wait on device-side counters for
// what a task depends on.  The list is a topological order of the dependency graph (every task only
// waits for tasks in front of it), so in-order fetching cannot deadlock whatever the number of CTAs.
enum DagTaskType : int {
  kTaskFactor = 0,   // a = panel k, b = slot(k,k)
  kTaskTrsm = 1,     // a = row tile i, b = panel k, c = part (rows 32c..32c+31), d = slot(i,k), e = slot(k,k),
                     // f = forward-substitution slot (position of k in row i's list)
  kTaskUpdate = 2,   // a = target slot(i,j), b = quadrant 2 qi + qj (48 x 48), c = how many groups of this
                     // target run before this one, d = first source in `sources`, e = number of sources
  kTaskBackFin = 3,  // a = panel k, b = slot(k,k), c = first backward slot of column k, d = tiles below k
  kTaskBackTile = 4, // a = row tile i, b = panel k, c = slot(i,k), d = backward slot
};
struct DagTask { int type, a, b, c, d, e, f, g; };

constexpr int kTrsmParts = 3;       // a tile's triangular solve is cut into 3 slabs of 32 rows

struct DagPlan {
  std::vector<DagTask> tasks;       // topological order; the factorisation first, then the backward substitution
  std::vector<int2> sources;        // per update task: (slot(i,k), slot(j,k)) of S(i,j) -= L(i,k) L(j,k)^T
  std::vector<int> need;            // [n_nz * 4] update groups that must have run on quadrant q of a tile
  int n_factor_tasks = 0;           // tasks[0 .. n_factor_tasks) factorise (and substitute forward)
};

// merge_levels: sources of one target whose elimination levels lie more than one level below the
// target's column are applied `merge_levels` levels at a time (one read-modify-write of the target for
// all of them); the sources of the level right below the target's column always run on their own.
void build_dag_plan(const TilePlan& plan, int merge_levels, DagPlan* dag);

}  // namespace rsba
