// One-off host analysis of a scene's structure: the point-major view of the observations, the
// (sub-tile, point) incidences and sub-tile pair / work-item lists of the Schur SYRK, and the tile
// plan of the reduced camera system.  This is the analogue of Ceres' program reordering and
// SchurEliminator block-structure detection plus CHOLMOD's analyse phase (third-party, reached
// through ceres::Solve with SPARSE_SCHUR, CeresHandler.h:403,419).  Pure host code: it touches no
// device state, so `rsba_cuda_analyze_structure` (include/rsba_cuda.h) and the CPU tests can run it
// without a GPU.
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "host_threads.h"
#include "lm.cuh"

namespace rsba {

// What the analysis reads (all host pointers; observations sorted by frame).
struct SceneTopology {
  long n_obs = 0;                 // this rank's share (== the whole scene on one GPU)
  int n_frames = 0, n_points = 0;
  const int* obs_frame = nullptr;
  const int* obs_point = nullptr;
  const unsigned char* point_const = nullptr;   // [points]
  bool free_cam = false;          // intrinsics are parameters 0..8 of a pseudo-frame behind the real frames
  bool free_ratio = false;        // interFrameRatio is parameter 9 of it (only with motion priors)
  std::vector<std::pair<int, int>> prior_pairs; // (frame, previous frame) of every motion prior
  // multi-GPU: the tile plan is a function of the WHOLE scene so that every rank derives the same one
  int world = 1;
  long n_obs_global = 0;
  const int* g_obs_frame = nullptr;
  const int* g_obs_point = nullptr;
  bool dense = false, reorder = true, sparse_keys = false;
  int n_cam_frames() const { return n_frames + ((free_cam || free_ratio) ? 1 : 0); }
};

struct HostStructure {
  std::vector<int> pt_ptr;                                           // point-major CSR
  HostVec<int> pt_obs;
  std::vector<int> chunk_frame, chunk_beg, chunk_cnt, frame_chunk_ptr;   // frame chunks of <= 128 observations
  int T = 0;                                                         // Cholesky tiles per dimension
  int n_inc = 0;
  HostVec<int> inc_point, inc_tile, slot_beg;
  std::vector<int> pt_inc_ptr, cam_inc;
  HostVec<unsigned char> slot_cnt;
  std::vector<unsigned char> inc_half;
  HostVec<int> obs_phi_off;
  std::vector<int> dup_inc;
  std::vector<int> pair_a, pair_b, pair_item_ptr;
  std::vector<int4> items;
  HostVec<int2> entries;
  int n_items = 0;
  // groups of whole points for the thread-per-observation back-substitution (lm.cuh, PointGroups): (first point, one past
  // the last), <= kPointGroupObs observations and <= kPointGroupPoints points each; point_big = the points with more
  // observations than one group holds (they go through the warp-per-point kernel)
  std::vector<int2> point_groups;
  std::vector<int> point_big;
  TilePlan plan;
  std::vector<int> fwd_slot;   // where each trsm tile (i, k) leaves its forward-substitution term in row i's list
};

// `lap(what)` is called after each phase (RSBA_CUDA_TRACE timing); may be empty.  Returns an RSBA_* code.
int analyze_structure(const SceneTopology& sc, HostStructure* out, std::string* error,
                      void (*lap)(const char* what, void* ctx) = nullptr, void* lap_ctx = nullptr);

}  // namespace rsba
