// Device-side state of the LM solver and the launchers of K2 (normal equations + Schur
// complement), K3 (tile Cholesky of the reduced camera system) and K4 (back-substitution and
// step bookkeeping).  Everything here supersedes third-party Ceres internals that the
// reference reaches through ceres::Solve (CeresHandler.h:419): SchurEliminator,
// SparseSchurComplementSolver/CHOLMOD, LevenbergMarquardtStrategy, TrustRegionMinimizer.
#pragma once
#include "common.cuh"
#include "tile_plan.cuh"

namespace rsba {

// The Jacobian as K1 leaves it for the solver: compact records (rsba_reproj_math.h) in observation order.
struct JacView {
  const double* rec;     // [N][kJacCompact]  jx row 0 | jx row 1 | jr row 0 | jr row 1
  const double* tau;     // [N]
  int rot_interp;        // shutter != GLOBAL && interpolateRotation: rotation columns weigh (1-tau, tau), else (1, 0)
  int point_major;       // records / tau are indexed by the position e in the point-major list (k2_fused.cu), not by observation
};
// the full 30-double record of observation i (kernels off the hot path)
__device__ __forceinline__ void load_full_jacobian(const JacView& jv, long i, double* __restrict__ J) {
  double rec[kJacCompact];
  const double2* src = reinterpret_cast<const double2*>(jv.rec + i * kJacCompact);
#pragma unroll
  for (int k = 0; k < kJacCompact / 2; ++k) { const double2 v = src[k]; rec[2 * k] = v.x; rec[2 * k + 1] = v.y; }
  expand_jacobian(rec, jv.tau[i], jv.rot_interp != 0, J);
}

constexpr int kTile = 96;          // Cholesky tile = 8 frames x 12 parameters
constexpr int kFramesPerTile = kTile / kFrameParams;

// ---- structure (built once per scene on the host, lm_structure.cu) ----------------------
struct SchurStructure {
  // point-major CSR over the frame-sorted observations
  const int* pt_ptr;      // [P+1]
  const int* pt_obs;      // [N]
  const double* pt_tau;   // [N] tau of observation pt_obs[e] (point-major copy: coalesced for the point kernels)
  const int* pt_frame;    // [N] its frame
  // frame chunks for the per-frame reductions: chunk c covers obs [chunk_beg[c], chunk_beg[c]+chunk_cnt[c]) of chunk_frame[c]
  const int* chunk_frame; // [n_chunks]
  const int* chunk_beg;   // [n_chunks]
  const int* chunk_cnt;   // [n_chunks]
  const int* frame_chunk_ptr;  // [F+1] chunks of frame f
  int n_chunks;
  // ---- Schur complement as a tile-level SYRK  S -= Phi Phi^T  (k2_schur.cu)
  // The SYRK works on SUB-TILES of kSubFrames = 4 frames (48 rows, a quadrant of a Cholesky tile):
  // incidence = (sub-tile h, point p) with at least one observation of p in frames 4h..4h+3;
  // its panel Phi[inc] is [3][kPanelLd] doubles (k-major, 48 rows = 4 frame slots x 12).
  int n_inc;
  const int* inc_point;    // [n_inc]
  const int* inc_tile;     // [n_inc] sub-tile index h
  const int* slot_beg;     // [n_inc*4] first index into pt_obs of the observations in that slot, -1 if none
  const unsigned char* slot_cnt;  // [n_inc*4] number of observations in the slot (duplicates in one frame)
  // The panels are written by frame_blocks_kernel while it holds the Jacobian records in shared
  // memory: obs_phi_off[i] = offset (doubles) of observation i's 12 rows inside Phi, or -1 for
  // observations of constant points and of incidences with two observations in one frame slot;
  // the latter (dup_inc) are rebuilt by phi_build_kernel, which sums the slot's observations.
  const int* cam_inc;      // [P] uncalibrated variant: the point's incidence that holds the pseudo-frame rows, or -1
  const int* obs_phi_off;  // [N]
  const int* dup_inc;      // [n_dup]
  int n_dup;
  // sub-tile pair (a <= b in frame order): output block rows = frames of b, columns = frames of a,
  // i.e. quadrant (b % 2, a % 2) of Cholesky tile (b / 2, a / 2)
  int n_pairs;
  const int* pair_a;       // [n_pairs]
  const int* pair_b;       // [n_pairs]
  const int* pair_item_ptr;  // [n_pairs+1] work items (K segments) of the pair
  // work item = up to kSchurSegPoints common points of one pair; entries padded to a multiple of 8
  // with the all-zero panel (index n_inc)
  int n_items;
  const int4* items;       // [n_items] (pair, first entry, entry count, A == B)
  const int2* entries;     // (incidence on the row side B, incidence on the column side A)
  long n_entries;
  long n_cam_params;       // 12 * frames (rows of the last tile beyond it are padding)
  const int* tile_pos;     // [T] position of frame tile A in the (permuted) reduced system
  const int* pos_tile;     // [T] inverse
};

constexpr int kSubFrames = 4;                 // frames per SYRK sub-tile
constexpr int kSub = kSubFrames * kFrameParams;   // 48 rows
constexpr int kPanelLd = kSub + 4;            // 48 rows + 4 pad: conflict-free DMMA fragment loads (ld % 16 == 4)
constexpr int kPanelDoubles = 3 * kPanelLd;   // one bulk copy of 1248 bytes
constexpr int kSchurSegPoints = 512;
constexpr int kPointRec = 12;

struct NormalEq {
  // unscaled blocks of J^T J and J^T r
  double* B;        // [F][144] camera diagonal blocks (full 12x12, row-major)
  // gc | wf | diagB are contiguous and sit behind S in the buffer that the multi-GPU path
  // all-reduces (every rank holds partial sums over the observations of the points it owns)
  double* gc;       // [F][12]
  double* wf;       // [F][12]   sum_i Jc_i^T Jx_i t_p   (rhs correction)
  double* diagB;    // [F][12]   diagonal of B (Jacobi scaling and LM diagonal of the cameras)
  double* C;        // [P][6]    point blocks, packed xx xy xz yy yz zz
  double* gp;       // [P][3]
  double* Cinv;     // [P][6]    s_p (s_p C s_p + D^2)^-1 s_p  (zero for constant points)
  double* tp;       // [P][3]    Cinv * gp
  double* Bcam;     // [F][144] uncalibrated variant: coupling of frame f (rows) with the intrinsics pseudo-frame (cols 0..8)
  double* Minv;     // [P][6]    L^-1 of the damped scaled point block (m00 m10 m11 m20 m21 m22)
  double* prec;     // [P][kPointRec] what frame_blocks gathers per observation: W = s_p L^-T as (s0 m00, s0 m10, s1 m11,
                    //           s0 m20, s1 m21, s2 m22) | t_p (3) | pad -- three full 32-byte sectors
  double* Phi;      // [n_inc+1][3][kPanelLd]  panels s_c Jc^T (Jx s_p) L^-T; last panel all zero
  double* partial;  // [n_items][48*48] per-item partial products of the Schur SYRK
  double* scale_c;  // [12F] Jacobi scaling (1 for constant parameters)
  double* scale_p;  // [3P]
  double* d2_c;     // [12F] LM diagonal (scaled space) of the current solve
  double* d2_p;     // [3P]
  double* partials; // [n_chunks][104] per-chunk partial sums of the frame kernel
  const unsigned short* pose_mask;  // [F] constant-scalar bits
  const unsigned char* point_const; // [P]
  const unsigned char* point_owned; // [P] 1 if this rank eliminates the point (all 1 on one GPU)
  // the point kernels (blocks, scaling, inverse, back-substitution, norms) walk the points this rank OWNS:
  // point id = owned_ids[k], k < n_owned; owned_ids == NULL means the identity (one GPU: n_owned == P)
  const int* owned_ids;
  int n_owned;
};

__device__ __forceinline__ int owned_point(const NormalEq& ne, int k) { return ne.owned_ids ? ne.owned_ids[k] : k; }

struct LmOptionsDev {
  double radius, min_diag, max_diag;
};

// ---- K2 ---------------------------------------------------------------------------------
void launch_point_blocks(const SchurStructure& st, const JacView& jv, const double* res, NormalEq ne, cudaStream_t s);
void launch_frame_blocks(const SchurStructure& st, const ObsView& obs, const JacView& jv, const double* res,
                         int n_frames, NormalEq ne, cudaStream_t s);
// point-major copies of tau / frame (tau == NULL: frames only)
void launch_point_major_obs(const SchurStructure& st, const ObsView& obs, const double* tau, long n, double* pt_tau,
                            int* pt_frame, cudaStream_t s);
void launch_frame_reduce(const SchurStructure& st, NormalEq ne, int n_frames, cudaStream_t s);
// fused linearisation passes (k2_fused.cu): calibrated scenes
// `packed`: the observations' constants in point-major order ([N] records of point_pass_record_bytes() bytes,
// filled once per scene by launch_pack_point_major)
size_t point_pass_record_bytes();
void launch_pack_point_major(const SchurStructure& st, const ObsView& obs, long n, void* packed, cudaStream_t s);
// Groups of whole points for the thread-per-observation back-substitution (k4_update.cu): groups[g] = (first point, one
// past the last), <= 256 observations and <= 128 points each; big_ids = the points with more than 256 observations.
struct PointGroups {
  const int2* groups = nullptr;
  int n_groups = 0;
  const int* big_ids = nullptr;
  int n_big = 0;
};
constexpr int kPointGroupObs = 256, kPointGroupPoints = 128;
void launch_point_pass(const CameraModel& cm, const SchurStructure& st, const void* packed, const double* poses,
                       const double* points, NormalEq ne, LmOptionsDev o, bool compute_scale, bool jacobi,
                       double* rec_pt, double* tau_pt, double* xt, bool write_phi /* Schur panel rows */, long n_obs,
                       int n_points, cudaStream_t s);
void launch_frame_pass(const CameraModel& cm, const SchurStructure& st, const ObsView& obs, const double* poses,
                       const double* xt, NormalEq ne, double* cost_partials, int* invalid_count, cudaStream_t s);
// n_frames > 0: the camera parameters; points: the owned points (two calls: the point part runs before the
// all-reduce, the camera part after it)
void launch_jacobi_scale(int n_frames, bool points, NormalEq ne, bool enabled, cudaStream_t s);
void launch_point_invert(NormalEq ne, LmOptionsDev o, cudaStream_t s);
// uncalibrated variant (k2_cam.cu): blocks of the intrinsics pseudo-frame (frame index n_frames) and its panel rows
void launch_cam_blocks(const SchurStructure& st, const ObsView& obs, const JacView& jv, const double* jac_cam,
                       const double* res, NormalEq ne, int n_frames, double* partials, double* scratch,
                       cudaStream_t s);
void launch_phi_cam(const SchurStructure& st, const JacView& jv, const double* jac_cam, NormalEq ne, int n_points,
                    int n_frames, cudaStream_t s);

// adds the priors' J^T J / J^T r to B, gc, diagB and writes the frame-to-previous-frame couplings
void launch_prior_blocks(const PriorView& pv, NormalEq ne, int n_frames, cudaStream_t s);

// pose priors (k2_pose_priors.cu): inverse of every damped prior block; with `add` also the block's
// contribution to B, diag(B), gc, wf of its control pose, Schur term included (B += w^2 - w^4 cinv)
void launch_pose_prior_blocks(const PosePriorView& pv, NormalEq ne, LmOptionsDev o, bool jacobi, bool add,
                              cudaStream_t s);
// back-substitution of the prior blocks, trial values; adds their part to scalars[0..2]
void launch_pose_prior_step(const PosePriorView& pv, const double* delta_c, double* scalars, cudaStream_t s);
// adds |x|^2 of the free prior blocks to scalars[3], max|g| to scalars[4]
void launch_pose_prior_norms(const PosePriorView& pv, double* scalars, cudaStream_t s);

// Schur complement (k2_schur.cu).  S is tile-packed (see TileSchedule: structurally non-zero lower
// tiles, diagonal tiles stored as full squares); rhs/d2_c in the permuted order given by tile_pos.
void launch_phi_build(const SchurStructure& st, const ObsView& obs, const JacView& jv, NormalEq ne,
                      cudaStream_t s);
void launch_schur_syrk(const SchurStructure& st, NormalEq ne, cudaStream_t s);
// writes the UNSCALED tiles  B - Phi Phi^T  (partial sums on a multi-GPU rank)
// cam_frame: index of the intrinsics pseudo-frame (uncalibrated variant) or -1
void launch_schur_reduce(const SchurStructure& st, NormalEq ne, const PriorView& pv, int cam_frame, double* S,
                         const int* tile_slot, int n_tiles, cudaStream_t s);
// after the (optional) all-reduce: Jacobi scaling, LM diagonal, constant rows, padding; d2_c and rhs
struct TileSchedule;
void launch_schur_finalize(const SchurStructure& st, NormalEq ne, LmOptionsDev o, double* S,
                           const TileSchedule& ts, int n_frames, double* rhs, cudaStream_t s);

// ---- K3 ---------------------------------------------------------------------------------
struct TileSchedule {      // device view of the TilePlan (tile_plan.cuh); tile indices are POSITIONS
  int n_tiles;              // tiles per dimension (T)
  // S is tile-packed: slot s holds tile nz_tiles[s] = (row tile i, col tile j), i >= j, all
  // structurally non-zero lower tiles after symbolic fill; tile_slot[i*T + j] = s or -1
  const int2* nz_tiles;     // [n_nz]
  const int* tile_slot;     // [T*T]
  int n_nz;
  const int* panels;        // [T] panels sorted by elimination level
  const int2* trsm;         // (row tile i, panel k) sorted by level of k
  const int4* upd;          // (i, j, k, -) sorted by (level, conflict-free group)
  // per panel k: rows i > k with tile (i,k) non-zero: rows[row_ptr[k] .. row_ptr[k+1])
  const int* row_ptr;       // [n_tiles+1]
  const int* rows;
  // per tile row i: the non-zero tiles (i, j), j < i, for the triangular solves
  const int* lrow_ptr;      // [n_tiles+1]
  const int* lrow_cols;
  double* Dinv;             // [n_tiles][96*96] inverses of the diagonal factors (scratch)
  double* solve_partials;   // [n_tiles][16][96] split matrix-vector partial sums of the solves
  // forward substitution folded into the factorisation: tile (i, k) leaves  L_ik z_k  in slot
  // fwd_slot[its trsm entry] = its index in the lrow lists; panel i sums its slots in list order
  const int* fwd_slot;      // [n_trsm]
  double* fwd_partials;     // [n off-diagonal tiles][96]
  long n_real;              // 12 * frames (rows beyond it are identity padding)
};

void k3_prepare();   // one-off kernel attribute setup (call before capturing the launches in a graph)
void launch_clear_tiles(double* S, const TileSchedule& ts, cudaStream_t s);
// return the number of kernel launches issued; info[0] != 0 on a non-positive pivot
// x (rhs on entry) leaves as z = L^-1 rhs: the forward substitution rides in the factorisation's launches
int launch_tile_cholesky(double* S, const TileSchedule& ts, const TilePlan& plan, double* x, int* info, cudaStream_t s);
int launch_tile_solve(const double* S, const TileSchedule& ts, const TilePlan& plan,
                      double* x /* in: z = L^-1 rhs (from launch_tile_cholesky), out: solution */, cudaStream_t s);

// task-graph form of the same factorisation + both substitutions (k3_dag.cu): one persistent kernel
struct DagDevice {
  const DagTask* tasks;     // [n_tasks] topological order (tile_plan.cuh); the first n_factor_tasks factorise
  int n_tasks, n_factor_tasks;
  const int2* sources;
  const int* need;          // [n_nz * 4]
  int* counters;            // [dag_counter_ints] device-side dependency counters (cleared by the launcher)
  double* bwd_partials;     // [off-diagonal tiles][96]
  long long* trace = nullptr;  // optional [n_tasks][16] per-task time stamps (rsba_cuda_reduced_solve's trace_out)
};
size_t dag_counter_ints(const TileSchedule& ts);
// factor: L, inverses of the diagonal factors, z = L^-1 x;  solve: x = L^-T z.  Returns the launches issued.
int launch_tile_dag(double* S, const TileSchedule& ts, const DagDevice& dd, double* x, int* info, bool factor,
                    bool solve, cudaStream_t s);

// ---- K4 ---------------------------------------------------------------------------------
struct StepScalars {  // device doubles, filled by launch_step_update
  double g_dot_delta, d2_delta2, step_norm2, x_norm2, gmax;
};
// delta_c = -scale_c * y ; delta_p by back-substitution ; trial = x + delta ; scalars
// jac_cam / cam_frame: uncalibrated variant (NULL / -1 otherwise)
void launch_step_update(const SchurStructure& st, const ObsView& obs, const JacView& jv, const double* jac_cam,
                        int cam_frame, NormalEq ne,
                        const double* y_c, int n_frames, int n_points, const double* poses,
                        const double* points, double* delta_c, double* delta_p, double* trial_poses,
                        double* trial_points, double* scalars /* [16]: [0..2] camera part, [8..10] point part */,
                        double* scratch,
                        int bounded_param /* index into the camera parameters with a lower bound, or -1 */,
                        double lower_bound, const PointGroups& pg /* groups == NULL: one warp per point */, cudaStream_t s);
// |x|^2 and max|g| split into the point part (local to the rank; before the all-reduce) and the
// camera part (replicated; after it).  out_points: [0] += |x_p|^2 over owned free points, [1] = max|g_p|
void launch_point_norms(NormalEq ne, int n_points, const double* points, double* out_xx, double* out_gmax,
                        double* scratch, cudaStream_t s);
void launch_camera_norms(NormalEq ne, int n_frames, const double* poses, double* scalars /* writes [3], [4] */,
                         cudaStream_t s);

}  // namespace rsba
