/* rsba_cuda.h -- C ABI of the B200-native rolling-shutter bundle-adjustment inner loop.
 *
 * This is the drop-in boundary for the Evaluator + LinearSolver that henrique/rsba obtains
 * from Ceres.  Every entry point states the reference interface it replaces (paths relative
 * to /root/reference/src/rsba/).  Plain pointers and sizes only; no C++ or torch types.
 * All functions return RSBA_OK (0) or a negative error code; rsba_cuda_last_error() gives
 * the message.  No exceptions cross this boundary.  A handle is not thread-safe (like
 * ceres::Problem).  There is no CPU fallback: without a CUDA device every call fails.
 *
 * Parameter-block conventions (mat/cam.h:19-34, VideoSfmBaRs.h:25-35):
 *   pose   = [angle-axis r(3), camera centre c(3)]                (NUM_POSE_PARAMS 6)
 *   frame  = pose0[6] | pose1[6]   (first / last scan-line control pose, f.poses[0..1])
 *   point  = X[3]                                                 (NUM_POINT_PARAMS 3)
 *   cam    = fx fy k1 k2 p1 p2 k3 cx cy
 *   shutter: 0 GLOBAL, 1 HORIZONTAL, 2 VERTICAL (mat/cam.h:37-41)
 * Jacobian layout per observation (30 doubles), Ceres' per-block row-major contract:
 *   J_pose0[2][6] | J_pose1[2][6] | J_point[2][3]
 */
#ifndef RSBA_CUDA_H_
#define RSBA_CUDA_H_

#include "rsba_ceres_constants.h"

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct rsba_problem rsba_problem;

enum {
  RSBA_OK = 0,
  RSBA_ERR_INVALID_ARGUMENT = -1,
  RSBA_ERR_CUDA = -2,
  RSBA_ERR_NO_DEVICE = -3,
  RSBA_ERR_STATE = -4,
  RSBA_ERR_EVALUATION_FAILED = -5, /* a functor returned false (mat/cam.h:410-412) */
  RSBA_ERR_LINEAR_SOLVER = -6,     /* reduced camera matrix not positive definite */
  RSBA_ERR_NCCL = -7,
  RSBA_ERR_INTERNAL = -8           /* a C++ exception of the host code (e.g. out of host memory), caught at the boundary */
};

/* Solver::Options as used by CeresHandler::solve (CeresHandler.h:394-419) and
 * VideoSfMHandler::BA (VideoSfMHandler.cc:579-583).  rsba_cuda_default_options() fills the
 * Ceres 1.9.0 defaults. */
typedef struct rsba_solve_options {
  int max_num_iterations;           /* 50 (CeresHandler.h:405); callers pass maxIter */
  double initial_trust_region_radius; /* 1e4 */
  double max_trust_region_radius;     /* 1e16 */
  double min_trust_region_radius;     /* 1e-32 */
  double min_relative_decrease;       /* 1e-3 */
  double min_lm_diagonal;             /* 1e-6 */
  double max_lm_diagonal;             /* 1e32 */
  double function_tolerance;          /* 1e-6 */
  double gradient_tolerance;          /* 1e-10 */
  double parameter_tolerance;         /* 1e-8 */
  int jacobi_scaling;                 /* 1 */
  double huber_loss;                  /* > 0: same as rsba_cuda_set_loss(h, huber_loss) before solving;
                                         0 = keep the problem's loss (SfmOptions.h:64, CeresHandler.h:85-90) */
  int verbose;                        /* minimizer_progress_to_stdout (CeresHandler.h:404) */
  int dense_cholesky;                 /* 1: ignore the tile occupancy map, factor S fully dense */
  int reorder_tiles;                  /* 1: nested-dissection ordering of the reduced system (default);
                                         0: natural (frame) order */
  int max_num_consecutive_invalid_steps; /* 5: a failed linear solve (reduced camera matrix not positive
                                         definite, non-finite step) or a step with model_cost_change <= 0
                                         is an INVALID step -- the radius shrinks and the loop goes on;
                                         the solve fails only after this many in a row (Ceres 1.9
                                         trust_region_minimizer.cc; values <= 0 mean 5) */
} rsba_solve_options;

/* Solver::Summary fields the reference reads (VideoSfMHandler.cc:593-596, 627-630). */
typedef struct rsba_solve_summary {
  int usable;                  /* Summary::IsSolutionUsable() */
  int termination;             /* 0 convergence, 1 no convergence (max iter), 2 failure */
  int iterations;              /* LM iterations performed (excluding the initial evaluation) */
  int num_successful_steps;
  int num_unsuccessful_steps;
  int num_jacobian_evaluations;
  int num_residual_evaluations;
  long num_residual_blocks;
  long num_parameters_reduced; /* scalar parameters that are not constant */
  double initial_cost;
  double final_cost;
  double final_radius;
  double final_gradient_max_norm;
  double time_total_ms;        /* wall clock of rsba_cuda_solve, copies included */
  double time_jacobian_ms;     /* device time: residual+Jacobian kernel */
  double time_residual_ms;     /* device time: cost-only kernel */
  double time_schur_ms;        /* device time: normal equations + Schur complement */
  double time_cholesky_ms;     /* device time: reduced-system factorisation + solves */
  double time_update_ms;       /* device time: back-substitution, step, bookkeeping */
  double time_allreduce_ms;    /* device time: NCCL allreduce of the reduced system */
  char message[128];
  double tile_flops;           /* FP64 work of one reduced-system factorisation at tile granularity (what the
                                  symbolic analysis of this scene executes: potrf n^3/3, trsm n^3, update 2 n^3 per tile) */
  int reduced_levels;          /* elimination levels (length of the dependent panel chain) */
  int reduced_tiles;           /* non-zero 96 x 96 tiles of the factor after fill */
  /* Bounded problems (the free interFrameRatio, SetParameterLowerBound at CeresHandler.h:161,172): Ceres 1.9 follows
   * the projected trust-region step with an Armijo line search that starts at step size 1 and only contracts when
   * f(x + d) > f(x) + 1e-4 g.d.  The loop here takes the projected step as it is and counts the steps on which that
   * search would NOT have accepted step size 1: 0 means the missing line search was a no-op for this solve. */
  int num_armijo_violations;
} rsba_solve_summary;

/* ------------------------------------------------------------------ lifecycle */
/* Replaces: construction of CeresHandler / ceres::Problem (CeresHandler.h:83-91). */
int rsba_cuda_create(rsba_problem** out, int device);
void rsba_cuda_destroy(rsba_problem* h);
const char* rsba_cuda_last_error(void);
/* Optional: run on the caller's CUDA stream (a cudaStream_t); default is a private stream. */
int rsba_cuda_set_stream(rsba_problem* h, void* cuda_stream);
void rsba_cuda_default_options(rsba_solve_options* o);

/* ------------------------------------------------------------------ problem construction */
/* Replaces: the per-residual constants captured by RsBundleAdjustment::Create /
 * ReprojectionError ctor (VideoSfmBaRs.h:16-22,53-64; video_bundler_free.h:21-29): the
 * session intrinsics sess.cam, sess.rs, sess.scanlines and opt.model.interpolateRotation. */
int rsba_cuda_set_camera(rsba_problem* h, const double cam9[9], int shutter,
                         const int scanlines[2], int interpolate_rotation);

/* Uncalibrated variant.  Replaces: RsBundleAdjustment::CreateWithCam(sess, opt, obs) + problem.AddResidualBlock(
 * cost, loss, sess.cam.data(), f.poses[0], f.poses[1], t->pt)  (VideoSfmBaRs.h:38-49,68-80; CeresHandler.h:256-264,
 * opt.model.calibrated == false): the session's nine intrinsics (fx fy k1 k2 p1 p2 k3 cx cy) become one shared
 * 9-wide PARAMETER block of every residual block -- a dense border of the reduced camera system.  Call with 1
 * before solving; rsba_cuda_set_camera supplies the initial values, rsba_cuda_get_camera returns the optimised
 * ones; rsba_cuda_get_intrinsics_jacobian returns d residual / d intrinsics [N][2][9] of the last
 * rsba_cuda_evaluate (HOST pointer, caller's observation order).  Per-frame intrinsics blocks (f.cam) are
 * not supported. */
int rsba_cuda_set_intrinsics_free(rsba_problem* h, int free_intrinsics);
/* Pointer-identity form of the same: problem.AddResidualBlock(RsBundleAdjustment::CreateWithCam(sess, opt, obs), loss,
 * sess.cam.data(), f.poses[0].data(), f.poses[1].data(), t->pt.data())  (CeresHandler.h:256-264).  `intrinsics`
 * (9 doubles, caller-owned, updated in place by rsba_cuda_solve) must be the same block in every call; its
 * values at solve / evaluate time are the starting point (rsba_cuda_set_camera still supplies shutter,
 * scan lines and interpolateRotation). */
int rsba_cuda_add_rs_residual_with_intrinsics(rsba_problem* h, const double observed[2], double* intrinsics,
                                              double* pose0, double* pose1, double* point);
int rsba_cuda_get_camera(rsba_problem* h, double cam9[9]);
int rsba_cuda_get_intrinsics_jacobian(rsba_problem* h, double* jacobian_cam);

/* Replaces: lossFunction = new ceres::HuberLoss(opt.ceres.huberLoss), handed to every
 * AddResidualBlock (CeresHandler.h:85-90, 252).  huber_a = 0 removes the loss (the default,
 * SfmOptions.h:64).  Applied the way Ceres' Corrector does: cost = 1/2 rho(|r|^2), residual and
 * Jacobian rescaled by sqrt(rho') (rho'' <= 0 for Huber); rsba_cuda_evaluate returns the
 * corrected values, as problem.Evaluate does with apply_loss_function = true. */
int rsba_cuda_set_loss(rsba_problem* h, double huber_a);

/* Replaces: RsBundleAdjustment::Create(sess,opt,obs) + problem.AddResidualBlock(cost, loss,
 * f.poses[0].data(), f.poses[1].data(), t->pt.data())  (CeresHandler.h:250-255).
 * Block identity is pointer identity, as in Ceres; the caller owns the parameter memory,
 * which must stay put until the handle is destroyed.  Results are written back in place. */
int rsba_cuda_add_rs_residual(rsba_problem* h, const double observed[2], double* pose0,
                              double* pose1, double* point);
/* ceres::Problem::AddParameterBlock for one frame: registers the two control-pose blocks of a frame that no
 * rolling-shutter residual block or motion prior has introduced yet -- a frame that (so far) only carries
 * GoodPosePrior blocks (CeresHandler.h:188-204), which name ONE pose block each and cannot tell the library
 * which blocks form a frame.  Ceres accepts such a problem; so does rsba_cuda_solve.  Idempotent. */
int rsba_cuda_add_frame_blocks(rsba_problem* h, double* pose0, double* pose1);

/* Replaces: RsConstVeloPrior::Create(scale) / RsConstAccelerationPrior::Create(scale) +
 * problem.AddResidualBlock(cost, loss, &opt.ceres.interFrameRatio, f.poses[0], f.poses[1],
 * f_1.poses[0], f_1.poses[1])  (CeresHandler.h:148-186; functors video_bundler_rs_inter.h:55-173):
 * the 12-residual constant-velocity (kind 1) / constant-acceleration (kind 2) prior between frame k
 * (pose0, end0) and frame k-1 (pose1, end1).  inter_frame_ratio is the value of a CONSTANT ratio block --
 * the reference fixes it whenever opt.ceres.interFrameRatio != 1 (CeresHandler.h:178-180); with the
 * reference's default (== 1) the block is a free, lower-bounded parameter: see
 * rsba_cuda_set_inter_frame_ratio_block / _free below (the per-prior value is then ignored).  A frame may be the current
 * frame of one prior and the previous frame of one prior.  The problem's loss (rsba_cuda_set_loss)
 * applies to the prior blocks too, as in the reference.  Rejects ratios for which the functor
 * returns false (velocity: ratio < 0; acceleration: ratio < DBL_EPSILON). */
int rsba_cuda_add_motion_prior(rsba_problem* h, int kind, double scale, double inter_frame_ratio,
                               double* pose0, double* end0, double* pose1, double* end1);
/* Bulk form (after rsba_cuda_set_scene, which clears the list): frame[i] / prev_frame[i] index poses. */
int rsba_cuda_set_motion_priors(rsba_problem* h, int n, const int* kind, const double* scale,
                                const double* inter_frame_ratio, const int* frame, const int* prev_frame);
/* Replaces: the parameter block `&opt.ceres.interFrameRatio` that every motion prior shares, left variable
 * (CeresHandler.h:156,167 -- SetParameterBlockConstant only when the value != 1, :178-180) with
 * problem.SetParameterLowerBound(&ratio, 0, _EPS) for acceleration priors / (.., 0, 0.0) for velocity priors
 * (:161,172).  The scalar becomes one more column of the reduced camera system (parameter 9 of the pseudo-frame
 * that also carries free intrinsics); the trial point is projected onto the bound as Ceres' Plus does.
 * _block: pointer API -- *ratio is read at solve/evaluate entry and written back after the solve.
 * _free : bulk API -- free_ratio = 0 returns to the constant per-prior values; read the result with _get. */
int rsba_cuda_set_inter_frame_ratio_block(rsba_problem* h, double* ratio);
int rsba_cuda_set_inter_frame_ratio_free(rsba_problem* h, int free_ratio, double value);
int rsba_cuda_get_inter_frame_ratio(rsba_problem* h, double* value);
/* d residual / d interFrameRatio [12 n] (loss-corrected) of the priors at the last residual+Jacobian
 * evaluation, free ratio only (HOST pointer, may be NULL); returns the number of priors or -1. */
long rsba_cuda_get_prior_ratio_jacobian(rsba_problem* h, double* d_residual_d_ratio);
/* Loss-corrected residuals [12 n] of the priors at the last residual+Jacobian evaluation (HOST pointer,
 * may be NULL); returns the number of priors. */
long rsba_cuda_get_prior_residuals(rsba_problem* h, double* residuals);

/* Replaces: GoodPosePrior::Create(opt.ceres.trustPriorCamRotation, opt.ceres.trustPriorCamPosition) +
 * problem.AddResidualBlock(cost, nullptr, f.priorPoses[i].data(), f.poses[i].data())  (CeresHandler.h:55-73,
 * 188-204): 6 residuals  diag(rotation x3, position x3) (prior - pose), no loss; the functor returns false
 * when the first residual is >= 1.  BOTH blocks are parameter blocks -- the reference never fixes the prior
 * block, so it moves too (it is eliminated in closed form on the device and written back after the solve);
 * rsba_cuda_set_block_constant(prior_block) fixes it.  pose_block must be one of the control-pose blocks of
 * a frame that also appears in a reprojection residual block.  One prior per control pose. */
int rsba_cuda_add_pose_prior(rsba_problem* h, double rotation, double position, double* prior_block,
                             double* pose_block);
/* Bulk form (after rsba_cuda_set_scene, which clears the list): prior i ties prior_values[6 i .. 6 i + 5] to
 * control pose which_pose[i] (0 | 1) of frame[i]; prior_constant may be NULL (all free, as in the reference). */
int rsba_cuda_set_pose_priors(rsba_problem* h, int n, const int* frame, const int* which_pose,
                              const double* rotation, const double* position, const double* prior_values,
                              const unsigned char* prior_constant);
/* Current values [6 n] of the prior blocks and the trial values of the last LM step (HOST pointers, either
 * may be NULL); returns the number of pose priors or -1. */
long rsba_cuda_get_pose_priors(rsba_problem* h, double* prior_values, double* trial_values);

/* Replaces: problem.SetParameterBlockConstant(double*)  (CeresHandler.h:283,299,344-345). */
int rsba_cuda_set_block_constant(rsba_problem* h, double* block);
/* Replaces: problem.SetParameterization(pose, new SubsetParameterization(6, constant))
 * (CeresHandler.h:350-381): the listed components of a 6-wide pose block are held fixed. */
int rsba_cuda_set_subset_constant(rsba_problem* h, double* pose_block, int n_constant,
                                  const int* constant_components);

/* Bulk form of the same construction for callers that already hold flat arrays (the
 * session <-> SoA marshaller).  Observation i ties frame obs_frame[i] (poses + 12*frame)
 * to point obs_point[i] (points + 3*point).  Arrays are HOST memory and are copied.
 * Observations are stably sorted by frame on upload (== the reference's insertion order,
 * CeresHandler.h:208); outputs are always reported in the caller's order.
 * const_pose_mask[frame] bit k set => scalar k (0..11) of the frame is constant (may be NULL).
 * const_point[point] != 0 => point block constant (may be NULL). */
int rsba_cuda_set_scene(rsba_problem* h, long n_obs, const double* obs_xy, const int* obs_frame,
                        const int* obs_point, int n_frames, int n_points,
                        const unsigned short* const_pose_mask, const unsigned char* const_point);
/* Host -> device copy of the current parameter values (bulk form). */
int rsba_cuda_set_parameters(rsba_problem* h, const double* poses, const double* points);
/* Device -> host copy of the current parameter values (bulk form). */
int rsba_cuda_get_parameters(rsba_problem* h, double* poses, double* points);

/* ------------------------------------------------------------------ evaluation */
/* Replaces: problem.Evaluate(EvaluateOptions(), &cost, residuals, NULL, jacobian)
 * (CeresHandler.h:386) == what ceres::ProgramEvaluator does each LM iteration with
 * AutoDiffCostFunction<RsBundleAdjustment,2,6,6,3> (VideoSfmBaRs.h:58-63).
 * Outputs are HOST pointers, any may be NULL: cost = 1/2 sum rho(r^2) over the reprojection blocks AND
 * the motion priors; residuals[2N] / jacobian[30N] / valid[N] cover the reprojection blocks;
 * jacobian[30N]; valid[N] (the functor's bool).  Pointer-API problems read the parameter
 * values from the caller's blocks first.  Returns RSBA_ERR_EVALUATION_FAILED (outputs still
 * written, invalid rows zero) if any functor returned false. */
int rsba_cuda_evaluate(rsba_problem* h, double* cost, double* residuals, double* jacobian,
                       unsigned char* valid);

/* Track-validation sweep.  Replaces: validate(sess, f, opt, t.pt, obs) (struct/VideoSfM.cc:159-169) as
 * evalTracks / createTracks run it over every observation after a BA (VideoSfMHandler.cc:377-410,
 * 599-600): ok[i] = |c(tau_i) - X| >= min_distance_to_camera (opt.tracks.minDistanceToCamera) and the
 * projection succeeds (z >= 1e-8) and |proj - obs|^2 < sqrd_threshold (opt.tracks.sqrdThreshold).
 * Unlike the BA functor the pose is interpolated at the observation's own scan line: x for a
 * HORIZONTAL, y for a VERTICAL shutter (getPose, struct/VideoSfM.cc:108-111).  HOST outputs in the
 * caller's observation order, either may be NULL; sqrd_error[i] = -1 where the projection failed. */
int rsba_cuda_validate(rsba_problem* h, double sqrd_threshold, double min_distance_to_camera,
                       unsigned char* ok, double* sqrd_error);

/* Iterative re-projection.  Replaces: reproject(sess, f, opt, t.pt, obs) (struct/VideoSfM.cc:139-155) for a batch
 * of (frame, point) pairs at the CURRENT parameters: the rolling-shutter scan line of a projection is unknown, so
 * the reference starts at the principal point and repeats  pose = getPose(proj), proj = w2i(cam, pose, pt)  until
 * the projection moves by less than 1e-3 px, at most 49 times.  ok[i] = 0 when w2i fails (z < 1e-8), the limit is
 * hit or sqrd_threshold <= 0 (the final validate() compares the projection with itself); proj_xy[i] is the last
 * iterate either way.  HOST arrays: frame[n], point[n] in, proj_xy[2 n], ok[n] out. */
int rsba_cuda_reproject(rsba_problem* h, long n, const int* frame, const int* point, double sqrd_threshold,
                        double* proj_xy, unsigned char* ok);

/* HBM-resident form: evaluates on the device and leaves everything there.  The returned
 * device pointers (sorted-by-frame observation order) stay valid until the next scene
 * change.  with_jacobian = 0 runs the cost-only kernel. */
int rsba_cuda_evaluate_device(rsba_problem* h, int with_jacobian, double* cost,
                              long* num_invalid);
int rsba_cuda_device_buffers(rsba_problem* h, void** residuals, void** jacobian, void** valid,
                             void** poses, void** points);
/* Sorted position -> caller's observation index (HOST array of n_obs longs, may be NULL to
 * query the count). */
long rsba_cuda_observation_order(rsba_problem* h, long* order);

/* Host only, needs no device: the order above for a given observation list -- a STABLE sort by frame, which
 * keeps the caller's within-frame order (the reference's insertion order, CeresHandler.h:208-255).  Returns
 * n_obs, or -1 (index out of range; message in rsba_cuda_last_error).  order may be NULL (validation only). */
long rsba_cuda_sort_observations(long n_obs, const int* obs_frame, const int* obs_point, int n_frames, int n_points,
                                 long* order);

/* ------------------------------------------------------------------ solve */
/* Replaces: ceres::Solve(options, &problem, &summary) with linear_solver_type = SPARSE_SCHUR
 * (CeresHandler.h:403,419): Levenberg-Marquardt trust region; point blocks eliminated by a
 * Schur complement; reduced camera system factorised by Cholesky on the device; parameters
 * updated in place (pointer API) or on the device (bulk API; read back with
 * rsba_cuda_get_parameters). */
int rsba_cuda_solve(rsba_problem* h, const rsba_solve_options* opt, rsba_solve_summary* summary);

/* One linearisation at the current parameters, for parity tests and profiling: builds the
 * normal equations, applies Jacobi scaling + the LM diagonal for `radius`, eliminates the
 * points, factorises and solves.  HOST outputs, any may be NULL:
 *   (n = 12*frames; with free intrinsics n = 12*(frames+1): the intrinsics are parameters 0..8 of a last
 *    pseudo-frame whose parameters 9..11 are constant)
 *   S[n*n] (row-major, full symmetric), rhs[n]  -- reduced system in the
 *     scaled space, constant parameters replaced by identity rows;
 *   delta_poses[n], delta_points[3*points]                    -- unscaled LM step;
 *   model_cost_change.
 * Does not move the parameters. */
int rsba_cuda_linearize_and_step(rsba_problem* h, const rsba_solve_options* opt, double radius,
                                 double* S, double* rhs, double* delta_poses,
                                 double* delta_points, double* model_cost_change);

/* Host-only introspection of the symbolic analysis that precedes the factorisation (what CHOLMOD's
 * analyse phase does for Ceres' SparseSchurComplementSolver); needs no device.  The reduced
 * system is cut into n_tiles tiles of 96 rows (8 frames); pair_a/pair_b list the structurally
 * non-zero tile pairs in frame order.  Two-call pattern: with all output pointers NULL only
 * counts[] is filled: {levels, non-zero tiles after fill, trsm tiles, update triples, update
 * groups, tile-level flops}.  Outputs (tile indices are positions after reordering):
 *   tile_pos[n_tiles] frame tile -> position; nz_tiles[2*n_nz] (i, j); panels[n_tiles] sorted by
 *   level; panel_ptr[levels+1]; trsm[2*n_trsm] (i, k); trsm_ptr[levels+1]; upd[3*n_upd] (i, j, k);
 *   group_ptr[groups+1] (as long); level_group_ptr[levels+1]. */
int rsba_cuda_plan_reduced_system(int n_tiles, int n_pairs, const int* pair_a, const int* pair_b,
                                  int dense, int reorder, long counts[6], int* tile_pos, int* nz_tiles,
                                  int* panels, int* panel_ptr, int* trsm, int* trsm_ptr, int* upd,
                                  long* group_ptr, int* level_group_ptr);

/* Host-only: the numeric phase of the reduced-system solve (Cholesky factorisation, forward and backward
 * substitution) as the static task graph that the persistent kernel of rsba_b200/csrc/k3_dag.cu executes --
 * the stand-in for CHOLMOD's numeric factorisation + solve behind Ceres' SPARSE_SCHUR (CeresHandler.h:403).
 * Same plan inputs as rsba_cuda_plan_reduced_system; merge_levels >= 1 = elimination levels per merged update
 * group.  Two-call pattern: counts[] = {tasks, update sources, non-zero tiles, factorisation tasks}; then
 *   tasks[8*n_tasks]   records {type, a..g} in topological (execution) order, see tile_plan.cuh
 *   sources[2*n_src]   (slot(i,k), slot(j,k)) of the updates;  need[4*n_nz] update groups per tile quadrant. */
int rsba_cuda_plan_task_graph(int n_tiles, int n_pairs, const int* pair_a, const int* pair_b, int dense,
                              int reorder, int merge_levels, long counts[4], int* tasks, int* sources,
                              int* need);

/* K3 on its own (test and measurement hook; no counterpart in the reference, where CHOLMOD sits behind
 * ceres::Solve): solves A x = rhs on `device` for a symmetric positive definite A (n x n row-major,
 * n = 96 n_tiles) whose block pattern is given by the tile pairs (or dense).  mode 0 = task-graph kernel,
 * 1 = level-batched launches.  Optional outputs: L (n x n, in the PERMUTED tile order tile_pos_out[n_tiles]
 * describes), info (0, or 1 + index of the first non-positive pivot), device time of the numeric phase
 * (the fastest of `repeats` runs on the same data), and -- mode 0, trace_out[16 * tasks] -- per task of the
 * last run {globaltimer ns at fetch, inputs ready, end; clock64 at the same three points; SM id; type;
 * FACTOR tasks also clock64 after load / factorisation / inversion / stores and two phase sums}. */
int rsba_cuda_reduced_solve(int device, int n_tiles, int n_pairs, const int* pair_a, const int* pair_b, int dense,
                            int reorder, int mode, int merge_levels, int repeats, const double* A,
                            const double* rhs, double* x_out, double* L_out, int* tile_pos_out, int* info_out,
                            float* ms_out, long long* trace_out);

/* Measurement hook: the FP64 peak of `device` in TFLOP/s, by register-resident DFMA chains and by
 * mma.sync.m8n8k4.f64 chains (the two FP64 paths of sm_100a; tcgen05 has no FP64 kind).  ~30 ms.  bench.py quotes
 * the K2 / K3 rooflines against the larger of the two, measured in the same run. */
int rsba_cuda_measure_fp64_peak(int device, double* dfma_tflops, double* dmma_tflops);

/* Host-only introspection of the WHOLE one-off structure analysis that rsba_cuda_solve runs before its first
 * linearisation -- what Ceres does in Program reordering + SchurEliminator block-structure detection + CHOLMOD's
 * analyse phase (third-party; reached through ceres::Solve, CeresHandler.h:403,419).  Needs no device: the CPU
 * tests execute the returned work lists in numpy against a direct Schur complement.
 *   obs_frame / obs_point [n_obs]  observations sorted by frame (the order rsba_cuda_observation_order reports)
 *   const_point [n_points] or NULL; free_intrinsics / free_ratio: the pseudo-frame behind the real frames
 *   prior_frame / prior_prev [n_priors]: the (frame, previous frame) couplings of the motion priors
 *   dense / reorder: as rsba_solve_options.dense_cholesky / reorder_tiles; sparse_keys: force the sorted-key path
 *   rank / world_size: world_size > 1 analyses the share that rank keeps of the scene (all observations of the
 *     points it owns, rsba_cuda_point_owners); the tile plan is derived from the whole scene on every rank.  Extra
 *     arrays then: local_ids (long: the rank's observations as positions in the given list), point_owned (bytes)
 * The result is an opaque object; rsba_cuda_structure_array returns the element count of the named array (or -1)
 * and, through data / elem_bytes, a pointer into the object (valid until rsba_cuda_structure_free).  Names:
 *   pt_ptr pt_obs chunk_frame chunk_beg chunk_cnt frame_chunk_ptr inc_point inc_tile slot_beg slot_cnt pt_inc_ptr
 *   cam_inc inc_half obs_phi_off dup_inc pair_a pair_b pair_item_ptr items(int4) entries(int2) fwd_slot
 *   plan.tile_pos plan.pos_tile plan.nz_tiles(int2) plan.tile_slot plan.panels plan.panel_ptr plan.trsm(int2)
 *   plan.trsm_ptr plan.upd(int4) plan.lrow_ptr plan.lrow_cols; the scalars T, n_inc, n_items come back as the count. */
typedef struct rsba_structure rsba_structure;
int rsba_cuda_analyze_structure(long n_obs, const int* obs_frame, const int* obs_point, int n_frames, int n_points,
                                const unsigned char* const_point, int free_intrinsics, int free_ratio, int n_priors,
                                const int* prior_frame, const int* prior_prev, int dense, int reorder,
                                int sparse_keys, int rank, int world_size, rsba_structure** out);
long rsba_cuda_structure_array(const rsba_structure* s, const char* name, const void** data, int* elem_bytes);
void rsba_cuda_structure_free(rsba_structure* s);

/* ------------------------------------------------------------------ batched RS-PnP */
/* Replaces: the inner ceres::Solve of vision::solveRsPnP (solveRSpnp.cpp:100-192: RsBA<float> residual
 * blocks <2; 6, 6> on the two control poses of one frame, 3-D points fixed, w2i without the
 * in-front-of-camera test, max_num_iterations = 10) for ALL RANSAC hypotheses at once (pnpTask,
 * solveRSpnp.cpp:265-335), and the inlier count each refined hypothesis is scored with
 * (project3dPoints + |obs - proj| < reprojectionError, solveRSpnp.cpp:226-263, 318-323).
 *   points3d[n][3], obs_xy[n][2]   the frame's 2-D/3-D correspondences (the reference passes them
 *                                  through float; quantise before the call to reproduce that)
 *   sample_idx[n_hyp][sample_size] the points of each hypothesis' minimal sample (sample_size <= 32)
 *   poses[n_hyp][12]               in: initial pose0|pose1 (the reference starts every hypothesis from
 *                                  the same guess); out: refined
 *   options                        LM constants (max_num_iterations = 10 in the reference)
 *   final_cost / usable / iterations / inlier_count [n_hyp]   HOST outputs, any may be NULL
 * All arrays are HOST memory.  Larger point sets (the final refinement on all inliers, the const3d
 * PnP-BA of VideoSfMHandler.cc:434-485) are a one-frame problem with constant points for rsba_cuda_solve. */
int rsba_cuda_pnp_batch(rsba_problem* h, const double cam9[9], int shutter, const int scanlines[2],
                        int n_points, const double* points3d, const double* obs_xy, int n_hyp,
                        int sample_size, const int* sample_idx, double* poses,
                        const rsba_solve_options* options, double inlier_threshold, double* final_cost,
                        int* usable, int* iterations, int* inlier_count);

/* ------------------------------------------------------------------ multi-GPU */
/* One process per GPU.  Rank 0 obtains an id, the host framework broadcasts the 128 bytes,
 * every rank calls comm_init and then passes the SAME whole scene to set_scene / the pointer
 * API; the library keeps this rank's share (all observations of the points it owns) on the
 * device.  rsba_cuda_solve then performs one NCCL all-reduce of the unscaled reduced system
 * [S | gradient terms | diag(B) | cost, invalid count, norms] per linearisation, plus an
 * 8-double all-reduce of the trial-step scalars per iteration and one of the points at the end;
 * the factorisation is replicated.  Single-GPU use never loads NCCL (dlopen at comm_init). */
int rsba_cuda_nccl_unique_id(unsigned char id[128]);
int rsba_cuda_comm_init(rsba_problem* h, int rank, int world_size, const unsigned char id[128]);
/* One HOST THREAD, N GPUs of one node -- for a driver that is single-threaded like the reference's
 * (VideoSfMHandler::BA builds one CeresHandler and calls ceres::Solve from the request thread,
 * VideoSfMHandler.cc:574-631, CeresHandler.h:408-419).  rsba_cuda_create_multi creates one problem handle per
 * device and connects them (ncclCommInitRank from worker threads of this process; n_devices == 1 never loads
 * NCCL).  The caller forwards every builder call (set_camera, set_scene / add_*, set_parameters, priors ...) to
 * each rsba_cuda_multi_handle(m, r) with the SAME arguments -- every rank keeps its share -- and then calls
 * rsba_cuda_multi_solve once: the ranks' LM loops run on worker threads and are joined before it returns;
 * `summary` is rank 0's (costs and counts are global).  Results are read from handle 0 (get_parameters); with
 * the pointer API rank 0 writes the caller's blocks back.  include/rsba_cuda_handler.hpp does the forwarding when
 * it is constructed with a device list. */
typedef struct rsba_multi rsba_multi;
int rsba_cuda_device_count(void);   /* CUDA devices visible to this process (0 without a driver / device) */
int rsba_cuda_create_multi(rsba_multi** out, const int* devices, int n_devices);
int rsba_cuda_multi_size(const rsba_multi* m);
rsba_problem* rsba_cuda_multi_handle(rsba_multi* m, int rank);
int rsba_cuda_multi_solve(rsba_multi* m, const rsba_solve_options* options, rsba_solve_summary* summary);
void rsba_cuda_destroy_multi(rsba_multi* m);

/* Host-only (no device): the sharding rule.  owner[p] = rank that eliminates point p and therefore
 * evaluates ALL its observations -- the rank whose contiguous range of frame tiles (8 frames)
 * holds the point's median observation.  obs_frame must be non-decreasing. */
int rsba_cuda_point_owners(int n_frames, int n_points, long n_obs, const int* obs_frame,
                           const int* obs_point, int world_size, int* owner);

/* ------------------------------------------------------------------ introspection */
/* Number of kernel launches issued by this handle so far (bench.py's gpu_launches). */
long rsba_cuda_launch_count(rsba_problem* h);
/* Device time (ms) of the last call of each stage, measured with CUDA events on the
 * launching stream: 0 jacobian, 1 residual, 2 schur, 3 cholesky, 4 update, 5 allreduce; and of
 * single kernels inside them: 6 point blocks, 7 frame blocks, 8 Schur panels, 9 Schur SYRK,
 * 10 Schur reduce, 11 factorisation, 12 triangular solves, 13 point back-substitution,
 * 14 finalize, 15 batched PnP. */
double rsba_cuda_stage_ms(rsba_problem* h, int stage);
const char* rsba_cuda_version(void);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* RSBA_CUDA_H_ */
