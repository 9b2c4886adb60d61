// Host driver of the Levenberg-Marquardt loop and the one-off structure analysis.
//
// Supersedes ceres::Solve(options, &problem, &summary) as CeresHandler::solve calls it
// (CeresHandler.h:394-426; options set at VideoSfMHandler.cc:579-583): trust-region LM with
// Jacobi scaling, Schur elimination of the point blocks, Cholesky of the reduced camera system.
// The control flow restates Ceres 1.9.0's TrustRegionMinimizer / LevenbergMarquardtStrategy
// (third-party, not in the reference tree; constants in rsba_cuda_default_options).  All
// state stays in HBM; per iteration the host reads back a handful of scalars.
#include "lm.cuh"
#include "problem.cuh"
#include "structure.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace rsba {

struct LmState {
  // ---- structure (device)
  DeviceBuffer<int> pt_ptr, pt_obs, chunk_frame, chunk_beg, chunk_cnt, frame_chunk_ptr;
  DeviceBuffer<int> inc_point, inc_tile, slot_beg, pair_a, pair_b, pair_item_ptr, tile_pos, pos_tile;
  DeviceBuffer<int> obs_phi_off, dup_inc, cam_inc;
  DeviceBuffer<double> Bcam, cam_partials, cam_scratch;
  bool free_cam = false, free_ratio = false;
  int n_cam_frames = 0;          // frames of the camera system: real frames (+ the intrinsics pseudo-frame)
  DeviceBuffer<unsigned char> slot_cnt, point_owned;
  DeviceBuffer<int> owned_ids;
  DeviceBuffer<int4> items;
  DeviceBuffer<int2> entries;

  DeviceBuffer<int2> nz_tiles, trsm;
  DeviceBuffer<int4> upd;
  DeviceBuffer<int> tile_slot, row_ptr, rows, lrow_ptr, lrow_cols, panels;
  DeviceBuffer<unsigned short> pose_mask;
  DeviceBuffer<unsigned char> point_const;
  TilePlan plan;                              // host copy of the symbolic analysis
  // the numeric phase of K3 as one persistent task-graph kernel (k3_dag.cu); RSBA_CUDA_K3=levels selects the
  // level-batched launch sequence instead
  bool use_dag = true, dag_split = false;
  int n_dag_tasks = 0, n_dag_factor_tasks = 0;
  DeviceBuffer<DagTask> dag_tasks;
  DeviceBuffer<int2> dag_sources;
  DeviceBuffer<int> dag_need, dag_counters;
  DeviceBuffer<double> bwd_partials;
  DagDevice dd{};
  SchurStructure st{};
  TileSchedule ts{};
  bool dense = false, reorder = true;
  // ---- numeric state (device)
  DeviceBuffer<double> B, C, gp, Cinv, tp, Minv, Phi, partial, scale_c, scale_p, d2_c, d2_p, partials;
  DeviceBuffer<double> prec, pt_tau;      // per-point gather record; point-major tau (refreshed once per solve)
  DeviceBuffer<int> pt_frame;
  DeviceBuffer<unsigned char> pt_packed;  // fused point pass: (x, y, frame, phi_off, obs) per observation, point-major
  bool pt_major_valid = false;
  // calibrated scenes: the linearisation is the two fused passes of k2_fused.cu (RSBA_CUDA_FUSED=0: the
  // K1 -> point_blocks -> frame_blocks sequence, which the uncalibrated variant always uses)
  bool fused = false;
  DeviceBuffer<double> rec_pt, xt;        // point-major compact records [N][12]; (X_p | t_p) [P][6]
  DeviceBuffer<int2> pt_groups;           // thread-per-observation back-substitution: groups of whole points (lm.cuh)
  DeviceBuffer<int> pt_big;
  PointGroups pg{};
  // S is followed by the tail  gc | wf | diagB | misc  -- one buffer, one all-reduce (multi-GPU)
  DeviceBuffer<double> solve_partials, fwd_partials;
  DeviceBuffer<int> fwd_slot;
  DeviceBuffer<double> S, misc_local, Dinv, rhs, y, delta_c, delta_p, trial_poses, trial_points, scalars, scratch;
  // the launch sequences of the factorisation and of the triangular solves are static per scene:
  // captured once into CUDA graphs, replayed every LM iteration (launch gaps matter here -- ~170
  // short dependent kernels); nullptr = capture not available on this stream (legacy default stream)
  cudaGraphExec_t graph_factor = nullptr, graph_solve = nullptr;
  bool graph_tried = false;
  int graph_factor_launches = 0, graph_solve_launches = 0;
  ~LmState() {
    if (graph_factor) cudaGraphExecDestroy(graph_factor);
    if (graph_solve) cudaGraphExecDestroy(graph_solve);
  }
  double* misc = nullptr;      // tail: [0] cost [1] invalid [2] |x_p|^2 ... [8 + r] max|g_p| of rank r
  size_t comm_count = 0;       // doubles in S + tail
  int misc_count = 0;
  DeviceBuffer<int> info;
  NormalEq ne{};
  long n_pad = 0;
  long num_free_params = 0;
};

void lm_state_free(LmState* s) { delete s; }

namespace {

int fail(int code, const std::string& msg) {
  set_last_error(msg);
  return code;
}

template <typename T, typename A>
int upload(DeviceBuffer<T>& d, const std::vector<T, A>& h, cudaStream_t s) {
  RSBA_CUDA_TRY(d.resize(std::max<size_t>(h.size(), 1)));
  if (!h.empty()) RSBA_CUDA_TRY(cudaMemcpyAsync(d.ptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  return RSBA_OK;
}

// ------------------------------------------------------------------ structure analysis
// The analogue of Ceres' program reordering + symbolic factorisation, done once per scene.
int build_structure(rsba_problem* h, LmState* lm, bool dense) {
  // RSBA_CUDA_TRACE=1: wall-clock of the one-off host analysis, to stderr
  const bool trace = getenv("RSBA_CUDA_TRACE") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[rsba_cuda] structure: %-28s %8.1f ms\n", what,
            std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  const long N = h->n_obs;
  const int F = h->n_frames, P = h->n_points;
  const bool free_cam = h->free_cam, free_ratio = h->free_ratio && !h->priors.empty();
  const bool pseudo = h->has_pseudo_frame();
  const int Fc = h->n_cam_frames();    // + the pseudo-frame (free intrinsics / free interFrameRatio)
  cudaStream_t s = h->stream;

  // ---- host analysis (structure.cu; no device involved)
  SceneTopology topo;
  topo.n_obs = N; topo.n_frames = F; topo.n_points = P;
  topo.obs_frame = h->h_obs_frame.data(); topo.obs_point = h->h_obs_point.data();
  topo.point_const = h->point_const.data();
  topo.free_cam = free_cam; topo.free_ratio = h->free_ratio;
  for (const auto& pr : h->priors) topo.prior_pairs.emplace_back(pr.frame, pr.prev);
  topo.world = h->world; topo.n_obs_global = h->n_obs_global;
  topo.g_obs_frame = h->g_obs_frame.data(); topo.g_obs_point = h->g_obs_point.data();
  topo.dense = dense; topo.reorder = h->reorder_tiles;
  topo.sparse_keys = getenv("RSBA_CUDA_SPARSE_KEYS") != nullptr;   // (env: test hook)
  HostStructure hs;
  {
    std::string err;
    auto lap_cb = [](const char* what, void* ctx) { (*static_cast<decltype(lap)*>(ctx))(what); };
    const int arc = analyze_structure(topo, &hs, &err, lap_cb, &lap);
    if (arc) return fail(arc, err);
  }
  lm->plan = std::move(hs.plan);
  const TilePlan& plan = lm->plan;
  const int T = hs.T, n_inc = hs.n_inc, n_items = hs.n_items;
  const std::vector<int>& chunk_frame = hs.chunk_frame;
  const std::vector<int>& pair_a = hs.pair_a;
  const std::vector<int2>& nz_tiles = plan.nz_tiles;

  // ---- upload
  int rc;
#define UP(dev, host) if ((rc = upload(lm->dev, host, s))) return rc
  UP(pt_ptr, hs.pt_ptr); UP(pt_obs, hs.pt_obs); UP(chunk_frame, hs.chunk_frame); UP(chunk_beg, hs.chunk_beg);
  UP(chunk_cnt, hs.chunk_cnt); UP(frame_chunk_ptr, hs.frame_chunk_ptr);
  UP(inc_point, hs.inc_point); UP(inc_tile, hs.inc_tile); UP(slot_beg, hs.slot_beg); UP(slot_cnt, hs.slot_cnt);
  UP(obs_phi_off, hs.obs_phi_off); UP(dup_inc, hs.dup_inc); UP(cam_inc, hs.cam_inc);
  UP(pair_a, hs.pair_a); UP(pair_b, hs.pair_b);
  UP(pair_item_ptr, hs.pair_item_ptr); UP(items, hs.items); UP(entries, hs.entries);
  UP(tile_pos, plan.tile_pos); UP(pos_tile, plan.pos_tile); UP(point_owned, h->point_owned);
  UP(nz_tiles, plan.nz_tiles); UP(upd, plan.upd); UP(tile_slot, plan.tile_slot);
  UP(row_ptr, plan.row_ptr); UP(rows, plan.rows); UP(lrow_ptr, plan.lrow_ptr); UP(lrow_cols, plan.lrow_cols);
  UP(panels, plan.panels); UP(trsm, plan.trsm); UP(fwd_slot, hs.fwd_slot);
  {
    const char* mode = getenv("RSBA_CUDA_K3");
    lm->use_dag = !(mode && strcmp(mode, "levels") == 0);
    lm->dag_split = getenv("RSBA_CUDA_K3_SPLIT") != nullptr;
    const char* merge = getenv("RSBA_CUDA_K3_MERGE");   // (env: experiment hook) levels per merged update group
    DagPlan dag;
    build_dag_plan(plan, merge ? atoi(merge) : 4, &dag);
    lap("task graph");
    lm->n_dag_tasks = (int)dag.tasks.size();
    lm->n_dag_factor_tasks = dag.n_factor_tasks;
    UP(dag_tasks, dag.tasks); UP(dag_sources, dag.sources); UP(dag_need, dag.need);
  }
  {
    std::vector<unsigned short> mask(h->pose_mask);
    mask.resize(Fc, 0);
    if (pseudo)   // parameters 0..8 = intrinsics, 9 = interFrameRatio; what is not a parameter is constant
      mask[F] = (unsigned short)(0xFFF & ~(free_cam ? 0x1FF : 0) & ~(free_ratio ? 0x200 : 0));
    UP(pose_mask, mask);
  }
  UP(point_const, h->point_const);
#undef UP
  lm->dense = dense;
  lm->reorder = h->reorder_tiles;
  lm->n_pad = (long)T * kTile;

  SchurStructure& st = lm->st;
  st.pt_ptr = lm->pt_ptr.ptr; st.pt_obs = lm->pt_obs.ptr;
  st.chunk_frame = lm->chunk_frame.ptr; st.chunk_beg = lm->chunk_beg.ptr; st.chunk_cnt = lm->chunk_cnt.ptr;
  st.frame_chunk_ptr = lm->frame_chunk_ptr.ptr; st.n_chunks = (int)chunk_frame.size();
  st.n_inc = n_inc; st.inc_point = lm->inc_point.ptr; st.inc_tile = lm->inc_tile.ptr;
  st.slot_beg = lm->slot_beg.ptr; st.slot_cnt = lm->slot_cnt.ptr;
  st.obs_phi_off = lm->obs_phi_off.ptr; st.dup_inc = lm->dup_inc.ptr; st.n_dup = (int)hs.dup_inc.size();
  st.cam_inc = lm->cam_inc.ptr;
  st.n_pairs = (int)pair_a.size(); st.pair_a = lm->pair_a.ptr; st.pair_b = lm->pair_b.ptr;
  st.pair_item_ptr = lm->pair_item_ptr.ptr; st.n_items = n_items; st.items = lm->items.ptr;
  st.entries = lm->entries.ptr; st.n_entries = (long)hs.entries.size(); st.tile_pos = lm->tile_pos.ptr;
  st.pos_tile = lm->pos_tile.ptr; st.n_cam_params = 12L * Fc;
  lm->free_cam = free_cam;
  lm->free_ratio = h->free_ratio;
  lm->n_cam_frames = Fc;

  // ---- numeric buffers
  const size_t Fz = std::max(Fc, 1), Pz = std::max(P, 1);
  if (pseudo) {
    RSBA_CUDA_TRY(lm->Bcam.resize(Fz * 144));
    RSBA_CUDA_TRY(cudaMemsetAsync(lm->Bcam.ptr, 0, lm->Bcam.bytes(), s));
  }
  if (free_cam) {
    RSBA_CUDA_TRY(lm->cam_partials.resize(std::max<size_t>(chunk_frame.size(), 1) * 207));
    RSBA_CUDA_TRY(lm->cam_scratch.resize(Fz * 99));
  }
  RSBA_CUDA_TRY(lm->B.resize(Fz * 144));
  RSBA_CUDA_TRY(lm->C.resize(Pz * 6)); RSBA_CUDA_TRY(lm->gp.resize(Pz * 3)); RSBA_CUDA_TRY(lm->Cinv.resize(Pz * 6));
  RSBA_CUDA_TRY(lm->tp.resize(Pz * 3)); RSBA_CUDA_TRY(lm->Minv.resize(Pz * 6));
  RSBA_CUDA_TRY(lm->prec.resize(Pz * kPointRec));
  RSBA_CUDA_TRY(cudaMemsetAsync(lm->prec.ptr, 0, lm->prec.bytes(), s));
  RSBA_CUDA_TRY(lm->pt_tau.resize(std::max<size_t>((size_t)h->n_obs, 1)));
  RSBA_CUDA_TRY(lm->pt_frame.resize(std::max<size_t>((size_t)h->n_obs, 1)));
  lm->st.pt_tau = lm->pt_tau.ptr; lm->st.pt_frame = lm->pt_frame.ptr;
  {
    const char* e = getenv("RSBA_CUDA_FUSED");
    lm->fused = !free_cam && !(e && e[0] == '0');
  }
  if (lm->fused) {
    const size_t Nz = std::max<size_t>((size_t)h->n_obs, 1);
    RSBA_CUDA_TRY(lm->pt_packed.resize(Nz * point_pass_record_bytes()));
    launch_pack_point_major(lm->st, h->obs_view(), h->n_obs, lm->pt_packed.ptr, s);
    RSBA_CUDA_TRY(lm->rec_pt.resize(Nz * kJacCompact));
    RSBA_CUDA_TRY(lm->xt.resize(Pz * 6));
    RSBA_CUDA_TRY(cudaMemsetAsync(lm->xt.ptr, 0, lm->xt.bytes(), s));
    {   // groups of whole points for the back-substitution (structure.cu); long tracks of OWNED points: warp-per-point
      std::vector<int> big;
      for (int p : hs.point_big)
        if (h->point_owned[p]) big.push_back(p);
      if ((rc = upload(lm->pt_groups, hs.point_groups, s))) return rc;
      if ((rc = upload(lm->pt_big, big, s))) return rc;
      lm->pg.groups = lm->pt_groups.ptr; lm->pg.n_groups = (int)hs.point_groups.size();
      lm->pg.big_ids = lm->pt_big.ptr; lm->pg.n_big = (int)big.size();
    }
    // the frames in point-major order are a constant of the scene (tau follows from the first point pass)
    launch_point_major_obs(lm->st, h->obs_view(), nullptr, h->n_obs, nullptr, lm->pt_frame.ptr, s);
  }
  RSBA_CUDA_TRY(lm->Phi.resize((size_t)(n_inc + 1) * kPanelDoubles));
  RSBA_CUDA_TRY(cudaMemsetAsync(lm->Phi.ptr, 0, lm->Phi.bytes(), s));   // pad columns + the zero panel
  RSBA_CUDA_TRY(lm->partial.resize((size_t)std::max(n_items, 1) * kSub * kSub));
  RSBA_CUDA_TRY(lm->scale_c.resize(Fz * 12)); RSBA_CUDA_TRY(lm->scale_p.resize(Pz * 3));
  RSBA_CUDA_TRY(lm->d2_c.resize(Fz * 12)); RSBA_CUDA_TRY(lm->d2_p.resize(Pz * 3));
  RSBA_CUDA_TRY(lm->partials.resize(std::max<size_t>(chunk_frame.size(), 1) * 168));
  const size_t s_count = std::max<size_t>(nz_tiles.size(), 1) * kTile * kTile;
  lm->misc_count = 8 + h->world;
  lm->comm_count = s_count + 3 * Fz * 12 + lm->misc_count;
  RSBA_CUDA_TRY(lm->S.resize(lm->comm_count));
  RSBA_CUDA_TRY(lm->misc_local.resize(lm->misc_count));
  RSBA_CUDA_TRY(cudaMemsetAsync(lm->S.ptr, 0, lm->S.bytes(), s));
  RSBA_CUDA_TRY(cudaMemsetAsync(lm->misc_local.ptr, 0, lm->misc_local.bytes(), s));
  RSBA_CUDA_TRY(lm->Dinv.resize((size_t)std::max(T, 1) * kTile * kTile));
  RSBA_CUDA_TRY(lm->solve_partials.resize((size_t)std::max(T, 1) * 16 * kTile));
  RSBA_CUDA_TRY(lm->fwd_partials.resize(std::max<size_t>(plan.lrow_cols.size(), 1) * kTile));
  RSBA_CUDA_TRY(lm->rhs.resize(std::max<long>(lm->n_pad, 1))); RSBA_CUDA_TRY(lm->y.resize(std::max<long>(lm->n_pad, 1)));
  RSBA_CUDA_TRY(lm->delta_c.resize(Fz * 12)); RSBA_CUDA_TRY(lm->delta_p.resize(Pz * 3));
  RSBA_CUDA_TRY(lm->trial_poses.resize(Fz * 12)); RSBA_CUDA_TRY(lm->trial_points.resize(Pz * 3));
  RSBA_CUDA_TRY(lm->scalars.resize(16));
  RSBA_CUDA_TRY(cudaMemsetAsync(lm->scalars.ptr, 0, lm->scalars.bytes(), s));
  // (step partials: one triple per CTA of the point back-substitution -- per 8 points, or per group of whole points)
  RSBA_CUDA_TRY(lm->scratch.resize(std::max<size_t>(3 + 3 * ((Pz + 7) / 8) + 3 * (size_t)(lm->pg.n_groups + 8), 1024)));
  RSBA_CUDA_TRY(lm->info.resize(4));
  RSBA_CUDA_TRY(cudaMemsetAsync(lm->rhs.ptr, 0, lm->rhs.bytes(), s));
  RSBA_CUDA_TRY(cudaMemsetAsync(lm->d2_c.ptr, 0, lm->d2_c.bytes(), s));

  NormalEq& ne = lm->ne;
  ne.B = lm->B.ptr; ne.C = lm->C.ptr; ne.gp = lm->gp.ptr; ne.Bcam = lm->Bcam.ptr;
  ne.gc = lm->S.ptr + s_count; ne.wf = ne.gc + Fz * 12; ne.diagB = ne.wf + Fz * 12;
  lm->misc = ne.diagB + Fz * 12;
  ne.Cinv = lm->Cinv.ptr; ne.tp = lm->tp.ptr; ne.Minv = lm->Minv.ptr; ne.Phi = lm->Phi.ptr; ne.partial = lm->partial.ptr;
  ne.prec = lm->prec.ptr;
  ne.scale_c = lm->scale_c.ptr; ne.scale_p = lm->scale_p.ptr;
  ne.d2_c = lm->d2_c.ptr; ne.d2_p = lm->d2_p.ptr; ne.partials = lm->partials.ptr;
  ne.pose_mask = lm->pose_mask.ptr; ne.point_const = lm->point_const.ptr; ne.point_owned = lm->point_owned.ptr;
  ne.owned_ids = nullptr;
  ne.n_owned = P;
  if (h->world > 1) {   // the point kernels only walk the points this rank eliminates
    std::vector<int> ids;
    for (int p = 0; p < P; ++p)
      if (h->point_owned[p]) ids.push_back(p);
    ne.n_owned = (int)ids.size();
    if ((rc = upload(lm->owned_ids, ids, s))) return rc;
    ne.owned_ids = lm->owned_ids.ptr;
    // what this rank never computes (blocks, steps, trial values of other ranks' points) must still be defined
    RSBA_CUDA_TRY(cudaMemsetAsync(lm->C.ptr, 0, lm->C.bytes(), s)); RSBA_CUDA_TRY(cudaMemsetAsync(lm->gp.ptr, 0, lm->gp.bytes(), s));
    RSBA_CUDA_TRY(cudaMemsetAsync(lm->Cinv.ptr, 0, lm->Cinv.bytes(), s)); RSBA_CUDA_TRY(cudaMemsetAsync(lm->tp.ptr, 0, lm->tp.bytes(), s));
    RSBA_CUDA_TRY(cudaMemsetAsync(lm->Minv.ptr, 0, lm->Minv.bytes(), s)); RSBA_CUDA_TRY(cudaMemsetAsync(lm->delta_p.ptr, 0, lm->delta_p.bytes(), s));
    RSBA_CUDA_TRY(cudaMemsetAsync(lm->scale_p.ptr, 0, lm->scale_p.bytes(), s)); RSBA_CUDA_TRY(cudaMemsetAsync(lm->d2_p.ptr, 0, lm->d2_p.bytes(), s));
    RSBA_CUDA_TRY(cudaMemcpyAsync(lm->trial_points.ptr, h->d_points.ptr, 3L * P * sizeof(double), cudaMemcpyDeviceToDevice, s));
  }

  TileSchedule& ts = lm->ts;
  ts.n_tiles = T; ts.nz_tiles = lm->nz_tiles.ptr; ts.tile_slot = lm->tile_slot.ptr; ts.n_nz = (int)nz_tiles.size();
  ts.row_ptr = lm->row_ptr.ptr; ts.rows = lm->rows.ptr; ts.upd = lm->upd.ptr; ts.panels = lm->panels.ptr; ts.trsm = lm->trsm.ptr;
  ts.lrow_ptr = lm->lrow_ptr.ptr; ts.lrow_cols = lm->lrow_cols.ptr; ts.Dinv = lm->Dinv.ptr; ts.solve_partials = lm->solve_partials.ptr; ts.n_real = 12L * Fc;
  ts.fwd_slot = lm->fwd_slot.ptr; ts.fwd_partials = lm->fwd_partials.ptr;
  RSBA_CUDA_TRY(lm->dag_counters.resize(dag_counter_ints(ts)));
  RSBA_CUDA_TRY(lm->bwd_partials.resize(std::max<size_t>(plan.rows.size(), 1) * kTile));
  lm->dd = DagDevice{lm->dag_tasks.ptr, lm->n_dag_tasks, lm->n_dag_factor_tasks, lm->dag_sources.ptr, lm->dag_need.ptr,
                     lm->dag_counters.ptr, lm->bwd_partials.ptr};

  long free_params = 0;
  for (int f = 0; f < F; ++f) free_params += 12 - __builtin_popcount(h->pose_mask[f] & 0xFFF);
  if (free_cam) free_params += 9;
  if (free_ratio) free_params += 1;
  for (int p = 0; p < P; ++p) free_params += h->point_const[p] ? 0 : 3;
  lm->num_free_params = free_params;
  RSBA_CUDA_TRY(cudaStreamSynchronize(s));
  lap("upload + allocation + memset");
  return RSBA_OK;
}

int ensure_lm(rsba_problem* h, bool dense) {
  if (h->lm && h->lm->dense == dense && h->lm->reorder == h->reorder_tiles && h->lm->free_cam == h->free_cam &&
      h->lm->free_ratio == h->free_ratio) return RSBA_OK;
  if (h->lm) { lm_state_free(h->lm); h->lm = nullptr; }
  LmState* lm = new LmState;
  int rc = build_structure(h, lm, dense);
  if (rc) { delete lm; return rc; }
  h->lm = lm;
  return RSBA_OK;
}

// ------------------------------------------------------------------ pipeline stages
__global__ void pack_cost_kernel(const double* __restrict__ cost, const int* __restrict__ invalid,
                                 double* __restrict__ out) {
  out[0] = cost[0];
  out[1] = (double)invalid[0];
}

JacView jac_view(const rsba_problem* h) {
  const int rot = (h->cm.shutter != 0 && h->cm.interp_rot) ? 1 : 0;
  if (h->lm && h->lm->fused) return JacView{h->lm->rec_pt.ptr, h->lm->pt_tau.ptr, rot, 1};
  return JacView{h->d_jacc.ptr, h->d_tau.ptr, rot, 0};
}

// Normal equations + Schur complement for `radius` from the compact Jacobian in h->d_jacc (and, when
// new_jacobian, the cost / invalid count of the evaluation that produced it, in h->d_scalars).
// want_S = false: only what the convergence tests of the LAST iteration need (cost, gradient, norms) -- the Schur
// complement of a linearisation that no step will use is not formed (2.3 of 4.7 ms at C3).
int linearize(rsba_problem* h, LmState* lm, const rsba_solve_options& opt, double radius, bool new_jacobian,
              bool compute_scale, bool want_S = true) {
  cudaStream_t s = h->stream;
  const ObsView obs = h->obs_view();
  const LmOptionsDev o{radius, opt.min_lm_diagonal, opt.max_lm_diagonal};
  const JacView jv = jac_view(h);
  if (lm->fused) {
    // the two fused passes re-evaluate the functor at the current point (also for a new damping of the same
    // point: rejected steps are rare, and nothing of the old Jacobian is kept in observation order)
    cudaMemsetAsync(h->d_invalid.ptr, 0, sizeof(int), s);
    stage_begin(h, kStageJacobian);
    launch_point_pass(h->cm, lm->st, lm->pt_packed.ptr, h->d_poses.ptr, h->d_points.ptr, lm->ne, o, compute_scale,
                      opt.jacobi_scaling != 0, lm->rec_pt.ptr, lm->pt_tau.ptr, lm->xt.ptr, want_S, h->n_obs, h->n_points, s);
    stage_end(h, kStageJacobian);
    stage_begin(h, kStageSchur);
    stage_begin(h, kStageFrameBlocks);
    launch_frame_pass(h->cm, lm->st, obs, h->d_poses.ptr, lm->xt.ptr, lm->ne, h->d_cost_partials.ptr, h->d_invalid.ptr, s);
    launch_frame_reduce(lm->st, lm->ne, lm->n_cam_frames, s);
    stage_end(h, kStageFrameBlocks);
    int rc = eval_tail(h, true, h->d_poses.ptr, lm->st.n_chunks);   // priors' residuals + cost -> d_scalars[0]
    if (rc) return rc;
    h->launches += 4;
  } else {
  stage_begin(h, kStageSchur);
  if (new_jacobian && !lm->pt_major_valid) {   // tau is a constant of the observation: once per solve
    launch_point_major_obs(lm->st, obs, h->d_tau.ptr, h->n_obs, lm->pt_tau.ptr, lm->pt_frame.ptr, s);
    lm->pt_major_valid = true;
    h->launches += 1;
  }
  if (new_jacobian) {
    stage_begin(h, kStagePointBlocks);
    launch_point_blocks(lm->st, jv, h->d_res.ptr, lm->ne, s);
    stage_end(h, kStagePointBlocks);
    h->launches += 1;
  }
  if (compute_scale) { launch_jacobi_scale(0, true, lm->ne, opt.jacobi_scaling != 0, s); h->launches += 1; }
  launch_point_invert(lm->ne, o, s);
  stage_begin(h, kStageFrameBlocks);
  launch_frame_blocks(lm->st, obs, jv, h->d_res.ptr, lm->n_cam_frames, lm->ne, s);
  stage_end(h, kStageFrameBlocks);
  h->launches += 3;
  }
  if (lm->free_cam) {   // blocks of the intrinsics pseudo-frame and its couplings with the frames
    launch_cam_blocks(lm->st, obs, jv, h->d_jac_cam.ptr, h->d_res.ptr, lm->ne, h->n_frames,
                      lm->cam_partials.ptr, lm->cam_scratch.ptr, s);
    h->launches += 3;
  }
  const PriorView pv = h->prior_view();
  if (pv.n > 0 && h->rank == 0) {   // camera-only residual blocks: added once, not sharded
    launch_prior_blocks(pv, lm->ne, h->n_frames, s);
    h->launches += 1;
  }
  const PosePriorView ppv = h->pose_prior_view();
  if (ppv.n > 0) {   // GoodPosePrior blocks: eliminated in closed form; their terms are added once (rank 0)
    launch_pose_prior_blocks(ppv, lm->ne, o, opt.jacobi_scaling != 0, h->rank == 0, s);
    h->launches += 1;
  }
  PriorView pvr = pv;
  if (h->rank != 0) pvr.n = 0;
  if (want_S) {
  launch_clear_tiles(lm->S.ptr, lm->ts, s);
  stage_begin(h, kStagePhiBuild);
  launch_phi_build(lm->st, obs, jv, lm->ne, s);
  stage_end(h, kStagePhiBuild);
  if (lm->free_cam) {
    launch_phi_cam(lm->st, jv, h->d_jac_cam.ptr, lm->ne, h->n_points, h->n_frames, s);
    h->launches += 1;
  }
  stage_begin(h, kStageSchurSyrk);
  launch_schur_syrk(lm->st, lm->ne, s);
  stage_end(h, kStageSchurSyrk);
  stage_begin(h, kStageSchurReduce);
  const int cam_frame_idx = lm->n_cam_frames > h->n_frames ? h->n_frames : -1;
  launch_schur_reduce(lm->st, lm->ne, pvr, cam_frame_idx, lm->S.ptr, lm->ts.tile_slot, lm->ts.n_tiles, s);
  stage_end(h, kStageSchurReduce);
  h->launches += 4;
  }
  if (new_jacobian) {   // scalars of the current point that ride in the same buffer
    pack_cost_kernel<<<1, 1, 0, s>>>(h->d_scalars.ptr, h->d_invalid.ptr, lm->misc_local.ptr);
    launch_point_norms(lm->ne, h->n_points, h->d_points.ptr, lm->misc_local.ptr + 2, lm->misc_local.ptr + 8 + h->rank,
                       lm->scratch.ptr, s);
    h->launches += 3;
  }
  cudaMemcpyAsync(lm->misc, lm->misc_local.ptr, lm->misc_count * sizeof(double), cudaMemcpyDeviceToDevice, s);
  stage_end(h, kStageSchur);
  if (h->world > 1) {   // the one exchange step: partial S | gc | wf | diag(B) | scalars
    stage_begin(h, kStageAllreduce);
    // (without S: only the tail  gc | wf | diag(B) | scalars)
    const size_t tail = 3 * (size_t)std::max(lm->n_cam_frames, 1) * 12 + lm->misc_count;
    int rc = want_S ? allreduce_sum(h, lm->S.ptr, lm->comm_count) : allreduce_sum(h, lm->S.ptr + (lm->comm_count - tail), tail);
    stage_end(h, kStageAllreduce);
    if (rc) return rc;
  }
  stage_begin(h, kStageFinalize);
  if (compute_scale) { launch_jacobi_scale(lm->n_cam_frames, false, lm->ne, opt.jacobi_scaling != 0, s); h->launches += 1; }
  if (want_S) launch_schur_finalize(lm->st, lm->ne, o, lm->S.ptr, lm->ts, lm->n_cam_frames, lm->rhs.ptr, s);
  launch_camera_norms(lm->ne, lm->n_cam_frames, h->d_poses.ptr, lm->scalars.ptr, s);
  if (ppv.n > 0) { launch_pose_prior_norms(ppv, lm->scalars.ptr, s); h->launches += 1; }
  h->launches += 3;
  stage_end(h, kStageFinalize);
  return RSBA_OK;
}

// Capture `body` (a fixed sequence of launches on stream s) into an executable graph.
template <typename F>
cudaGraphExec_t capture_graph(cudaStream_t s, F body, int* launches) {
  if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  *launches = body();
  cudaGraph_t g = nullptr;
  cudaGraphExec_t exec = nullptr;
  if (cudaStreamEndCapture(s, &g) != cudaSuccess || !g) {
    cudaGetLastError();
    return nullptr;
  }
  if (cudaGraphInstantiate(&exec, g, 0) != cudaSuccess) {
    cudaGetLastError();
    exec = nullptr;
  }
  cudaGraphDestroy(g);
  return exec;
}

void factor_and_solve(rsba_problem* h, LmState* lm) {
  cudaStream_t s = h->stream;
  if (lm->use_dag) {   // one persistent kernel: factorisation, forward and backward substitution
    stage_begin(h, kStageCholesky);
    cudaMemsetAsync(lm->info.ptr, 0, sizeof(int), s);
    cudaMemcpyAsync(lm->y.ptr, lm->rhs.ptr, lm->n_pad * sizeof(double), cudaMemcpyDeviceToDevice, s);
    if (lm->dag_split) {   // RSBA_CUDA_K3_SPLIT=1: two launches, so that the stage timers can tell the parts apart
      stage_begin(h, kStageFactor);
      h->launches += launch_tile_dag(lm->S.ptr, lm->ts, lm->dd, lm->y.ptr, lm->info.ptr, true, false, s);
      stage_end(h, kStageFactor);
      stage_begin(h, kStageTriSolve);
      h->launches += launch_tile_dag(lm->S.ptr, lm->ts, lm->dd, lm->y.ptr, lm->info.ptr, false, true, s);
      stage_end(h, kStageTriSolve);
    } else {
      h->launches += launch_tile_dag(lm->S.ptr, lm->ts, lm->dd, lm->y.ptr, lm->info.ptr, true, true, s);
    }
    stage_end(h, kStageCholesky);
    return;
  }
  if (!lm->graph_tried) {
    lm->graph_tried = true;
    k3_prepare();   // the kernels' one-off attribute setup stays outside the capture
    lm->graph_factor = capture_graph(s, [&] { return launch_tile_cholesky(lm->S.ptr, lm->ts, lm->plan, lm->y.ptr, lm->info.ptr, s); },
                                     &lm->graph_factor_launches);
    if (lm->graph_factor)
      lm->graph_solve = capture_graph(s, [&] { return launch_tile_solve(lm->S.ptr, lm->ts, lm->plan, lm->y.ptr, s); },
                                      &lm->graph_solve_launches);
  }
  stage_begin(h, kStageCholesky);
  cudaMemsetAsync(lm->info.ptr, 0, sizeof(int), s);
  cudaMemcpyAsync(lm->y.ptr, lm->rhs.ptr, lm->n_pad * sizeof(double), cudaMemcpyDeviceToDevice, s);
  stage_begin(h, kStageFactor);
  if (lm->graph_factor && cudaGraphLaunch(lm->graph_factor, s) == cudaSuccess) h->launches += lm->graph_factor_launches;
  else h->launches += launch_tile_cholesky(lm->S.ptr, lm->ts, lm->plan, lm->y.ptr, lm->info.ptr, s);
  stage_end(h, kStageFactor);
  stage_begin(h, kStageTriSolve);
  if (lm->graph_solve && cudaGraphLaunch(lm->graph_solve, s) == cudaSuccess) h->launches += lm->graph_solve_launches;
  else h->launches += launch_tile_solve(lm->S.ptr, lm->ts, lm->plan, lm->y.ptr, s);
  stage_end(h, kStageTriSolve);
  stage_end(h, kStageCholesky);
}

void step_update(rsba_problem* h, LmState* lm) {
  stage_begin(h, kStageUpdate);
  launch_step_update(lm->st, h->obs_view(), jac_view(h), lm->free_cam ? h->d_jac_cam.ptr : nullptr,
                     lm->free_cam ? h->n_frames : -1, lm->ne, lm->y.ptr, lm->n_cam_frames, h->n_points,
                     h->d_poses.ptr, h->d_points.ptr, lm->delta_c.ptr, lm->delta_p.ptr, lm->trial_poses.ptr,
                     lm->trial_points.ptr, lm->scalars.ptr, lm->scratch.ptr,
                     lm->free_ratio ? 12 * h->n_frames + 9 : -1, h->ratio_lower_bound(), lm->pg, h->stream);
  h->launches += 3;
  const PosePriorView ppv = h->pose_prior_view();
  if (ppv.n > 0) { launch_pose_prior_step(ppv, lm->delta_c.ptr, lm->scalars.ptr, h->stream); h->launches += 1; }
  stage_end(h, kStageUpdate);
}

// cost at the trial point (K1r) + the step scalars, summed over the ranks
int trial_cost(rsba_problem* h, LmState* lm) {
  int rc = run_evaluate(h, false, lm->trial_poses.ptr, lm->trial_points.ptr, nullptr, nullptr);
  if (rc) return rc;
  pack_cost_kernel<<<1, 1, 0, h->stream>>>(h->d_scalars.ptr, h->d_invalid.ptr, lm->scalars.ptr + 11);
  h->launches += 1;
  return allreduce_sum(h, lm->scalars.ptr + 8, 8);
}

struct HostScalars {
  double s[16];     // lm->scalars: [0..2] camera g.d, D^2 d^2, |d|^2  [3] |x_c|^2 [4] max|g_c|
                    //              [8..10] the same three over the points  [11] trial cost [12] trial invalid
  double misc[8];   // tail of the reduced-system buffer: [0] cost [1] invalid [2] |x_p|^2
  double gmax_p;    // max over the ranks' slots
  int info;
  // derived
  double cost, x_norm, gmax, g_dot_delta, d2_delta2, step_norm, trial_cost;
  long invalid, trial_invalid;
};

int fetch(rsba_problem* h, LmState* lm, HostScalars* out) {
  cudaStream_t s = h->stream;
  std::vector<double> misc(lm->misc_count);
  RSBA_CUDA_TRY(cudaMemcpyAsync(out->s, lm->scalars.ptr, 16 * sizeof(double), cudaMemcpyDeviceToHost, s));
  RSBA_CUDA_TRY(cudaMemcpyAsync(misc.data(), lm->misc, lm->misc_count * sizeof(double), cudaMemcpyDeviceToHost, s));
  RSBA_CUDA_TRY(cudaMemcpyAsync(&out->info, lm->info.ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
  RSBA_CUDA_TRY(cudaStreamSynchronize(s));
  RSBA_CUDA_TRY(cudaGetLastError());
  for (int k = 0; k < 8; ++k) out->misc[k] = misc[k];
  out->gmax_p = 0.0;
  for (int r = 0; r < h->world; ++r) out->gmax_p = std::max(out->gmax_p, misc[8 + r]);
  out->cost = misc[0];
  out->invalid = (long)(misc[1] + 0.5);
  out->x_norm = std::sqrt(out->s[3] + misc[2]);
  out->gmax = std::max(out->s[4], out->gmax_p);
  out->g_dot_delta = out->s[0] + out->s[8];
  out->d2_delta2 = out->s[1] + out->s[9];
  out->step_norm = std::sqrt(out->s[2] + out->s[10]);
  out->trial_cost = out->s[11];
  out->trial_invalid = (long)(out->s[12] + 0.5);
  for (int k = 0; k < kNumStages; ++k) stage_collect(h, (Stage)k);
  return RSBA_OK;
}

__global__ void mask_unowned_kernel(const double* __restrict__ points, const unsigned char* __restrict__ owned,
                                    long n, double* __restrict__ out) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = owned[t / 3] ? points[t] : 0.0;
}

// Every rank moves only the points it owns; at the end of a solve the owners' values are summed
// into every rank's copy (each point has exactly one owner).
int gather_points(rsba_problem* h, LmState* lm) {
  if (h->world <= 1 || h->n_points == 0) return RSBA_OK;
  const long n = 3L * h->n_points;
  mask_unowned_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_points.ptr, lm->point_owned.ptr, n,
                                                                       lm->trial_points.ptr);
  h->launches += 1;
  int rc = allreduce_sum(h, lm->trial_points.ptr, (size_t)n);
  if (rc) return rc;
  RSBA_CUDA_TRY(cudaMemcpyAsync(h->d_points.ptr, lm->trial_points.ptr, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  return RSBA_OK;
}

int prepare_solve(rsba_problem* h, const rsba_solve_options* opt) {
  if (!h) return fail(RSBA_ERR_INVALID_ARGUMENT, "handle is NULL");
  if (!opt) return fail(RSBA_ERR_INVALID_ARGUMENT, "options are NULL");
  if (!h->camera_set) return fail(RSBA_ERR_STATE, "rsba_cuda_set_camera has not been called");
  if (opt->huber_loss < 0.0) return fail(RSBA_ERR_INVALID_ARGUMENT, "huber_loss must be >= 0");
  if (opt->huber_loss > 0.0) h->cm.huber = opt->huber_loss;   // same as rsba_cuda_set_loss
  RSBA_CUDA_TRY(cudaSetDevice(h->device));
  if (h->ptr_mode) {
    int rc = finalize_pointer_problem(h);
    if (rc) return rc;
    rc = gather_pointer_parameters(h);
    if (rc) return rc;
  }
  if (!h->scene_set) return fail(RSBA_ERR_STATE, "no residual blocks");
  if (!h->params_set) return fail(RSBA_ERR_STATE, "rsba_cuda_set_parameters has not been called");
  int rc = upload_priors(h);
  if (rc) return rc;
  rc = upload_pose_priors(h);
  if (rc) return rc;
  h->reorder_tiles = opt->reorder_tiles != 0;
  rc = ensure_eval_buffers(h, false);
  if (rc) return rc;
  rc = ensure_lm(h, opt->dense_cholesky != 0);
  if (rc) return rc;
  h->lm->pt_major_valid = false;   // the camera model (hence tau) may have changed since the last solve
  return ensure_eval_buffers(h, !h->lm->fused, true);
}

}  // namespace
}  // namespace rsba

using namespace rsba;

extern "C" {

int rsba_cuda_solve(rsba_problem* h, const rsba_solve_options* opt, rsba_solve_summary* sum) {
  return rsba::api_guard([&]() -> int {
  const auto t_begin = std::chrono::steady_clock::now();
  rsba_solve_summary local;
  if (!sum) sum = &local;
  memset(sum, 0, sizeof(*sum));
  sum->termination = 2;
  int rc = prepare_solve(h, opt);
  if (rc) { snprintf(sum->message, sizeof(sum->message), "%s", rsba_cuda_last_error()); return rc; }
  LmState* lm = h->lm;
  h->fine_timers = getenv("RSBA_CUDA_TRACE") != nullptr || getenv("RSBA_CUDA_FINE_TIMERS") != nullptr;
  for (auto& t : h->timers) t.total_ms = 0.0;
  sum->num_residual_blocks = h->n_obs;
  sum->num_parameters_reduced = lm->num_free_params + h->free_pose_prior_params();
  sum->tile_flops = lm->plan.flops;
  sum->reduced_levels = lm->plan.n_levels;
  sum->reduced_tiles = (int)lm->plan.nz_tiles.size();

  auto finish = [&](int term, const char* msg, double cost, double radius, double gmax) {
    sum->termination = term;
    sum->usable = term != 2;
    sum->final_cost = cost;
    sum->final_radius = radius;
    sum->final_gradient_max_norm = gmax;
    snprintf(sum->message, sizeof(sum->message), "%s", msg);
    sum->time_jacobian_ms = h->timers[kStageJacobian].total_ms;
    sum->time_residual_ms = h->timers[kStageResidual].total_ms;
    sum->time_schur_ms = h->timers[kStageSchur].total_ms;
    sum->time_cholesky_ms = h->timers[kStageCholesky].total_ms;
    sum->time_update_ms = h->timers[kStageUpdate].total_ms;
    sum->time_allreduce_ms = h->timers[kStageAllreduce].total_ms;
  };
  auto wall = [&]() {
    sum->time_total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  };

  HostScalars hs{};
  double radius = opt->initial_trust_region_radius, decrease = RSBA_CERES_LM_INITIAL_DECREASE_FACTOR;
  double cost = 0.0, x_norm = 0.0, gmax = 0.0;
  int invalid_in_a_row = 0, fail_rc = RSBA_OK;
  const int max_invalid = opt->max_num_consecutive_invalid_steps > 0 ? opt->max_num_consecutive_invalid_steps
                                                                      : RSBA_CERES_MAX_NUM_CONSECUTIVE_INVALID_STEPS;
  // ---- iteration 0: evaluate, linearise (fixes the Jacobi scaling), gradient check
  if (!lm->fused) {
    rc = run_evaluate(h, true, h->d_poses.ptr, h->d_points.ptr, nullptr, nullptr, true);
    if (rc) return rc;
  }
  sum->num_jacobian_evaluations = 1;
  if ((rc = linearize(h, lm, *opt, radius, true, true))) return rc;
  if ((rc = fetch(h, lm, &hs))) return rc;
  cost = hs.cost; x_norm = hs.x_norm; gmax = hs.gmax;
  sum->initial_cost = cost;
  if (hs.invalid > 0) {
    finish(2, "FAILURE: residual evaluation failed at the initial point (point behind a camera)", cost, radius, gmax);
    wall();
    return fail(RSBA_ERR_EVALUATION_FAILED, sum->message);
  }
  if (opt->verbose && h->rank == 0)
    printf("iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius\n%4d % .6e                % .2e                        % .2e\n",
           0, cost, gmax, radius);
  if (gmax <= opt->gradient_tolerance) {
    finish(0, "CONVERGENCE: gradient tolerance reached", cost, radius, gmax);
  } else {
    int it = 0;
    while (true) {
      if (it >= opt->max_num_iterations) {
        finish(1, "NO_CONVERGENCE: maximum number of iterations reached", cost, radius, gmax);
        break;
      }
      ++it;
      sum->iterations = it;
      factor_and_solve(h, lm);
      step_update(h, lm);
      if ((rc = trial_cost(h, lm))) return rc;
      sum->num_residual_evaluations++;
      if ((rc = fetch(h, lm, &hs))) return rc;
      const double mcc = -0.5 * hs.g_dot_delta + 0.5 * hs.d2_delta2;
      const double step_norm = hs.step_norm;
      // Ceres 1.9 (trust_region_minimizer.cc): a failed linear solve, a non-finite step or a step that does not
      // decrease the model is an INVALID step -- radius shrinks like after a rejected step and the loop goes
      // on; only max_num_consecutive_invalid_steps of them in a row end the solve
      const bool invalid = hs.info != 0 || !std::isfinite(step_norm) || !std::isfinite(mcc) || !(mcc > 0.0);
      if (invalid) {
        ++invalid_in_a_row;
        sum->num_unsuccessful_steps++;
        if (opt->verbose && h->rank == 0)
          printf("%4d % .6e  invalid step (%s), radius % .2e\n", it, cost,
                 hs.info != 0 ? "reduced camera matrix not positive definite" : "model cost does not decrease", radius);
        if (invalid_in_a_row >= max_invalid) {
          finish(2, hs.info != 0 ? "FAILURE: reduced camera matrix is not positive definite (successive invalid steps)"
                                 : "FAILURE: number of successive invalid steps exceeds max_num_consecutive_invalid_steps",
                 cost, radius, gmax);
          fail_rc = fail(RSBA_ERR_LINEAR_SOLVER, sum->message);
          break;
        }
        radius /= decrease;
        decrease *= 2.0;
        if (radius < opt->min_trust_region_radius) {
          finish(0, "CONVERGENCE: trust region radius below minimum", cost, radius, gmax);
          break;
        }
        if ((rc = linearize(h, lm, *opt, radius, false, false))) return rc;   // same Jacobian, new damping
        continue;
      }
      invalid_in_a_row = 0;
      bool accepted = false;
      double rho = 0.0;
      const double new_cost = hs.trial_cost;
      if (hs.trial_invalid == 0) {
        if (step_norm <= opt->parameter_tolerance * (x_norm + opt->parameter_tolerance)) {
          finish(0, "CONVERGENCE: parameter tolerance reached", cost, radius, gmax);
          break;
        }
        if (std::fabs(cost - new_cost) < opt->function_tolerance * cost) {
          finish(0, "CONVERGENCE: function tolerance reached", cost, radius, gmax);
          break;
        }
        rho = (cost - new_cost) / mcc;
        accepted = rho > opt->min_relative_decrease;
        // what Ceres' bounded-problem line search tests at step size 1 (rsba_solve_summary::num_armijo_violations)
        if (lm->free_ratio && !(new_cost <= cost + RSBA_CERES_ARMIJO_SUFFICIENT_DECREASE * hs.g_dot_delta))
          sum->num_armijo_violations++;
      }
      if (accepted) {
        sum->num_successful_steps++;
        cudaMemcpyAsync(h->d_poses.ptr, lm->trial_poses.ptr, 12L * lm->n_cam_frames * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
        cudaMemcpyAsync(h->d_points.ptr, lm->trial_points.ptr, h->d_points.bytes(), cudaMemcpyDeviceToDevice, h->stream);
        if (!h->pose_priors.empty())
          cudaMemcpyAsync(h->d_pp_val.ptr, h->d_pp_trial.ptr, h->d_pp_val.bytes(), cudaMemcpyDeviceToDevice, h->stream);
        radius = std::min(opt->max_trust_region_radius,
                          radius / std::max(RSBA_CERES_LM_MIN_RADIUS_SHRINK, 1.0 - std::pow(2.0 * rho - 1.0, 3)));
        decrease = RSBA_CERES_LM_INITIAL_DECREASE_FACTOR;
        if (!lm->fused) {
          rc = run_evaluate(h, true, h->d_poses.ptr, h->d_points.ptr, nullptr, nullptr, true);
          if (rc) return rc;
        }
        sum->num_jacobian_evaluations++;
        if ((rc = linearize(h, lm, *opt, radius, true, false, it < opt->max_num_iterations))) return rc;
        if ((rc = fetch(h, lm, &hs))) return rc;
        const double old = cost;
        cost = hs.cost; x_norm = hs.x_norm; gmax = hs.gmax;
        if (opt->verbose && h->rank == 0)
          printf("%4d % .6e  % .2e  % .2e  % .2e  % .2e  % .2e\n", it, cost, old - cost, gmax, step_norm, rho, radius);
        if (gmax <= opt->gradient_tolerance) {
          finish(0, "CONVERGENCE: gradient tolerance reached", cost, radius, gmax);
          break;
        }
      } else {
        sum->num_unsuccessful_steps++;
        radius /= decrease;
        decrease *= 2.0;
        if (opt->verbose && h->rank == 0)
          printf("%4d % .6e  % .2e  % .2e  % .2e  % .2e  % .2e (rejected)\n", it, cost, 0.0, gmax, step_norm, rho, radius);
        if (radius < opt->min_trust_region_radius) {
          finish(0, "CONVERGENCE: trust region radius below minimum", cost, radius, gmax);
          break;
        }
        if ((rc = linearize(h, lm, *opt, radius, false, false))) return rc;   // same Jacobian, new damping
      }
    }
  }
  if ((rc = gather_points(h, lm))) return rc;
  if (lm->free_cam)   // the optimised intrinsics are also the host copy used by validate / pnp and returned by get_camera
    RSBA_CUDA_TRY(cudaMemcpyAsync(h->cm.cam, h->d_poses.ptr + 12L * h->n_frames, 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (lm->free_ratio)
    RSBA_CUDA_TRY(cudaMemcpyAsync(&h->ratio_value, h->d_poses.ptr + 12L * h->n_frames + 9, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  RSBA_CUDA_TRY(cudaStreamSynchronize(h->stream));
  sum->time_schur_ms += h->timers[kStageFinalize].total_ms;
  if (h->ptr_mode && h->scatter_owner) {
    rc = scatter_pointer_parameters(h);
    if (rc) return rc;
  }
  wall();
  return fail_rc;   // a failed solve still leaves the last accepted iterate in the caller's blocks
  });
}

int rsba_cuda_linearize_and_step(rsba_problem* h, const rsba_solve_options* opt, double radius, double* S_out,
                                 double* rhs_out, double* delta_poses, double* delta_points,
                                 double* model_cost_change) {
  return rsba::api_guard([&]() -> int {
  int rc = prepare_solve(h, opt);
  if (rc) return rc;
  if (!(radius > 0.0)) return fail(RSBA_ERR_INVALID_ARGUMENT, "radius must be positive");
  LmState* lm = h->lm;
  h->fine_timers = true;
  if (!lm->fused) {
    rc = run_evaluate(h, true, h->d_poses.ptr, h->d_points.ptr, nullptr, nullptr, true);
    if (rc) return rc;
  }
  if ((rc = linearize(h, lm, *opt, radius, true, true))) return rc;
  const long n = 12L * lm->n_cam_frames;   // uncalibrated variant: the intrinsics pseudo-frame comes last
  RSBA_CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (S_out) {
    std::vector<double> tile((size_t)kTile * kTile);
    memset(S_out, 0, sizeof(double) * n * n);
    for (size_t sidx = 0; sidx < lm->plan.nz_tiles.size(); ++sidx) {
      const int2 t = lm->plan.nz_tiles[sidx];
      RSBA_CUDA_TRY(cudaMemcpy(tile.data(), lm->S.ptr + sidx * kTile * kTile, tile.size() * sizeof(double), cudaMemcpyDeviceToHost));
      for (int r = 0; r < kTile; ++r)
        for (int c = 0; c < kTile; ++c) {
          // tile positions -> frame tiles (the caller sees the un-permuted system)
          const long gr = (long)lm->plan.pos_tile[t.x] * kTile + r, gcol = (long)lm->plan.pos_tile[t.y] * kTile + c;
          if (gr >= n || gcol >= n) continue;
          if (t.x == t.y && c > r) continue;   // diagonal tiles: the factorisation reads the lower triangle
          S_out[gr * n + gcol] = tile[r * kTile + c];
          S_out[gcol * n + gr] = tile[r * kTile + c];
        }
    }
  }
  if (rhs_out) {
    std::vector<double> tmp((size_t)lm->n_pad);
    RSBA_CUDA_TRY(cudaMemcpy(tmp.data(), lm->rhs.ptr, lm->n_pad * sizeof(double), cudaMemcpyDeviceToHost));
    for (long k = 0; k < n; ++k)   // S delta_c' = rhs, un-permuted
      rhs_out[k] = -tmp[(size_t)lm->plan.tile_pos[k / kTile] * kTile + k % kTile];
  }
  factor_and_solve(h, lm);
  step_update(h, lm);
  if ((rc = allreduce_sum(h, lm->scalars.ptr + 8, 8))) return rc;
  HostScalars hs{};
  if ((rc = fetch(h, lm, &hs))) return rc;
  if (hs.info != 0) return fail(RSBA_ERR_LINEAR_SOLVER, "reduced camera matrix is not positive definite");
  if (delta_poses) RSBA_CUDA_TRY(cudaMemcpy(delta_poses, lm->delta_c.ptr, n * sizeof(double), cudaMemcpyDeviceToHost));
  if (delta_points)
    RSBA_CUDA_TRY(cudaMemcpy(delta_points, lm->delta_p.ptr, 3L * h->n_points * sizeof(double), cudaMemcpyDeviceToHost));
  if (model_cost_change) *model_cost_change = -0.5 * hs.g_dot_delta + 0.5 * hs.d2_delta2;
  return RSBA_OK;
  });
}

int rsba_cuda_plan_reduced_system(int n_tiles, int n_pairs, const int* pair_a, const int* pair_b, int dense,
                                  int reorder, long counts[6], int* tile_pos, int* nz_tiles, int* panels,
                                  int* panel_ptr, int* trsm, int* trsm_ptr, int* upd, long* group_ptr,
                                  int* level_group_ptr) {
  return rsba::api_guard([&]() -> int {
  if (n_tiles < 0 || n_pairs < 0 || (n_pairs > 0 && (!pair_a || !pair_b)) || !counts)
    return fail(RSBA_ERR_INVALID_ARGUMENT, "bad plan arguments");
  std::vector<std::pair<int, int>> tp;
  for (int k = 0; k < n_pairs; ++k) {
    if (pair_a[k] < 0 || pair_b[k] < 0 || pair_a[k] >= n_tiles || pair_b[k] >= n_tiles)
      return fail(RSBA_ERR_INVALID_ARGUMENT, "tile index out of range");
    tp.emplace_back(std::min(pair_a[k], pair_b[k]), std::max(pair_a[k], pair_b[k]));
  }
  TilePlan plan;
  build_tile_plan(n_tiles, tp, dense != 0, reorder != 0, -1, &plan);
  counts[0] = plan.n_levels; counts[1] = (long)plan.nz_tiles.size(); counts[2] = (long)plan.trsm.size();
  counts[3] = (long)plan.upd.size(); counts[4] = (long)plan.group_ptr.size() - 1; counts[5] = (long)plan.flops;
  if (tile_pos) std::copy(plan.tile_pos.begin(), plan.tile_pos.begin() + n_tiles, tile_pos);
  if (nz_tiles) for (size_t k = 0; k < plan.nz_tiles.size(); ++k) { nz_tiles[2 * k] = plan.nz_tiles[k].x; nz_tiles[2 * k + 1] = plan.nz_tiles[k].y; }
  if (panels) std::copy(plan.panels.begin(), plan.panels.begin() + n_tiles, panels);
  if (panel_ptr) std::copy(plan.panel_ptr.begin(), plan.panel_ptr.end(), panel_ptr);
  if (trsm) for (size_t k = 0; k < plan.trsm.size(); ++k) { trsm[2 * k] = plan.trsm[k].x; trsm[2 * k + 1] = plan.trsm[k].y; }
  if (trsm_ptr) std::copy(plan.trsm_ptr.begin(), plan.trsm_ptr.end(), trsm_ptr);
  if (upd) for (size_t k = 0; k < plan.upd.size(); ++k) { upd[3 * k] = plan.upd[k].x; upd[3 * k + 1] = plan.upd[k].y; upd[3 * k + 2] = plan.upd[k].z; }
  if (group_ptr) std::copy(plan.group_ptr.begin(), plan.group_ptr.end(), group_ptr);
  if (level_group_ptr) std::copy(plan.level_group_ptr.begin(), plan.level_group_ptr.end(), level_group_ptr);
  return RSBA_OK;
  });
}


}  // extern "C"
