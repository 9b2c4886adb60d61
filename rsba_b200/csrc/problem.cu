// C ABI: lifecycle, problem construction, evaluation (include/rsba_cuda.h).
#include "problem.cuh"

#include <mutex>
#include "nccl_dl.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

namespace rsba {

static thread_local std::string g_last_error;

void set_last_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file,
           line, what);
  set_last_error(buf);
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? RSBA_ERR_NO_DEVICE
                                                                      : RSBA_ERR_CUDA;
}

static int fail(int code, const std::string& msg) {
  set_last_error(msg);
  return code;
}

// RSBA_CUDA_TRACE=1: wall-clock of the one-off host phases, to stderr (as lm_solver.cu prints the structure analysis)
struct TraceLap {
  const bool on = getenv("RSBA_CUDA_TRACE") != nullptr;
  std::chrono::steady_clock::time_point prev = std::chrono::steady_clock::now();
  void operator()(const char* what) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[rsba_cuda] scene:     %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - prev).count());
    prev = now;
  }
};

const NcclApi* nccl_api() {
  static NcclApi api;
  static int state = 0;   // 0 untried, 1 ok, -1 failed
  if (state == 0) {
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
      set_last_error(std::string("cannot load libnccl.so.2: ") + dlerror());
      state = -1;
    } else {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(lib, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(lib, "ncclCommDestroy");
      api.AllReduce = (decltype(api.AllReduce))dlsym(lib, "ncclAllReduce");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(lib, "ncclGetErrorString");
      const bool ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
      if (!ok) set_last_error("libnccl.so.2 lacks a required symbol");
      state = ok ? 1 : -1;
    }
  }
  return state == 1 ? &api : nullptr;
}

bool device_pool_enabled() {
  static const bool on = getenv("RSBA_CUDA_NO_POOL") == nullptr;
  if (!on) return false;
  static bool prepared[64] = {};
  static std::mutex mu;      // rsba_cuda_multi_solve runs the ranks on worker threads
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  std::lock_guard<std::mutex> lock(mu);
  if (!prepared[dev & 63]) {
    int supported = 0;
    cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, dev);
    cudaMemPool_t pool;
    if (!supported || cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    unsigned long long keep = ~0ull;    // never trim: freed blocks serve the next handle
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    prepared[dev & 63] = true;
  }
  return true;
}

int allreduce_sum(rsba_problem* h, double* buf, size_t count) {
  if (h->world <= 1 || count == 0) return RSBA_OK;
  const NcclApi* api = nccl_api();
  if (!api || !h->nccl_comm) return fail(RSBA_ERR_NCCL, "no NCCL communicator (rsba_cuda_comm_init)");
  ncclResult_t r = api->AllReduce(buf, buf, count, ncclDouble, ncclSum, (ncclComm_t)h->nccl_comm, h->stream);
  if (r != ncclSuccess) return fail(RSBA_ERR_NCCL, std::string("ncclAllReduce: ") + api->GetErrorString(r));
  return RSBA_OK;
}

static bool stage_on(const rsba_problem* h, Stage s) { return s < kStagePointBlocks || s == kStagePnp || s == kStageFinalize || h->fine_timers; }

void stage_begin(rsba_problem* h, Stage s) {
  if (!stage_on(h, s)) return;
  StageTimer& t = h->timers[s];
  if (!t.beg) {
    cudaEventCreate(&t.beg);
    cudaEventCreate(&t.end);
  }
  cudaEventRecord(t.beg, h->stream);
}

void stage_end(rsba_problem* h, Stage s) {
  if (!stage_on(h, s)) return;
  StageTimer& t = h->timers[s];
  cudaEventRecord(t.end, h->stream);
  t.pending = true;
}

double stage_collect(rsba_problem* h, Stage s) {
  StageTimer& t = h->timers[s];
  if (t.pending) {
    cudaEventSynchronize(t.end);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t.beg, t.end);
    t.last_ms = ms;
    t.total_ms += ms;
    t.pending = false;
  }
  return t.last_ms;
}

int ensure_eval_buffers(rsba_problem* h, bool jac, bool compact) {
  const size_t n = (size_t)h->n_obs;
  RSBA_CUDA_TRY(h->d_res.resize(2 * n));
  RSBA_CUDA_TRY(h->d_valid.resize(n));
  // one partial per K1 CTA, or per frame chunk of the fused frame pass (<= one extra chunk per frame), + the priors' two
  RSBA_CUDA_TRY(h->d_cost_partials.resize((size_t)k1_num_partials(h->n_obs) + (size_t)h->n_frames + 4));
  RSBA_CUDA_TRY(h->d_scalars.resize(16));
  RSBA_CUDA_TRY(h->d_invalid.resize(4));
  if (jac && !compact) RSBA_CUDA_TRY(h->d_jac.resize((size_t)kJacDoubles * n));
  if (jac && compact) {
    RSBA_CUDA_TRY(h->d_jacc.resize((size_t)kJacCompact * n));
    RSBA_CUDA_TRY(h->d_tau.resize(n));
  }
  if (jac && h->free_cam) RSBA_CUDA_TRY(h->d_jac_cam.resize((size_t)18 * n));
  return RSBA_OK;
}

// Camera-only blocks and the final sum of an evaluation whose observation kernel has left `np` cost partials in
// d_cost_partials: the priors' cost rides in the two extra slots, then one fixed-order sum -> d_scalars[0].
int eval_tail(rsba_problem* h, bool store, const double* poses, int np) {
  // the priors' cost rides in the extra partial slot (rank 0 only: they are not sharded)
  const PriorView pv = h->prior_view();
  if (pv.n > 0 && h->rank == 0) {
    launch_prior_eval(pv, poses, h->cm.huber, h->d_cost_partials.ptr + np, store, h->d_invalid.ptr, h->stream);
    h->launches += 1;
  } else {
    RSBA_CUDA_TRY(cudaMemsetAsync(h->d_cost_partials.ptr + np, 0, sizeof(double), h->stream));
  }
  // ... and the pose priors' (GoodPosePrior) in the next one; they go with the point the poses belong to
  const PosePriorView ppv = h->pose_prior_view();
  if (ppv.n > 0) {
    // every rank needs the residuals (the prior blocks are back-substituted everywhere, identically); the cost
    // is counted once, by rank 0
    launch_pose_prior_eval(ppv, poses, poses == h->d_poses.ptr ? ppv.val : ppv.trial, h->d_cost_partials.ptr + np + 1,
                           store, h->d_invalid.ptr, h->stream);
    h->launches += 1;
  }
  if (ppv.n == 0 || h->rank != 0)
    RSBA_CUDA_TRY(cudaMemsetAsync(h->d_cost_partials.ptr + np + 1, 0, sizeof(double), h->stream));
  launch_reduce_partials(h->d_cost_partials.ptr, np + 2, h->d_scalars.ptr, h->stream);
  return RSBA_OK;
}

int run_evaluate(rsba_problem* h, bool jac, const double* poses, const double* points,
                 double* cost_out_host, long* invalid_out_host, bool compact, bool store_residuals) {
  int rc = ensure_eval_buffers(h, jac, compact);
  if (rc) return rc;
  RSBA_CUDA_TRY(cudaMemsetAsync(h->d_invalid.ptr, 0, sizeof(int), h->stream));
  const Stage st = jac ? kStageJacobian : kStageResidual;
  stage_begin(h, st);
  if (jac && compact) {
    launch_k1_compact(h->cm, h->obs_view(), poses, points, h->d_res.ptr, h->d_jacc.ptr, h->d_tau.ptr, h->d_jac_cam.ptr,
                      h->d_cost_partials.ptr, h->d_invalid.ptr, h->stream);
  } else if (jac) {
    launch_k1(h->cm, h->obs_view(), poses, points, h->d_res.ptr, h->d_jac.ptr, h->d_jac_cam.ptr, h->d_valid.ptr,
              h->d_cost_partials.ptr, h->d_invalid.ptr, h->stream);
  } else {
    launch_k1r(h->cm, h->obs_view(), poses, points, h->d_cost_partials.ptr, h->d_invalid.ptr, h->stream,
               store_residuals ? h->d_res.ptr : nullptr, store_residuals ? h->d_valid.ptr : nullptr);
  }
  rc = eval_tail(h, jac, poses, k1_num_partials(h->n_obs));
  if (rc) return rc;
  stage_end(h, st);
  h->launches += 2;
  RSBA_CUDA_TRY(cudaGetLastError());
  if (cost_out_host || invalid_out_host) {
    double c = 0.0;
    int bad = 0;
    RSBA_CUDA_TRY(cudaMemcpyAsync(&c, h->d_scalars.ptr, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    RSBA_CUDA_TRY(cudaMemcpyAsync(&bad, h->d_invalid.ptr, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    RSBA_CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (h->world > 1) {   // cost and invalid count are sums over the ranks' shares
      double pair[2] = {c, (double)bad};
      RSBA_CUDA_TRY(cudaMemcpyAsync(h->d_scalars.ptr + 8, pair, sizeof(pair), cudaMemcpyHostToDevice, h->stream));
      int rc2 = allreduce_sum(h, h->d_scalars.ptr + 8, 2);
      if (rc2) return rc2;
      RSBA_CUDA_TRY(cudaMemcpyAsync(pair, h->d_scalars.ptr + 8, sizeof(pair), cudaMemcpyDeviceToHost, h->stream));
      RSBA_CUDA_TRY(cudaStreamSynchronize(h->stream));
      c = pair[0];
      bad = (int)(pair[1] + 0.5);
    }
    if (cost_out_host) *cost_out_host = c;
    if (invalid_out_host) *invalid_out_host = bad;
  }
  return RSBA_OK;
}

void compute_point_owners(int n_frames, int n_points, long n_obs, const int* obs_frame_sorted,
                          const int* obs_point, int world, std::vector<int>* owner) {
  owner->assign(n_points, 0);
  const int T = std::max(1, (n_frames + 7) / 8);
  std::vector<long> ptr(n_points + 1, 0);
  for (long i = 0; i < n_obs; ++i) ptr[obs_point[i] + 1]++;
  for (int p = 0; p < n_points; ++p) ptr[p + 1] += ptr[p];
  std::vector<int> frames(n_obs);
  {
    std::vector<long> cur(ptr.begin(), ptr.end() - 1);
    for (long i = 0; i < n_obs; ++i) frames[cur[obs_point[i]]++] = obs_frame_sorted[i];  // ascending per point
  }
  for (int p = 0; p < n_points; ++p) {
    const long k = ptr[p + 1] - ptr[p];
    if (k == 0) { (*owner)[p] = p % world; continue; }
    const int home_tile = frames[ptr[p] + k / 2] / 8;
    (*owner)[p] = std::min(world - 1, (int)((long)home_tile * world / T));
  }
}

int materialize_local_share(rsba_problem* h) {
  const long N = h->n_obs_global;
  const HostThreads pool(N, h->world);
  h->point_owned.assign(h->n_points, 1);
  const double2* src_xy = h->g_obs_xy.data();   // what goes to the device: the whole scene on one GPU ...
  HostVec<double2> sxy;                         // ... this rank's share otherwise
  if (h->world > 1) {
    std::vector<int> owner;
    compute_point_owners(h->n_frames, h->n_points, N, h->g_obs_frame.data(), h->g_obs_point.data(), h->world, &owner);
    for (int p = 0; p < h->n_points; ++p) h->point_owned[p] = owner[p] == h->rank;
    h->local_ids.clear();
    for (long i = 0; i < N; ++i)
      if (h->point_owned[h->g_obs_point[i]]) h->local_ids.push_back(i);
    const long n = (long)h->local_ids.size();
    sxy.resize(n);
    h->h_obs_frame.resize(n);
    h->h_obs_point.resize(n);
    pool.split(n, [&](int, long b, long e) {
      for (long i = b; i < e; ++i) {
        const long g = h->local_ids[i];
        sxy[i] = h->g_obs_xy[g];
        h->h_obs_frame[i] = h->g_obs_frame[g];
        h->h_obs_point[i] = h->g_obs_point[g];
      }
    });
    src_xy = sxy.data();
  } else {
    h->local_ids.resize(N);
    h->h_obs_frame.resize(N);
    h->h_obs_point.resize(N);
    pool.split(N, [&](int, long b, long e) {
      for (long i = b; i < e; ++i) h->local_ids[i] = i;
      std::copy(h->g_obs_frame.begin() + b, h->g_obs_frame.begin() + e, h->h_obs_frame.begin() + b);
      std::copy(h->g_obs_point.begin() + b, h->g_obs_point.begin() + e, h->h_obs_point.begin() + b);
    });
  }
  const long n = (long)h->local_ids.size();
  h->n_obs = n;
  RSBA_CUDA_TRY(h->d_obs_xy.resize(n));
  RSBA_CUDA_TRY(h->d_obs_frame.resize(n));
  RSBA_CUDA_TRY(h->d_obs_point.resize(n));
  if (n > 0) {
    RSBA_CUDA_TRY(cudaMemcpyAsync(h->d_obs_xy.ptr, src_xy, n * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
    RSBA_CUDA_TRY(cudaMemcpyAsync(h->d_obs_frame.ptr, h->h_obs_frame.data(), n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    RSBA_CUDA_TRY(cudaMemcpyAsync(h->d_obs_point.ptr, h->h_obs_point.data(), n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  }
  RSBA_CUDA_TRY(cudaStreamSynchronize(h->stream));  // sxy is a local
  if (h->lm) {
    lm_state_free(h->lm);
    h->lm = nullptr;
  }
  return RSBA_OK;
}

// order[i] = the caller's index of the i-th observation after a STABLE sort by frame (keeps the caller's
// within-frame order, which is the reference's insertion order, CeresHandler.h:208).  Host only.
static int sort_by_frame(const HostThreads& pool, long n, const int* fr, const int* pt, int n_frames, int n_points,
                         HostVec<long>* order) {
  // range check and "already sorted by frame" in one parallel pass (bit 0: index out of range, bit 1: unsorted)
  std::vector<int> flags(pool.n, 0);
  pool.split(n, [&](int t, long b, long e) {
    int f = 0;
    for (long i = b; i < e; ++i) {
      if (fr[i] < 0 || fr[i] >= n_frames || pt[i] < 0 || pt[i] >= n_points) f |= 1;
      if (i > 0 && fr[i - 1] > fr[i]) f |= 2;
    }
    flags[t] = f;
  });
  int flag = 0;
  for (int f : flags) flag |= f;
  if (flag & 1) return fail(RSBA_ERR_INVALID_ARGUMENT, "observation index out of range");
  order->resize(n);
  long* o = order->data();
  pool.split(n, [&](int, long b, long e) { for (long i = b; i < e; ++i) o[i] = i; });
  if (flag & 2) std::stable_sort(order->begin(), order->end(), [&](long a, long b) { return fr[a] < fr[b]; });
  return RSBA_OK;
}

// Sort observations by frame and upload this rank's share of the SoA.
static int upload_scene(rsba_problem* h, long n, const double* xy, const int* fr, const int* pt,
                        int n_frames, int n_points) {
  const HostThreads pool(n, h->world);
  TraceLap lap;
  int rc0 = sort_by_frame(pool, n, fr, pt, n_frames, n_points, &h->order);
  if (rc0) return rc0;
  lap("check + sort by frame");
  h->g_obs_xy.resize(n);
  h->g_obs_frame.resize(n);
  h->g_obs_point.resize(n);
  pool.split(n, [&](int, long b, long e) {
    for (long i = b; i < e; ++i) {
      const long s = h->order[i];
      h->g_obs_xy[i] = make_double2(xy[2 * s], xy[2 * s + 1]);
      h->g_obs_frame[i] = fr[s];
      h->g_obs_point[i] = pt[s];
    }
  });
  lap("SoA copy");
  h->n_obs_global = n;
  h->n_frames = n_frames;
  h->n_points = n_points;
  RSBA_CUDA_TRY(h->d_poses.resize((size_t)kFrameParams * (n_frames + 1)));   // + the intrinsics pseudo-frame
  RSBA_CUDA_TRY(cudaMemsetAsync(h->d_poses.ptr, 0, h->d_poses.bytes(), h->stream));
  h->cm.cam_offset = h->free_cam ? (long)kFrameParams * n_frames : -1;
  RSBA_CUDA_TRY(h->d_points.resize((size_t)kPointParams * n_points));
  int rc = materialize_local_share(h);
  if (rc) return rc;
  lap("local share + H2D");
  h->scene_set = true;
  h->params_set = false;
  return RSBA_OK;
}

// Closed-form coefficients of the two functors with a constant interFrameRatio
// (video_bundler_rs_inter.h:69-84 velocity, :127-148 acceleration); block order pose0, end0, pose1, end1.
static void prior_coefficients(int kind, double r, double c[8]) {
  if (kind == 1) {
    c[0] = 1.0; c[1] = 0.0; c[2] = r; c[3] = -(1.0 + r);
    if (r > 2.220446049250313e-16) { c[4] = -(1.0 + 1.0 / r); c[5] = 1.0; c[6] = 0.0; c[7] = 1.0 / r; }
    else                           { c[4] = -1.0; c[5] = 1.0; c[6] = 1.0; c[7] = -1.0; }
  } else {
    c[0] = 0.5; c[1] = 0.0; c[2] = 0.5 * r; c[3] = -0.5 * (1.0 + r);
    c[4] = -0.5 * (1.0 + 1.0 / r); c[5] = 0.5; c[6] = 0.0; c[7] = 0.5 / r;
  }
}

int upload_priors(rsba_problem* h) {
  if (!h->priors_dirty) return RSBA_OK;
  const int n = (int)h->priors.size(), F = h->n_frames;
  std::vector<int> frame(std::max(n, 1)), prev(std::max(n, 1)), cur_of(std::max(F, 1), -1), prev_of(std::max(F, 1), -1);
  std::vector<double> coef(8 * (size_t)std::max(n, 1)), scale(std::max(n, 1));
  std::vector<int> kind(std::max(n, 1), 1);
  for (int i = 0; i < n; ++i) {
    const auto& p = h->priors[i];
    if (p.frame < 0 || p.frame >= F || p.prev < 0 || p.prev >= F || p.frame == p.prev)
      return fail(RSBA_ERR_INVALID_ARGUMENT, "motion prior: frame index out of range");
    if (cur_of[p.frame] >= 0) return fail(RSBA_ERR_INVALID_ARGUMENT, "motion prior: a frame has two priors");
    if (prev_of[p.prev] >= 0) return fail(RSBA_ERR_INVALID_ARGUMENT, "motion prior: a frame precedes two priors");
    cur_of[p.frame] = i;
    prev_of[p.prev] = i;
    frame[i] = p.frame;
    prev[i] = p.prev;
    scale[i] = p.scale;
    kind[i] = p.kind;
    // (with a free ratio the device refreshes the coefficients from the ratio parameter at every evaluation)
    prior_coefficients(p.kind, h->free_ratio ? h->ratio_value : p.ratio, &coef[8 * (size_t)i]);
  }
  auto up = [&](auto& dev, const auto& host) -> cudaError_t {
    cudaError_t e = dev.resize(host.size());
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(dev.ptr, host.data(), host.size() * sizeof(host[0]), cudaMemcpyHostToDevice, h->stream);
  };
  RSBA_CUDA_TRY(up(h->d_prior_frame, frame));
  RSBA_CUDA_TRY(up(h->d_prior_prev, prev));
  RSBA_CUDA_TRY(up(h->d_prior_cur_of, cur_of));
  RSBA_CUDA_TRY(up(h->d_prior_prev_of, prev_of));
  RSBA_CUDA_TRY(up(h->d_prior_coef, coef));
  RSBA_CUDA_TRY(up(h->d_prior_scale, scale));
  RSBA_CUDA_TRY(up(h->d_prior_kind, kind));
  RSBA_CUDA_TRY(h->d_prior_jr.resize(12 * (size_t)std::max(n, 1)));
  RSBA_CUDA_TRY(h->d_prior_r.resize(12 * (size_t)std::max(n, 1)));
  RSBA_CUDA_TRY(h->d_prior_w2.resize(std::max(n, 1)));
  RSBA_CUDA_TRY(h->d_prior_Bx.resize(24 * (size_t)std::max(F, 1)));
  RSBA_CUDA_TRY(cudaMemsetAsync(h->d_prior_Bx.ptr, 0, h->d_prior_Bx.bytes(), h->stream));
  RSBA_CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->priors_dirty = false;
  if (h->lm) {   // the priors add structurally non-zero tile pairs
    lm_state_free(h->lm);
    h->lm = nullptr;
  }
  return RSBA_OK;
}

int upload_pose_priors(rsba_problem* h) {
  if (!h->pose_priors_dirty) return RSBA_OK;
  const int n = (int)h->pose_priors.size(), F = h->n_frames;
  const size_t nz = (size_t)std::max(n, 1);
  std::vector<int> slot(nz, 0);
  std::vector<unsigned char> cst(nz, 0), used((size_t)2 * std::max(F, 1), 0);
  std::vector<double> w(2 * nz, 0.0), val(6 * nz, 0.0);
  for (int i = 0; i < n; ++i) {
    auto& p = h->pose_priors[i];
    if (h->ptr_mode) {   // the control pose is known by its block address only
      auto a = h->pose0_to_frame.find(p.pose);
      if (a != h->pose0_to_frame.end()) p.slot = 2 * a->second;
      else {
        auto b = h->pose1_to_frame.find(p.pose);
        if (b == h->pose1_to_frame.end()) return fail(RSBA_ERR_INVALID_ARGUMENT, "pose prior on a pose block that no residual block uses");
        p.slot = 2 * b->second + 1;
      }
      memcpy(p.val, p.prior, sizeof(p.val));
    }
    if (p.slot < 0 || p.slot >= 2 * F) return fail(RSBA_ERR_INVALID_ARGUMENT, "pose prior: control pose out of range");
    if (used[p.slot]) return fail(RSBA_ERR_INVALID_ARGUMENT, "pose prior: a control pose has two priors");
    used[p.slot] = 1;
    slot[i] = p.slot; cst[i] = p.constant; w[2 * i] = p.rot; w[2 * i + 1] = p.pos;
    memcpy(&val[6 * (size_t)i], p.val, sizeof(p.val));
  }
  auto up = [&](auto& dev, const auto& host) -> cudaError_t {
    cudaError_t e = dev.resize(host.size());
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(dev.ptr, host.data(), host.size() * sizeof(host[0]), cudaMemcpyHostToDevice, h->stream);
  };
  RSBA_CUDA_TRY(up(h->d_pp_slot, slot));
  RSBA_CUDA_TRY(up(h->d_pp_const, cst));
  RSBA_CUDA_TRY(up(h->d_pp_w, w));
  RSBA_CUDA_TRY(up(h->d_pp_val, val));
  RSBA_CUDA_TRY(up(h->d_pp_trial, val));
  RSBA_CUDA_TRY(h->d_pp_r.resize(6 * nz));
  RSBA_CUDA_TRY(h->d_pp_cinv.resize(6 * nz));
  RSBA_CUDA_TRY(h->d_pp_d2.resize(6 * nz));
  RSBA_CUDA_TRY(cudaMemsetAsync(h->d_pp_r.ptr, 0, h->d_pp_r.bytes(), h->stream));
  RSBA_CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->pose_priors_dirty = false;
  return RSBA_OK;
}

static int check_prior_args(int kind, double scale, double ratio) {
  if (kind != 1 && kind != 2) return fail(RSBA_ERR_INVALID_ARGUMENT, "motion prior kind must be 1 (velocity) or 2 (acceleration)");
  if (!(scale > 0.0)) return fail(RSBA_ERR_INVALID_ARGUMENT, "motion prior scale must be positive");
  // the functors return false below these bounds (video_bundler_rs_inter.h:92, 156)
  if (kind == 1 && !(ratio >= 0.0)) return fail(RSBA_ERR_INVALID_ARGUMENT, "velocity prior needs interFrameRatio >= 0");
  if (kind == 2 && !(ratio >= 2.220446049250313e-16)) return fail(RSBA_ERR_INVALID_ARGUMENT, "acceleration prior needs interFrameRatio >= eps");
  return RSBA_OK;
}

int finalize_pointer_problem(rsba_problem* h) {
  if (!h->ptr_mode || !h->ptr_dirty) return RSBA_OK;
  const long n = (long)h->ptr_obs.size();
  std::vector<double> xy(2 * n);
  std::vector<int> fr(n), pt(n);
  for (long i = 0; i < n; ++i) {
    xy[2 * i] = h->ptr_obs[i].x;
    xy[2 * i + 1] = h->ptr_obs[i].y;
    fr[i] = h->ptr_obs[i].frame;
    pt[i] = h->ptr_obs[i].point;
  }
  int rc = upload_scene(h, n, xy.data(), fr.data(), pt.data(), (int)h->frame_pose0.size(),
                        (int)h->point_ptr.size());
  if (rc) return rc;
  h->pose_mask = h->ptr_pose_mask;
  h->point_const = h->ptr_point_const;
  h->pose_mask.resize(h->n_frames, 0);
  h->point_const.resize(h->n_points, 0);
  h->ptr_dirty = false;
  h->priors_dirty = true;
  h->pose_priors_dirty = true;
  return upload_priors(h);
}

int gather_pointer_parameters(rsba_problem* h) {
  std::vector<double> poses((size_t)kFrameParams * h->n_frames), points((size_t)kPointParams * h->n_points);
  for (int f = 0; f < h->n_frames; ++f) {
    memcpy(&poses[(size_t)12 * f], h->frame_pose0[f], 6 * sizeof(double));
    memcpy(&poses[(size_t)12 * f + 6], h->frame_pose1[f], 6 * sizeof(double));
  }
  for (int p = 0; p < h->n_points; ++p) memcpy(&points[(size_t)3 * p], h->point_ptr[p], 3 * sizeof(double));
  if (h->ptr_cam) memcpy(h->cm.cam, h->ptr_cam, 9 * sizeof(double));   // the intrinsics block's current values
  if (h->ptr_ratio) h->ratio_value = *h->ptr_ratio;
  if (!h->pose_priors.empty()) h->pose_priors_dirty = true;   // re-read the caller's prior blocks
  return rsba_cuda_set_parameters(h, poses.data(), points.data());
}

int scatter_pointer_parameters(rsba_problem* h) {
  std::vector<double> poses((size_t)kFrameParams * h->n_frames), points((size_t)kPointParams * h->n_points);
  int rc = rsba_cuda_get_parameters(h, poses.data(), points.data());
  if (rc) return rc;
  for (int f = 0; f < h->n_frames; ++f) {
    memcpy(h->frame_pose0[f], &poses[(size_t)12 * f], 6 * sizeof(double));
    memcpy(h->frame_pose1[f], &poses[(size_t)12 * f + 6], 6 * sizeof(double));
  }
  for (int p = 0; p < h->n_points; ++p) memcpy(h->point_ptr[p], &points[(size_t)3 * p], 3 * sizeof(double));
  if (h->ptr_cam) {
    double cam[9];
    rc = rsba_cuda_get_camera(h, cam);
    if (rc) return rc;
    memcpy(h->ptr_cam, cam, sizeof(cam));
  }
  if (h->ptr_ratio) {
    rc = rsba_cuda_get_inter_frame_ratio(h, h->ptr_ratio);
    if (rc) return rc;
  }
  if (!h->pose_priors.empty()) {
    std::vector<double> vals(6 * h->pose_priors.size());
    if (rsba_cuda_get_pose_priors(h, vals.data(), nullptr) < 0) return RSBA_ERR_CUDA;
    for (size_t i = 0; i < h->pose_priors.size(); ++i)
      if (h->pose_priors[i].prior) memcpy(h->pose_priors[i].prior, &vals[6 * i], 6 * sizeof(double));
  }
  return RSBA_OK;
}

}  // namespace rsba

using namespace rsba;

extern "C" {

const char* rsba_cuda_last_error(void) { return g_last_error.c_str(); }
const char* rsba_cuda_version(void) { return "rsba_b200 0.1 (sm_100a)"; }

int rsba_cuda_create(rsba_problem** out, int device) {
  return rsba::api_guard([&]() -> int {
  if (!out) return fail(RSBA_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(RSBA_ERR_NO_DEVICE, "no CUDA device: rsba_cuda has no CPU fallback");
  }
  if (device < 0 || device >= count) return fail(RSBA_ERR_INVALID_ARGUMENT, "device index out of range");
  RSBA_CUDA_TRY(cudaSetDevice(device));
  rsba_problem* h = new rsba_problem;
  h->device = device;
  e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete h;
    return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__);
  }
  h->own_stream = true;
  *out = h;
  return RSBA_OK;
  });
}

void rsba_cuda_destroy(rsba_problem* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->lm) lm_state_free(h->lm);
  for (auto& t : h->timers) {
    if (t.beg) cudaEventDestroy(t.beg);
    if (t.end) cudaEventDestroy(t.end);
  }
  if (h->nccl_comm) {
    const NcclApi* api = nccl_api();
    if (api) api->CommDestroy((ncclComm_t)h->nccl_comm);
  }
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int rsba_cuda_nccl_unique_id(unsigned char id[128]) {
  return rsba::api_guard([&]() -> int {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (!id) return fail(RSBA_ERR_INVALID_ARGUMENT, "id is NULL");
  const NcclApi* api = nccl_api();
  if (!api) return RSBA_ERR_NCCL;
  ncclUniqueId uid;
  ncclResult_t r = api->GetUniqueId(&uid);
  if (r != ncclSuccess) return fail(RSBA_ERR_NCCL, std::string("ncclGetUniqueId: ") + api->GetErrorString(r));
  memcpy(id, &uid, 128);
  return RSBA_OK;
  });
}

int rsba_cuda_point_owners(int n_frames, int n_points, long n_obs, const int* obs_frame, const int* obs_point,
                           int world_size, int* owner) {
  return rsba::api_guard([&]() -> int {
  if (n_frames < 0 || n_points < 0 || n_obs < 0 || world_size < 1 || !owner || (n_obs > 0 && (!obs_frame || !obs_point)))
    return fail(RSBA_ERR_INVALID_ARGUMENT, "bad arguments");
  for (long i = 0; i < n_obs; ++i) {
    if (obs_frame[i] < 0 || obs_frame[i] >= n_frames || obs_point[i] < 0 || obs_point[i] >= n_points)
      return fail(RSBA_ERR_INVALID_ARGUMENT, "observation index out of range");
    if (i > 0 && obs_frame[i - 1] > obs_frame[i]) return fail(RSBA_ERR_INVALID_ARGUMENT, "obs_frame must be sorted");
  }
  std::vector<int> o;
  compute_point_owners(n_frames, n_points, n_obs, obs_frame, obs_point, world_size, &o);
  std::copy(o.begin(), o.end(), owner);
  return RSBA_OK;
  });
}

long rsba_cuda_sort_observations(long n_obs, const int* obs_frame, const int* obs_point, int n_frames, int n_points,
                                 long* order) {
  if (n_obs < 0 || n_frames < 0 || n_points < 0 || (n_obs > 0 && (!obs_frame || !obs_point))) {
    fail(RSBA_ERR_INVALID_ARGUMENT, "bad arguments");
    return -1;
  }
  HostVec<long> o;
  if (sort_by_frame(HostThreads(n_obs), n_obs, obs_frame, obs_point, n_frames, n_points, &o)) return -1;
  if (order) std::copy(o.begin(), o.end(), order);
  return n_obs;
}

int rsba_cuda_comm_init(rsba_problem* h, int rank, int world_size, const unsigned char id[128]) {
  return rsba::api_guard([&]() -> int {
  if (!h || !id || world_size < 1 || rank < 0 || rank >= world_size)
    return fail(RSBA_ERR_INVALID_ARGUMENT, "bad communicator arguments");
  if (h->nccl_comm) return fail(RSBA_ERR_STATE, "communicator already initialised");
  RSBA_CUDA_TRY(cudaSetDevice(h->device));
  if (world_size > 1) {
    const NcclApi* api = nccl_api();
    if (!api) return RSBA_ERR_NCCL;
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    ncclComm_t comm;
    ncclResult_t r = api->CommInitRank(&comm, world_size, uid, rank);
    if (r != ncclSuccess) return fail(RSBA_ERR_NCCL, std::string("ncclCommInitRank: ") + api->GetErrorString(r));
    h->nccl_comm = comm;
    h->rank = rank;
    h->world = world_size;
    // NCCL connects its channels lazily, at the first collective of each protocol: pay that here (one small
    // and one multi-megabyte all-reduce) and not inside the first solve, where it was ~1 s at 8 ranks
    DeviceBuffer<double> warm;
    RSBA_CUDA_TRY(warm.resize((size_t)1 << 20));
    RSBA_CUDA_TRY(cudaMemsetAsync(warm.ptr, 0, warm.bytes(), h->stream));
    int rc = allreduce_sum(h, warm.ptr, 8);
    if (!rc) rc = allreduce_sum(h, warm.ptr, warm.count);
    if (rc) return rc;
    RSBA_CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  h->rank = rank;
  h->world = world_size;
  if (h->scene_set) return materialize_local_share(h);   // re-shard a scene that was set first
  return RSBA_OK;
  });
}

int rsba_cuda_set_stream(rsba_problem* h, void* cuda_stream) {
  return rsba::api_guard([&]() -> int {
  if (!h) return fail(RSBA_ERR_INVALID_ARGUMENT, "handle is NULL");
  cudaStreamSynchronize(h->stream);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  h->stream = (cudaStream_t)cuda_stream;
  h->own_stream = false;
  return RSBA_OK;
  });
}

int rsba_cuda_set_camera(rsba_problem* h, const double cam9[9], int shutter, const int scanlines[2],
                         int interpolate_rotation) {
  return rsba::api_guard([&]() -> int {
  if (!h || !cam9 || !scanlines) return fail(RSBA_ERR_INVALID_ARGUMENT, "NULL argument");
  if (shutter < 0 || shutter > 2) return fail(RSBA_ERR_INVALID_ARGUMENT, "shutter must be 0, 1 or 2");
  memcpy(h->cm.cam, cam9, sizeof(h->cm.cam));
  h->cm.shutter = shutter;
  h->cm.scan0 = (double)scanlines[0];
  h->cm.scan_span = (double)(scanlines[1] - scanlines[0]);
  h->cm.interp_rot = interpolate_rotation ? 1 : 0;   // (cm.huber is set by rsba_cuda_set_loss and kept)
  h->cm.cam_offset = (h->free_cam && h->scene_set) ? (long)kFrameParams * h->n_frames : -1;
  h->camera_set = true;
  if (h->free_cam && h->scene_set) {   // the intrinsics are parameters: refresh their device copy
    RSBA_CUDA_TRY(cudaSetDevice(h->device));
    RSBA_CUDA_TRY(cudaMemcpyAsync(h->d_poses.ptr + h->cm.cam_offset, h->cm.cam, 9 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    RSBA_CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  return RSBA_OK;
  });
}

int rsba_cuda_set_intrinsics_free(rsba_problem* h, int free_intrinsics) {
  return rsba::api_guard([&]() -> int {
  if (!h) return fail(RSBA_ERR_INVALID_ARGUMENT, "handle is NULL");
  const bool want = free_intrinsics != 0;
  if (want == h->free_cam) return RSBA_OK;
  h->free_cam = want;
  h->cm.cam_offset = (want && h->scene_set) ? (long)kFrameParams * h->n_frames : -1;
  if (h->lm) {   // one more block in the reduced system
    lm_state_free(h->lm);
    h->lm = nullptr;
  }
  if (want && h->scene_set && h->camera_set) {
    RSBA_CUDA_TRY(cudaSetDevice(h->device));
    RSBA_CUDA_TRY(cudaMemcpyAsync(h->d_poses.ptr + h->cm.cam_offset, h->cm.cam, 9 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    RSBA_CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  return RSBA_OK;
  });
}

int rsba_cuda_get_camera(rsba_problem* h, double cam9[9]) {
  return rsba::api_guard([&]() -> int {
  if (!h || !cam9) return fail(RSBA_ERR_INVALID_ARGUMENT, "NULL argument");
  if (!h->camera_set) return fail(RSBA_ERR_STATE, "rsba_cuda_set_camera has not been called");
  if (h->free_cam && h->scene_set && h->params_set) {
    RSBA_CUDA_TRY(cudaSetDevice(h->device));
    RSBA_CUDA_TRY(cudaMemcpyAsync(h->cm.cam, h->d_poses.ptr + (size_t)kFrameParams * h->n_frames, 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    RSBA_CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  memcpy(cam9, h->cm.cam, 9 * sizeof(double));
  return RSBA_OK;
  });
}

int rsba_cuda_get_intrinsics_jacobian(rsba_problem* h, double* jacobian_cam) {
  return rsba::api_guard([&]() -> int {
  if (!h || !jacobian_cam) return fail(RSBA_ERR_INVALID_ARGUMENT, "NULL argument");
  if (!h->free_cam || h->d_jac_cam.count == 0) return fail(RSBA_ERR_STATE, "no intrinsics Jacobian: set_intrinsics_free + evaluate first");
  RSBA_CUDA_TRY(cudaSetDevice(h->device));
  const long n = h->n_obs, ng = h->n_obs_global;
  std::vector<double> tmp((size_t)18 * n);
  RSBA_CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (n) RSBA_CUDA_TRY(cudaMemcpy(tmp.data(), h->d_jac_cam.ptr, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
  memset(jacobian_cam, 0, (size_t)18 * ng * sizeof(double));
  for (long i = 0; i < n; ++i)
    memcpy(jacobian_cam + 18 * h->order[h->local_ids[i]], &tmp[18 * (size_t)i], 18 * sizeof(double));
  return RSBA_OK;
  });
}

int rsba_cuda_set_loss(rsba_problem* h, double huber_a) {
  return rsba::api_guard([&]() -> int {
  if (!h || !(huber_a >= 0.0)) return fail(RSBA_ERR_INVALID_ARGUMENT, "huber_a must be >= 0");
  h->cm.huber = huber_a;
  return RSBA_OK;
  });
}

// pointer API: the frame made of the two control-pose blocks (registered on first sight); < 0 = error
static int frame_of_blocks(rsba_problem* h, double* pose0, double* pose1) {
  auto it = h->pose0_to_frame.find(pose0);
  if (it == h->pose0_to_frame.end()) {
    if (h->pose1_to_frame.count(pose1) || h->pose0_to_frame.count(pose1) || h->pose1_to_frame.count(pose0))
      return fail(RSBA_ERR_INVALID_ARGUMENT, "pose block already paired with a different frame");
    const int f = (int)h->frame_pose0.size();
    h->pose0_to_frame[pose0] = f;
    h->pose1_to_frame[pose1] = f;
    h->frame_pose0.push_back(pose0);
    h->frame_pose1.push_back(pose1);
    h->ptr_pose_mask.push_back(0);
    return f;
  }
  if (h->frame_pose1[it->second] != pose1)
    return fail(RSBA_ERR_INVALID_ARGUMENT, "pose0 block already paired with a different pose1 block");
  return it->second;
}

int rsba_cuda_add_frame_blocks(rsba_problem* h, double* pose0, double* pose1) {
  return rsba::api_guard([&]() -> int {
  if (!h || !pose0 || !pose1) return fail(RSBA_ERR_INVALID_ARGUMENT, "NULL argument");
  if (h->scene_set && !h->ptr_mode) return fail(RSBA_ERR_STATE, "handle already holds a bulk scene");
  h->ptr_mode = true;
  const int f = frame_of_blocks(h, pose0, pose1);
  if (f < 0) return f;
  h->ptr_dirty = true;
  return RSBA_OK;
  });
}

int rsba_cuda_add_motion_prior(rsba_problem* h, int kind, double scale, double inter_frame_ratio, double* pose0,
                               double* end0, double* pose1, double* end1) {
  return rsba::api_guard([&]() -> int {
  if (!h || !pose0 || !end0 || !pose1 || !end1) return fail(RSBA_ERR_INVALID_ARGUMENT, "NULL argument");
  if (h->scene_set && !h->ptr_mode) return fail(RSBA_ERR_STATE, "handle already holds a bulk scene");
  int rc = check_prior_args(kind, scale, inter_frame_ratio);
  if (rc) return rc;
  h->ptr_mode = true;
  const int prev = frame_of_blocks(h, pose1, end1);
  if (prev < 0) return prev;
  const int cur = frame_of_blocks(h, pose0, end0);
  if (cur < 0) return cur;
  h->priors.push_back({kind, scale, inter_frame_ratio, cur, prev});
  h->priors_dirty = true;
  h->ptr_dirty = true;
  return RSBA_OK;
  });
}

int rsba_cuda_set_motion_priors(rsba_problem* h, int n, const int* kind, const double* scale,
                                const double* inter_frame_ratio, const int* frame, const int* prev_frame) {
  return rsba::api_guard([&]() -> int {
  if (!h || n < 0 || (n > 0 && (!kind || !scale || !inter_frame_ratio || !frame || !prev_frame)))
    return fail(RSBA_ERR_INVALID_ARGUMENT, "bad prior arguments");
  if (h->ptr_mode) return fail(RSBA_ERR_STATE, "handle holds pointer-API blocks: use rsba_cuda_add_motion_prior");
  if (!h->scene_set) return fail(RSBA_ERR_STATE, "rsba_cuda_set_scene first");
  for (int i = 0; i < n; ++i) {
    int rc = check_prior_args(kind[i], scale[i], inter_frame_ratio[i]);
    if (rc) return rc;
  }
  h->priors.clear();
  for (int i = 0; i < n; ++i) h->priors.push_back({kind[i], scale[i], inter_frame_ratio[i], frame[i], prev_frame[i]});
  h->priors_dirty = true;
  RSBA_CUDA_TRY(cudaSetDevice(h->device));
  return upload_priors(h);
  });
}

long rsba_cuda_get_prior_residuals(rsba_problem* h, double* residuals) {
  if (!h) return -1;
  const long n = h->priors_dirty ? 0 : (long)h->priors.size();
  if (residuals && n > 0) {
    cudaSetDevice(h->device);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return -1;   // written on the handle's (non-blocking) stream
    if (cudaMemcpy(residuals, h->d_prior_r.ptr, 12 * n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  }
  return n;
}

int rsba_cuda_add_pose_prior(rsba_problem* h, double rotation, double position, double* prior_block,
                             double* pose_block) {
  return rsba::api_guard([&]() -> int {
  if (!h || !prior_block || !pose_block) return fail(RSBA_ERR_INVALID_ARGUMENT, "NULL argument");
  if (h->scene_set && !h->ptr_mode) return fail(RSBA_ERR_STATE, "handle already holds a bulk scene");
  if (h->pose_prior_of_block.count(prior_block)) return fail(RSBA_ERR_INVALID_ARGUMENT, "prior block already used");
  h->ptr_mode = true;
  rsba_problem::PosePriorHost p{};
  p.slot = -1; p.rot = rotation; p.pos = position; p.constant = 0; p.prior = prior_block; p.pose = pose_block;
  h->pose_prior_of_block[prior_block] = (int)h->pose_priors.size();
  h->pose_priors.push_back(p);
  h->pose_priors_dirty = true;
  h->ptr_dirty = true;
  return RSBA_OK;
  });
}

int rsba_cuda_set_pose_priors(rsba_problem* h, int n, const int* frame, const int* which_pose, const double* rotation,
                              const double* position, const double* prior_values, const unsigned char* prior_constant) {
  return rsba::api_guard([&]() -> int {
  if (!h || n < 0 || (n > 0 && (!frame || !which_pose || !rotation || !position || !prior_values)))
    return fail(RSBA_ERR_INVALID_ARGUMENT, "bad pose prior arguments");
  if (h->ptr_mode) return fail(RSBA_ERR_STATE, "handle holds pointer-API blocks: use rsba_cuda_add_pose_prior");
  if (!h->scene_set) return fail(RSBA_ERR_STATE, "rsba_cuda_set_scene first");
  std::vector<rsba_problem::PosePriorHost> list;
  for (int i = 0; i < n; ++i) {
    if (which_pose[i] != 0 && which_pose[i] != 1) return fail(RSBA_ERR_INVALID_ARGUMENT, "which_pose must be 0 or 1");
    if (frame[i] < 0 || frame[i] >= h->n_frames) return fail(RSBA_ERR_INVALID_ARGUMENT, "pose prior: frame out of range");
    rsba_problem::PosePriorHost p{};
    p.slot = 2 * frame[i] + which_pose[i]; p.rot = rotation[i]; p.pos = position[i];
    p.constant = prior_constant && prior_constant[i] ? 1 : 0;
    memcpy(p.val, prior_values + 6 * (size_t)i, sizeof(p.val));
    list.push_back(p);
  }
  h->pose_priors.swap(list);
  h->pose_priors_dirty = true;
  RSBA_CUDA_TRY(cudaSetDevice(h->device));
  return upload_pose_priors(h);
  });
}

long rsba_cuda_get_pose_priors(rsba_problem* h, double* prior_values, double* trial_values) {
  if (!h) return -1;
  const long n = h->pose_priors_dirty ? 0 : (long)h->pose_priors.size();
  if (n > 0 && (prior_values || trial_values)) {
    cudaSetDevice(h->device);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return -1;
    if (prior_values && cudaMemcpy(prior_values, h->d_pp_val.ptr, 6 * n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (trial_values && cudaMemcpy(trial_values, h->d_pp_trial.ptr, 6 * n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  }
  return n;
}

static int set_ratio_free(rsba_problem* h, bool want, double value) {
  if (want) {
    // the functors return false below these bounds (video_bundler_rs_inter.h:92, 157)
    if (!(value >= h->ratio_lower_bound())) return fail(RSBA_ERR_INVALID_ARGUMENT, "interFrameRatio below its lower bound");
    h->ratio_value = value;
  }
  if (want != h->free_ratio) {
    h->free_ratio = want;
    h->priors_dirty = true;   // coefficients + structure (the ratio couples with every prior frame)
    if (h->lm) { lm_state_free(h->lm); h->lm = nullptr; }
  }
  if (h->scene_set) {
    RSBA_CUDA_TRY(cudaSetDevice(h->device));
    const double v = want ? value : 0.0;
    RSBA_CUDA_TRY(cudaMemcpyAsync(h->d_poses.ptr + (size_t)kFrameParams * h->n_frames + 9, &v, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    RSBA_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return upload_priors(h);
  }
  return RSBA_OK;
}

int rsba_cuda_set_inter_frame_ratio_free(rsba_problem* h, int free_ratio, double value) {
  return rsba::api_guard([&]() -> int {
  if (!h) return fail(RSBA_ERR_INVALID_ARGUMENT, "handle is NULL");
  if (h->ptr_ratio && !free_ratio) h->ptr_ratio = nullptr;
  return set_ratio_free(h, free_ratio != 0, value);
  });
}

int rsba_cuda_set_inter_frame_ratio_block(rsba_problem* h, double* ratio) {
  return rsba::api_guard([&]() -> int {
  if (!h || !ratio) return fail(RSBA_ERR_INVALID_ARGUMENT, "NULL argument");
  h->ptr_ratio = ratio;
  return set_ratio_free(h, true, *ratio);
  });
}

long rsba_cuda_get_prior_ratio_jacobian(rsba_problem* h, double* d_residual_d_ratio) {
  if (!h) return -1;
  if (!h->free_ratio) { fail(RSBA_ERR_STATE, "the interFrameRatio is not a free parameter"); return -1; }
  const long n = h->priors_dirty ? 0 : (long)h->priors.size();
  if (d_residual_d_ratio && n > 0) {
    cudaSetDevice(h->device);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return -1;
    if (cudaMemcpy(d_residual_d_ratio, h->d_prior_jr.ptr, 12 * n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  }
  return n;
}

int rsba_cuda_get_inter_frame_ratio(rsba_problem* h, double* value) {
  return rsba::api_guard([&]() -> int {
  if (!h || !value) return fail(RSBA_ERR_INVALID_ARGUMENT, "NULL argument");
  if (h->free_ratio && h->scene_set && h->params_set) {
    RSBA_CUDA_TRY(cudaSetDevice(h->device));
    RSBA_CUDA_TRY(cudaMemcpyAsync(&h->ratio_value, h->d_poses.ptr + (size_t)kFrameParams * h->n_frames + 9, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    RSBA_CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  *value = h->ratio_value;
  return RSBA_OK;
  });
}

int rsba_cuda_add_rs_residual_with_intrinsics(rsba_problem* h, const double observed[2], double* intrinsics,
                                              double* pose0, double* pose1, double* point) {
  return rsba::api_guard([&]() -> int {
  if (!h || !intrinsics) return fail(RSBA_ERR_INVALID_ARGUMENT, "NULL argument");
  if (h->ptr_cam && h->ptr_cam != intrinsics)
    return fail(RSBA_ERR_INVALID_ARGUMENT, "only ONE shared intrinsics block is supported (sess.cam); per-frame f.cam blocks are not");
  if (!h->ptr_cam && !h->ptr_obs.empty())
    return fail(RSBA_ERR_INVALID_ARGUMENT, "calibrated and uncalibrated residual blocks cannot be mixed");
  int rc = rsba_cuda_add_rs_residual(h, observed, pose0, pose1, point);
  if (rc) return rc;
  h->ptr_cam = intrinsics;
  if (!h->free_cam) {
    h->free_cam = true;
    if (h->lm) { lm_state_free(h->lm); h->lm = nullptr; }
  }
  return RSBA_OK;
  });
}

int rsba_cuda_add_rs_residual(rsba_problem* h, const double observed[2], double* pose0, double* pose1,
                              double* point) {
  return rsba::api_guard([&]() -> int {
  if (!h || !observed || !pose0 || !pose1 || !point) return fail(RSBA_ERR_INVALID_ARGUMENT, "NULL argument");
  if (h->scene_set && !h->ptr_mode) return fail(RSBA_ERR_STATE, "handle already holds a bulk scene");
  h->ptr_mode = true;
  const int f = frame_of_blocks(h, pose0, pose1);
  if (f < 0) return f;
  int p;
  auto ip = h->point_to_id.find(point);
  if (ip == h->point_to_id.end()) {
    p = (int)h->point_ptr.size();
    h->point_to_id[point] = p;
    h->point_ptr.push_back(point);
    h->ptr_point_const.push_back(0);
  } else {
    p = ip->second;
  }
  h->ptr_obs.push_back({observed[0], observed[1], f, p});
  h->ptr_dirty = true;
  return RSBA_OK;
  });
}

int rsba_cuda_set_block_constant(rsba_problem* h, double* block) {
  return rsba::api_guard([&]() -> int {
  if (!h || !block) return fail(RSBA_ERR_INVALID_ARGUMENT, "NULL argument");
  auto a = h->pose0_to_frame.find(block);
  if (a != h->pose0_to_frame.end()) { h->ptr_pose_mask[a->second] |= 0x03F; h->ptr_dirty = true; return RSBA_OK; }
  auto b = h->pose1_to_frame.find(block);
  if (b != h->pose1_to_frame.end()) { h->ptr_pose_mask[b->second] |= 0xFC0; h->ptr_dirty = true; return RSBA_OK; }
  auto c = h->point_to_id.find(block);
  if (c != h->point_to_id.end()) { h->ptr_point_const[c->second] = 1; h->ptr_dirty = true; return RSBA_OK; }
  auto d = h->pose_prior_of_block.find(block);
  if (d != h->pose_prior_of_block.end()) { h->pose_priors[d->second].constant = 1; h->pose_priors_dirty = true; return RSBA_OK; }
  return fail(RSBA_ERR_INVALID_ARGUMENT, "unknown parameter block (Ceres would abort here too)");
  });
}

int rsba_cuda_set_subset_constant(rsba_problem* h, double* pose_block, int n_constant,
                                  const int* constant_components) {
  return rsba::api_guard([&]() -> int {
  if (!h || !pose_block || (n_constant > 0 && !constant_components))
    return fail(RSBA_ERR_INVALID_ARGUMENT, "NULL argument");
  int shift, f;
  auto a = h->pose0_to_frame.find(pose_block);
  if (a != h->pose0_to_frame.end()) { f = a->second; shift = 0; }
  else {
    auto b = h->pose1_to_frame.find(pose_block);
    if (b == h->pose1_to_frame.end()) return fail(RSBA_ERR_INVALID_ARGUMENT, "unknown pose block");
    f = b->second; shift = 6;
  }
  for (int i = 0; i < n_constant; ++i) {
    if (constant_components[i] < 0 || constant_components[i] >= 6)
      return fail(RSBA_ERR_INVALID_ARGUMENT, "constant component out of range");
    h->ptr_pose_mask[f] |= (unsigned short)(1u << (shift + constant_components[i]));
  }
  h->ptr_dirty = true;
  return RSBA_OK;
  });
}

int rsba_cuda_set_scene(rsba_problem* h, long n_obs, const double* obs_xy, const int* obs_frame,
                        const int* obs_point, int n_frames, int n_points,
                        const unsigned short* const_pose_mask, const unsigned char* const_point) {
  return rsba::api_guard([&]() -> int {
  if (!h || n_obs < 0 || n_frames < 0 || n_points < 0 || (n_obs > 0 && (!obs_xy || !obs_frame || !obs_point)))
    return fail(RSBA_ERR_INVALID_ARGUMENT, "bad scene arguments");
  if (h->ptr_mode) return fail(RSBA_ERR_STATE, "handle already holds pointer-API residual blocks");
  RSBA_CUDA_TRY(cudaSetDevice(h->device));
  int rc = upload_scene(h, n_obs, obs_xy, obs_frame, obs_point, n_frames, n_points);
  if (rc) return rc;
  h->priors.clear();
  h->priors_dirty = true;
  if ((rc = upload_priors(h))) return rc;
  h->pose_priors.clear();
  h->pose_priors_dirty = true;
  if ((rc = upload_pose_priors(h))) return rc;
  h->pose_mask.assign(n_frames, 0);
  h->point_const.assign(n_points, 0);
  if (const_pose_mask) for (int f = 0; f < n_frames; ++f) h->pose_mask[f] = const_pose_mask[f] & 0xFFF;
  if (const_point) for (int p = 0; p < n_points; ++p) h->point_const[p] = const_point[p] ? 1 : 0;
  return RSBA_OK;
  });
}

int rsba_cuda_set_parameters(rsba_problem* h, const double* poses, const double* points) {
  return rsba::api_guard([&]() -> int {
  if (!h || !h->scene_set) return fail(RSBA_ERR_STATE, "no scene");
  if ((h->n_frames && !poses) || (h->n_points && !points)) return fail(RSBA_ERR_INVALID_ARGUMENT, "NULL argument");
  RSBA_CUDA_TRY(cudaSetDevice(h->device));
  if (h->n_frames)
    RSBA_CUDA_TRY(cudaMemcpyAsync(h->d_poses.ptr, poses, (size_t)kFrameParams * h->n_frames * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  // the pseudo-frame behind the real frames: intrinsics (0..8) and interFrameRatio (9) when they are parameters
  RSBA_CUDA_TRY(cudaMemsetAsync(h->d_poses.ptr + (size_t)kFrameParams * h->n_frames, 0, kFrameParams * sizeof(double), h->stream));
  if (h->free_cam) {
    if (!h->camera_set) return fail(RSBA_ERR_STATE, "rsba_cuda_set_camera before rsba_cuda_set_parameters (free intrinsics)");
    RSBA_CUDA_TRY(cudaMemcpyAsync(h->d_poses.ptr + (size_t)kFrameParams * h->n_frames, h->cm.cam, 9 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  if (h->free_ratio)
    RSBA_CUDA_TRY(cudaMemcpyAsync(h->d_poses.ptr + (size_t)kFrameParams * h->n_frames + 9, &h->ratio_value, sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (h->n_points)
    RSBA_CUDA_TRY(cudaMemcpyAsync(h->d_points.ptr, points, h->d_points.bytes(), cudaMemcpyHostToDevice, h->stream));
  RSBA_CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->params_set = true;
  return RSBA_OK;
  });
}

int rsba_cuda_get_parameters(rsba_problem* h, double* poses, double* points) {
  return rsba::api_guard([&]() -> int {
  if (!h || !h->scene_set || !h->params_set) return fail(RSBA_ERR_STATE, "no parameters on the device");
  RSBA_CUDA_TRY(cudaSetDevice(h->device));
  if (poses && h->n_frames)
    RSBA_CUDA_TRY(cudaMemcpyAsync(poses, h->d_poses.ptr, (size_t)kFrameParams * h->n_frames * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (points && h->n_points)
    RSBA_CUDA_TRY(cudaMemcpyAsync(points, h->d_points.ptr, h->d_points.bytes(), cudaMemcpyDeviceToHost, h->stream));
  RSBA_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return RSBA_OK;
  });
}

static int prepare(rsba_problem* h) {
  if (!h) return fail(RSBA_ERR_INVALID_ARGUMENT, "handle is NULL");
  if (!h->camera_set) return fail(RSBA_ERR_STATE, "rsba_cuda_set_camera has not been called");
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
  return upload_pose_priors(h);
}

int rsba_cuda_evaluate_device(rsba_problem* h, int with_jacobian, double* cost, long* num_invalid) {
  return rsba::api_guard([&]() -> int {
  int rc = prepare(h);
  if (rc) return rc;
  return run_evaluate(h, with_jacobian != 0, h->d_poses.ptr, h->d_points.ptr, cost, num_invalid);
  });
}

int rsba_cuda_evaluate(rsba_problem* h, double* cost, double* residuals, double* jacobian,
                       unsigned char* valid) {
  return rsba::api_guard([&]() -> int {
  int rc = prepare(h);
  if (rc) return rc;
  const bool jac = jacobian != nullptr;
  long bad = 0;
  double c = 0.0;
  // only a requested Jacobian runs the Jacobian kernel (and allocates its 240 bytes per observation)
  rc = run_evaluate(h, jac, h->d_poses.ptr, h->d_points.ptr, &c, &bad, false, residuals || valid);
  if (rc) return rc;
  if (cost) *cost = c;
  const long n = h->n_obs, ng = h->n_obs_global;
  // outputs go back in the caller's observation order; on a multi-GPU rank only the rows of
  // this rank's observations are filled, the others are zero
  bool identity = n == ng;
  for (long i = 0; i < n && identity; ++i) identity = h->order[i] == i;
  if (!identity) {
    if (residuals) memset(residuals, 0, 2 * ng * sizeof(double));
    if (jacobian) memset(jacobian, 0, (size_t)kJacDoubles * ng * sizeof(double));
    if (valid) memset(valid, 0, ng);
  }
  auto dst_of = [&](long i) { return h->order[h->local_ids[i]]; };
  if (residuals && n) {
    if (identity) {
      RSBA_CUDA_TRY(cudaMemcpy(residuals, h->d_res.ptr, 2 * n * sizeof(double), cudaMemcpyDeviceToHost));
    } else {
      std::vector<double> tmp(2 * n);
      RSBA_CUDA_TRY(cudaMemcpy(tmp.data(), h->d_res.ptr, 2 * n * sizeof(double), cudaMemcpyDeviceToHost));
      for (long i = 0; i < n; ++i) memcpy(residuals + 2 * dst_of(i), &tmp[2 * i], 2 * sizeof(double));
    }
  }
  if (jacobian && n) {
    if (identity) {
      RSBA_CUDA_TRY(cudaMemcpy(jacobian, h->d_jac.ptr, (size_t)kJacDoubles * n * sizeof(double), cudaMemcpyDeviceToHost));
    } else {
      std::vector<double> tmp((size_t)kJacDoubles * n);
      RSBA_CUDA_TRY(cudaMemcpy(tmp.data(), h->d_jac.ptr, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
      for (long i = 0; i < n; ++i)
        memcpy(jacobian + (size_t)kJacDoubles * dst_of(i), &tmp[(size_t)kJacDoubles * i], kJacDoubles * sizeof(double));
    }
  }
  if (valid && n) {
    if (identity) {
      RSBA_CUDA_TRY(cudaMemcpy(valid, h->d_valid.ptr, n, cudaMemcpyDeviceToHost));
    } else {
      std::vector<unsigned char> tmp(n);
      RSBA_CUDA_TRY(cudaMemcpy(tmp.data(), h->d_valid.ptr, n, cudaMemcpyDeviceToHost));
      for (long i = 0; i < n; ++i) valid[dst_of(i)] = tmp[i];
    }
  }
  if (bad > 0) return fail(RSBA_ERR_EVALUATION_FAILED, "a cost functor returned false (point behind camera)");
  return RSBA_OK;
  });
}

int rsba_cuda_validate(rsba_problem* h, double sqrd_threshold, double min_distance_to_camera,
                       unsigned char* ok, double* sqrd_error) {
  return rsba::api_guard([&]() -> int {
  int rc = prepare(h);
  if (rc) return rc;
  const long n = h->n_obs, ng = h->n_obs_global;
  rc = ensure_eval_buffers(h, false);
  if (rc) return rc;
  // d_valid / d_res double as the output buffers of the sweep
  launch_validate(h->cm, h->obs_view(), h->d_poses.ptr, h->d_points.ptr, sqrd_threshold, min_distance_to_camera,
                  h->d_valid.ptr, h->d_res.ptr, h->stream);
  h->launches += 1;
  RSBA_CUDA_TRY(cudaGetLastError());
  std::vector<unsigned char> tv(n);
  std::vector<double> te(n);
  if (n) {
    RSBA_CUDA_TRY(cudaMemcpyAsync(tv.data(), h->d_valid.ptr, n, cudaMemcpyDeviceToHost, h->stream));
    RSBA_CUDA_TRY(cudaMemcpyAsync(te.data(), h->d_res.ptr, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  RSBA_CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (ok) memset(ok, 0, ng);
  if (sqrd_error) for (long i = 0; i < ng; ++i) sqrd_error[i] = -1.0;
  for (long i = 0; i < n; ++i) {
    const long dst = h->order[h->local_ids[i]];
    if (ok) ok[dst] = tv[i];
    if (sqrd_error) sqrd_error[dst] = te[i];
  }
  return RSBA_OK;
  });
}

int rsba_cuda_reproject(rsba_problem* h, long n, const int* frame, const int* point, double sqrd_threshold,
                        double* proj_xy, unsigned char* ok) {
  return rsba::api_guard([&]() -> int {
  int rc = prepare(h);
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!frame || !point || !proj_xy || !ok))) return fail(RSBA_ERR_INVALID_ARGUMENT, "bad reproject arguments");
  for (long i = 0; i < n; ++i)
    if (frame[i] < 0 || frame[i] >= h->n_frames || point[i] < 0 || point[i] >= h->n_points)
      return fail(RSBA_ERR_INVALID_ARGUMENT, "reproject: frame / point index out of range");
  if (n == 0) return RSBA_OK;
  DeviceBuffer<int> d_f, d_p;
  DeviceBuffer<double> d_xy;
  DeviceBuffer<unsigned char> d_ok;
  RSBA_CUDA_TRY(d_f.resize(n)); RSBA_CUDA_TRY(d_p.resize(n)); RSBA_CUDA_TRY(d_xy.resize(2 * n)); RSBA_CUDA_TRY(d_ok.resize(n));
  RSBA_CUDA_TRY(cudaMemcpyAsync(d_f.ptr, frame, n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  RSBA_CUDA_TRY(cudaMemcpyAsync(d_p.ptr, point, n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  launch_reproject(h->cm, n, d_f.ptr, d_p.ptr, h->d_poses.ptr, h->d_points.ptr, sqrd_threshold, d_xy.ptr, d_ok.ptr, h->stream);
  h->launches += 1;
  RSBA_CUDA_TRY(cudaGetLastError());
  RSBA_CUDA_TRY(cudaMemcpyAsync(proj_xy, d_xy.ptr, 2 * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  RSBA_CUDA_TRY(cudaMemcpyAsync(ok, d_ok.ptr, n, cudaMemcpyDeviceToHost, h->stream));
  RSBA_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return RSBA_OK;
  });
}

int rsba_cuda_device_buffers(rsba_problem* h, void** residuals, void** jacobian, void** valid,
                             void** poses, void** points) {
  return rsba::api_guard([&]() -> int {
  if (!h) return fail(RSBA_ERR_INVALID_ARGUMENT, "handle is NULL");
  if (residuals) *residuals = h->d_res.ptr;
  if (jacobian) *jacobian = h->d_jac.ptr;
  if (valid) *valid = h->d_valid.ptr;
  if (poses) *poses = h->d_poses.ptr;
  if (points) *points = h->d_points.ptr;
  return RSBA_OK;
  });
}

long rsba_cuda_observation_order(rsba_problem* h, long* order) {
  if (!h) return -1;
  if (order) memcpy(order, h->order.data(), h->order.size() * sizeof(long));
  return (long)h->order.size();
}

long rsba_cuda_launch_count(rsba_problem* h) { return h ? h->launches : 0; }

double rsba_cuda_stage_ms(rsba_problem* h, int stage) {
  if (!h || stage < 0 || stage >= kNumStages) return -1.0;
  return stage_collect(h, (Stage)stage);
}

void rsba_cuda_default_options(rsba_solve_options* o) {
  if (!o) return;
  memset(o, 0, sizeof(*o));
  // recalled Ceres 1.9.0 defaults: include/rsba_ceres_constants.h
  o->max_num_iterations = RSBA_CERES_MAX_NUM_ITERATIONS;
  o->initial_trust_region_radius = RSBA_CERES_INITIAL_TRUST_REGION_RADIUS;
  o->max_trust_region_radius = RSBA_CERES_MAX_TRUST_REGION_RADIUS;
  o->min_trust_region_radius = RSBA_CERES_MIN_TRUST_REGION_RADIUS;
  o->min_relative_decrease = RSBA_CERES_MIN_RELATIVE_DECREASE;
  o->min_lm_diagonal = RSBA_CERES_MIN_LM_DIAGONAL;
  o->max_lm_diagonal = RSBA_CERES_MAX_LM_DIAGONAL;
  o->function_tolerance = RSBA_CERES_FUNCTION_TOLERANCE;
  o->gradient_tolerance = RSBA_CERES_GRADIENT_TOLERANCE;
  o->parameter_tolerance = RSBA_CERES_PARAMETER_TOLERANCE;
  o->jacobi_scaling = RSBA_CERES_JACOBI_SCALING;
  o->huber_loss = 0.0;
  o->verbose = 0;
  o->dense_cholesky = 0;
  o->reorder_tiles = 1;
  o->max_num_consecutive_invalid_steps = RSBA_CERES_MAX_NUM_CONSECUTIVE_INVALID_STEPS;
}

}  // extern "C"
