// C-ABI access to K3 on its own: the task graph of the numeric phase (host only) and a solve of a
// caller-supplied reduced system on the device.  Neither is on the reference's path -- Ceres hides
// CHOLMOD behind ceres::Solve (CeresHandler.h:403,419) -- they exist so that the tests can pin the
// factorisation kernels against LAPACK on arbitrary block-sparsity patterns, dense ones included.
#include "api_guard.h"
#include "lm.cuh"
#include "problem.cuh"

#include <cstring>
#include <vector>

using namespace rsba;

namespace {

// ---- FP64 peak of the device this process runs on (the denominator of the K2/K3 rooflines): chains of
// register-resident DFMAs and of mma.sync.m8n8k4.f64, the two FP64 paths of sm_100a (tcgen05 has no FP64 kind)
template <int ILP>
__global__ void peak_dfma_kernel(double* out, int iters) {
  double a[ILP], b = 1.0000001, c = 1e-9;
  for (int k = 0; k < ILP; ++k) a[k] = threadIdx.x + k;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int k = 0; k < ILP; ++k) a[k] = fma(a[k], b, c);
  double s = 0;
  for (int k = 0; k < ILP; ++k) s += a[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void peak_dmma_kernel(double* out, int iters) {
  double c0[ILP], c1[ILP];
  const double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int k = 0; k < ILP; ++k) c0[k] = c1[k] = 0.0;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int k = 0; k < ILP; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[k]), "+d"(c1[k]) : "d"(a), "d"(b));
  double s = 0;
  for (int k = 0; k < ILP; ++k) s += c0[k] + c1[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int fail(int code, const std::string& msg) {
  set_last_error(msg);
  return code;
}

int make_plan(int n_tiles, int n_pairs, const int* pair_a, const int* pair_b, int dense, int reorder, TilePlan* plan) {
  if (n_tiles < 0 || n_pairs < 0 || (n_pairs > 0 && (!pair_a || !pair_b)))
    return fail(RSBA_ERR_INVALID_ARGUMENT, "bad plan arguments");
  std::vector<std::pair<int, int>> tp;
  for (int k = 0; k < n_pairs; ++k) {
    if (pair_a[k] < 0 || pair_b[k] < 0 || pair_a[k] >= n_tiles || pair_b[k] >= n_tiles)
      return fail(RSBA_ERR_INVALID_ARGUMENT, "tile index out of range");
    tp.emplace_back(std::min(pair_a[k], pair_b[k]), std::max(pair_a[k], pair_b[k]));
  }
  build_tile_plan(n_tiles, tp, dense != 0, reorder != 0, -1, plan);
  return RSBA_OK;
}

template <typename T>
int upload(DeviceBuffer<T>& d, const std::vector<T>& h) {
  RSBA_CUDA_TRY(d.resize(std::max<size_t>(h.size(), 1)));
  if (!h.empty()) RSBA_CUDA_TRY(cudaMemcpy(d.ptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return RSBA_OK;
}

}  // namespace

extern "C" {

int rsba_cuda_plan_task_graph(int n_tiles, int n_pairs, const int* pair_a, const int* pair_b, int dense, int reorder,
                              int merge_levels, long counts[4], int* tasks, int* sources, int* need) {
  return rsba::api_guard([&]() -> int {
    if (!counts) return fail(RSBA_ERR_INVALID_ARGUMENT, "counts is NULL");
    TilePlan plan;
    int rc = make_plan(n_tiles, n_pairs, pair_a, pair_b, dense, reorder, &plan);
    if (rc) return rc;
    DagPlan dag;
    build_dag_plan(plan, merge_levels, &dag);
    counts[0] = (long)dag.tasks.size(); counts[1] = (long)dag.sources.size(); counts[2] = (long)plan.nz_tiles.size();
    counts[3] = dag.n_factor_tasks;
    static_assert(sizeof(DagTask) == 8 * sizeof(int), "task records are 8 ints");
    if (tasks && !dag.tasks.empty()) memcpy(tasks, dag.tasks.data(), dag.tasks.size() * sizeof(DagTask));
    if (sources && !dag.sources.empty()) memcpy(sources, dag.sources.data(), dag.sources.size() * sizeof(int2));
    if (need) memcpy(need, dag.need.data(), plan.nz_tiles.size() * 4 * sizeof(int));
    return RSBA_OK;
  });
}

int rsba_cuda_reduced_solve(int device, int n_tiles, int n_pairs, const int* pair_a, const int* pair_b, int dense,
                            int reorder, int mode, int merge_levels, int repeats, const double* A,
                            const double* rhs, double* x_out, double* L_out, int* tile_pos_out, int* info_out,
                            float* ms_out, long long* trace_out) {
  return rsba::api_guard([&]() -> int {
    if (n_tiles <= 0 || !A || !rhs || !x_out) return fail(RSBA_ERR_INVALID_ARGUMENT, "bad arguments");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
      cudaGetLastError();
      return fail(RSBA_ERR_NO_DEVICE, "no such CUDA device");
    }
    RSBA_CUDA_TRY(cudaSetDevice(device));
    TilePlan plan;
    int rc = make_plan(n_tiles, n_pairs, pair_a, pair_b, dense, reorder, &plan);
    if (rc) return rc;
    DagPlan dag;
    build_dag_plan(plan, merge_levels, &dag);
    const int T = n_tiles;
    const long n = (long)T * kTile;
    // ---- pack: tile (i, j) of the permuted matrix, row-major; diagonal tiles as full squares
    const size_t n_nz = plan.nz_tiles.size();
    std::vector<double> packed(n_nz * kTile * kTile);
    for (size_t s = 0; s < n_nz; ++s) {
      const int2 t = plan.nz_tiles[s];
      const long r0 = (long)plan.pos_tile[t.x] * kTile, c0 = (long)plan.pos_tile[t.y] * kTile;
      for (int r = 0; r < kTile; ++r)
        for (int c = 0; c < kTile; ++c) packed[(s * kTile + r) * kTile + c] = A[(r0 + r) * n + c0 + c];
    }
    std::vector<double> b((size_t)n);
    for (long k = 0; k < n; ++k) b[(size_t)plan.tile_pos[k / kTile] * kTile + k % kTile] = rhs[k];
    std::vector<int> fwd_slot(std::max<size_t>(plan.trsm.size(), 1), 0);
    for (size_t t = 0; t < plan.trsm.size(); ++t) {
      const int i = plan.trsm[t].x, k = plan.trsm[t].y;
      for (int q = plan.lrow_ptr[i]; q < plan.lrow_ptr[i + 1]; ++q)
        if (plan.lrow_cols[q] == k) fwd_slot[t] = q;
    }
    // ---- device copies
    DeviceBuffer<double> S, x, Dinv, solve_partials, fwd_partials, bwd_partials;
    DeviceBuffer<int2> nz_tiles, trsm, sources;
    DeviceBuffer<int4> upd;
    DeviceBuffer<int> tile_slot, row_ptr, rows, lrow_ptr, lrow_cols, panels, d_fwd_slot, need, counters, info;
    DeviceBuffer<DagTask> tasks;
    if ((rc = upload(S, packed)) || (rc = upload(x, b)) || (rc = upload(nz_tiles, plan.nz_tiles)) ||
        (rc = upload(trsm, plan.trsm)) || (rc = upload(upd, plan.upd)) || (rc = upload(tile_slot, plan.tile_slot)) ||
        (rc = upload(row_ptr, plan.row_ptr)) || (rc = upload(rows, plan.rows)) || (rc = upload(lrow_ptr, plan.lrow_ptr)) ||
        (rc = upload(lrow_cols, plan.lrow_cols)) || (rc = upload(panels, plan.panels)) ||
        (rc = upload(d_fwd_slot, fwd_slot)) || (rc = upload(tasks, dag.tasks)) || (rc = upload(sources, dag.sources)) ||
        (rc = upload(need, dag.need)))
      return rc;
    RSBA_CUDA_TRY(Dinv.resize((size_t)T * kTile * kTile));
    RSBA_CUDA_TRY(solve_partials.resize((size_t)T * 16 * kTile));
    RSBA_CUDA_TRY(fwd_partials.resize(std::max<size_t>(plan.lrow_cols.size(), 1) * kTile));
    RSBA_CUDA_TRY(bwd_partials.resize(std::max<size_t>(plan.rows.size(), 1) * kTile));
    RSBA_CUDA_TRY(info.resize(4));
    RSBA_CUDA_TRY(cudaMemset(info.ptr, 0, 4 * sizeof(int)));
    TileSchedule ts{};
    ts.n_tiles = T; ts.nz_tiles = nz_tiles.ptr; ts.tile_slot = tile_slot.ptr; ts.n_nz = (int)n_nz;
    ts.row_ptr = row_ptr.ptr; ts.rows = rows.ptr; ts.upd = upd.ptr; ts.panels = panels.ptr; ts.trsm = trsm.ptr;
    ts.lrow_ptr = lrow_ptr.ptr; ts.lrow_cols = lrow_cols.ptr; ts.Dinv = Dinv.ptr; ts.solve_partials = solve_partials.ptr;
    ts.n_real = n; ts.fwd_slot = d_fwd_slot.ptr; ts.fwd_partials = fwd_partials.ptr;
    RSBA_CUDA_TRY(counters.resize(dag_counter_ints(ts)));
    DagDevice dd{tasks.ptr, (int)dag.tasks.size(), dag.n_factor_tasks, sources.ptr, need.ptr, counters.ptr,
                 bwd_partials.ptr};
    DeviceBuffer<long long> trace;
    if (trace_out && mode == 0) {
      RSBA_CUDA_TRY(trace.resize(std::max<size_t>(dag.tasks.size(), 1) * 16));
      RSBA_CUDA_TRY(cudaMemset(trace.ptr, 0, trace.bytes()));
      dd.trace = trace.ptr;
    }
    cudaStream_t s = nullptr;
    RSBA_CUDA_TRY(cudaStreamCreate(&s));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k3_prepare();
    cudaError_t err = cudaSuccess;
    float ms = 0.f;
    for (int rep = 0; rep < std::max(repeats, 1) && err == cudaSuccess; ++rep) {   // fastest of `repeats` runs
      if (rep > 0) {
        cudaMemcpyAsync(S.ptr, packed.data(), packed.size() * sizeof(double), cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(x.ptr, b.data(), b.size() * sizeof(double), cudaMemcpyHostToDevice, s);
        cudaMemsetAsync(info.ptr, 0, 4 * sizeof(int), s);
      }
      cudaEventRecord(e0, s);
      if (mode == 0) {
        launch_tile_dag(S.ptr, ts, dd, x.ptr, info.ptr, true, true, s);
      } else {
        launch_tile_cholesky(S.ptr, ts, plan, x.ptr, info.ptr, s);
        launch_tile_solve(S.ptr, ts, plan, x.ptr, s);
      }
      cudaEventRecord(e1, s);
      err = cudaStreamSynchronize(s);
      float t = 0.f;
      if (err == cudaSuccess) cudaEventElapsedTime(&t, e0, e1);
      if (rep == 0 || t < ms) ms = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(s);
    if (err != cudaSuccess) return fail(RSBA_ERR_CUDA, std::string("reduced solve: ") + cudaGetErrorString(err));
    RSBA_CUDA_TRY(cudaGetLastError());
    if (ms_out) *ms_out = ms;
    if (dd.trace)   // time stamps of the LAST run
      RSBA_CUDA_TRY(cudaMemcpy(trace_out, trace.ptr, dag.tasks.size() * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
    int h_info = 0;
    RSBA_CUDA_TRY(cudaMemcpy(&h_info, info.ptr, sizeof(int), cudaMemcpyDeviceToHost));
    if (info_out) *info_out = h_info;
    RSBA_CUDA_TRY(cudaMemcpy(b.data(), x.ptr, n * sizeof(double), cudaMemcpyDeviceToHost));
    for (long k = 0; k < n; ++k) x_out[k] = b[(size_t)plan.tile_pos[k / kTile] * kTile + k % kTile];
    if (tile_pos_out) std::copy(plan.tile_pos.begin(), plan.tile_pos.begin() + T, tile_pos_out);
    if (L_out) {   // the factor in PERMUTED order (position tiles), dense n x n, zeros outside the pattern
      RSBA_CUDA_TRY(cudaMemcpy(packed.data(), S.ptr, packed.size() * sizeof(double), cudaMemcpyDeviceToHost));
      memset(L_out, 0, sizeof(double) * n * n);
      for (size_t sl = 0; sl < n_nz; ++sl) {
        const int2 t = plan.nz_tiles[sl];
        for (int r = 0; r < kTile; ++r)
          for (int c = 0; c < kTile; ++c) {
            if (t.x == t.y && c > r) continue;
            L_out[((long)t.x * kTile + r) * n + (long)t.y * kTile + c] = packed[(sl * kTile + r) * kTile + c];
          }
      }
    }
    return RSBA_OK;
  });
}

int rsba_cuda_measure_fp64_peak(int device, double* dfma_tflops, double* dmma_tflops) {
  return rsba::api_guard([&]() -> int {
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
      cudaGetLastError();
      return fail(RSBA_ERR_NO_DEVICE, "no such CUDA device");
    }
    RSBA_CUDA_TRY(cudaSetDevice(device));
    int sms = 0;
    RSBA_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int threads = 1024, grid = 2 * sms, iters = 4000;
    DeviceBuffer<double> out;
    RSBA_CUDA_TRY(out.resize((size_t)grid * threads));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best[2] = {0.0, 0.0};
    for (int rep = 0; rep < 4; ++rep)          // first round = warm-up (clocks), best of the rest
      for (int which = 0; which < 2; ++which) {
        cudaEventRecord(e0);
        if (which == 0) peak_dfma_kernel<8><<<grid, threads>>>(out.ptr, iters);
        else peak_dmma_kernel<8><<<grid, threads>>>(out.ptr, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = which == 0 ? 2.0 * 8 * iters * (double)threads * grid
                                       : 2.0 * 8 * 8 * 4 * 8 * iters * (double)(threads / 32) * grid;
        if (rep > 0 && ms > 0.f) best[which] = std::max(best[which], flop / ms / 1e9);
      }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    RSBA_CUDA_TRY(cudaGetLastError());
    if (dfma_tflops) *dfma_tflops = best[0];
    if (dmma_tflops) *dmma_tflops = best[1];
    return RSBA_OK;
  });
}

}  // extern "C"
