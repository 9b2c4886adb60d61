// One host thread, N GPUs of one node.
//
// The reference's driver is single-threaded: VideoSfMHandler::BA builds ONE CeresHandler and calls
// ceres::Solve from the request thread (VideoSfMHandler.cc:574-631, CeresHandler.h:408-419).  A drop-in that
// wants every GPU of the box therefore cannot ask that thread to become N ranks.  rsba_multi keeps the
// rank-per-GPU design of the solver (point-owner sharding, one NCCL all-reduce of the reduced system per
// linearisation) and hides the ranks behind one handle: one rsba_problem per device, the communicator built by
// ncclCommInitRank from N short-lived worker threads of THIS process, builder calls forwarded to every rank by
// the caller's single thread (rsba_cuda_multi_handle), and rsba_cuda_multi_solve running the ranks' LM loops on
// N worker threads and joining them before it returns.  NCCL is only loaded when n_devices > 1.
#include "problem.cuh"

#include <memory>
#include <thread>

struct rsba_multi {
  std::vector<rsba_problem*> ranks;
};

namespace rsba {
namespace {
int mfail(int code, const std::string& msg) {
  set_last_error(msg);
  return code;
}
}  // namespace
}  // namespace rsba

using namespace rsba;

extern "C" {

int rsba_cuda_create_multi(rsba_multi** out, const int* devices, int n_devices) {
  return rsba::api_guard([&]() -> int {
  if (!out) return mfail(RSBA_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (!devices || n_devices < 1 || n_devices > 64) return mfail(RSBA_ERR_INVALID_ARGUMENT, "bad device list");
  for (int a = 0; a < n_devices; ++a)
    for (int b = 0; b < a; ++b)
      if (devices[a] == devices[b]) return mfail(RSBA_ERR_INVALID_ARGUMENT, "a device may hold one rank only");
  std::unique_ptr<rsba_multi> m(new rsba_multi);
  auto destroy_all = [&]() { for (rsba_problem* h : m->ranks) rsba_cuda_destroy(h); m->ranks.clear(); };
  for (int r = 0; r < n_devices; ++r) {
    rsba_problem* h = nullptr;
    const int rc = rsba_cuda_create(&h, devices[r]);
    if (rc) { destroy_all(); return rc; }
    h->scatter_owner = r == 0;   // pointer API: one rank writes the caller's blocks back
    m->ranks.push_back(h);
  }
  if (n_devices > 1) {
    unsigned char id[128];
    int rc = rsba_cuda_nccl_unique_id(id);
    if (rc) { destroy_all(); return rc; }
    // ncclCommInitRank blocks until every rank has joined: one thread per rank
    std::vector<int> rcs(n_devices, 0);
    std::vector<std::string> errs(n_devices);
    std::vector<std::thread> th;
    for (int r = 0; r < n_devices; ++r)
      th.emplace_back([&, r] {
        rcs[r] = rsba_cuda_comm_init(m->ranks[r], r, n_devices, id);
        if (rcs[r]) errs[r] = rsba_cuda_last_error();
      });
    for (auto& t : th) t.join();
    for (int r = 0; r < n_devices; ++r)
      if (rcs[r]) { rc = rcs[r]; set_last_error(errs[r]); destroy_all(); return rc; }
  }
  *out = m.release();
  return RSBA_OK;
  });
}

int rsba_cuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int rsba_cuda_multi_size(const rsba_multi* m) { return m ? (int)m->ranks.size() : 0; }

rsba_problem* rsba_cuda_multi_handle(rsba_multi* m, int rank) {
  if (!m || rank < 0 || rank >= (int)m->ranks.size()) return nullptr;
  return m->ranks[rank];
}

int rsba_cuda_multi_solve(rsba_multi* m, const rsba_solve_options* options, rsba_solve_summary* summary) {
  return rsba::api_guard([&]() -> int {
  if (!m || m->ranks.empty()) return mfail(RSBA_ERR_INVALID_ARGUMENT, "handle is NULL");
  const int n = (int)m->ranks.size();
  if (n == 1) return rsba_cuda_solve(m->ranks[0], options, summary);
  std::vector<int> rcs(n, 0);
  std::vector<std::string> errs(n);
  std::vector<rsba_solve_summary> sums(n);
  std::vector<std::thread> th;
  // every rank runs the same LM loop on its share; the all-reduces inside keep them in step
  for (int r = 0; r < n; ++r)
    th.emplace_back([&, r] {
      rcs[r] = rsba_cuda_solve(m->ranks[r], options, &sums[r]);
      if (rcs[r]) errs[r] = rsba_cuda_last_error();
    });
  for (auto& t : th) t.join();
  if (summary) *summary = sums[0];
  for (int r = 0; r < n; ++r)
    if (rcs[r]) { set_last_error(errs[r]); return rcs[r]; }
  return RSBA_OK;
  });
}

void rsba_cuda_destroy_multi(rsba_multi* m) {
  if (!m) return;
  // ncclCommDestroy of one rank may wait for its peers: tear the ranks down concurrently
  std::vector<std::thread> th;
  for (rsba_problem* h : m->ranks) th.emplace_back([h] { rsba_cuda_destroy(h); });
  for (auto& t : th) t.join();
  delete m;
}

}  // extern "C"
