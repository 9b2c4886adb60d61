// A handful of host threads for the one-off host work (scene intake, structure analysis): std::thread per phase,
// the phases take milliseconds each.  RSBA_CUDA_HOST_THREADS overrides the count; small inputs run on the calling
// thread alone.  Plus a std::vector whose resize() leaves trivially constructible elements uninitialised: the big
// host arrays are written exactly once by those threads, and a serial zero fill of 20-80 MB each would cost as much
// as the phase that fills them.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <exception>
#include <memory>
#include <system_error>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

namespace rsba {

struct HostThreads {
  int n = 1;
  // sharers: processes that do the same work at the same time on this host (the ranks of a multi-GPU run)
  explicit HostThreads(long work, int sharers = 1) {
    int want = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency() / (unsigned)std::max(1, sharers)));
    if (const char* e = getenv("RSBA_CUDA_HOST_THREADS")) want = std::max(1, std::min(64, atoi(e)));
    else if (work < 100000) want = 1;
    n = want;
  }
  // fn(t) on threads t = 0 .. n-1 (t = 0 is the caller); returns when all are done
  // An exception of any share (std::bad_alloc in a worker, ...) is rethrown on the caller AFTER every thread
  // has been joined: it reaches api_guard as RSBA_ERR_INTERNAL instead of std::terminate.
  template <typename Fn>
  void run(Fn&& fn) const {
    if (n == 1) { fn(0); return; }
    std::vector<std::thread> th;
    std::vector<std::exception_ptr> err(n);
    auto guarded = [&fn, &err](int t) {
      try { fn(t); } catch (...) { err[t] = std::current_exception(); }
    };
    int started = 1;
    try {
      th.reserve(n - 1);
      for (int t = 1; t < n; ++t) { th.emplace_back(guarded, t); ++started; }
    } catch (...) {
      // no more threads to be had (pid / resource limit, memory): the caller takes the remaining shares itself
    }
    guarded(0);
    for (int t = started; t < n; ++t) guarded(t);
    for (auto& x : th) x.join();
    for (auto& e : err)
      if (e) std::rethrow_exception(e);
  }
  // fn(t, begin, end) over an even split of [0, count)
  template <typename Fn>
  void split(long count, Fn&& fn) const {
    const int nt = n;
    run([&](int t) { fn(t, count * t / nt, count * (t + 1) / nt); });
  }
  // v.resize(count) without a serial fill, then every thread writes `value` over its share
  template <typename Vec, typename T>
  void resize_fill(Vec& v, size_t count, T value) const {
    v.resize(count);
    auto* d = v.data();
    const int nt = n;
    run([&](int t) { std::fill(d + count * t / nt, d + count * (t + 1) / nt, value); });
  }
};

template <typename T>
struct DefaultInitAllocator : std::allocator<T> {
  template <typename U> struct rebind { using other = DefaultInitAllocator<U>; };
  using std::allocator<T>::allocator;
  template <typename U> void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void*>(p)) U; }
  template <typename U, typename... Args> void construct(U* p, Args&&... args) { ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...); }
};
template <typename T>
using HostVec = std::vector<T, DefaultInitAllocator<T>>;

}  // namespace rsba
