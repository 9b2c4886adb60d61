// "No exceptions cross this boundary" (include/rsba_cuda.h): every int-returning entry point of the C ABI runs its
// body inside api_guard, which turns a C++ exception of the host code (std::bad_alloc from a scene that does not fit
// in host memory, std::system_error, ...) into RSBA_ERR_INTERNAL + rsba_cuda_last_error().
#pragma once
#include <exception>
#include <new>
#include <string>

#include "../../include/rsba_cuda.h"

namespace rsba {

void set_last_error(const std::string& msg);

template <typename Fn>
int api_guard(Fn&& fn) noexcept {
  try {
    return fn();
  } catch (const std::bad_alloc&) {
    try { set_last_error("out of host memory"); } catch (...) {}
  } catch (const std::exception& e) {
    try { set_last_error(std::string("internal error: ") + e.what()); } catch (...) {}
  } catch (...) {
    try { set_last_error("internal error"); } catch (...) {}
  }
  return RSBA_ERR_INTERNAL;
}

}  // namespace rsba
