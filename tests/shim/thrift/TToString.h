// TEST INFRASTRUCTURE.  Stand-in for <thrift/TToString.h>: apache::thrift::to_string for the generated printTo().
#ifndef RSBA_TEST_SHIM_THRIFT_TTOSTRING_H_
#define RSBA_TEST_SHIM_THRIFT_TTOSTRING_H_
#include <sstream>
#include <string>
#include <vector>
namespace apache { namespace thrift {
template <typename T>
std::string to_string(const T& t) {
  std::ostringstream o;
  o << t;
  return o.str();
}
template <typename T>
std::string to_string(const std::vector<T>& v) {
  std::ostringstream o;
  o << "[";
  for (size_t i = 0; i < v.size(); ++i) o << (i ? ", " : "") << to_string(v[i]);
  o << "]";
  return o.str();
}
}}
#endif
