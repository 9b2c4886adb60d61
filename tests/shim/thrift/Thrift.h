// TEST INFRASTRUCTURE.  Stand-in for Apache Thrift 0.9.3's <thrift/Thrift.h> (absent from the build
// container): just enough of its declarations for the reference's generated gen-cpp/sfm_types.{h,cpp} to
// compile IN PLACE from /root/reference, so that include/rsba_cuda_handler.hpp can be instantiated with the
// reference's own gen::Session.  Nothing is (de)serialised: the protocol object below is inert.
#ifndef RSBA_TEST_SHIM_THRIFT_H_
#define RSBA_TEST_SHIM_THRIFT_H_

#include <stdint.h>

#include <exception>
#include <map>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

namespace apache { namespace thrift {

class TException : public std::exception {
 public:
  TException() {}
  explicit TException(const std::string& m) : message_(m) {}
  virtual ~TException() throw() {}
  virtual const char* what() const throw() { return message_.c_str(); }
 protected:
  std::string message_;
};

// iterator over (value, name) pairs that initialises the generated *_VALUES_TO_NAMES maps
class TEnumIterator : public std::iterator<std::forward_iterator_tag, std::pair<int, const char*> > {
 public:
  TEnumIterator(int n, int* enums, const char** names) : ii_(0), n_(n), enums_(enums), names_(names) {}
  int operator++() { return ++ii_; }
  bool operator!=(const TEnumIterator&) { return ii_ != n_; }
  std::pair<int, const char*> operator*() const { return std::make_pair(enums_[ii_], names_[ii_]); }
 private:
  int ii_;
  const int n_;
  int* enums_;
  const char** names_;
};

}}  // namespace apache::thrift

#endif
