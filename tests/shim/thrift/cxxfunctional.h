// TEST INFRASTRUCTURE.  Stand-in for <thrift/cxxfunctional.h>.
#ifndef RSBA_TEST_SHIM_THRIFT_CXXFUNCTIONAL_H_
#define RSBA_TEST_SHIM_THRIFT_CXXFUNCTIONAL_H_
#include <functional>
#endif
