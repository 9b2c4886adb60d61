// TEST INFRASTRUCTURE.  Stand-in for <thrift/TApplicationException.h> (only included, never used by sfm_types).
#ifndef RSBA_TEST_SHIM_THRIFT_TAPPEXC_H_
#define RSBA_TEST_SHIM_THRIFT_TAPPEXC_H_
#include <thrift/Thrift.h>
#endif
