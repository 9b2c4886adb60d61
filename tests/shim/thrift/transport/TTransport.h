// TEST INFRASTRUCTURE.  Stand-in for <thrift/transport/TTransport.h> (only included by sfm_types.h).
#ifndef RSBA_TEST_SHIM_THRIFT_TTRANSPORT_H_
#define RSBA_TEST_SHIM_THRIFT_TTRANSPORT_H_
#include <thrift/Thrift.h>
#endif
