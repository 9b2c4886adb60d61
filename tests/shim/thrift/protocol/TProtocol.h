// TEST INFRASTRUCTURE.  Stand-in for <thrift/protocol/TProtocol.h>: the reader / writer calls that the
// generated read() / write() methods of gen-cpp/sfm_types.cpp make, all inert (nothing here serialises).
#ifndef RSBA_TEST_SHIM_THRIFT_TPROTOCOL_H_
#define RSBA_TEST_SHIM_THRIFT_TPROTOCOL_H_
#include <thrift/Thrift.h>

namespace apache { namespace thrift { namespace protocol {

enum TType { T_STOP = 0, T_VOID = 1, T_BOOL = 2, T_BYTE = 3, T_I08 = 3, T_I16 = 6, T_I32 = 8, T_U64 = 9, T_I64 = 10,
             T_DOUBLE = 4, T_STRING = 11, T_UTF7 = 11, T_STRUCT = 12, T_MAP = 13, T_SET = 14, T_LIST = 15,
             T_UTF8 = 16, T_UTF16 = 17 };

class TProtocolException : public ::apache::thrift::TException {
 public:
  enum TProtocolExceptionType { UNKNOWN = 0, INVALID_DATA = 1, NEGATIVE_SIZE = 2, SIZE_LIMIT = 3, BAD_VERSION = 4,
                                NOT_IMPLEMENTED = 5, DEPTH_LIMIT = 6 };
  TProtocolException() {}
  explicit TProtocolException(TProtocolExceptionType) {}
};

class TProtocol {
 public:
  virtual ~TProtocol() {}
  uint32_t skip(TType) { return 0; }
  uint32_t readStructBegin(std::string&) { return 0; }
  uint32_t readStructEnd() { return 0; }
  uint32_t readFieldBegin(std::string&, TType& t, int16_t& id) { t = T_STOP; id = 0; return 0; }
  uint32_t readFieldEnd() { return 0; }
  uint32_t readListBegin(TType& t, uint32_t& n) { t = T_STOP; n = 0; return 0; }
  uint32_t readListEnd() { return 0; }
  uint32_t readBool(bool& v) { v = false; return 0; }
  uint32_t readBool(std::vector<bool>::reference v) { v = false; return 0; }
  uint32_t readI16(int16_t& v) { v = 0; return 0; }
  uint32_t readI32(int32_t& v) { v = 0; return 0; }
  uint32_t readI64(int64_t& v) { v = 0; return 0; }
  uint32_t readDouble(double& v) { v = 0; return 0; }
  uint32_t readString(std::string& v) { v.clear(); return 0; }
  uint32_t readBinary(std::string& v) { v.clear(); return 0; }
  uint32_t writeStructBegin(const char*) { return 0; }
  uint32_t writeStructEnd() { return 0; }
  uint32_t writeFieldBegin(const char*, TType, int16_t) { return 0; }
  uint32_t writeFieldEnd() { return 0; }
  uint32_t writeFieldStop() { return 0; }
  uint32_t writeListBegin(TType, uint32_t) { return 0; }
  uint32_t writeListEnd() { return 0; }
  uint32_t writeBool(bool) { return 0; }
  uint32_t writeI16(int16_t) { return 0; }
  uint32_t writeI32(int32_t) { return 0; }
  uint32_t writeI64(int64_t) { return 0; }
  uint32_t writeDouble(double) { return 0; }
  uint32_t writeString(const std::string&) { return 0; }
  uint32_t writeBinary(const std::string&) { return 0; }
};

struct TInputRecursionTracker { explicit TInputRecursionTracker(TProtocol&) {} };
struct TOutputRecursionTracker { explicit TOutputRecursionTracker(TProtocol&) {} };

}}}  // namespace apache::thrift::protocol

#endif
