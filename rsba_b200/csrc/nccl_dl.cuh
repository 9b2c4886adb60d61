// NCCL through dlopen: the product library has no link-time dependency on NCCL, single-GPU use
// never loads it, and inside a PyTorch process the already-loaded libnccl.so.2 is reused.
#pragma once
#include <nccl.h>

namespace rsba {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

// nullptr (and last_error set) if libnccl.so.2 cannot be loaded
const NcclApi* nccl_api();

}  // namespace rsba
