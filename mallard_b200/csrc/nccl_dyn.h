// NCCL, resolved at run time.  The library has no link-time dependency on libnccl: single-GPU hosts never load it, and a
// process that already carries an NCCL (PyTorch bundles its own) keeps using that one.  Only the native multi-GPU driver
// (mlb_comm_init / mlb_run_distributed, api.cu) calls through this table.
#pragma once
#include <nccl.h>

namespace mlb {

struct NcclApi {
    ncclResult_t (*GetVersion)(int *);
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, ncclConfig_t *);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    const char * (*GetErrorString)(ncclResult_t);
    const char * path;   // what was loaded
};

// dlopen on first use: $MLB_NCCL_LIB, else the libnccl.so.2 already in the process, else the system's.  Throws
// std::runtime_error when no NCCL can be loaded.
const NcclApi & nccl();

}  // namespace mlb
