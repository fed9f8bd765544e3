// lbm_nccl.h -- NCCL bound at run time (dlopen), so that liblbm_b200.so loads on a box without
// NCCL for single-GPU work and shares the process's libnccl.so.2 when torch already loaded one.
// Only the calls the halo exchange needs (replacing MPI_Isend/Irecv/Waitall of the reference,
// include/LBMGrid.h:255-280).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

namespace lbm {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    bool ok = false;
    const char* why = "";
};

// Loads libnccl.so.2 on first use; returns the same table afterwards.
const NcclApi& nccl_api();

}  // namespace lbm
