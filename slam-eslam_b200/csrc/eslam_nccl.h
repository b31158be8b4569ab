// eslam_nccl.h -- NCCL entry points resolved at run time (dlopen), so libeslam_gpu.so has no link-time dependency on
// NCCL and shares whichever libnccl.so.2 the process already loaded (PyTorch bundles one).  Used for the exchange steps
// of the sharded filter: all-gathers of per-tile partials / descriptors and the contiguous-range send/recv of resampling.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

typedef struct { char internal[128]; } eslam_ncclUniqueId;
typedef void* eslam_ncclComm_t;
enum { ESLAM_NCCL_UINT8 = 1 };   // ncclUint8 / ncclChar's unsigned sibling in nccl.h's ncclDataType_t

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(eslam_ncclUniqueId*) = nullptr;
    int (*CommInitRank)(eslam_ncclComm_t*, int, eslam_ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(eslam_ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, eslam_ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, eslam_ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, eslam_ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;

    bool load() {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);            // the one already in the process, if any
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
#define ESLAM_NCCL_SYM(field, name) *(void**)(&field) = dlsym(lib, name); if (!field) return false;
        ESLAM_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        ESLAM_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        ESLAM_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        ESLAM_NCCL_SYM(AllGather, "ncclAllGather")
        ESLAM_NCCL_SYM(Send, "ncclSend")
        ESLAM_NCCL_SYM(Recv, "ncclRecv")
        ESLAM_NCCL_SYM(GroupStart, "ncclGroupStart")
        ESLAM_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        ESLAM_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef ESLAM_NCCL_SYM
        return true;
    }
};
