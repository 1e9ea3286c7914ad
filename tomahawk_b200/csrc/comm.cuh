// comm.cuh -- NCCL data plane of libtwkb (host code): one communicator per context, used ONCE per load
// to complete the packed genotype matrix on every GPU over NVLink / NVSwitch.
//
// Reference analogue: twk_ld::LoadAllBlocks (lib/ld/ld.cpp:370-465) unpacks the whole file once and every
// slave thread of twk_ld::Compute reads it through shared host memory; the balancer
// (lib/ld/ld_balancing.h:23-80, 176-233) then deals block pairs. On one 8 x B200 box the "shared memory" is
// eight HBM stacks: every rank uploads (or decodes) 1/N of the variant rows over ITS OWN PCIe link and the
// ranks complete the matrix with ONE in-place ncclAllGather over NVLink (equal slices of ceil(M / N) rows; the
// resident buffer is padded to N slices). No collective runs after that: tiles are independent (BASELINE.json
// north_star (4)).
//
// libnccl.so.2 is resolved at run time (dlopen): a single-GPU caller never needs it, and inside a process that
// already carries PyTorch's NCCL the same library instance is reused.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>
#include <string>

namespace twkb {

struct NcclApi {
    bool ok = false;
    std::string why;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

inline const NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        if (const char* e = getenv("TWKB_NCCL_LIB")) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
        for (const char* n : names)
            if (!h) h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (!h) { api.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
        auto sym = [&](const char* name) -> void* {
            void* p = dlsym(h, name);
            if (!p && api.why.empty()) api.why = std::string("libnccl lacks ") + name;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.ok = api.why.empty();
    });
    return api;
}

// Rows [begin, end) of rank `rank`: equal slices of ceil(M / N) rows (the last ones may be short or empty).
inline void comm_slice(uint32_t n_variants, int rank, int n_ranks, uint32_t& begin, uint32_t& end) {
    const uint64_t per = ((uint64_t)n_variants + (uint64_t)n_ranks - 1) / (uint64_t)n_ranks;
    begin = (uint32_t)std::min<uint64_t>(n_variants, per * (uint64_t)rank);
    end = (uint32_t)std::min<uint64_t>(n_variants, per * (uint64_t)(rank + 1));
}

}  // namespace twkb
