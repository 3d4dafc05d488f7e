// kb2_comm.h — NCCL communicator behind the C ABI (one all-gather of per-shard candidates over NVLink).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2): a host process that already carries NCCL (e.g. through
// torch.distributed) shares that copy, a plain C++ host picks up the system library, and a single-GPU host never
// needs it.  Bootstrap follows NCCL's own model: rank 0 calls kb2_comm_unique_id(), the host application ships the
// 128 bytes to the other ranks by whatever means it has (Milvus: its RPC layer; the tests/bench: torch.distributed),
// every rank calls kb2_comm_create().  The reference has no multi-GPU path at all (one index per device,
// src/common/cuvs/integration/cuvs_knowhere_index.cuh:415-460).
#pragma once
#include <dlfcn.h>
#include <nccl.h>   // types and enums only; every function is looked up with dlsym

#include "kb2_common.cuh"

namespace kb2 {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    std::string why;

    static NcclApi&
    get() {
        static NcclApi api = [] {
            NcclApi a;
            void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy the host process already uses, if any
            if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            if (!h) {
                a.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
                return a;
            }
            bool all = true;
            auto sym = [&](const char* n) {
                void* p = dlsym(h, n);
                if (!p) { all = false; a.why = std::string("missing NCCL symbol ") + n; }
                return p;
            };
            a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
            a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
            a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
            a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
            a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
            a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
            a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
            a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
            a.ok = all;
            return a;
        }();
        return api;
    }
};

#define KB2_NCCL_CHECK(expr)                                                                               \
    do {                                                                                                   \
        ncclResult_t _r = (expr);                                                                          \
        if (_r != ncclSuccess)                                                                             \
            throw ::kb2::Error(KB2_CUDA_RUNTIME_ERROR, std::string(#expr) + ": " +                         \
                                                           ::kb2::NcclApi::get().GetErrorString(_r));      \
    } while (0)

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;

    ~Comm() {
        if (comm) NcclApi::get().CommDestroy(comm);
    }
    static NcclApi&
    api() {
        NcclApi& a = NcclApi::get();
        KB2_REQUIRE(a.ok, KB2_CUDA_RUNTIME_ERROR, "NCCL unavailable: " + a.why);
        return a;
    }
    // every rank contributes `bytes`; recv holds world * bytes (rank-major)
    void
    all_gather(const void* send, void* recv, size_t bytes, cudaStream_t st) const {
        KB2_NCCL_CHECK(api().AllGather(send, recv, bytes, ncclUint8, comm, st));
    }
    // two buffers in one fused launch
    void
    all_gather2(const void* s0, void* r0, size_t b0, const void* s1, void* r1, size_t b1, cudaStream_t st) const {
        NcclApi& a = api();
        KB2_NCCL_CHECK(a.GroupStart());
        KB2_NCCL_CHECK(a.AllGather(s0, r0, b0, ncclUint8, comm, st));
        KB2_NCCL_CHECK(a.AllGather(s1, r1, b1, ncclUint8, comm, st));
        KB2_NCCL_CHECK(a.GroupEnd());
    }
    void
    all_reduce_min_f32(const float* send, float* recv, size_t count, cudaStream_t st) const {
        KB2_NCCL_CHECK(api().AllReduce(send, recv, count, ncclFloat32, ncclMin, comm, st));
    }
    void
    all_reduce_max_u32(const uint32_t* send, uint32_t* recv, size_t count, cudaStream_t st) const {
        KB2_NCCL_CHECK(api().AllReduce(send, recv, count, ncclUint32, ncclMax, comm, st));
    }
};

}  // namespace kb2
