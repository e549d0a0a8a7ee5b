// nccl_dyn.h -- NCCL bound at run time with dlopen, so that single-GPU users of
// libwvb200.so carry no NCCL dependency and a process that already loaded a
// libnccl.so.2 (e.g. through torch) shares it instead of loading a second one.
// Only the handful of entry points the ghost-plane exchange needs.
#pragma once

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstddef>

namespace wvb {
namespace nccl {

typedef struct ncclComm* comm_t;
struct unique_id {
    char internal[128];
};
enum result_t { success = 0 };
enum data_t { t_int8 = 0, t_uint8 = 1, t_int32 = 2, t_uint64 = 5, t_float32 = 7, t_float64 = 8 };
enum red_t { op_sum = 0, op_prod = 1, op_max = 2, op_min = 3 };

struct api {
    void* handle = nullptr;
    int (*GetUniqueId)(unique_id*) = nullptr;
    int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};

inline api& get() {
    static api a = [] {
        api r;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            r.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (r.handle) break;
        }
        if (!r.handle) return r;
#define WVB_SYM(field, name) \
    r.field = reinterpret_cast<decltype(r.field)>(dlsym(r.handle, name))
        WVB_SYM(GetUniqueId, "ncclGetUniqueId");
        WVB_SYM(CommInitRank, "ncclCommInitRank");
        WVB_SYM(CommDestroy, "ncclCommDestroy");
        WVB_SYM(GroupStart, "ncclGroupStart");
        WVB_SYM(GroupEnd, "ncclGroupEnd");
        WVB_SYM(Send, "ncclSend");
        WVB_SYM(Recv, "ncclRecv");
        WVB_SYM(AllReduce, "ncclAllReduce");
        WVB_SYM(GetErrorString, "ncclGetErrorString");
#undef WVB_SYM
        r.ok = r.GetUniqueId && r.CommInitRank && r.CommDestroy && r.GroupStart && r.GroupEnd &&
               r.Send && r.Recv && r.AllReduce && r.GetErrorString;
        return r;
    }();
    return a;
}

}  // namespace nccl
}  // namespace wvb
