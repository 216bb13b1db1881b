// NCCL plumbing (one process per GPU). libnccl is opened at run time so that single-GPU use has
// no NCCL dependency; in a torch process this resolves to the libnccl.so.2 torch already loaded.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <string>
#include <cstring>
#include <stdint.h>

namespace rq {

struct NcclId { char internal[128]; };
typedef int (*nccl_get_unique_id_t)(NcclId*);
typedef int (*nccl_comm_init_rank_t)(void**, int, NcclId, int);
typedef int (*nccl_comm_destroy_t)(void*);
typedef int (*nccl_all_gather_t)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*nccl_all_reduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_broadcast_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_send_t)(const void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_recv_t)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_group_t)(void);
typedef const char* (*nccl_get_error_string_t)(int);

struct Dist {
    void* lib = nullptr;
    void* comm = nullptr;
    int rank = 0, world = 1;
    nccl_get_unique_id_t get_unique_id = nullptr;
    nccl_comm_init_rank_t comm_init_rank = nullptr;
    nccl_comm_destroy_t comm_destroy = nullptr;
    nccl_all_gather_t all_gather = nullptr;
    nccl_all_reduce_t all_reduce = nullptr;
    nccl_broadcast_t broadcast = nullptr;
    nccl_send_t send = nullptr;
    nccl_recv_t recv = nullptr;
    nccl_group_t group_start = nullptr, group_end = nullptr;
    nccl_get_error_string_t get_error_string = nullptr;
};

inline bool dist_load(Dist& d, std::string& err) {
    if (d.lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        d.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (d.lib) break;
    }
    if (!d.lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
    d.get_unique_id = (nccl_get_unique_id_t)dlsym(d.lib, "ncclGetUniqueId");
    d.comm_init_rank = (nccl_comm_init_rank_t)dlsym(d.lib, "ncclCommInitRank");
    d.comm_destroy = (nccl_comm_destroy_t)dlsym(d.lib, "ncclCommDestroy");
    d.all_gather = (nccl_all_gather_t)dlsym(d.lib, "ncclAllGather");
    d.all_reduce = (nccl_all_reduce_t)dlsym(d.lib, "ncclAllReduce");
    d.broadcast = (nccl_broadcast_t)dlsym(d.lib, "ncclBroadcast");
    d.send = (nccl_send_t)dlsym(d.lib, "ncclSend");
    d.recv = (nccl_recv_t)dlsym(d.lib, "ncclRecv");
    d.group_start = (nccl_group_t)dlsym(d.lib, "ncclGroupStart");
    d.group_end = (nccl_group_t)dlsym(d.lib, "ncclGroupEnd");
    d.get_error_string = (nccl_get_error_string_t)dlsym(d.lib, "ncclGetErrorString");
    if (!d.get_unique_id || !d.comm_init_rank || !d.all_gather || !d.all_reduce || !d.broadcast || !d.send || !d.recv || !d.group_start || !d.group_end) {
        err = "libnccl lacks required symbols";
        return false;
    }
    return true;
}

inline Dist& dist_singleton_for_id() { static Dist d; return d; }

inline bool dist_unique_id(uint8_t out[128], std::string& err) {
    Dist& d = dist_singleton_for_id();
    if (!dist_load(d, err)) return false;
    NcclId id;
    int rc = d.get_unique_id(&id);
    if (rc != 0) { err = "ncclGetUniqueId failed"; return false; }
    memcpy(out, id.internal, 128);
    return true;
}

inline bool dist_init(Dist& d, int rank, int world, const uint8_t idb[128], std::string& err) {
    if (!dist_load(d, err)) return false;
    NcclId id;
    memcpy(id.internal, idb, 128);
    int rc = d.comm_init_rank(&d.comm, world, id, rank);
    if (rc != 0) {
        err = std::string("ncclCommInitRank failed: ") + (d.get_error_string ? d.get_error_string(rc) : "?");
        return false;
    }
    d.rank = rank;
    d.world = world;
    return true;
}

inline void dist_shutdown(Dist& d) {
    if (d.comm && d.comm_destroy) d.comm_destroy(d.comm);
    d.comm = nullptr;
    d.world = 1;
    d.rank = 0;
}

}  // namespace rq
