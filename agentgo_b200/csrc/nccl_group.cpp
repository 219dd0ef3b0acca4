// nccl_group.cpp — single-process multi-GPU: N engines (one per GPU, shard i of N) joined by one
// NCCL communicator, so partial counters are combined with ONE ncclAllReduce (sum, int64) per
// evaluation over NVLink, issued on each engine's own stream right after its kernel; the batch path
// (positions sharded, agb_playouts_batch) gathers its per-position table with ONE ncclAllGather.
//
// libnccl is opened at run time (dlopen), so the library has no link-time dependency on it and
// single-GPU use never touches it.  Collectives from several devices of one process must be
// issued concurrently: call agb_playouts* / agb_genmove on each engine from its own host thread
// (agentgo_gtp --gpus N does exactly that).
#include <dlfcn.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "agentgo_b200.h"
#include "engine.h"

namespace {

typedef void* ncclComm_t;
typedef int (*InitAllFn)(ncclComm_t*, int, const int*);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*AllGatherFn)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
typedef int (*DestroyFn)(ncclComm_t);
typedef const char* (*ErrStrFn)(int);
typedef int (*GroupFn)();
constexpr int kNcclInt64 = 4, kNcclSum = 0;      // nccl.h: ncclInt64, ncclSum

struct Nccl {
    void* lib = nullptr;
    InitAllFn init_all = nullptr;
    AllReduceFn all_reduce = nullptr;
    AllGatherFn all_gather = nullptr;
    DestroyFn destroy = nullptr;
    ErrStrFn err_str = nullptr;
    std::string error;
    bool load() {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { error = std::string("cannot open libnccl: ") + dlerror(); return false; }
        init_all = (InitAllFn)dlsym(lib, "ncclCommInitAll");
        all_reduce = (AllReduceFn)dlsym(lib, "ncclAllReduce");
        all_gather = (AllGatherFn)dlsym(lib, "ncclAllGather");
        destroy = (DestroyFn)dlsym(lib, "ncclCommDestroy");
        err_str = (ErrStrFn)dlsym(lib, "ncclGetErrorString");
        if (!init_all || !all_reduce || !all_gather || !destroy) { error = "libnccl lacks the expected symbols"; return false; }
        return true;
    }
};

Nccl g_nccl;

struct Member {
    ncclComm_t comm;
};

int hook(void* ctx, void* dev_int64, int count, void* stream) {
    Member* m = (Member*)ctx;
    const int r = g_nccl.all_reduce(dev_int64, dev_int64, (size_t)count, kNcclInt64, kNcclSum, m->comm,
                                    (cudaStream_t)stream);
    if (r != 0) fprintf(stderr, "ncclAllReduce: %s\n", g_nccl.err_str ? g_nccl.err_str(r) : "error");
    return r;
}

// the batch path: every shard's [count] table into [shards][count], one ncclAllGather
int gather_hook(void* ctx, const void* send, void* recv, int count, void* stream) {
    Member* m = (Member*)ctx;
    const int r = g_nccl.all_gather(send, recv, (size_t)count, kNcclInt64, m->comm, (cudaStream_t)stream);
    if (r != 0) fprintf(stderr, "ncclAllGather: %s\n", g_nccl.err_str ? g_nccl.err_str(r) : "error");
    return r;
}

}  // namespace

struct agb_nccl_group {
    std::vector<Member> members;
    std::vector<agb_engine*> engines;
};

extern "C" int agb_attach_nccl(agb_engine** engines, int n, agb_nccl_group** out) {
    if (!engines || n < 1 || !out) return AGB_ERR_INVALID;
    *out = nullptr;
    std::vector<int> devs(n);
    for (int i = 0; i < n; i++) {
        if (!engines[i]) return AGB_ERR_INVALID;
        const agb_config& c = engines[i]->e->cfg;
        if (c.shard_count != n || c.shard_rank != i || c.device < 0)
            return engines[i]->e->fail(AGB_ERR_INVALID, "engine i of an NCCL group must be shard i of n on its own device");
        devs[i] = c.device;
    }
    if (!g_nccl.load()) return engines[0]->e->fail(AGB_ERR_STATE, g_nccl.error);
    std::vector<ncclComm_t> comms(n);
    const int r = g_nccl.init_all(comms.data(), n, devs.data());
    if (r != 0)
        return engines[0]->e->fail(AGB_ERR_CUDA, std::string("ncclCommInitAll: ") + (g_nccl.err_str ? g_nccl.err_str(r) : "error"));
    agb_nccl_group* g = new agb_nccl_group;
    g->members.resize(n);
    g->engines.assign(engines, engines + n);
    for (int i = 0; i < n; i++) {
        g->members[i].comm = comms[i];
        engines[i]->e->set_allreduce(hook, &g->members[i]);
        engines[i]->e->set_allgather(gather_hook, &g->members[i]);
    }
    *out = g;
    return AGB_OK;
}

extern "C" void agb_detach_nccl(agb_nccl_group* g) {
    if (!g) return;
    for (size_t i = 0; i < g->engines.size(); i++) {
        g->engines[i]->e->set_allreduce(nullptr, nullptr);
        g->engines[i]->e->set_allgather(nullptr, nullptr);
        g_nccl.destroy(g->members[i].comm);
    }
    delete g;
}
