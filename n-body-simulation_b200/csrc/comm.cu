// Multi-GPU plumbing: one context per process / GPU, NCCL over NVLink 5 / NVSwitch.
//
// New functionality (the reference is single device: one sycl::queue, NaiveAlgorithm.cpp:56-62).  Bodies are
// replicated on every rank; each rank evaluates accelerations for its contiguous slice of targets (body order for the
// naive path, sorted Morton/DFS order for Barnes-Hut) and the slices are re-assembled with an in-place ncclAllGather.
// The integrator then advances all N bodies redundantly on every rank (192 B/body/step of HBM traffic), which avoids
// migrating velocities when slice membership changes between steps.
//
// libnccl.so.2 is resolved lazily with dlopen so that (a) single-GPU users need no NCCL at all and (b) inside a
// process that already loaded torch's bundled NCCL the same library instance is shared.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace {

typedef int ncclResult_t_;
typedef struct ncclComm *ncclComm_t_;
struct ncclUniqueId_ { char internal[NB_COMM_ID_BYTES]; };
enum { NCCL_UINT8 = 1, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

struct nccl_api {
    void *handle = nullptr;
    ncclResult_t_ (*GetUniqueId)(ncclUniqueId_ *) = nullptr;
    ncclResult_t_ (*CommInitRank)(ncclComm_t_ *, int, ncclUniqueId_, int) = nullptr;
    ncclResult_t_ (*CommDestroy)(ncclComm_t_) = nullptr;
    ncclResult_t_ (*AllGather)(const void *, void *, size_t, int, ncclComm_t_, cudaStream_t) = nullptr;
    ncclResult_t_ (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
    ncclResult_t_ (*GroupStart)() = nullptr;
    ncclResult_t_ (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t_) = nullptr;
    bool ok = false;
};

nccl_api &api() {
    static nccl_api a;
    if (a.handle) return a;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        a.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) return a;
#define NB_SYM(field, name) *(void **) (&a.field) = dlsym(a.handle, name)
    NB_SYM(GetUniqueId, "ncclGetUniqueId");
    NB_SYM(CommInitRank, "ncclCommInitRank");
    NB_SYM(CommDestroy, "ncclCommDestroy");
    NB_SYM(AllGather, "ncclAllGather");
    NB_SYM(AllReduce, "ncclAllReduce");
    NB_SYM(GroupStart, "ncclGroupStart");
    NB_SYM(GroupEnd, "ncclGroupEnd");
    NB_SYM(GetErrorString, "ncclGetErrorString");
#undef NB_SYM
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.AllReduce && a.GroupStart &&
           a.GroupEnd && a.GetErrorString;
    return a;
}

}  // namespace

extern "C" int nb_comm_get_unique_id(uint8_t id[NB_COMM_ID_BYTES]) {
    nccl_api &a = api();
    if (!a.ok) return NB_ERR_COMM;
    ncclUniqueId_ u;
    if (a.GetUniqueId(&u) != 0) return NB_ERR_COMM;
    memcpy(id, u.internal, NB_COMM_ID_BYTES);
    return NB_OK;
}

extern "C" int nb_comm_init(nb_ctx *ctx, const uint8_t id[NB_COMM_ID_BYTES], int world_size, int rank) {
    if (!ctx || world_size < 1 || rank < 0 || rank >= world_size) return nb_fail(ctx, NB_ERR_INVALID, "nb_comm_init: bad rank/world");
    if (ctx->n > 0 && world_size != ctx->world)
        return nb_fail(ctx, NB_ERR_INVALID, "nb_comm_init must be called before nb_set_bodies (buffers are sized per world)");
    nbk_comm_destroy(ctx);
    ctx->world = world_size;
    ctx->rank = rank;
    ctx->cfg.world_size = world_size;
    ctx->cfg.rank = rank;
    if (world_size == 1) return NB_OK;
    nccl_api &a = api();
    if (!a.ok) return nb_fail(ctx, NB_ERR_COMM, "libnccl.so.2 not found or incomplete");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId_ u;
    memcpy(u.internal, id, NB_COMM_ID_BYTES);
    ncclComm_t_ comm = nullptr;
    ncclResult_t_ r = a.CommInitRank(&comm, world_size, u, rank);
    if (r != 0) return nb_fail(ctx, NB_ERR_COMM, "ncclCommInitRank: %s", a.GetErrorString(r));
    ctx->nccl_comm = comm;
    return NB_OK;
}

// ---- peer slabs over CUDA IPC ------------------------------------------------------------------------------------------
// Every rank's state arrays live in one allocation (api.cu ensure_capacity).  The ranks exchange the IPC handles of
// their slabs through the NCCL communicator they already share and map each other's memory; from then on a kernel on
// any GPU can store into any rank's arrays over NVLink (bh_traverse.cu stores the walk's results that way, which
// replaces the all-gather that followed the walk).  If anything here fails the ranks agree to stay on the NCCL
// all-gather path (p2p_ok == false on every rank).
int nbk_comm_map_peers(nb_ctx *ctx) {
    nbk_comm_unmap_peers(ctx);
    if (ctx->world <= 1 || !ctx->nccl_comm || !ctx->slab) return NB_OK;
    nccl_api &a = api();
    ncclComm_t_ comm = (ncclComm_t_) ctx->nccl_comm;
    const int W = ctx->world;
    const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;   // handle + "this rank can export" flag
    unsigned char *dev = nullptr;
    std::vector<unsigned char> host((size_t) W * rec, 0);
    bool mine = W <= NB_MAX_PEERS && !getenv("NB_DISABLE_P2P");
    cudaIpcMemHandle_t h;
    if (mine && cudaIpcGetMemHandle(&h, ctx->slab) != cudaSuccess) { cudaGetLastError(); mine = false; }
    if (mine) { memcpy(&host[(size_t) ctx->rank * rec], &h, sizeof h); host[(size_t) ctx->rank * rec + sizeof h] = 1; }
    NB_CUDA(ctx, cudaMalloc((void **) &dev, (size_t) W * rec));
    NB_CUDA(ctx, cudaMemcpyAsync(dev + (size_t) ctx->rank * rec, &host[(size_t) ctx->rank * rec], rec, cudaMemcpyHostToDevice, ctx->stream));
    ncclResult_t_ r = a.AllGather(dev + (size_t) ctx->rank * rec, dev, rec, NCCL_UINT8, comm, ctx->stream);
    if (r != 0) { cudaFree(dev); return nb_fail(ctx, NB_ERR_COMM, "ncclAllGather (IPC handles): %s", a.GetErrorString(r)); }
    NB_CUDA(ctx, cudaMemcpyAsync(host.data(), dev, (size_t) W * rec, cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    bool all = true;
    for (int p = 0; p < W; ++p) all = all && host[(size_t) p * rec + sizeof h] == 1;
    double ok = 1.0;
    if (all) {
        for (int p = 0; p < W && ok == 1.0; ++p) {
            if (p == ctx->rank) { ctx->peer_slab[p] = ctx->slab; continue; }
            cudaIpcMemHandle_t hp;
            memcpy(&hp, &host[(size_t) p * rec], sizeof hp);
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, hp, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0.0; break; }
            ctx->peer_slab[p] = (unsigned char *) ptr;
        }
    } else {
        ok = 0.0;
    }
    // every rank must take the same path: all-reduce the "mapped everything" flags (sum == world)
    if (!ctx->barrier_word) NB_CHECK(nb_alloc(ctx, &ctx->barrier_word, 2));
    NB_CUDA(ctx, cudaMemcpyAsync(ctx->barrier_word, &ok, sizeof ok, cudaMemcpyHostToDevice, ctx->stream));
    r = a.AllReduce(ctx->barrier_word, ctx->barrier_word, 1, NCCL_FLOAT64, NCCL_SUM, comm, ctx->stream);
    double sum = 0;
    NB_CUDA(ctx, cudaMemcpyAsync(&sum, ctx->barrier_word, sizeof sum, cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(dev);
    if (r != 0) return nb_fail(ctx, NB_ERR_COMM, "ncclAllReduce (IPC agreement): %s", a.GetErrorString(r));
    ctx->p2p_ok = sum == (double) W;
    if (!ctx->p2p_ok) {
        for (int p = 0; p < W; ++p) {
            if (p != ctx->rank && ctx->peer_slab[p]) cudaIpcCloseMemHandle(ctx->peer_slab[p]);
            ctx->peer_slab[p] = nullptr;
        }
    }
    return NB_OK;
}

void nbk_comm_unmap_peers(nb_ctx *ctx) {
    if (!ctx->p2p_ok) return;
    cudaStreamSynchronize(ctx->stream);
    for (int p = 0; p < ctx->world && p < NB_MAX_PEERS; ++p) {
        if (p != ctx->rank && ctx->peer_slab[p]) cudaIpcCloseMemHandle(ctx->peer_slab[p]);
        ctx->peer_slab[p] = nullptr;
    }
    ctx->p2p_ok = false;
    // the owners free their slabs after this returns: wait until every rank has let go of them
    if (ctx->nccl_comm && nbk_comm_barrier(ctx) == NB_OK) cudaStreamSynchronize(ctx->stream);
}

nb_peer_table nbk_peer_table(const nb_ctx *ctx) {
    nb_peer_table t;
    memset(&t, 0, sizeof t);
    t.world = ctx->p2p_ok ? ctx->world : 1;
    t.rank = ctx->p2p_ok ? ctx->rank : 0;
    if (ctx->p2p_ok) for (int p = 0; p < ctx->world; ++p) t.base[p] = ctx->peer_slab[p];
    else t.base[0] = ctx->slab;
    return t;
}

// stream-ordered barrier: an all-reduce of one word.  Kernels enqueued after it on any rank start after the kernels
// enqueued before it on every rank have finished (and their stores, peer stores included, are visible).
int nbk_comm_barrier(nb_ctx *ctx) {
    if (ctx->world <= 1) return NB_OK;
    if (!ctx->nccl_comm) return nb_fail(ctx, NB_ERR_COMM, "world_size > 1 but nb_comm_init was not called");
    if (!ctx->barrier_word) {
        NB_CHECK(nb_alloc(ctx, &ctx->barrier_word, 2));
        NB_CUDA(ctx, cudaMemsetAsync(ctx->barrier_word, 0, 2 * sizeof(double), ctx->stream));
    }
    nccl_api &a = api();
    nb_timer_scope t(ctx, NB_T_COMM);
    ncclResult_t_ r = a.AllReduce(ctx->barrier_word + 1, ctx->barrier_word + 1, 1, NCCL_FLOAT64, NCCL_SUM, (ncclComm_t_) ctx->nccl_comm, ctx->stream);
    if (r != 0) return nb_fail(ctx, NB_ERR_COMM, "ncclAllReduce (barrier): %s", a.GetErrorString(r));
    return NB_OK;
}

void nbk_comm_destroy(nb_ctx *ctx) {
    if (ctx->nccl_comm) {
        api().CommDestroy((ncclComm_t_) ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
}

// In-place all-gather of three SoA arrays: rank r contributed [r*chunk, (r+1)*chunk) with chunk = ceil(n / world).
// The arrays are allocated with world*chunk capacity.
int nbk_comm_allgather_accel(nb_ctx *ctx, double *ax, double *ay, double *az, uint64_t n) {
    if (ctx->world <= 1) return NB_OK;
    if (!ctx->nccl_comm) return nb_fail(ctx, NB_ERR_COMM, "world_size > 1 but nb_comm_init was not called");
    nccl_api &a = api();
    nb_timer_scope t(ctx, NB_T_COMM);
    const uint64_t chunk = (n + ctx->world - 1) / ctx->world;
    ncclComm_t_ comm = (ncclComm_t_) ctx->nccl_comm;
    double *arrs[3] = {ax, ay, az};
    a.GroupStart();
    for (int k = 0; k < 3; ++k) {
        ncclResult_t_ r = a.AllGather(arrs[k] + (uint64_t) ctx->rank * chunk, arrs[k], chunk, NCCL_FLOAT64, comm, ctx->stream);
        if (r != 0) { a.GroupEnd(); return nb_fail(ctx, NB_ERR_COMM, "ncclAllGather: %s", a.GetErrorString(r)); }
    }
    ncclResult_t_ r = a.GroupEnd();
    if (r != 0) return nb_fail(ctx, NB_ERR_COMM, "ncclGroupEnd: %s", a.GetErrorString(r));
    return NB_OK;
}

int nbk_comm_allreduce_sum(nb_ctx *ctx, double *buf, size_t count) {
    if (ctx->world <= 1) return NB_OK;
    if (!ctx->nccl_comm) return nb_fail(ctx, NB_ERR_COMM, "world_size > 1 but nb_comm_init was not called");
    nccl_api &a = api();
    ncclResult_t_ r = a.AllReduce(buf, buf, count, NCCL_FLOAT64, NCCL_SUM, (ncclComm_t_) ctx->nccl_comm, ctx->stream);
    if (r != 0) return nb_fail(ctx, NB_ERR_COMM, "ncclAllReduce: %s", a.GetErrorString(r));
    return NB_OK;
}

extern "C" int nb_comm_p2p_enabled(const nb_ctx *ctx) { return ctx && ctx->p2p_ok ? 1 : 0; }

extern "C" void nb_slice_bounds(uint64_t n, int world_size, int rank, uint64_t *begin, uint64_t *end) {
    if (world_size < 1) world_size = 1;
    const uint64_t chunk = (n + world_size - 1) / world_size;
    uint64_t b = (uint64_t) rank * chunk, e = b + chunk;
    if (b > n) b = n;
    if (e > n) e = n;
    *begin = b;
    *end = e;
}
