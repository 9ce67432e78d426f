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
#include <string.h>

#include "common.cuh"

namespace {

typedef int ncclResult_t_;
typedef struct ncclComm *ncclComm_t_;
struct ncclUniqueId_ { char internal[NB_COMM_ID_BYTES]; };
enum { NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

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

extern "C" void nb_slice_bounds(uint64_t n, int world_size, int rank, uint64_t *begin, uint64_t *end) {
    if (world_size < 1) world_size = 1;
    const uint64_t chunk = (n + world_size - 1) / world_size;
    uint64_t b = (uint64_t) rank * chunk, e = b + chunk;
    if (b > n) b = n;
    if (e > n) e = n;
    *begin = b;
    *end = e;
}
