// C ABI of libnbody_b200.so (include/nbody_b200.h): context, body upload, operator entry points, read-back,
// tree export for the parity tests.  No CPU fallback anywhere: every compute entry point launches sm_100a kernels.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

#include <cstring>

int nbk_comm_allreduce_sum(nb_ctx *ctx, double *buf, size_t count);

int nb_fail(nb_ctx *ctx, int status, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    return status;
}

extern "C" {

static void release_state(nb_ctx *ctx);

int nb_abi_version(void) { return NB_ABI_VERSION; }

const char *nb_status_string(int s) {
    switch (s) {
        case NB_OK: return "ok";
        case NB_ERR_INVALID: return "invalid argument";
        case NB_ERR_NO_DEVICE: return "no usable CUDA device (no CPU fallback)";
        case NB_ERR_CUDA: return "CUDA error";
        case NB_ERR_TREE_DEPTH: return "octree deeper than 42 levels (coincident bodies)";
        case NB_ERR_NODE_POOL: return "octree node pool overflow (raise storage_size_param)";
        case NB_ERR_COMM: return "NCCL error";
        case NB_ERR_UNSUPPORTED: return "unsupported";
        default: return "unknown";
    }
}

void nb_config_default(nb_config *c) {
    memset(c, 0, sizeof *c);
    c->struct_size = sizeof(nb_config);
    c->device = 0;
    // nBodyAlgorithm.hpp:55-61
    double G = 6.67428 * pow(10, -11);
    double meter_AU = 1.0 / (1.49597870691 * pow(10, 11));
    double second_Days = 1.0 / 86400;
    c->G = G * (pow(meter_AU, 3) / pow(second_Days, 2));
    c->epsilon2 = pow(10, -22);  // Configuration.cpp:6
    c->theta = 1.05;             // :18
    c->block_size = 64;          // :10
    c->opt_stage = 2;            // :11
    c->sort_bodies = 1;          // :21
    c->wg_size_barnes_hut = 64;  // :22
    c->storage_size_param = 16;  // main.cpp:122-127
    c->stack_size_param = 16;    // main.cpp:129-134
    c->num_wi_aabb = 1024;       // :14
    c->num_wi_octree = 640;      // :15
    c->num_wi_top_octree = 1024; // :16
    c->num_wi_com = 1024;        // :17
    c->max_level_top_octree = 7; // :19
    c->precise_rsqrt = 1;
    c->world_size = 1;
    c->rank = 0;
}

int nb_create(const nb_config *cfg, nb_ctx **out) {
    if (!cfg || !out || cfg->struct_size != sizeof(nb_config)) return NB_ERR_INVALID;
    if (cfg->opt_stage < 0 || cfg->opt_stage > 2) return NB_ERR_INVALID;  // main.cpp:154-157
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || cfg->device >= count) return NB_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return NB_ERR_NO_DEVICE;
    if (prop.major != 10) return NB_ERR_NO_DEVICE;  // the kernels are sm_100a cubins: they load on compute capability 10.x only
    nb_ctx *ctx = new nb_ctx();
    ctx->cfg = *cfg;
    ctx->device = cfg->device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->coop_launch = prop.cooperativeLaunch != 0;
    ctx->device_name = prop.name;
    ctx->world = cfg->world_size > 0 ? cfg->world_size : 1;
    ctx->rank = cfg->rank;
    if (cudaSetDevice(ctx->device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return NB_ERR_CUDA;
    }
    for (int i = 0; i < 2 * NB_T_COUNT; ++i) cudaEventCreate(&ctx->ev[i]);
    for (int i = 0; i < 8; ++i) cudaEventCreate(&ctx->user_ev[i]);
    *out = ctx;
    return NB_OK;
}

void nb_destroy(nb_ctx *ctx) {
    if (ctx && ctx->step_graph) { cudaGraphExecDestroy(ctx->step_graph); ctx->step_graph = nullptr; }
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    release_state(ctx);          // unmaps the peers' slabs first (collective, while the communicator is alive)
    nbk_comm_destroy(ctx);
    nbk_bh_release(ctx);
    nb_free(&ctx->src); nb_free(&ctx->naive_partial); nb_free(&ctx->barrier_word);
    for (int i = 0; i < 2 * NB_T_COUNT; ++i) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 8; ++i) cudaEventDestroy(ctx->user_ev[i]);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *nb_last_error(const nb_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

static int check_bh_flags(nb_ctx *ctx);

int nb_synchronize(nb_ctx *ctx) {
    if (!ctx) return NB_ERR_INVALID;
    NB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return check_bh_flags(ctx);
}

int nb_device_name(nb_ctx *ctx, char *buf, size_t buflen) {
    if (!ctx || !buf || !buflen) return NB_ERR_INVALID;
    snprintf(buf, buflen, "%s", ctx->device_name.c_str());
    return NB_OK;
}

int nb_set_theta(nb_ctx *ctx, double theta) { if (!ctx) return NB_ERR_INVALID; ctx->cfg.theta = theta; return NB_OK; }
int nb_set_block_size(nb_ctx *ctx, int bs) { if (!ctx || bs <= 0) return NB_ERR_INVALID; ctx->cfg.block_size = bs; return NB_OK; }
int nb_set_sort_bodies(nb_ctx *ctx, int s) { if (!ctx) return NB_ERR_INVALID; ctx->cfg.sort_bodies = s; return NB_OK; }
int nb_set_precise_rsqrt(nb_ctx *ctx, int p) { if (!ctx) return NB_ERR_INVALID; ctx->cfg.precise_rsqrt = p; return NB_OK; }

uint64_t nb_num_bodies(const nb_ctx *ctx) { return ctx ? ctx->n : 0; }
uint64_t nb_launch_count(const nb_ctx *ctx) { return ctx ? ctx->launches : 0; }

static void release_state(nb_ctx *ctx) {
    nbk_comm_unmap_peers(ctx);   // nobody may still address this slab when it is freed
    if (ctx->slab) cudaFree(ctx->slab);
    ctx->slab = nullptr;
    ctx->slab_bytes = 0;
    ctx->m = ctx->x = ctx->y = ctx->z = ctx->vx = ctx->vy = ctx->vz = ctx->ax = ctx->ay = ctx->az = ctx->anorm = nullptr;
    for (int k = 0; k < 10; ++k) ctx->alt[k] = nullptr;
    nb_free(&ctx->id); nb_free(&ctx->id_alt); nb_free(&ctx->e_partial); nb_free(&ctx->tile_start); nb_free(&ctx->dyn_bounds);
    ctx->tile_cost = nullptr;
    ctx->cap = 0;
}

static int ensure_capacity(nb_ctx *ctx, uint64_t n) {
    // arrays that take part in the in-place all-gather need world * ceil(n/world) elements
    const uint64_t chunk = (n + ctx->world - 1) / ctx->world;
    const uint64_t need = ((chunk * ctx->world + 32 + 31) / 32) * 32;   // 256-byte multiples: every array stays aligned
    if (need <= ctx->cap) return NB_OK;
    NB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    release_state(ctx);
    const size_t tile_words = ((size_t) need / 32 + 64 + 63) / 64 * 64;
    ctx->slab_bytes = (size_t) 21 * need * sizeof(double) + tile_words * sizeof(uint32_t);
    NB_CUDA(ctx, cudaMalloc((void **) &ctx->slab, ctx->slab_bytes));
    ctx->tile_cost = reinterpret_cast<uint32_t *>(ctx->slab + (size_t) 21 * need * sizeof(double));
    NB_CHECK(nb_alloc(ctx, &ctx->tile_start, tile_words));
    NB_CHECK(nb_alloc(ctx, &ctx->dyn_bounds, (size_t) NB_MAX_PEERS + 2 + need / (32 * 256) + 8));   // bounds + group sums
    ctx->bounds_valid = false;
    double *base = reinterpret_cast<double *>(ctx->slab);
    double **cur[10] = {&ctx->m, &ctx->x, &ctx->y, &ctx->z, &ctx->vx, &ctx->vy, &ctx->vz, &ctx->ax, &ctx->ay, &ctx->az};
    for (int k = 0; k < 10; ++k) { *cur[k] = base + (size_t) k * need; ctx->alt[k] = base + (size_t) (10 + k) * need; }
    ctx->anorm = base + (size_t) 20 * need;
    NB_CHECK(nb_alloc(ctx, &ctx->id, need));
    NB_CHECK(nb_alloc(ctx, &ctx->id_alt, need));
    NB_CHECK(nb_alloc(ctx, &ctx->e_partial, 2 * need + 8 + 2048));
    ctx->cap = need;
    // multi-GPU: let every rank address every rank's slab (collective: all ranks size their state in the same call)
    if (ctx->world > 1 && ctx->nccl_comm) NB_CHECK(nbk_comm_map_peers(ctx));
    return NB_OK;
}

static int h2d(nb_ctx *ctx, double *dst, const double *src, uint64_t n) {
    if (!src) return NB_OK;
    NB_CUDA(ctx, cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return NB_OK;
}
static int d2h(nb_ctx *ctx, double *dst, const double *src, uint64_t n) {
    if (!dst) return NB_OK;
    NB_CUDA(ctx, cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return NB_OK;
}

int nb_set_bodies(nb_ctx *ctx, uint64_t n, const double *mass, const double *x, const double *y, const double *z,
                  const double *vx, const double *vy, const double *vz) {
    if (!ctx) return NB_ERR_INVALID;
    if (n == 0 || !mass || !x || !y || !z) return nb_fail(ctx, NB_ERR_INVALID, "nb_set_bodies: need n > 0 and mass/x/y/z");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CHECK(ensure_capacity(ctx, n));
    ctx->n = n;
    ctx->bh.built = false;
    ctx->identity_order = true;  // storage order == body-id order until the next Barnes-Hut build
    ctx->a_fresh = false;        // zeros below: in order, but not the forces of these positions
    ctx->a_order_ok = true;
    ctx->bounds_valid = false;   // cost-weighted slices restart from equal counts for a new body set
    if (ctx->bh.dev_flags) NB_CUDA(ctx, cudaMemsetAsync(ctx->bh.dev_flags, 0, 8 * sizeof(uint32_t), ctx->stream));
    NB_CHECK(h2d(ctx, ctx->m, mass, n));
    NB_CHECK(h2d(ctx, ctx->x, x, n));
    NB_CHECK(h2d(ctx, ctx->y, y, n));
    NB_CHECK(h2d(ctx, ctx->z, z, n));
    if (vx && vy && vz) {
        NB_CHECK(h2d(ctx, ctx->vx, vx, n));
        NB_CHECK(h2d(ctx, ctx->vy, vy, n));
        NB_CHECK(h2d(ctx, ctx->vz, vz, n));
    } else {
        NB_CUDA(ctx, cudaMemsetAsync(ctx->vx, 0, n * sizeof(double), ctx->stream));
        NB_CUDA(ctx, cudaMemsetAsync(ctx->vy, 0, n * sizeof(double), ctx->stream));
        NB_CUDA(ctx, cudaMemsetAsync(ctx->vz, 0, n * sizeof(double), ctx->stream));
    }
    NB_CUDA(ctx, cudaMemsetAsync(ctx->ax, 0, n * sizeof(double), ctx->stream));
    NB_CUDA(ctx, cudaMemsetAsync(ctx->ay, 0, n * sizeof(double), ctx->stream));
    NB_CUDA(ctx, cudaMemsetAsync(ctx->az, 0, n * sizeof(double), ctx->stream));
    // the host arrays are caller-owned and only valid during the call
    NB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NB_OK;
}

int nb_set_positions(nb_ctx *ctx, const double *x, const double *y, const double *z) {
    if (!ctx || !ctx->n) return nb_fail(ctx, NB_ERR_INVALID, "nb_set_positions: no bodies");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!x || !y || !z) return nb_fail(ctx, NB_ERR_INVALID, "nb_set_positions: null array");
    if (ctx->identity_order) {
        NB_CHECK(h2d(ctx, ctx->x, x, ctx->n));
        NB_CHECK(h2d(ctx, ctx->y, y, ctx->n));
        NB_CHECK(h2d(ctx, ctx->z, z, ctx->n));
    } else {  // host arrays are in body-id order, the device state in storage order
        NB_CHECK(h2d(ctx, ctx->alt[1], x, ctx->n));
        NB_CHECK(h2d(ctx, ctx->alt[2], y, ctx->n));
        NB_CHECK(h2d(ctx, ctx->alt[3], z, ctx->n));
        const double *src[3] = {ctx->alt[1], ctx->alt[2], ctx->alt[3]};
        double *dst[3] = {ctx->x, ctx->y, ctx->z};
        NB_CHECK(nbk_permute_in(ctx, 3, src, dst));
    }
    ctx->bh.built = false;
    ctx->a_fresh = false;
    NB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NB_OK;
}

// ---- operators -------------------------------------------------------------------------------------------------------
int nb_naive_accel(nb_ctx *ctx) {
    if (!ctx || !ctx->n) return nb_fail(ctx, NB_ERR_INVALID, "nb_naive_accel: no bodies");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t b = 0, e = ctx->n;
    nb_slice_bounds(ctx->n, ctx->world, ctx->rank, &b, &e);
    {
        nb_timer_scope t(ctx, NB_T_ACCEL);
        NB_CHECK(nbk_naive_accel(ctx, b, e));
    }
    NB_CHECK(nbk_comm_allgather_accel(ctx, ctx->ax, ctx->ay, ctx->az, ctx->n));
    ctx->a_fresh = ctx->a_order_ok = true;
    return NB_OK;
}

int nb_bh_build(nb_ctx *ctx) {
    if (!ctx || !ctx->n) return nb_fail(ctx, NB_ERR_INVALID, "nb_bh_build: no bodies");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    return nbk_bh_build(ctx);
}

// The walk of this rank's slice with one of the epilogues of bh_traverse.cu.  Several GPUs with mapped peer slabs: the
// kernel stores its results into every rank's arrays, bracketed by two barriers -- before it, so that no rank still
// reads (build, integrator) what a faster rank is about to overwrite; after it, so that every rank's stores have
// landed before anybody goes on.  Otherwise (no IPC): accelerations only, then the NCCL all-gather.
// "Acceleration Kernel Time" is the walk alone; barriers / all-gather are "Allgather".
static int bh_walk(nb_ctx *ctx, int epilogue, double dt) {
    uint64_t b = 0, e = ctx->n;
    nb_slice_bounds(ctx->n, ctx->world, ctx->rank, &b, &e);
    const bool peers = ctx->world > 1 && ctx->p2p_ok && !ctx->bh.stats_enabled;
    if (ctx->world > 1 && !peers && epilogue != 0)
        return nb_fail(ctx, NB_ERR_INVALID, "fused walk on several GPUs needs mapped peer slabs");
    // Cost-weighted slices: with peer stores the slice a rank walks need not be the equal-count one of
    // nb_slice_bounds.  Every walk records what each 32-body tile cost (clock ticks), in every rank's copy; after the
    // barrier all ranks cut the sorted order into pieces of equal cost -- the same pieces, from the same numbers -- for
    // the next walk.  The accelerations do not depend on the cut.  cfg.reserved[5] = 1 keeps the equal-count slices.
    const bool dynamic = peers && ctx->cfg.reserved[5] != 1 && (ctx->cfg.reserved[3] == 50 || e - b >= (1ull << 19));
    if (dynamic && !ctx->bounds_valid) {
        // first walk of this body set: no costs yet -> equal tile counts.  BEFORE the barrier: once a rank has passed it, its
        // walk stores tile costs into this rank's copy, which a later memset would wipe (and the ranks' cuts would differ)
        NB_CUDA(ctx, cudaMemsetAsync(ctx->tile_cost, 0, ((ctx->n + 31) / 32) * sizeof(uint32_t), ctx->stream));
        NB_CHECK(nbk_bh_rebalance(ctx));
    }
    if (peers) NB_CHECK(nbk_comm_barrier(ctx));
    {
        nb_timer_scope t(ctx, NB_T_ACCEL);
        NB_CHECK(nbk_bh_accel_fused(ctx, b, e, epilogue, dt, peers, dynamic));
    }
    if (peers) NB_CHECK(nbk_comm_barrier(ctx));
    else NB_CHECK(nbk_comm_allgather_accel(ctx, ctx->ax, ctx->ay, ctx->az, ctx->n));
    if (dynamic) NB_CHECK(nbk_bh_rebalance(ctx));
    return NB_OK;
}

int nb_bh_accel(nb_ctx *ctx) {
    if (!ctx || !ctx->n) return nb_fail(ctx, NB_ERR_INVALID, "nb_bh_accel: no bodies");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CHECK(bh_walk(ctx, 0, 0.0));
    ctx->a_fresh = ctx->a_order_ok = true;
    return NB_OK;
}

int nb_bh_accel_range(nb_ctx *ctx, uint64_t slot_begin, uint64_t slot_end) {
    if (!ctx || !ctx->n) return nb_fail(ctx, NB_ERR_INVALID, "nb_bh_accel_range: no bodies");
    if (slot_begin > slot_end || slot_end > ctx->n) return nb_fail(ctx, NB_ERR_INVALID, "nb_bh_accel_range: bad slot range");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    nb_timer_scope t(ctx, NB_T_ACCEL);
    NB_CHECK(nbk_bh_accel(ctx, slot_begin, slot_end));
    ctx->a_fresh = ctx->a_order_ok = true;   // (for the slots evaluated so far: the caller assembles the slices)
    return NB_OK;
}

static int need_ordered_accel(nb_ctx *ctx, const char *who) {
    if (ctx->a_order_ok) return NB_OK;
    return nb_fail(ctx, NB_ERR_INVALID, "%s: the accelerations on the device belong to positions from before the last tree "
                   "build and were not carried through it; evaluate the forces first", who);
}

int nb_leapfrog_part1(nb_ctx *ctx, double dt) {
    if (!ctx) return NB_ERR_INVALID;
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CHECK(need_ordered_accel(ctx, "nb_leapfrog_part1"));
    nb_timer_scope t(ctx, NB_T_LEAPFROG1);
    ctx->bh.built = false;
    ctx->a_fresh = false;   // the positions move on
    return nbk_leapfrog_part1(ctx, dt);
}
int nb_leapfrog_part2(nb_ctx *ctx, double dt) {
    if (!ctx) return NB_ERR_INVALID;
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CHECK(need_ordered_accel(ctx, "nb_leapfrog_part2"));
    nb_timer_scope t(ctx, NB_T_LEAPFROG2);
    return nbk_leapfrog_part2(ctx, dt);
}
int nb_leapfrog_part2_part1(nb_ctx *ctx, double dt) {
    if (!ctx) return NB_ERR_INVALID;
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CHECK(need_ordered_accel(ctx, "nb_leapfrog_part2_part1"));
    nb_timer_scope t(ctx, NB_T_LEAPFROG1);
    ctx->bh.built = false;
    ctx->a_fresh = false;
    return nbk_leapfrog_part2_part1(ctx, dt);
}

// ---- nb_advance: batches of steps, inner steps replayed from a CUDA graph --------------------------------------------
// Two forms of a batch of `nsteps` leapfrog steps (the result is the same bit for bit):
//   unfused (naive; Barnes-Hut with the instrumented walk, or on several GPUs without IPC):
//       part 1; forces; { part 2 + part 1 in one pass; forces } x (nsteps - 1); part 2
//   fused (Barnes-Hut): the leapfrog half-steps ride in the epilogue of the walk (bh_traverse.cu), the state is updated
//   in place body by body and, on several GPUs, stored into every rank's arrays by the walk itself:
//       part 1; { build; walk + part 2 + part 1 } x (nsteps - 1); build; walk + part 2
// On one GPU the repeated unit is captured once as a CUDA graph of TWO units (a build swaps the state arrays with
// their ping-pong partners, so the pointer configuration has period two) and replayed.
static void drop_step_graph(nb_ctx *ctx) {
    if (ctx->step_graph) cudaGraphExecDestroy(ctx->step_graph);
    ctx->step_graph = nullptr;
    ctx->graph_algorithm = -1;
}

static void state_pointers(const nb_ctx *ctx, const void *p[8]) {
    p[0] = ctx->m; p[1] = ctx->x; p[2] = ctx->vx; p[3] = ctx->ax; p[4] = ctx->id; p[5] = ctx->bh.key_hi;
    p[6] = ctx->bh.perm; p[7] = ctx->src;
}

static bool advance_fused(const nb_ctx *ctx, int algorithm) {
    return algorithm == 1 && !ctx->bh.stats_enabled && (ctx->world == 1 || ctx->p2p_ok) && ctx->cfg.reserved[1] != 1;
}

// the repeated unit of a batch (see above)
static int inner_step(nb_ctx *ctx, int algorithm, double dt) {
    if (advance_fused(ctx, algorithm)) {
        NB_CHECK(nb_bh_build(ctx));
        NB_CHECK(bh_walk(ctx, 2, dt));
        ctx->bh.built = false;   // the positions moved on
        ctx->a_fresh = false;
        return NB_OK;
    }
    NB_CHECK(nb_leapfrog_part2_part1(ctx, dt));
    if (algorithm == 0) return nb_naive_accel(ctx);
    NB_CHECK(nb_bh_build(ctx));
    return nb_bh_accel(ctx);
}

static bool graph_matches(const nb_ctx *ctx, int algorithm, double dt) {
    if (!ctx->step_graph || ctx->graph_algorithm != algorithm || ctx->graph_dt != dt || ctx->graph_n != ctx->n) return false;
    if (memcmp(&ctx->graph_cfg, &ctx->cfg, sizeof(nb_config)) != 0) return false;
    // the graph froze the number of sort passes of its builds; the eager step just before this check chose from newer
    // run statistics
    if (algorithm == 1 && ctx->graph_sort_passes != ctx->bh.sort_passes) return false;
    const void *p[8];
    state_pointers(ctx, p);
    return memcmp(p, ctx->graph_ptrs, sizeof p) == 0;
}

// host-side roles of the device buffers that a step swaps (nothing a captured kernel touches: a failed capture must
// put them back, because none of the captured kernels has run)
struct pointer_roles {
    double *state[10], *alt[10];
    uint32_t *id, *id_alt, *perm, *perm_alt;
    uint64_t *key_hi, *key_hi_alt;
    bool identity_order, built, coop_launch;
};
static void save_roles(const nb_ctx *ctx, pointer_roles &r) {
    double *const cur[10] = {ctx->m, ctx->x, ctx->y, ctx->z, ctx->vx, ctx->vy, ctx->vz, ctx->ax, ctx->ay, ctx->az};
    for (int k = 0; k < 10; ++k) { r.state[k] = cur[k]; r.alt[k] = ctx->alt[k]; }
    r.id = ctx->id; r.id_alt = ctx->id_alt; r.perm = ctx->bh.perm; r.perm_alt = ctx->bh.perm_alt;
    r.key_hi = ctx->bh.key_hi; r.key_hi_alt = ctx->bh.key_hi_alt;
    r.identity_order = ctx->identity_order; r.built = ctx->bh.built; r.coop_launch = ctx->coop_launch;
}
static void restore_roles(nb_ctx *ctx, const pointer_roles &r) {
    double **cur[10] = {&ctx->m, &ctx->x, &ctx->y, &ctx->z, &ctx->vx, &ctx->vy, &ctx->vz, &ctx->ax, &ctx->ay, &ctx->az};
    for (int k = 0; k < 10; ++k) { *cur[k] = r.state[k]; ctx->alt[k] = r.alt[k]; }
    ctx->id = r.id; ctx->id_alt = r.id_alt; ctx->bh.perm = r.perm; ctx->bh.perm_alt = r.perm_alt;
    ctx->bh.key_hi = r.key_hi; ctx->bh.key_hi_alt = r.key_hi_alt;
    ctx->identity_order = r.identity_order; ctx->bh.built = r.built; ctx->coop_launch = r.coop_launch;
}

// captures two inner steps; on success the host-side state (pointer roles) is back where it started.  Any failure
// (a launch refused under capture, pointer roles that are not periodic over two steps, instantiation) restores the
// roles, marks the graph unusable and returns NB_OK: the caller continues on the eager path.
static int capture_step_graph(nb_ctx *ctx, int algorithm, double dt) {
    drop_step_graph(ctx);
    const void *before[8], *after[8];
    state_pointers(ctx, before);
    pointer_roles saved;
    save_roles(ctx, saved);
    const bool timers = ctx->timers_enabled;
    const uint64_t launches0 = ctx->launches;
    ctx->timers_enabled = false;   // no event records inside the graph
    ctx->capturing = true;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
    int rc = NB_OK;
    if (e == cudaSuccess) {
        rc = inner_step(ctx, algorithm, dt);
        if (rc == NB_OK) rc = inner_step(ctx, algorithm, dt);
        e = cudaStreamEndCapture(ctx->stream, &graph);
    }
    ctx->timers_enabled = timers;
    ctx->capturing = false;
    ctx->graph_launches = ctx->launches - launches0;
    ctx->launches = launches0;     // nothing ran yet
    state_pointers(ctx, after);
    if (e != cudaSuccess || rc != NB_OK || !graph || memcmp(before, after, sizeof before) != 0) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        restore_roles(ctx, saved);
        ctx->graph_unusable = true;
        return NB_OK;
    }
    e = cudaGraphInstantiate(&ctx->step_graph, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        cudaGetLastError();
        ctx->step_graph = nullptr;
        ctx->graph_unusable = true;
        return NB_OK;
    }
    ctx->graph_algorithm = algorithm;
    ctx->graph_dt = dt;
    ctx->graph_n = ctx->n;
    ctx->graph_cfg = ctx->cfg;
    ctx->graph_sort_passes = ctx->bh.sort_passes;
    memcpy(ctx->graph_ptrs, before, sizeof before);
    return NB_OK;
}

int nb_advance(nb_ctx *ctx, int algorithm, double dt, uint32_t nsteps, double *ms) {
    if (!ctx || !ctx->n) return nb_fail(ctx, NB_ERR_INVALID, "nb_advance: no bodies");
    if (algorithm != 0 && algorithm != 1) return nb_fail(ctx, NB_ERR_INVALID, "nb_advance: algorithm must be 0 (naive) or 1 (Barnes-Hut)");
    if (ms) for (int i = 0; i < NB_T_COUNT; ++i) ms[i] = 0;
    if (nsteps == 0) return NB_OK;
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool fused = advance_fused(ctx, algorithm);
    // the first unit runs eagerly and timed: its phase times are the sample reported for the batch
    NB_CHECK(nb_leapfrog_part1(ctx, dt));
    uint32_t rest;   // repeated units still to run
    if (fused) {
        rest = nsteps - 1;
        if (rest > 0) { NB_CHECK(inner_step(ctx, algorithm, dt)); --rest; }
    } else {
        if (algorithm == 0) {
            NB_CHECK(nb_naive_accel(ctx));
        } else {
            NB_CHECK(nb_bh_build(ctx));
            NB_CHECK(nb_bh_accel(ctx));
        }
        rest = nsteps - 1;
    }
    if (ms && ctx->timers_enabled && (!fused || nsteps > 1)) NB_CHECK(nb_get_timers(ctx, ms));
    // pairs of units from the graph: one GPU, no instrumentation, and enough steps to repay the capture
    const bool graph_ok = ctx->world == 1 && !ctx->bh.stats_enabled && !ctx->graph_unusable;
    if (graph_ok && rest >= 4) {
        if (!graph_matches(ctx, algorithm, dt)) NB_CHECK(capture_step_graph(ctx, algorithm, dt));
        while (ctx->step_graph && rest >= 2) {
            NB_CUDA(ctx, cudaGraphLaunch(ctx->step_graph, ctx->stream));
            ctx->launches += ctx->graph_launches;
            rest -= 2;
        }
    }
    const bool timers = ctx->timers_enabled;
    ctx->timers_enabled = false;   // keep the sample of the first unit
    int rc = NB_OK;
    for (; rest > 0 && rc == NB_OK; --rest) rc = inner_step(ctx, algorithm, dt);
    ctx->timers_enabled = timers && (fused && nsteps == 1);
    if (rc == NB_OK && fused) {   // the batch's last force evaluation carries the closing half-kick
        rc = nb_bh_build(ctx);
        if (rc == NB_OK) rc = bh_walk(ctx, 1, dt);
        if (rc == NB_OK) ctx->a_fresh = ctx->a_order_ok = true;
    }
    ctx->timers_enabled = timers;
    NB_CHECK(rc);
    if (fused) {
        if (ms && timers && nsteps == 1) NB_CHECK(nb_get_timers(ctx, ms));
        return NB_OK;
    }
    NB_CHECK(nb_leapfrog_part2(ctx, dt));
    if (ms && ctx->timers_enabled) {   // Leapfrog Part 2 of the batch's last step
        double tail[NB_T_COUNT];
        NB_CHECK(nb_get_timers(ctx, tail));
        ms[NB_T_LEAPFROG2] = tail[NB_T_LEAPFROG2];
    }
    return NB_OK;
}

int nb_energy(nb_ctx *ctx, double out[4]) {
    if (!ctx || !ctx->n || !out) return nb_fail(ctx, NB_ERR_INVALID, "nb_energy: no bodies");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t n = ctx->n;
    {
        nb_timer_scope t(ctx, NB_T_ENERGY);
        // triangular work: rank r takes targets [n*sqrt(r/P), n*sqrt((r+1)/P)) so pair counts balance
        uint64_t jb = 0, je = n;
        if (ctx->world > 1) {
            jb = (uint64_t) floor((double) n * sqrt((double) ctx->rank / ctx->world));
            je = ctx->rank + 1 == ctx->world ? n : (uint64_t) floor((double) n * sqrt((double) (ctx->rank + 1) / ctx->world));
        }
        NB_CHECK(nbk_energy(ctx, jb, je));
        NB_CHECK(nbk_comm_allreduce_sum(ctx, ctx->e_partial + 2 * n, 2));
    }
    double h[2];
    NB_CUDA(ctx, cudaMemcpyAsync(h, ctx->e_partial + 2 * n, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // nBodyAlgorithm.cpp:77-85
    double E_kin_result = h[0], E_pot_result = h[1];
    E_pot_result *= -1;
    out[0] = E_kin_result;
    out[1] = E_pot_result;
    out[2] = E_kin_result + E_pot_result;
    out[3] = (2.0 * E_kin_result) / fabs(E_pot_result);
    return NB_OK;
}

// ---- read-back ----------------------------------------------------------------------------------------------------------
// Host arrays are always in body-id order.  After a Barnes-Hut build the device state is in storage (sorted) order:
// the requested arrays are un-permuted on the device into the idle ping-pong buffers and copied from there.
static int read_back(nb_ctx *ctx, int count, const double *const *dev, double *const *host) {
    if (!ctx || !ctx->n) return nb_fail(ctx, NB_ERR_INVALID, "no bodies");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->identity_order) {
        for (int k = 0; k < count; ++k) NB_CHECK(d2h(ctx, host[k], dev[k], ctx->n));
    } else {
        const double *src[10];
        double *dst[10];
        int m = 0;
        for (int k = 0; k < count; ++k)
            if (host[k]) { src[m] = dev[k]; dst[m] = ctx->alt[m]; ++m; }
        if (m) NB_CHECK(nbk_unpermute(ctx, m, src, dst));
        m = 0;
        for (int k = 0; k < count; ++k)
            if (host[k]) { NB_CHECK(d2h(ctx, host[k], ctx->alt[m], ctx->n)); ++m; }
    }
    return nb_synchronize(ctx);
}
int nb_get_positions(nb_ctx *ctx, double *x, double *y, double *z) {
    if (!ctx) return NB_ERR_INVALID;
    const double *dev[3] = {ctx->x, ctx->y, ctx->z};
    double *host[3] = {x, y, z};
    return read_back(ctx, 3, dev, host);
}
int nb_get_velocities(nb_ctx *ctx, double *vx, double *vy, double *vz) {
    if (!ctx) return NB_ERR_INVALID;
    const double *dev[3] = {ctx->vx, ctx->vy, ctx->vz};
    double *host[3] = {vx, vy, vz};
    return read_back(ctx, 3, dev, host);
}
int nb_get_accelerations(nb_ctx *ctx, double *ax, double *ay, double *az) {
    if (!ctx) return NB_ERR_INVALID;
    NB_CHECK(need_ordered_accel(ctx, "nb_get_accelerations"));
    const double *dev[3] = {ctx->ax, ctx->ay, ctx->az};
    double *host[3] = {ax, ay, az};
    return read_back(ctx, 3, dev, host);
}
int nb_get_acceleration_norms(nb_ctx *ctx, double *anorm) {
    if (!ctx || !ctx->n || !anorm) return nb_fail(ctx, NB_ERR_INVALID, "no bodies");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CHECK(need_ordered_accel(ctx, "nb_get_acceleration_norms"));
    NB_CHECK(nbk_accel_norm(ctx));
    const double *dev[1] = {ctx->anorm};
    double *host[1] = {anorm};
    return read_back(ctx, 1, dev, host);
}

int nb_device_pointers(nb_ctx *ctx, void *p[10]) {
    if (!ctx || !p) return NB_ERR_INVALID;
    p[0] = ctx->x; p[1] = ctx->y; p[2] = ctx->z; p[3] = ctx->vx; p[4] = ctx->vy; p[5] = ctx->vz;
    p[6] = ctx->ax; p[7] = ctx->ay; p[8] = ctx->az; p[9] = ctx->m;
    return NB_OK;
}

// ---- one-call operator forms with host buffers --------------------------------------------------------------------------
static int upload_for_op(nb_ctx *ctx, uint64_t n, const double *mass, const double *x, const double *y, const double *z) {
    if (n == 0 || !mass || !x || !y || !z) return nb_fail(ctx, NB_ERR_INVALID, "operator: need n > 0 and mass/x/y/z");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CHECK(ensure_capacity(ctx, n));
    if (n != ctx->n) ctx->bounds_valid = false;   // the operator form is called again and again with the same N
    ctx->n = n;
    ctx->bh.built = false;
    ctx->identity_order = true;
    ctx->a_fresh = false;
    ctx->a_order_ok = true;
    NB_CHECK(h2d(ctx, ctx->m, mass, n));
    NB_CHECK(h2d(ctx, ctx->x, x, n));
    NB_CHECK(h2d(ctx, ctx->y, y, n));
    NB_CHECK(h2d(ctx, ctx->z, z, n));
    return NB_OK;
}

int nb_op_naive_accelerations(nb_ctx *ctx, uint64_t n, const double *mass, const double *x, const double *y,
                              const double *z, double *ax, double *ay, double *az) {
    if (!ctx) return NB_ERR_INVALID;
    NB_CHECK(upload_for_op(ctx, n, mass, x, y, z));
    NB_CHECK(nb_naive_accel(ctx));
    return nb_get_accelerations(ctx, ax, ay, az);
}

int nb_op_barnes_hut_accelerations(nb_ctx *ctx, uint64_t n, const double *mass, const double *x, const double *y,
                                   const double *z, double *ax, double *ay, double *az) {
    if (!ctx) return NB_ERR_INVALID;
    NB_CHECK(upload_for_op(ctx, n, mass, x, y, z));
    NB_CHECK(nb_bh_build(ctx));
    NB_CHECK(nb_bh_accel(ctx));
    return nb_get_accelerations(ctx, ax, ay, az);
}

// ---- timers ------------------------------------------------------------------------------------------------------------------
int nb_enable_timers(nb_ctx *ctx, int enable) {
    if (!ctx) return NB_ERR_INVALID;
    ctx->timers_enabled = enable != 0;
    return NB_OK;
}
int nb_get_timers(nb_ctx *ctx, double ms[NB_T_COUNT]) {
    if (!ctx || !ms) return NB_ERR_INVALID;
    NB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < NB_T_COUNT; ++i) {
        ms[i] = 0;
        if (ctx->ev_valid[i]) {
            float f = 0;
            if (cudaEventElapsedTime(&f, ctx->ev[2 * i], ctx->ev[2 * i + 1]) == cudaSuccess) ms[i] = f;
        }
    }
    return NB_OK;
}
const char *nb_timer_name(int t) {
    static const char *names[NB_T_COUNT] = {"Acceleration Kernel Time", "Leapfrog Part 1", "Leapfrog Part 2",
                                            "AABB creation", "Sort bodies for subtrees", "Build subtrees",
                                            "Compute center of mass", "Octree creation", "Energy", "Allgather"};
    return (t >= 0 && t < NB_T_COUNT) ? names[t] : "";
}

int nb_event_record(nb_ctx *ctx, int slot) {
    if (!ctx || slot < 0 || slot >= 8) return NB_ERR_INVALID;
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CUDA(ctx, cudaEventRecord(ctx->user_ev[slot], ctx->stream));
    return NB_OK;
}
int nb_event_elapsed_ms(nb_ctx *ctx, int a, int b, double *ms) {
    if (!ctx || !ms || a < 0 || a >= 8 || b < 0 || b >= 8) return NB_ERR_INVALID;
    NB_CUDA(ctx, cudaEventSynchronize(ctx->user_ev[b]));
    float f = 0;
    NB_CUDA(ctx, cudaEventElapsedTime(&f, ctx->user_ev[a], ctx->user_ev[b]));
    *ms = f;
    return NB_OK;
}

int nb_measure_fp64_peak(nb_ctx *ctx, double *tflops) {
    if (!ctx || !tflops) return NB_ERR_INVALID;
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    return nbk_fp64_peak(ctx, tflops);
}

// ---- Barnes-Hut inspection -------------------------------------------------------------------------------------------------
static int check_bh_flags(nb_ctx *ctx) {
    nb_bh_state &b = ctx->bh;
    if (!b.dev_flags) return NB_OK;
    uint32_t f[5];
    NB_CUDA(ctx, cudaMemcpy(f, b.dev_flags, sizeof f, cudaMemcpyDeviceToHost));
    if (b.built) {
        b.num_internal = f[1];
        b.num_nodes = ctx->n + f[1];
        b.max_depth = f[2];
    }
    // f[0]: the latest build; f[4]: every build since the last report (a build that failed in the middle of a batch of
    // steps left zero accelerations for that step, and the builds after it cleared f[0])
    const uint32_t bits = f[0] | f[4];
    if (!bits) return NB_OK;
    if (f[4]) NB_CUDA(ctx, cudaMemset(b.dev_flags + 4, 0, sizeof(uint32_t)));   // reported once
    if (f[0]) b.built = false;
    const char *when = f[0] ? "" : " (in an earlier step of this batch; that step used zero accelerations)";
    if (bits & 2u)
        return nb_fail(ctx, NB_ERR_NODE_POOL, "octree needs %llu internal nodes, pool holds %llu (storage_size_param=%d)%s",
                       (unsigned long long) f[1], (unsigned long long) (b.cap_nodes - ctx->n), ctx->cfg.storage_size_param, when);
    return nb_fail(ctx, NB_ERR_TREE_DEPTH, "octree deeper than %d levels: coincident bodies are not supported (reference: unbounded splitting)%s",
                   NB_MAX_TREE_DEPTH, when);
}

int nb_bh_aabb(nb_ctx *ctx, double out[7]) {
    if (!ctx || !ctx->n || !out) return nb_fail(ctx, NB_ERR_INVALID, "nb_bh_aabb: no bodies");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CHECK(nbk_bh_reserve(ctx));
    NB_CHECK(nbk_bh_aabb(ctx));
    NB_CUDA(ctx, cudaMemcpyAsync(out, ctx->bh.aabb_dev, 7 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NB_OK;
}

int nb_bh_tree_info(nb_ctx *ctx, nb_tree_info *info) {
    if (!ctx || !info) return NB_ERR_INVALID;
    if (!ctx->bh.built) return nb_fail(ctx, NB_ERR_INVALID, "nb_bh_tree_info: no tree");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CHECK(nb_synchronize(ctx));
    nb_bh_state &b = ctx->bh;
    NB_CUDA(ctx, cudaMemcpy(b.aabb, b.aabb_dev, 7 * sizeof(double), cudaMemcpyDeviceToHost));
    memset(info, 0, sizeof *info);
    info->num_bodies = ctx->n;
    info->num_nodes_materialised = b.num_nodes;
    info->num_internal = b.num_internal;
    info->num_nodes_canonical = 1 + 8 * b.num_internal;
    info->max_depth = b.max_depth;
    for (int k = 0; k < 3; ++k) { info->aabb_min[k] = b.aabb[k]; info->aabb_max[k] = b.aabb[3 + k]; }
    info->aabb_edge = b.aabb[6];
    return NB_OK;
}

int nb_bh_enable_stats(nb_ctx *ctx, int enable) {
    if (!ctx) return NB_ERR_INVALID;
    ctx->bh.stats_enabled = enable != 0;
    return NB_OK;
}

int nb_bh_get_stats(nb_ctx *ctx, uint64_t *total_visits, uint64_t *total_accepts, uint32_t *visits_per_body) {
    if (!ctx || !ctx->bh.built) return nb_fail(ctx, NB_ERR_INVALID, "nb_bh_get_stats: no tree");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CHECK(nb_synchronize(ctx));
    unsigned long long t[8];
    NB_CUDA(ctx, cudaMemcpy(t, ctx->bh.stat_totals, sizeof t, cudaMemcpyDeviceToHost));
    if (getenv("NB_DEBUG_STATS"))
        fprintf(stderr, "[nb stats] visits %llu accepts %llu rounds %llu items %llu mixed %llu ilist %llu\n", t[0], t[1], t[2], t[3], t[4], t[5]);
    if (total_visits) *total_visits = t[0];
    if (total_accepts) *total_accepts = t[1];
    if (visits_per_body) {
        // device counters are in sorted order; return them by body id
        std::vector<uint32_t> v(ctx->n), ids(ctx->n);
        NB_CUDA(ctx, cudaMemcpy(v.data(), ctx->bh.visits, ctx->n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        NB_CUDA(ctx, cudaMemcpy(ids.data(), ctx->id, ctx->n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        for (uint64_t s = 0; s < ctx->n; ++s) visits_per_body[ids[s]] = v[s];
    }
    return NB_OK;
}

// Host-side expansion of the device tree (DFS pre-order, visit-rank child order, empty leaves implied) into the
// reference's canonical node set.  Test / inspection path only.
namespace {
struct HostTree {
    uint64_t n = 0, M = 0;
    std::vector<uint2> meta;
    std::vector<double> msum;
    std::vector<uint32_t> body_count, perm;
    double aabb[7];
};
int fetch_tree(nb_ctx *ctx, HostTree &t) {
    nb_bh_state &b = ctx->bh;
    NB_CHECK(nb_synchronize(ctx));
    t.n = ctx->n;
    t.M = b.num_nodes;
    t.meta.resize(t.M); t.msum.resize(4 * t.M);
    t.body_count.resize(t.M); t.perm.resize(t.n);
    NB_CUDA(ctx, cudaMemcpy(t.meta.data(), b.meta, t.M * sizeof(uint2), cudaMemcpyDeviceToHost));
    NB_CUDA(ctx, cudaMemcpy(t.msum.data(), b.msum, 4 * t.M * sizeof(double), cudaMemcpyDeviceToHost));
    NB_CUDA(ctx, cudaMemcpy(t.body_count.data(), b.body_count, t.M * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    NB_CUDA(ctx, cudaMemcpy(t.perm.data(), ctx->id, t.n * sizeof(uint32_t), cudaMemcpyDeviceToHost));  // slot -> body id
    NB_CUDA(ctx, cudaMemcpy(t.aabb, b.aabb_dev, 7 * sizeof(double), cudaMemcpyDeviceToHost));
    return NB_OK;
}
inline uint32_t rank_to_octant(uint32_t r) {  // rank = 4u + 2b + (1-r)  ->  octant = 4u + 2r + b
    const uint32_t u = r >> 2, bk = (r >> 1) & 1, rt = 1 - (r & 1);
    return 4 * u + 2 * rt + bk;
}
struct CanonSink {
    uint32_t *depth; uint64_t *path_hi, *path_lo; uint32_t *kind, *body, *count;
    double *edge, *minx, *miny, *minz, *mass, *comx, *comy, *comz;
    size_t k = 0;
    std::vector<uint32_t> *sorted = nullptr;
};
// returns false when the node array is inconsistent (never expected; guards the host walk against garbage)
bool canon_rec(const HostTree &t, uint32_t node, int depth, uint64_t phi, uint64_t plo, double edge, double mnx,
               double mny, double mnz, CanonSink &s) {
    if (node >= t.M || depth > NB_MAX_TREE_DEPTH + 1) return false;
    const uint2 m = t.meta[node];
    const bool leaf = (m.y & NB_LEAF_FLAG) != 0;
    if (s.depth) {
        const size_t k = s.k;
        s.depth[k] = depth; s.path_hi[k] = phi; s.path_lo[k] = plo;
        s.kind[k] = leaf ? 1 : 2;
        s.body[k] = leaf ? t.perm[m.y & NB_PAYLOAD_MASK] : (uint32_t) t.n;
        s.count[k] = t.body_count[node];
        s.edge[k] = edge; s.minx[k] = mnx; s.miny[k] = mny; s.minz[k] = mnz;
        s.mass[k] = t.msum[4 * (size_t) node + 3];
        s.comx[k] = t.msum[4 * (size_t) node]; s.comy[k] = t.msum[4 * (size_t) node + 1]; s.comz[k] = t.msum[4 * (size_t) node + 2];
    }
    s.k++;
    if (leaf) {
        if ((m.y & NB_PAYLOAD_MASK) >= t.n) return false;
        if (s.sorted) s.sorted->push_back(t.perm[m.y & NB_PAYLOAD_MASK]);
        return true;
    }
    if (m.x > t.M || m.x <= node) return false;
    uint32_t child_of_octant[8];
    for (int o = 0; o < 8; ++o) child_of_octant[o] = NB_LEAF_FLAG;  // marker: empty
    for (uint32_t c = node + 1; c < m.x;) {
        const uint2 mc = t.meta[c];
        if (mc.x <= c || mc.x > m.x) return false;
        child_of_octant[rank_to_octant(nb_meta_rank(mc.y))] = c;
        c = mc.x;
    }
    const double h = edge / 2;  // ParallelOctreeTopDownSubtrees.cpp:256
    for (uint64_t o = 0; o < 8; ++o) {
        uint64_t h2 = phi, l2 = plo;
        if (depth < 21) h2 |= o << (60 - 3 * depth);
        else if (depth < 42) l2 |= o << (60 - 3 * (depth - 21));
        // child bounds :277-315: +h in y for bit 2, +h in x for bit 1, +h in z when bit 0 is CLEAR
        const double cx = (o & 2) ? mnx + h : mnx;
        const double cy = (o & 4) ? mny + h : mny;
        const double cz = (o & 1) ? mnz : mnz + h;
        if (child_of_octant[o] == NB_LEAF_FLAG) {
            if (s.depth) {
                const size_t k = s.k;
                s.depth[k] = depth + 1; s.path_hi[k] = h2; s.path_lo[k] = l2;
                s.kind[k] = 0; s.body[k] = (uint32_t) t.n; s.count[k] = 0;
                s.edge[k] = h; s.minx[k] = cx; s.miny[k] = cy; s.minz[k] = cz;
                s.mass[k] = 0; s.comx[k] = 0; s.comy[k] = 0; s.comz[k] = 0;
            }
            s.k++;
        } else if (!canon_rec(t, child_of_octant[o], depth + 1, h2, l2, h, cx, cy, cz, s)) {
            return false;
        }
    }
    return true;
}
}  // namespace

int nb_bh_export_canonical(nb_ctx *ctx, uint32_t *depth, uint64_t *path_hi, uint64_t *path_lo, uint32_t *kind,
                           uint32_t *body, uint32_t *count, double *edge, double *minx, double *miny, double *minz,
                           double *mass, double *comx, double *comy, double *comz) {
    if (!ctx || !ctx->bh.built) return nb_fail(ctx, NB_ERR_INVALID, "nb_bh_export_canonical: no tree");
    if (!depth || !path_hi || !path_lo || !kind || !body || !count || !edge || !minx || !miny || !minz || !mass ||
        !comx || !comy || !comz)
        return nb_fail(ctx, NB_ERR_INVALID, "nb_bh_export_canonical: null output");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    HostTree t;
    NB_CHECK(fetch_tree(ctx, t));
    CanonSink s{depth, path_hi, path_lo, kind, body, count, edge, minx, miny, minz, mass, comx, comy, comz};
    if (!canon_rec(t, 0, 0, 0, 0, t.aabb[6], t.aabb[0], t.aabb[1], t.aabb[2], s))
        return nb_fail(ctx, NB_ERR_INVALID, "nb_bh_export_canonical: inconsistent node array");
    return NB_OK;
}

int nb_bh_sorted_bodies(nb_ctx *ctx, uint32_t *sorted_bodies) {
    if (!ctx || !ctx->bh.built || !sorted_bodies) return nb_fail(ctx, NB_ERR_INVALID, "nb_bh_sorted_bodies: no tree");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    HostTree t;
    NB_CHECK(fetch_tree(ctx, t));
    std::vector<uint32_t> order;
    order.reserve(t.n);
    CanonSink s{};
    s.sorted = &order;
    if (!canon_rec(t, 0, 0, 0, 0, t.aabb[6], t.aabb[0], t.aabb[1], t.aabb[2], s) || order.size() != t.n)
        return nb_fail(ctx, NB_ERR_INVALID, "nb_bh_sorted_bodies: inconsistent node array");
    memcpy(sorted_bodies, order.data(), t.n * sizeof(uint32_t));
    return NB_OK;
}

}  // extern "C"
