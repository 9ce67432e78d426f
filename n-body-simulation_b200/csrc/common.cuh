// Shared definitions of the sm_100a gravity kernels: context, error handling, small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/nbody_b200.h"

#define NB_SM_COUNT_FALLBACK 148
#define NB_MAX_PEERS 8

// Packed source record of the all-pairs kernel: one 48-byte AoS entry per source body so a whole tile is a
// single contiguous TMA bulk copy.  c1 = 1.5*m and c2 = 1.875*m are the Taylor coefficients of
// m*(1-e)^(-3/2) (see naive.cu).
struct __align__(16) nb_src_rec {
    double x, y, z, m;
    double c1, c2;
};

// Barnes-Hut node record (DFS pre-order array, children of a node in the reference's visit order
// [2,0,3,1,6,4,7,5], BarnesHutAlgorithm.cpp:370-385).  40 bytes of traversal payload per node:
//   com[4*n+0..2] = centre of mass (already divided by the mass), com[4*n+3] = mass     (32 B)
//   meta[n].x     = index of the first node AFTER this node's subtree ("skip link")
//   meta[n].y     = body leaf:     bit 31 | visit rank of the node inside its parent << 28 | sorted index of the body
//                   internal node: depth << 21 | visit rank (bits 0-2).  The walk adds this word to the high word of
//                   a squared distance: depth << 21 is the exponent shift of 4^depth, the rank only widens the
//                   undecided band of the acceptance test (bh_traverse.cu)                             ( 8 B)
#define NB_LEAF_FLAG 0x80000000u
#define NB_DIGIT_SHIFT 28
#define NB_PAYLOAD_MASK 0x0fffffffu
#define NB_DEPTH_SHIFT 21
static __host__ __device__ __forceinline__ uint32_t nb_meta_rank(uint32_t y) {
    return (y & NB_LEAF_FLAG) ? (y >> NB_DIGIT_SHIFT) & 7u : y & 7u;
}
static __host__ __device__ __forceinline__ uint32_t nb_meta_depth(uint32_t y) { return y >> NB_DEPTH_SHIFT; }  // internal nodes

struct nb_bh_state {
    // per-body, sorted order
    uint64_t *key_hi = nullptr;                              // octant-path key, levels 0..20 (visit-rank digits)
    uint64_t *key_hi_alt = nullptr;                          // radix sort ping-pong; after the sort: key word of levels 21..41 (where needed)
    uint64_t *word_a = nullptr, *word_b = nullptr;           // packed sort words {upper key bits | storage slot}, ping-pong
    uint32_t *perm = nullptr, *perm_alt = nullptr;           // sorted index -> storage slot before this build's reorder
    int32_t *delta = nullptr;                                // common-prefix digits of sorted neighbours (i, i+1)
    uint32_t *chain_cnt = nullptr, *chain_base = nullptr;    // internal nodes starting at body i, and their scan
    uint32_t *leaf_node = nullptr;                           // node index of the leaf of sorted body i
    // per-node
    double *com = nullptr;                                   // 4 doubles per node
    double *msum = nullptr;                                  // 4 doubles per node: {sum m*x, sum m*y, sum m*z, sum m} (reference's massCenters_* / sumOfMasses)
    uint2 *meta = nullptr;
    uint32_t *level = nullptr, *level_list = nullptr;        // per-depth internal-node lists (count / start / cursor, nodes)
    uint32_t *body_count = nullptr;
    // sort scratch
    uint32_t *hist = nullptr;
    // results
    uint32_t *visits = nullptr;                              // per-body visit counters (stats)
    unsigned long long *stat_totals = nullptr;               // {visits, accepts}
    int walk_ctas_per_sm = 0, walk_ctas_threads = 0;         // occupancy of the persistent walk (cached)
    int com_ctas_per_sm = 0;                                 // occupancy of the cooperative centre-of-mass kernel (cached)
    // device scalars
    double *aabb_dev = nullptr;                              // min xyz, max xyz, edge, unused, then the walk's acceptance threshold of
                                                             // every depth (bh_accept.cuh), computed for accept_theta
    double accept_theta = 0;
    double *aabb_partial = nullptr;
    uint32_t *dev_flags = nullptr;                           // [0] error bits of the latest build (1 depth, 2 pool), [1] internal node count,
                                                             // [2] max depth, [3] longest run the packed sort leaves undecided (capped),
                                                             // [4] error bits of all builds since they were last reported (sticky)
    uint32_t *stat_host = nullptr;                           // pinned copies of dev_flags[3] and [6] of the latest finished build
    cudaEvent_t stat_event = nullptr;
    bool stat_pending = false, stat_known = false, long_runs = false, long_runs32 = false;
    int sort_passes = 0;                                     // passes of the latest build's sort (8 = the full sort)
    uint64_t cap_bodies = 0, cap_nodes = 0;
    uint64_t num_nodes = 0, num_internal = 0;
    uint32_t max_depth = 0;
    double aabb[7] = {0, 0, 0, 0, 0, 0, 0};
    bool built = false;
    bool stats_enabled = false;
};

struct nb_ctx {
    nb_config cfg;
    int device = 0;
    int sm_count = NB_SM_COUNT_FALLBACK;
    bool coop_launch = false;   // device supports cooperative launches (grid-wide barrier in the centre-of-mass pass)
    cudaStream_t stream = nullptr;
    std::string last_error;
    std::string device_name;
    uint64_t n = 0, cap = 0;
    // SoA state in STORAGE order.  Storage order is body-id order until the first Barnes-Hut build; every build then
    // physically permutes the whole state into the new sorted (Morton / DFS) order, so the tree kernels stream it and
    // consecutive steps permute by a near-identity map.  id[slot] = body id of the slot (valid when !identity_order).
    double *m = nullptr, *x = nullptr, *y = nullptr, *z = nullptr;
    double *vx = nullptr, *vy = nullptr, *vz = nullptr;
    double *ax = nullptr, *ay = nullptr, *az = nullptr;
    double *alt[10] = {};            // ping-pong partners of m,x,y,z,vx,vy,vz,ax,ay,az (also scratch for read-back)
    // All 21 arrays above (10 + 10 partners + anorm) are carved out of ONE allocation, so that a peer GPU's copy of any of
    // them is "peer slab base + the same byte offset" (multi-GPU: the walk stores its results straight into every
    // rank's arrays over NVLink, bh_traverse.cu).
    unsigned char *slab = nullptr;
    size_t slab_bytes = 0;
    // cost-weighted Morton slices (several GPUs): clock ticks every 32-body tile of the sorted order took in the latest
    // walk (written into every rank's copy by the walk itself; lives in the slab), tile_start = scratch for the tile's
    // start time, dyn_bounds[r] .. dyn_bounds[r + 1] = the slots rank r walks next (multiples of 32, identical on
    // every rank because computed from identical costs)
    uint32_t *tile_cost = nullptr, *tile_start = nullptr;
    unsigned long long *dyn_bounds = nullptr;
    bool bounds_valid = false;
    // accelerations: computed for the current positions and stored in the current storage order?  Both turn false when
    // the positions advance / a build leaves them behind (the permutation skips arrays nobody will read).
    bool a_fresh = false, a_order_ok = true;
    uint32_t *id = nullptr, *id_alt = nullptr;
    bool identity_order = true;
    double *anorm = nullptr;
    // naive
    nb_src_rec *src = nullptr;
    uint64_t src_cap = 0;
    double *naive_partial = nullptr;  // per-segment partial accelerations when the source range is split
    size_t naive_partial_cap = 0;
    // energy
    double *e_partial = nullptr;  // 2*n doubles: kinetic, potential per body
    nb_bh_state bh;
    // timers
    bool timers_enabled = false;
    cudaEvent_t ev[2 * NB_T_COUNT] = {};
    bool ev_valid[NB_T_COUNT] = {};
    uint64_t launches = 0;
    cudaEvent_t user_ev[8] = {};
    // nb_advance: CUDA graph of two consecutive inner steps (the Barnes-Hut build swaps the state with its ping-pong
    // partners once per step, so the pointer configuration has period two)
    cudaGraphExec_t step_graph = nullptr;
    int graph_algorithm = -1;
    double graph_dt = 0;
    uint64_t graph_n = 0;
    nb_config graph_cfg = {};
    int graph_sort_passes = 0;        // sort passes of the builds inside the graph (chosen at capture time)
    const void *graph_ptrs[8] = {};
    uint64_t graph_launches = 0;      // kernel launches one replay stands for
    bool graph_unusable = false;      // the period-two assumption did not hold: stay on the eager path
    bool capturing = false;           // inside the stream capture of nb_advance: no host-visible side effects
    // comm
    void *nccl_comm = nullptr;
    int world = 1, rank = 0;
    // peer slabs mapped with CUDA IPC (index = rank; own entry = slab); p2p_ok: every rank mapped every peer
    unsigned char *peer_slab[NB_MAX_PEERS] = {};
    bool p2p_ok = false;
    double *barrier_word = nullptr;   // one device double for the NCCL barrier (all-reduce of nothing)
};

// what a kernel needs to store into every rank's copy of an array: p-th copy of local pointer q is
// base[p] + ((unsigned char *) q - base[rank])
struct nb_peer_table {
    int world, rank;
    unsigned char *base[NB_MAX_PEERS];
};

int nb_fail(nb_ctx *ctx, int status, const char *fmt, ...);

#define NB_CUDA(ctx, expr)                                                                              \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return nb_fail((ctx), NB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                           __FILE__, __LINE__);                                                         \
    } while (0)

#define NB_CHECK(expr)                \
    do {                              \
        int _s = (expr);              \
        if (_s != NB_OK) return _s;   \
    } while (0)

#define NB_LAUNCH_CHECK(ctx)                                                                            \
    do {                                                                                                \
        (ctx)->launches++;                                                                              \
        cudaError_t _e = cudaGetLastError();                                                            \
        if (_e != cudaSuccess)                                                                          \
            return nb_fail((ctx), NB_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                                         \
    } while (0)

template <typename T>
static inline int nb_alloc(nb_ctx *ctx, T **p, size_t count) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (count == 0) return NB_OK;
    NB_CUDA(ctx, cudaMalloc((void **) p, count * sizeof(T)));
    return NB_OK;
}
template <typename T>
static inline void nb_free(T **p) {
    if (*p) { cudaFree(*p); *p = nullptr; }
}

struct nb_timer_scope {
    nb_ctx *ctx; int id;
    nb_timer_scope(nb_ctx *c, int i) : ctx(c), id(i) {
        if (ctx->timers_enabled) cudaEventRecord(ctx->ev[2 * id], ctx->stream);
    }
    ~nb_timer_scope() {
        if (ctx->timers_enabled) { cudaEventRecord(ctx->ev[2 * id + 1], ctx->stream); ctx->ev_valid[id] = true; }
    }
};

// ---- internal entry points implemented across the .cu files ------------------------------------------------
int nbk_naive_accel(nb_ctx *ctx, uint64_t i_begin, uint64_t i_end);
int nbk_leapfrog_part1(nb_ctx *ctx, double dt);
int nbk_leapfrog_part2(nb_ctx *ctx, double dt);
int nbk_leapfrog_part2_part1(nb_ctx *ctx, double dt);
int nbk_accel_norm(nb_ctx *ctx);
int nbk_energy(nb_ctx *ctx, uint64_t j_begin, uint64_t j_end);
int nbk_fp64_peak(nb_ctx *ctx, double *tflops);
int nbk_bh_reserve(nb_ctx *ctx);
void nbk_bh_release(nb_ctx *ctx);
int nbk_bh_aabb(nb_ctx *ctx);
int nbk_bh_accept_table(nb_ctx *ctx);           // acceptance thresholds for the current theta (no-op when up to date)
int nbk_bh_build(nb_ctx *ctx);
int nbk_bh_accel(nb_ctx *ctx, uint64_t s_begin, uint64_t s_end);
int nbk_unpermute(nb_ctx *ctx, int count, const double *const *src, double *const *dst);
int nbk_permute_in(nb_ctx *ctx, int count, const double *const *src, double *const *dst);
int nbk_comm_allgather_accel(nb_ctx *ctx, double *ax, double *ay, double *az, uint64_t n);
void nbk_comm_destroy(nb_ctx *ctx);
int nbk_comm_barrier(nb_ctx *ctx);              // stream-ordered barrier over all ranks (no-op on one rank)
int nbk_comm_map_peers(nb_ctx *ctx);            // collective: exchange the slab handles, map the peers' slabs
void nbk_comm_unmap_peers(nb_ctx *ctx);         // collective when peers were mapped (ends with a barrier)
nb_peer_table nbk_peer_table(const nb_ctx *ctx);
int nbk_bh_accel_fused(nb_ctx *ctx, uint64_t s_begin, uint64_t s_end, int epilogue, double dt, bool to_peers, bool dynamic_slices = false);
int nbk_bh_rebalance(nb_ctx *ctx);              // new slice bounds from the tile costs of the latest walk
int nbk_comm_allreduce_sum(nb_ctx *ctx, double *buf, size_t count);

// ---- device helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ double nb_rsqrt_seed(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));  // MUFU.RSQ64H, ~2^-21 relative
    return y;
}
