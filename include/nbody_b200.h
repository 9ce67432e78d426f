/* ============================================================================
 * nbody_b200.h -- C ABI of the B200-native gravity hot path.
 *
 * Drop-in boundary for the force / tree / integrator / energy operators of
 * TimThuering/N-Body-Simulation.  Every entry point names the reference interface it replaces
 * (file:line relative to the reference tree).  The reference passes sycl::queue& + sycl::buffer<double>&
 * per SoA component; here an opaque context owns all device memory (one CUDA device, one stream, optionally
 * one NCCL communicator) and host arrays are plain caller-owned `const double*` that are only touched
 * during the call.
 *
 * Conventions
 *   - all functions return NB_OK (0) or a negative nb_status; nb_last_error() gives the text.
 *   - no exceptions cross the ABI; no torch / STL types in any signature.
 *   - calls are stream-ordered on the context's stream and return after enqueue, except the nb_get_*,
 *     nb_energy, nb_bh_export_* and nb_synchronize calls which synchronise.
 *   - one host thread drives one context.
 *   - there is NO CPU fallback: nb_create fails with NB_ERR_NO_DEVICE when no sm_100 GPU is usable.
 * ==========================================================================*/
#ifndef NBODY_B200_H
#define NBODY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB_ABI_VERSION 1

typedef enum nb_status {
    NB_OK = 0,
    NB_ERR_INVALID = -1,       /* bad argument / call order                                   */
    NB_ERR_NO_DEVICE = -2,     /* no usable CUDA device (there is no CPU fallback)            */
    NB_ERR_CUDA = -3,          /* CUDA runtime error, text in nb_last_error                   */
    NB_ERR_TREE_DEPTH = -4,    /* octree deeper than NB_MAX_TREE_DEPTH: coincident bodies     */
    NB_ERR_NODE_POOL = -5,     /* node pool overflow (reference: silent UB, README.md:117-118)*/
    NB_ERR_COMM = -6,          /* NCCL error                                                  */
    NB_ERR_UNSUPPORTED = -7
} nb_status;

#define NB_MAX_TREE_DEPTH 42 /* 2 x 63-bit octant-path keys */

typedef struct nb_ctx nb_ctx;

/* Plain-old-data mirror of namespace configuration (src/utility/Configuration.hpp:12-90) plus the
 * constants of nBodyAlgorithm (G, nBodyAlgorithm.hpp:55-61).  Fill with nb_config_default() first. */
typedef struct nb_config {
    uint32_t struct_size;        /* = sizeof(nb_config), ABI check                                         */
    int32_t device;              /* CUDA device ordinal                                                    */
    double G;                    /* gravitational constant in AU^3 kg^-1 day^-2                            */
    double epsilon2;             /* configuration::epsilon2, Configuration.cpp:6                           */
    double theta;                /* barnes_hut_algorithm::theta, Configuration.cpp:18 (default 1.05)       */
    int32_t block_size;          /* naive_algorithm::blockSize (:10) -> shared-memory source tile length   */
    int32_t opt_stage;           /* naive_algorithm::optimization_stage (:11), 0..2; one kernel serves all */
    int32_t sort_bodies;         /* barnes_hut_algorithm::sortBodies (:21); traversal order only           */
    int32_t wg_size_barnes_hut;  /* barnes_hut_algorithm::workGroupSize (:22) -> traversal CTA size        */
    int32_t storage_size_param;  /* main.cpp:122-127; node pool = N + param*N/8 nodes (same budget as the ref.) */
    int32_t stack_size_param;    /* main.cpp:129-134; accepted, traversal is stackless                     */
    int32_t num_wi_aabb;         /* AABBWorkItemCount (:14); result is independent of it (min/max)         */
    int32_t num_wi_octree;       /* accepted, ignored (lock-free build)                                    */
    int32_t num_wi_top_octree;   /* accepted, ignored                                                      */
    int32_t num_wi_com;          /* accepted, ignored                                                      */
    int32_t max_level_top_octree;/* accepted, ignored                                                      */
    int32_t precise_rsqrt;       /* 1 (default): cubic rsqrt refinement (~1 ulp); 0: linear (~2e-12 rel)   */
    int32_t world_size;          /* ranks sharing the bodies (1 = single GPU)                              */
    int32_t rank;                /* this context's rank                                                    */
    /* Tuning and A/B switches (0 = production default everywhere; the Python binding names them as given here):
     *   [0] ipt               naive: target bodies per consumer thread (register blocking); 0 -> 4
     *   [1] unfused_advance   1: nb_advance keeps separate integrator kernels for Barnes-Hut (default: fused into the walk)
     *   [2] naive_variant     naive: 1 round-1 form with a producer warp, 2 / 3 three / four CTAs per SM (A/B, see naive.cu)
     *   [3] walk_variant      Barnes-Hut walk: 20 grid-mapped, 50 persistent (SM-local tile queues); 0 picks 50 from 2^19
     *                         bodies per call, else 20
     *   [4] naive_segments    naive: number of source segments per target tile (0: chosen from the grid size)
     *   [5] static_slices     1: several GPUs keep the equal-count slices of nb_slice_bounds for the Barnes-Hut walk
     *                         (default: slices of equal cost, from the clock ticks each 32-body tile took in the latest walk)
     *   [6] sort_variant      tree build: 1 full 8-pass (key, slot) sort, 2 / 3 packed 5-pass / 4-pass sort; 0: per build (see
     *                         bh_build.cu)
     *   [7] com_variant       centre of mass: 1 one launch per level instead of one cooperative launch
     * Environment (developer A/B only): NB_DISABLE_P2P=1 keeps the NCCL all-gather path; NB_EMIT_PER_BODY=1 the round-1
     * node emission kernel.                                                                                          */
    int32_t reserved[8];
} nb_config;

/* Reference defaults (Configuration.cpp:5-22, nBodyAlgorithm.hpp:55-61). */
void nb_config_default(nb_config *cfg);

int nb_abi_version(void);
const char *nb_status_string(int status);

/* ---- context ---------------------------------------------------------------------------------------- */
int nb_create(const nb_config *cfg, nb_ctx **out);
void nb_destroy(nb_ctx *ctx);
const char *nb_last_error(const nb_ctx *ctx);
int nb_synchronize(nb_ctx *ctx);
/* device name as the reference stores it in times.json (NaiveAlgorithm.cpp:76-77). */
int nb_device_name(nb_ctx *ctx, char *buf, size_t buflen);
/* change theta / knobs between calls (configuration::setTheta etc., Configuration.cpp:35-80). */
int nb_set_theta(nb_ctx *ctx, double theta);
int nb_set_block_size(nb_ctx *ctx, int block_size);
int nb_set_sort_bodies(nb_ctx *ctx, int sort_bodies);
int nb_set_precise_rsqrt(nb_ctx *ctx, int precise);

/* ---- bodies: replaces the sycl::buffer wrapping of the SimulationData vectors
 *      (NaiveAlgorithm.cpp:39-66, BarnesHutAlgorithm.cpp:42-69).  Host SoA, fp64, N bodies.
 *      Velocities are the UNADJUSTED input velocities.  May be called again with the same
 *      or a different N; device buffers are reallocated only when N grows.                               */
int nb_set_bodies(nb_ctx *ctx, uint64_t n, const double *mass, const double *x, const double *y, const double *z,
                  const double *vx, const double *vy, const double *vz);
int nb_set_positions(nb_ctx *ctx, const double *x, const double *y, const double *z);
uint64_t nb_num_bodies(const nb_ctx *ctx);

/* ---- force operators ---------------------------------------------------------------------------------- */
/* NaiveAlgorithm::computeAccelerations_opt_{0,1,2} (NaiveAlgorithm.hpp:31-53, .cpp:262-482):
 * a_i = G * sum_j m_j r_ij (|r_ij|^2 + eps2)^(-3/2), self term included; j ascending inside up to 32 source segments
 * whose partial sums are added in ascending order (the segment count depends on N only: the bits of the result do not
 * depend on the device or on world_size).
 * With world_size > 1 the rank computes its target slice and all-gathers the accelerations.            */
int nb_naive_accel(nb_ctx *ctx);

/* BarnesHutOctree::buildOctree (BarnesHutOctree.hpp:102-103; ParallelOctreeTopDownSubtrees.cpp:15-93):
 * AABB (incl. origin) -> octant-path keys -> radix sort -> node construction -> centre of mass -> order. */
int nb_bh_build(nb_ctx *ctx);
/* BarnesHutAlgorithm::computeAccelerations (BarnesHutAlgorithm.hpp:43-46, .cpp:280-401), theta criterion
 * edge*rsqrt(d^2) < theta OR body leaf; requires nb_bh_build on the current positions.                    */
int nb_bh_accel(nb_ctx *ctx);
/* The same traversal for the bodies in storage slots [slot_begin, slot_end) only (storage order is the sorted
 * Morton / DFS order after nb_bh_build) and without the all-gather: what one rank of a world_size-P run executes
 * for its slice (nb_slice_bounds), callable on a single GPU -- used to profile a rank's share.               */
int nb_bh_accel_range(nb_ctx *ctx, uint64_t slot_begin, uint64_t slot_end);

/* ---- integrator ------------------------------------------------------------------------------------------ */
/* Leapfrog part 1 (NaiveAlgorithm.cpp:140-164 = BarnesHutAlgorithm.cpp:157-182): v += a*(dt/2); x += v*dt. */
int nb_leapfrog_part1(nb_ctx *ctx, double dt);
/* Leapfrog part 2 (NaiveAlgorithm.cpp:206-222 = BarnesHutAlgorithm.cpp:223-239): v += a*(dt/2).            */
int nb_leapfrog_part2(nb_ctx *ctx, double dt);
/* part 2 of step k immediately followed by part 1 of step k+1 in one pass (non-visualised steps).         */
int nb_leapfrog_part2_part1(nb_ctx *ctx, double dt);
/* `nsteps` complete leapfrog steps of the reference's time loop without output in between
 * (NaiveAlgorithm.cpp:131-258 / BarnesHutAlgorithm.cpp:149-276 for steps that are not visualised):
 *     per step: part 1; force evaluation (algorithm 0 = naive, 1 = Barnes-Hut build + traversal); part 2.
 * Accelerations of the current positions must be on the device (as after any force call).  Bit-identical to issuing
 * the calls one by one.  On one GPU the inner steps are replayed from a CUDA graph (the small-N time loop is launch
 * bound: one Barnes-Hut step is ~90 kernel launches).  ms (optional, NB_T_COUNT entries): phase times of the first
 * step of the batch, measured with events when timers are enabled -- a sample, the replayed steps are not timed
 * individually.  Asynchronous unless ms is requested.                                                            */
int nb_advance(nb_ctx *ctx, int algorithm, double dt, uint32_t nsteps, double *ms);

/* ---- energy ------------------------------------------------------------------------------------------------ */
/* nBodyAlgorithm::computeEnergy (nBodyAlgorithm.hpp:91-97, .cpp:11-86).
 * out = { kinetic, potential (negative), total, virial 2*Ekin/|Epot| }.  Synchronises.                    */
int nb_energy(nb_ctx *ctx, double out[4]);

/* ---- read-back (synchronising).  Any pointer may be NULL.  Replaces the host_accessor copies at
 *      NaiveAlgorithm.cpp:172-180,233-243 and nBodyAlgorithm::storeAccelerations (.cpp:88-102).          */
int nb_get_positions(nb_ctx *ctx, double *x, double *y, double *z);
int nb_get_velocities(nb_ctx *ctx, double *vx, double *vy, double *vz);
int nb_get_accelerations(nb_ctx *ctx, double *ax, double *ay, double *az);
int nb_get_acceleration_norms(nb_ctx *ctx, double *anorm);

/* ---- one-call operator forms with HOST buffers (the reference operator signature minus the queue):
 *      H2D copy, kernel(s), D2H copy.  Used for end-to-end timing and by bindings that keep state on the host. */
int nb_op_naive_accelerations(nb_ctx *ctx, uint64_t n, const double *mass, const double *x, const double *y,
                              const double *z, double *ax, double *ay, double *az);
int nb_op_barnes_hut_accelerations(nb_ctx *ctx, uint64_t n, const double *mass, const double *x, const double *y,
                                   const double *z, double *ax, double *ay, double *az);

/* ---- Barnes-Hut inspection (parity tests; synchronising) --------------------------------------------------- */
typedef struct nb_tree_info {
    uint64_t num_bodies;
    uint64_t num_nodes_materialised; /* internal + body-leaf nodes held on the device                     */
    uint64_t num_internal;
    uint64_t num_nodes_canonical;    /* 1 + 8*internal = the reference's nextFreeNodeID                   */
    uint32_t max_depth;
    uint32_t reserved;
    double aabb_min[3], aabb_max[3], aabb_edge; /* BarnesHutOctree::min_x.. / AABB_EdgeLength            */
} nb_tree_info;
int nb_bh_tree_info(nb_ctx *ctx, nb_tree_info *info);
/* computeMinMaxValuesAABB only (BarnesHutOctree.cpp:45-191): out = min xyz, max xyz, edge.               */
int nb_bh_aabb(nb_ctx *ctx, double out[7]);
/* Canonical node set, num_nodes_canonical records sorted by (path_hi, path_lo, depth), i.e. DFS order with
 * children in ascending octant code; same record as the oracle's orc_tree_canonical.
 * kind: 0 empty leaf, 1 body leaf, 2 internal.  body = N for "no body".  com* are mass-weighted SUMS.   */
int nb_bh_export_canonical(nb_ctx *ctx, uint32_t *depth, uint64_t *path_hi, uint64_t *path_lo, uint32_t *kind,
                           uint32_t *body, uint32_t *count, double *edge, double *minx, double *miny, double *minz,
                           double *mass, double *comx, double *comy, double *comz);
/* BarnesHutOctree::sortedBodiesInOrder (BarnesHutOctree.cpp:550-613): ascending octant-code order.       */
int nb_bh_sorted_bodies(nb_ctx *ctx, uint32_t *sorted_bodies);
/* Per-body traversal statistics of the last nb_bh_accel: non-empty visits as counted by the
 * `SUM_MASSES != 0 && BODY_OF_NODE != i` branch (BarnesHutAlgorithm.cpp:349).  Enable before the call.  */
int nb_bh_enable_stats(nb_ctx *ctx, int enable);
int nb_bh_get_stats(nb_ctx *ctx, uint64_t *total_visits, uint64_t *total_accepts, uint32_t *visits_per_body);

/* helpers of the reference's default builder, kept for its golden tests (tests/BarnesHutTest.cpp:129-220):
 * prepareSubtrees + sortBodiesForSubtrees (ParallelOctreeTopDownSubtrees.cpp:436-534) on the device.     */
int nb_util_group_by_subtree(nb_ctx *ctx, uint32_t n, const uint32_t *subtree_of_body, uint32_t node_count,
                             uint32_t *body_count_subtree, uint32_t *subtrees, uint32_t *subtree_count,
                             uint32_t *start_index, uint32_t *sorted_bodies);

/* ---- timers: per-phase milliseconds of the most recent call of each phase, CUDA-event timed.
 *      Names follow times.json (BarnesHutAlgorithm.cpp:79-100, ParallelOctreeTopDownSubtrees.cpp:76-92). */
typedef enum nb_timer {
    NB_T_ACCEL = 0,        /* "Acceleration Kernel Time" */
    NB_T_LEAPFROG1,        /* "Leapfrog Part 1" */
    NB_T_LEAPFROG2,        /* "Leapfrog Part 2" */
    NB_T_AABB,             /* "AABB creation" */
    NB_T_KEYS_SORT,        /* "Sort bodies for subtrees" (octant keys + radix sort + state reorder) */
    NB_T_BUILD,            /* "Build subtrees" (node construction + per-depth node lists) */
    NB_T_COM,              /* "Compute center of mass" */
    NB_T_TREE_TOTAL,       /* "Octree creation" */
    NB_T_ENERGY,
    NB_T_COMM,             /* all-gather */
    NB_T_COUNT
} nb_timer;
int nb_enable_timers(nb_ctx *ctx, int enable);
int nb_get_timers(nb_ctx *ctx, double ms[NB_T_COUNT]);
const char *nb_timer_name(int timer);

/* ---- multi-GPU: one context per process/GPU; bodies replicated, targets sharded by contiguous ranges of
 *      the sorted order, accelerations all-gathered over NCCL each evaluation (new functionality: the
 *      reference is single device, NaiveAlgorithm.cpp:56-62).                                            */
#define NB_COMM_ID_BYTES 128
int nb_comm_get_unique_id(uint8_t id[NB_COMM_ID_BYTES]); /* rank 0 creates, host broadcasts               */
int nb_comm_init(nb_ctx *ctx, const uint8_t id[NB_COMM_ID_BYTES], int world_size, int rank);
/* 1 when every rank has mapped every other rank's state arrays (CUDA IPC over NVLink): the Barnes-Hut walk then stores
 * its results straight into all ranks' arrays (two barriers instead of the all-gather) and nb_advance runs its fused
 * form on several GPUs.  0: NCCL all-gather path (IPC unavailable, NB_DISABLE_P2P set, or more than 8 ranks).        */
int nb_comm_p2p_enabled(const nb_ctx *ctx);
/* slice of targets [begin, end) a rank owns (host-side logic; no GPU needed)                             */
void nb_slice_bounds(uint64_t n, int world_size, int rank, uint64_t *begin, uint64_t *end);

/* user events on the context's stream (CUDA events; bench.py brackets its timed region with them).  slot 0..7.  */
int nb_event_record(nb_ctx *ctx, int slot);
int nb_event_elapsed_ms(nb_ctx *ctx, int slot_begin, int slot_end, double *ms); /* synchronises on slot_end */

/* ---- measurement helpers ---------------------------------------------------------------------------------- */
/* DFMA-chain microbenchmark on the context's device: achieved fp64 TFLOP/s (2 flops per DFMA).           */
int nb_measure_fp64_peak(nb_ctx *ctx, double *tflops);
/* number of kernel launches issued by this context so far (bench.py's gpu_launches).                     */
uint64_t nb_launch_count(const nb_ctx *ctx);
/* raw device pointers (x,y,z,vx,vy,vz,ax,ay,az,mass) for zero-copy wrappers.  The arrays are in STORAGE order: body-id
 * order until the first nb_bh_build, the sorted (Morton) order of the latest build afterwards, and every build swaps
 * the buffers -- re-query after nb_bh_build / nb_set_bodies.                                              */
int nb_device_pointers(nb_ctx *ctx, void *ptrs[10]);

#ifdef __cplusplus
}
#endif
#endif /* NBODY_B200_H */
