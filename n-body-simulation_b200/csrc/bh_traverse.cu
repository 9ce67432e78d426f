// Barnes-Hut force evaluation for sm_100a: warp-cooperative, stackless traversal of the DFS pre-order node array.
//
// Replaces BarnesHutAlgorithm::computeAccelerations (reference src/simulationBackend/BarnesHutAlgorithm.cpp:280-401):
// per body an explicit stack in GLOBAL memory (stackSize*N uint32, :8-15), 8 child-id loads + 8 stack writes per opened
// node, three fp64 divides per visit (:351-353).  Here:
//   * nodes live in DFS pre-order with children in the reference's pop order [2,0,3,1,6,4,7,5] (:370-385), so the
//     traversal is "next = n+1 (open) or skip[n] (accept / leaf)" -- no stack, no child table;
//   * one warp walks the UNION of its 32 lanes' node lists: cursor = min over lanes of their next node; only lanes whose
//     next == cursor interact.  Every lane therefore sees exactly the nodes the reference's per-body walk sees, in the
//     same order (no warp-vote widening of the acceptance test), but node loads are warp-uniform broadcasts;
//   * bodies are processed in sorted (Morton / DFS) order so neighbouring lanes share almost all of their lists;
//   * centre of mass is pre-divided in the COM pass (same IEEE quotient the reference computes per visit);
//   * acceptance test edge*rsqrt(d2) < theta is decided without memory access from the node's depth: the kernel
//     compares the high word of d2 + eps2 with hiword((edge_0/theta)^2) - (depth << 21) on the integer pipe; the
//     vanishing band around the threshold is decided by ONE fp64 compare with the depth's exact threshold
//     (bh_accept.cuh: the reference's expression is monotone in d2, the threshold is found with the reference's own
//     correctly rounded operations once per build), so the interaction set is identical;
//   * production launches are persistent: warps draw 32-body tiles from SM-local queues (see the kernel's comment);
//   * force: MUFU.RSQ64H seed + cubic Taylor refinement of (d2+eps2)^(-3/2) (see naive.cu).
// Payload per visited node: 32 B {com xyz, mass} + 8 B {skip, leaf|body / depth} = 40 B.
#include "common.cuh"
#include "bh_accept.cuh"

#include <algorithm>
#include <type_traits>

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// Production walk: acceptance test on the integer pipe, SM-local tile queues.
//
//   * D = dx^2 + dy^2 + dz^2 + eps2 is ONE fma chain (eps2 folded into the first term) and serves both the acceptance
//     test and the force;
//   * positive doubles order like their bit patterns, so "D > (edge_d/theta)^2" is decided by comparing HIGH WORDS:
//     W_d = hiword((edge_0/theta)^2) - (depth << 21)  (the edge halves exactly per level); accept when
//     hiword(D) >= W_d + 2, open when hiword(D) <= W_d - 2 (evaluated as V = hiword(D) + (depth << 21 | rank) against
//     W_0 + 8 and W_0 - 1).  The band in between (relative width 10^-5) and every depth for which eps2 is not negligible
//     against (edge_d/theta)^2 take exact_accept() below, so the decision is the reference's;
//   * PERSIST: the grid only fills the machine and every WARP draws 32-body tiles from a queue that belongs to the SM
//     it runs on (queue c owns runs of RUN consecutive tiles of the sorted order, interleaved with the other queues'
//     runs; a warp whose queue is empty steals from the following queues).  The 48 warps resident on an SM therefore
//     walk neighbouring bodies at the same time and share the node records they pull into that SM's L1 -- the hardware
//     block scheduler would hand an SM blocks that are 148 blocks apart -- while all SMs stay inside one moving window
//     of the tree (L2), and no CTA is ever re-launched.  ncu at N = 2^24: L1 hit rate 65 % -> 73 %, warps active
//     57 % -> 62 %, 89.6 ms -> 84.2 ms.  tile_counters: one uint32 per queue, zeroed before the launch.
//     The queue bookkeeping must not add live values to the cursor loop: with RUN as a run-time argument ptxas spilled
//     two loop invariants and the walk fell back to 90 ms, hence the template constant.
//   * the cursor step is bound by issue slots and by the latency the resident warps can hide, so both every instruction
//     taken out of it and every warp added show.  Round 1: the acceptance test is ONE add V = hiword(D) + (depth << 21)
//     and two compares against W_0 + 8 and W_0 - 1 (kept opaque so they are not re-derived per node), the skip link is
//     used as it is (it always points behind the node's subtree, no max), the depth guard t >= t_lim is compiled out of
//     the loop when the deepest level of THIS tree (flags[2]) cannot reach it: 56 -> 45 SASS instructions per accepted
//     cell.  Round 2: accept / open / undecided as three paths with their own copy of the force arithmetic (41); then
//       - the node's {com, mass} is one 256-bit load (LDG.E.256, sm_100) instead of two 128-bit ones            (40)
//       - the skip link is loaded straight into `next`; the paths that do not follow it overwrite it              (39)
//       - the exact re-test is a compare with a table entry instead of sqrt + division: no slow-path calls, six
//         registers fewer, nothing spills, and the register allocator stops re-loading loop invariants
//       - the bases of the two node arrays and the constant 1.875 are made values that only registers can hold (an
//         offset that is always zero but comes from memory): 3 x LDC per step gone                                (36)
//       - what is left fits 40 registers (__launch_bounds__(256, 6)): 48 instead of 40 resident warps per SM.
//     N = 2^24, theta = 0.5: 72.7 ms (41 instructions, 40 warps) -> 69.5 ms (36, 40) -> 65.4 ms (36, 48), bit-identical
//     results (profiles/walk36_ab_r02.log).
// Explored on top of this and rejected (N = 2^24, theta = 0.5, all parity-green): issuing the next node's loads before
// the force arithmetic (software pipelining: 95 ms, the 12 extra live registers cost 20 % of the resident warps);
// prefetch.global.L1 of the next node (98-102 ms); ticketed one-tile-per-warp assignment on a full grid (87 ms); one
// straight-line accept/skip decision for leaves and cells instead of the leaf branch (88.9 ms).
// ---------------------------------------------------------------------------------------------------------------------
// The undecided band, and every depth where eps2 is not negligible: the reference's expression (BarnesHutAlgorithm.cpp:
// 355-359) is monotone in the squared distance, so it is ONE comparison with the depth's exact threshold (bh_accept.cuh;
// the table is computed with the reference's own operations once per build).  d2 is summed as the reference sums it,
// without eps2 and without contraction.
__device__ __forceinline__ bool exact_accept(double dx, double dy, double dz, const double *__restrict__ accept_thr, uint32_t depth) {
    const double d2o = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    return d2o > accept_thr[depth];
}

// What a warp does with the accelerations of its 32 bodies once their walk is finished (EPI).  The walk never reads
// another body's position or velocity (leaves carry their own copy of the body in the node array), so the leapfrog
// half-steps that follow the force evaluation in the time loop can be applied to a body in place the moment its
// acceleration is known -- no separate pass over the state, and on several GPUs no all-gather: the results are stored
// straight into every rank's arrays over NVLink (peer slabs, comm.cu).  Same operations in the same order as
// integrator.cu, so the trajectory is bit-identical to separate calls.
//   NB_EPI_ACCEL      a = G * sum                                                   (BarnesHutAlgorithm.cpp:388-390)
//   NB_EPI_KICK       a as above; v += a*(dt/2)                           leapfrog part 2 (BarnesHutAlgorithm.cpp:223-239)
//   NB_EPI_KICK_DRIFT part 2 of this step and part 1 of the next: v = (v + a*(dt/2)) + a*(dt/2); x += v*dt   (:157-182)
enum { NB_EPI_ACCEL = 0, NB_EPI_KICK = 1, NB_EPI_KICK_DRIFT = 2 };

struct nb_walk_out {
    double *ax, *ay, *az;      // local arrays (this rank's slab)
    double *vx, *vy, *vz;
    double *x, *y, *z;
    double dt;
    nb_peer_table peers;       // world == 1: local stores only
    // cost-weighted slices (null when unused): the slots this rank walks are bounds[rank] .. bounds[rank + 1] instead of
    // the kernel arguments; tile_cost (in the slab: stored to every rank) receives the clock ticks of each tile,
    // tile_start is local scratch (the start time is parked in memory: no register lives across the cursor loop)
    const unsigned long long *bounds;
    uint32_t *tile_cost, *tile_start;
    int rank;
};

// store v into slot b of `local` on every rank (the own rank included)
__device__ __forceinline__ void store_everywhere(const nb_peer_table &pt, double *local, uint64_t b, double v) {
    if (pt.world <= 1) { local[b] = v; return; }
    const ptrdiff_t off = reinterpret_cast<unsigned char *>(local) - pt.base[pt.rank];
#pragma unroll
    for (int p = 0; p < NB_MAX_PEERS; ++p)
        if (p < pt.world) reinterpret_cast<double *>(pt.base[p] + off)[b] = v;
}

// Called once per 32-body tile; kept out of line so that the cursor loop of the walk is compiled exactly as it is
// without an epilogue (inlined, ptxas wrapped the cursor step into an extra convergence barrier: +5 instructions per
// node).
template <int EPI>
__device__ __noinline__ void walk_epilogue(const nb_walk_out &out, uint64_t b, double G, double ax, double ay, double az,
                                           double px, double py, double pz) {
    if (out.tile_cost && (threadIdx.x & 31) == 0) {   // lane 0 is the first body of the tile: b is a multiple of 32
        const uint32_t ticks = (uint32_t) clock() - out.tile_start[b >> 5];
        const nb_peer_table &pt = out.peers;
        const ptrdiff_t off = reinterpret_cast<unsigned char *>(out.tile_cost) - pt.base[pt.rank];
#pragma unroll
        for (int p = 0; p < NB_MAX_PEERS; ++p)
            if (p < pt.world) reinterpret_cast<uint32_t *>(pt.base[p] + off)[b >> 5] = ticks | 1u;
    }
    const double gx = __dmul_rn(ax, G), gy = __dmul_rn(ay, G), gz = __dmul_rn(az, G);
    if (EPI != NB_EPI_KICK_DRIFT) {
        store_everywhere(out.peers, out.ax, b, gx);
        store_everywhere(out.peers, out.ay, b, gy);
        store_everywhere(out.peers, out.az, b, gz);
    }
    if (EPI != NB_EPI_ACCEL) {
        const double h = out.dt / 2.0;
        const double kx = __dmul_rn(gx, h), ky = __dmul_rn(gy, h), kz = __dmul_rn(gz, h);
        double wx = __dadd_rn(out.vx[b], kx), wy = __dadd_rn(out.vy[b], ky), wz = __dadd_rn(out.vz[b], kz);
        if (EPI == NB_EPI_KICK_DRIFT) {
            wx = __dadd_rn(wx, kx); wy = __dadd_rn(wy, ky); wz = __dadd_rn(wz, kz);
            store_everywhere(out.peers, out.x, b, __dadd_rn(px, __dmul_rn(wx, out.dt)));
            store_everywhere(out.peers, out.y, b, __dadd_rn(py, __dmul_rn(wy, out.dt)));
            store_everywhere(out.peers, out.z, b, __dadd_rn(pz, __dmul_rn(wz, out.dt)));
        }
        store_everywhere(out.peers, out.vx, b, wx);
        store_everywhere(out.peers, out.vy, b, wy);
        store_everywhere(out.peers, out.vz, b, wz);
    }
    if (out.peers.world > 1) __threadfence_system();   // the peers read these stores after the next barrier
}

// resident 256-thread CTAs per SM the register budget is sized for: 5 -> 48 registers (40 warps), 6 -> 40 (48 warps)
#ifndef NB_WALK_MIN_CTAS
#define NB_WALK_MIN_CTAS 6
#endif

template <bool STATS, bool PERSIST, uint32_t RUN, int EPI>
__global__ void __launch_bounds__(256, NB_WALK_MIN_CTAS)
bh_traverse_iw_kernel(const double4 *__restrict__ com, const uint2 *__restrict__ meta, const uint32_t *__restrict__ flags,
                      uint64_t n_bodies, const double *__restrict__ aabb, const double *__restrict__ sx, const double *__restrict__ sy,
                      const double *__restrict__ sz, uint64_t s_begin, uint64_t s_end, double theta, double eps2, double G,
                      const __grid_constant__ nb_walk_out out, uint32_t *__restrict__ visits,
                      unsigned long long *__restrict__ totals, uint32_t *__restrict__ tile_counters, uint32_t n_chunks,
                      double k1875) {
    // The node array bases are used in every cursor step.  As plain kernel parameters ptxas re-loads them from the
    // constant bank in every step (LDC.64 x 2); offset by a word that is always zero but comes from memory they are
    // values only a register can hold.
    {
        const unsigned long long zero = flags[5];
        asm volatile("add.u64 %0, %0, %2;\n\tadd.u64 %1, %1, %2;" : "+l"(com), "+l"(meta) : "l"(zero));
        asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(k1875) : "d"(__longlong_as_double((long long) zero)));   // + 0.0
    }
    const double edge0 = aabb[6];
    const double ratio0 = (edge0 / theta) * (edge0 / theta);
    const bool scalable = ratio0 > 1e-200 && ratio0 < 1e200;   // theta == 0 or absurd boxes: always the exact branch
    const int W0 = __double2hiint(ratio0);
    int Whi = W0 + 8, Wlo = W0 - 1;
    asm("" : "+r"(Whi), "+r"(Wlo));   // opaque: two plain compare operands, not W_0 plus an add per node
    // fast decisions need eps2 <= 2^-24 (edge_d/theta)^2 and a normal threshold: depth << 21 must stay below t_lim
    int w_min = __double2hiint(eps2) + (24 << 20);
    if (w_min < (64 << 20)) w_min = 64 << 20;
    const uint32_t t_lim = (scalable && W0 > w_min) ? (uint32_t) (W0 - w_min) : 0u;
    const uint32_t n_nodes = (uint32_t) n_bodies + flags[1];
    // a failed build (depth / pool flag) leaves no valid tree: produce zeros instead of walking garbage
    const bool tree_ok = flags[0] == 0;
    const int lane = threadIdx.x & 31;
    if (out.bounds) {   // cost-weighted slice of this rank (multiples of 32) instead of the equal-count one
        s_begin = out.bounds[out.rank];
        s_end = out.bounds[out.rank + 1];
    }
    const uint64_t tiles_total = (s_end - s_begin + 31) >> 5;
    const uint32_t T = PERSIST ? (uint32_t) ((tiles_total + n_chunks - 1) / n_chunks) : 0u;
    uint32_t home = 0, probe = 0;
    if (PERSIST) {
        asm("mov.u32 %0, %%smid;" : "=r"(home));
        home %= n_chunks;
    }
    uint64_t tile = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned long long nvis_w = 0, nacc_w = 0;
    // The depth guard "t >= t_lim" (eps2 not negligible against (edge_d/theta)^2) cannot fire when even the deepest level
    // of THIS tree (flags[2], written by the build) is below the limit: the tile loop exists in two instantiations and
    // the common one has no guard in the cursor step.
    const bool guard_needed = ((flags[2] + 1u) << 21) >= t_lim || flags[2] >= 2047u;
    auto run = [&](auto guard_tag) {
    constexpr bool GUARD = decltype(guard_tag)::value;
    for (;;) {
        if constexpr (PERSIST) {
            bool have = false;
            while (probe < n_chunks) {
                uint32_t cidx = home + probe;
                if (cidx >= n_chunks) cidx -= n_chunks;
                uint32_t t = 0;
                if (lane == 0) t = atomicAdd(&tile_counters[cidx], 1u);
                t = __shfl_sync(0xffffffffu, t, 0);
                // RUN == 0: queue c is one contiguous chunk of T tiles; otherwise queue c owns runs of RUN consecutive
                // tiles interleaved with the other queues' runs (its run r is run r*n_chunks + c of the sorted order),
                // so all SMs stay inside one moving window of the tree (L2) while the warps of one SM walk
                // neighbouring bodies (L1)
                tile = RUN ? ((uint64_t) (t / (RUN ? RUN : 1u)) * n_chunks + cidx) * RUN + (t % (RUN ? RUN : 1u))
                           : (uint64_t) cidx * T + t;
                if ((RUN || t < T) && tile < tiles_total) { have = true; break; }
                ++probe;   // this queue is empty for good: own queue first, then steal from the following ones
            }
            if (!have) break;
        } else {
            if (tile >= tiles_total) break;
        }
        const uint64_t b = s_begin + tile * 32 + lane;
        const bool valid = b < s_end;
        if (out.tile_cost && lane == 0) out.tile_start[b >> 5] = (uint32_t) clock();
        const uint32_t me = (uint32_t) b;
        double px = 0, py = 0, pz = 0;
        if (valid) { px = sx[b]; py = sy[b]; pz = sz[b]; }
        double ax = 0, ay = 0, az = 0;
        uint32_t next = (valid && tree_ok) ? 0u : 0xffffffffu;
        uint32_t nvis = 0, nacc = 0;

        uint32_t cur = __reduce_min_sync(0xffffffffu, next);
        while (cur < n_nodes) {
            if (next == cur) {
                // warp-uniform address: broadcast transactions; sm_100 has 256-bit global loads (LDG.E.256)
                double4 c;
                asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(c.x), "=d"(c.y), "=d"(c.z), "=d"(c.w) : "l"(com + cur));
                uint2 mt;
                asm("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(mt.x), "=r"(mt.y) : "l"(meta + cur));
                next = mt.x;   // accepted cells keep it; every other path overwrites it with cur + 1
                const double dx = c.x - px, dy = c.y - py, dz = c.z - pz;
                const double D = fma(dz, dz, fma(dy, dy, fma(dx, dx, eps2)));
                // SUM_MASSES == 0 nodes are invisible in the reference (BarnesHutAlgorithm.cpp:349): massless bodies,
                // and cells that hold only massless bodies, are neither counted nor opened.  Their contribution is
                // exactly 0.0 either way (the build stores a finite record for them), so only the instrumented build
                // pays for the check that keeps the visit counts identical to the reference's.
                const bool massless = STATS && (__double2hiint(c.w) | __double2loint(c.w)) == 0;
                auto force = [&]() {
                    if (STATS) nacc += 1u;
                    const double y0 = nb_rsqrt_seed(D);
                    const double y2 = y0 * y0;
                    const double e = fma(-D, y2, 1.0);
                    const double y3 = y2 * y0;
                    const double p = fma(k1875, e, 1.5);   // 1.875 arrives as a kernel parameter: one LDC instead of two moves
                    const double q = fma(p, e, 1.0);
                    const double s = (y3 * c.w) * q;
                    ax = fma(dx, s, ax);
                    ay = fma(dy, s, ay);
                    az = fma(dz, s, az);
                };
                if (mt.y & NB_LEAF_FLAG) {
                    const bool interact = (mt.y & NB_PAYLOAD_MASK) != me && !massless;  // own leaf skipped (:349)
                    next = cur + 1;
                    if (STATS) nvis += interact ? 1u : 0u;
                    if (interact) force();
                } else if (massless) {
                    next = max(mt.x, cur + 1);
                } else {
                    // mt.y = depth << 21 | visit rank (3 bits), so V = hiword(D) + (depth << 21) + rank is ONE add.  The
                    // rank only widens the undecided band: accept for sure when V - 7 >= W_0 + 2 (V > W_0 + 8), open
                    // for sure when V <= W_0 - 2
                    const int V = __double2hiint(D) + (int) mt.y;
                    if (STATS) nvis += 1u;
                    const bool guarded = GUARD && mt.y >= t_lim;
                    if (V > Whi && !guarded) {
                        force();          // next = skip link: points behind the node's subtree, always > cur
                    } else if (V < Wlo && !guarded) {
                        next = cur + 1;
                    } else {
                        // undecided, or a depth where eps2 matters: the oracle's exact expression, no contraction
                        const bool accept = exact_accept(dx, dy, dz, aabb + NB_ACCEPT_TABLE_OFFSET, mt.y >> NB_DEPTH_SHIFT);
                        if (accept) force();
                        else next = cur + 1;
                    }
                }
            }
            cur = __reduce_min_sync(0xffffffffu, next);
        }
        if (valid) {
            walk_epilogue<EPI>(out, b, G, ax, ay, az, px, py, pz);
            if (STATS) visits[b] = nvis;
        }
        if (STATS) { nvis_w += nvis; nacc_w += nacc; }
        if (!PERSIST) break;
    }
    };
    if (guard_needed) run(std::true_type{});
    else run(std::false_type{});
    if (STATS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            nvis_w += __shfl_xor_sync(0xffffffffu, nvis_w, o);
            nacc_w += __shfl_xor_sync(0xffffffffu, nacc_w, o);
        }
        if (lane == 0) { atomicAdd(&totals[0], nvis_w); atomicAdd(&totals[1], nacc_w); }
    }
}

// New slice bounds from the tile costs of the latest walk: rank r gets the tiles whose cumulative cost lies in
// [r, r + 1) / world of the total.  Every rank runs this on the same numbers (the walks stored each tile's cost into
// every rank's copy) and therefore obtains the same bounds.  Two small kernels: sums over groups of 256 tiles, then one
// block scans the group sums and one warp per cut finds its group and, inside the group, its tile.
#define NB_REB_GROUP 256
__global__ void __launch_bounds__(256)
rebalance_groups_kernel(const uint32_t *__restrict__ tile_cost, uint64_t n_tiles, unsigned long long *__restrict__ group_sum) {
    __shared__ unsigned long long ws[8];
    const uint64_t t = (uint64_t) blockIdx.x * NB_REB_GROUP + threadIdx.x;
    unsigned long long v = t < n_tiles ? tile_cost[t] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int w = 0; w < 8; ++w) s += ws[w];
        group_sum[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(1024)
rebalance_cuts_kernel(const uint32_t *__restrict__ tile_cost, uint64_t n_bodies, int world, unsigned long long *group_sum,
                      uint32_t n_groups, unsigned long long *__restrict__ bounds) {
    __shared__ unsigned long long carry;
    __shared__ unsigned long long part[1024];
    const uint64_t n_tiles = (n_bodies + 31) >> 5;
    // inclusive scan of the group sums in place, 1024 groups per round
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_groups; base += 1024) {
        const uint32_t g = base + threadIdx.x;
        part[threadIdx.x] = g < n_groups ? group_sum[g] : 0ull;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const unsigned long long v = threadIdx.x >= (unsigned) o ? part[threadIdx.x - o] : 0ull;
            __syncthreads();
            part[threadIdx.x] += v;
            __syncthreads();
        }
        if (g < n_groups) group_sum[g] = carry + part[threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 1023) carry += part[1023];
        __syncthreads();
    }
    const unsigned long long total = carry;
    if (threadIdx.x == 0) { bounds[0] = 0; bounds[world] = n_bodies; }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = warp + 1;                      // this warp's cut
    if (k >= world) return;
    if (total == 0) {                            // no costs recorded yet: equal tile counts
        if (lane == 0) bounds[k] = ((n_tiles * (uint64_t) k) / (uint64_t) world) << 5;
        return;
    }
    // the cut lies at cost c = total * k / world; it falls into the tile whose cost interval [before, before + cost)
    // contains c, and the slice boundary goes after that tile
    const unsigned long long c = (total * (unsigned long long) k) / (unsigned long long) world;
    uint32_t lo = 0, hi = n_groups - 1;          // first group whose inclusive sum exceeds c
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (group_sum[mid] > c) hi = mid; else lo = mid + 1;
    }
    unsigned long long before = lo ? group_sum[lo - 1] : 0ull;
    const uint64_t t0 = (uint64_t) lo * NB_REB_GROUP;
    uint64_t cut_tile = n_tiles - 1;
    for (int r = 0; r < NB_REB_GROUP / 32; ++r) {   // 32 tiles at a time: warp-inclusive scan
        const uint64_t t = t0 + (uint64_t) r * 32 + lane;
        unsigned long long v = t < n_tiles ? tile_cost[t] : 0ull, inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, before + inc > c);
        if (hit) { cut_tile = t0 + (uint64_t) r * 32 + (__ffs(hit) - 1); break; }
        before += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) {
        const uint64_t slot = (cut_tile + 1) << 5;
        bounds[k] = slot < n_bodies ? slot : n_bodies;
    }
}

}  // namespace

int nbk_bh_rebalance(nb_ctx *ctx) {
    const uint64_t n_tiles = (ctx->n + 31) / 32;
    const uint32_t n_groups = (uint32_t) ((n_tiles + NB_REB_GROUP - 1) / NB_REB_GROUP);
    unsigned long long *group_sum = ctx->dyn_bounds + NB_MAX_PEERS + 2;
    rebalance_groups_kernel<<<n_groups, 256, 0, ctx->stream>>>(ctx->tile_cost, n_tiles, group_sum);
    NB_LAUNCH_CHECK(ctx);
    rebalance_cuts_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->tile_cost, ctx->n, ctx->world, group_sum, n_groups, ctx->dyn_bounds);
    NB_LAUNCH_CHECK(ctx);
    ctx->bounds_valid = true;
    return NB_OK;
}

// Walk of the bodies in storage slots [s_begin, s_end) (storage order == sorted order after nb_bh_build) with one of the
// epilogues above; to_peers: store the results into every rank's arrays (requires mapped peer slabs).
int nbk_bh_accel_fused(nb_ctx *ctx, uint64_t s_begin, uint64_t s_end, int epilogue, double dt, bool to_peers, bool dynamic_slices) {
    nb_bh_state &b = ctx->bh;
    if (!b.built) return nb_fail(ctx, NB_ERR_INVALID, "nb_bh_accel: call nb_bh_build first");
    if (dynamic_slices && to_peers) { s_begin = 0; s_end = ctx->n; }   // the kernel reads its slice from dyn_bounds: size the grid for any
    if (s_end <= s_begin) return NB_OK;
    NB_CHECK(nbk_bh_accept_table(ctx));   // theta may have changed since the build
    if (to_peers && !ctx->p2p_ok) return nb_fail(ctx, NB_ERR_INVALID, "walk with peer stores: peer slabs are not mapped");
    int threads = ctx->cfg.wg_size_barnes_hut;  // --wg_size_barnes_hut -> CTA size (multiple of 32, <= 256)
    if (threads < 32) threads = 32;
    if (threads > 256) threads = 256;
    threads = (threads + 31) & ~31;
    const uint64_t count = s_end - s_begin;
    const unsigned grid = (unsigned) ((count + threads - 1) / threads);
    const double4 *com = reinterpret_cast<const double4 *>(b.com);
    nb_walk_out out;
    out.ax = ctx->ax; out.ay = ctx->ay; out.az = ctx->az;
    out.vx = ctx->vx; out.vy = ctx->vy; out.vz = ctx->vz;
    out.x = ctx->x; out.y = ctx->y; out.z = ctx->z;
    out.dt = dt;
    out.peers = nbk_peer_table(ctx);
    if (!to_peers) { out.peers.world = 1; out.peers.rank = 0; out.peers.base[0] = ctx->slab; }
    out.bounds = nullptr;
    out.tile_cost = out.tile_start = nullptr;
    out.rank = ctx->rank;
    if (dynamic_slices && to_peers) {
        if (!ctx->bounds_valid) return nb_fail(ctx, NB_ERR_INVALID, "walk with cost-weighted slices: no slice bounds yet");
        out.bounds = ctx->dyn_bounds;
        out.tile_cost = ctx->tile_cost;
        out.tile_start = ctx->tile_start;
    }
    // walk_variant (cfg.reserved[3]): 0 = SM-local tile queues from 2^19 bodies per call (below that the tail of the
    // persistent form costs more than its locality gains), the grid-mapped form otherwise; 20 / 50 force the
    // grid-mapped / persistent form (tests: the two forms are the same arithmetic per body).
#define NB_LAUNCH_IW(ST, PERSIST, RUN, EPI, GRID)                                                                       \
    bh_traverse_iw_kernel<ST, PERSIST, RUN, EPI><<<GRID, threads, 0, ctx->stream>>>(                                    \
        com, b.meta, b.dev_flags, ctx->n, b.aabb_dev, ctx->x, ctx->y, ctx->z, s_begin, s_end, ctx->cfg.theta,           \
        ctx->cfg.epsilon2, ctx->cfg.G, out, b.visits, b.stat_totals, b.dev_flags + 8,                                   \
        (uint32_t) std::min<int>(ctx->sm_count, 1024), 1.875)
#define NB_LAUNCH_EPI(PERSIST, RUN, GRID)                                                                               \
    do {                                                                                                                \
        if (epilogue == NB_EPI_ACCEL) NB_LAUNCH_IW(false, PERSIST, RUN, NB_EPI_ACCEL, GRID);                            \
        else if (epilogue == NB_EPI_KICK) NB_LAUNCH_IW(false, PERSIST, RUN, NB_EPI_KICK, GRID);                         \
        else NB_LAUNCH_IW(false, PERSIST, RUN, NB_EPI_KICK_DRIFT, GRID);                                                \
    } while (0)
    const int wv = ctx->cfg.reserved[3];
    if (b.stats_enabled) {
        if (epilogue != NB_EPI_ACCEL) return nb_fail(ctx, NB_ERR_INVALID, "the instrumented walk has no fused integrator");
        NB_CUDA(ctx, cudaMemsetAsync(b.stat_totals, 0, 8 * sizeof(unsigned long long), ctx->stream));
        NB_CUDA(ctx, cudaMemsetAsync(b.visits, 0, ctx->n * sizeof(uint32_t), ctx->stream));
        NB_LAUNCH_IW(true, false, 0, NB_EPI_ACCEL, grid);
    } else if (wv == 50 || dynamic_slices || (wv != 20 && count >= (1ull << 19))) {
        if (b.walk_ctas_threads != threads) {   // resident CTAs per SM for this CTA size (queried once)
            int per_sm = 0;
            NB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bh_traverse_iw_kernel<false, true, 160, NB_EPI_KICK_DRIFT>, threads, 0));
            b.walk_ctas_per_sm = per_sm < 1 ? 1 : per_sm;
            b.walk_ctas_threads = threads;
        }
        NB_CUDA(ctx, cudaMemsetAsync(b.dev_flags + 8, 0, 1024 * sizeof(uint32_t), ctx->stream));
        const unsigned pg = std::min<unsigned>(grid, (unsigned) (b.walk_ctas_per_sm * ctx->sm_count));
        // SM queues own interleaved runs of 160 tiles (round 1, N = 2^24: 40 / 80 / 160 / 320 tiles 84.4 / 84.2 / 84.2 /
        // 84.5 ms; one contiguous chunk per SM 85.2 ms: L1 hit rate up, but 148 distant windows at a time drop the L2
        // hit rate from 92 % to 71 %).  A rank's slice of an 8-GPU run (2^21 bodies) has a tail of ~7 %: a warp needs
        // ~1 ms per tile, so the last round runs on partly empty SMs
        NB_LAUNCH_EPI(true, 160, pg);
    } else {
        NB_LAUNCH_EPI(false, 0, grid);
    }
#undef NB_LAUNCH_EPI
#undef NB_LAUNCH_IW
    NB_LAUNCH_CHECK(ctx);
    return NB_OK;
}

int nbk_bh_accel(nb_ctx *ctx, uint64_t s_begin, uint64_t s_end) {
    return nbk_bh_accel_fused(ctx, s_begin, s_end, NB_EPI_ACCEL, 0.0, false);
}
