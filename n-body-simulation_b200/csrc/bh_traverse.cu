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
//   * acceptance test edge*rsqrt(d2) < theta is decided without memory access from the node's depth: the production
//     kernel (bh_traverse_iw_kernel) compares the high word of d2 + eps2 with hiword((edge_0/theta)^2) - (depth << 21)
//     on the integer pipe; the vanishing band around the threshold is re-evaluated with correctly rounded sqrt /
//     reciprocal / multiply, i.e. exactly the oracle's expression, so the interaction set is identical;
//   * production launches are persistent: warps draw 32-body tiles from SM-local queues (see the kernel's comment);
//   * force: MUFU.RSQ64H seed + cubic Taylor refinement of (d2+eps2)^(-3/2) (see naive.cu).
// Payload per visited node: 32 B {com xyz, mass} + 8 B {skip, leaf|body / depth} = 40 B.
#include "common.cuh"

#include <algorithm>
#include <type_traits>

#define NB_BH_MAX_LEVELS 64

namespace {

// scale a positive double by 4^-depth exactly (exponent arithmetic; thresholds are far from the subnormal range)
__device__ __forceinline__ double scale_pow4(double v, uint32_t depth) {
    return __hiloint2double(__double2hiint(v) - (int) (depth << 21), __double2loint(v));
}

__device__ __forceinline__ double scale_pow2(double v, uint32_t depth) {  // v * 2^-depth, exact
    return __hiloint2double(__double2hiint(v) - (int) (depth << 20), __double2loint(v));
}

// The earlier walk (walk_variant 5, kept for A/B runs): acceptance by two fp64 compares of d2 against per-depth
// thresholds ((edge/theta)^2 widened by 1e-12, scaled by 4^-depth with exponent arithmetic), grid-mapped tiles.
template <bool STATS>
__global__ void __launch_bounds__(256, 5)
bh_traverse_kernel(const double4 *__restrict__ com, const uint2 *__restrict__ meta, const uint32_t *__restrict__ flags,
                   uint64_t n_bodies, const double *__restrict__ aabb, const double *__restrict__ sx,
                   const double *__restrict__ sy, const double *__restrict__ sz, uint64_t s_begin, uint64_t s_end,
                   double theta, double eps2, double G, double *__restrict__ asx, double *__restrict__ asy,
                   double *__restrict__ asz, uint32_t *__restrict__ visits, unsigned long long *__restrict__ totals) {
    const double edge0 = aabb[6];
    // accept <=> d2 > (edge/theta)^2; a depth-d cell scales the root thresholds by 4^-d exactly (the edge is halved
    // exactly per level, ParallelOctreeTopDownSubtrees.cpp:256)
    const double ratio0 = (edge0 / theta) * (edge0 / theta);
    const bool scalable = ratio0 > 1e-200 && ratio0 < 1e200;  // theta == 0 or absurd boxes: always take the exact branch
    const double hi0 = scalable ? ratio0 * (1.0 + 1e-12) : __longlong_as_double(0x7ff0000000000000ll);
    const double lo0 = scalable ? ratio0 * (1.0 - 1e-12) : 0.0;
    const uint32_t n_nodes = (uint32_t) n_bodies + flags[1];
    const int lane = threadIdx.x & 31;
    const uint64_t warp_global = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t b = s_begin + warp_global * 32 + lane;
    const bool valid = b < s_end;
    const uint32_t me = (uint32_t) b;
    double px = 0, py = 0, pz = 0;
    if (valid) { px = sx[b]; py = sy[b]; pz = sz[b]; }
    double ax = 0, ay = 0, az = 0;
    // a failed build (depth / pool flag) leaves no valid tree: produce zeros instead of walking garbage
    uint32_t next = (valid && flags[0] == 0) ? 0u : 0xffffffffu;
    uint32_t nvis = 0, nacc = 0;

    uint32_t cur = __reduce_min_sync(0xffffffffu, next);
    while (cur < n_nodes) {
        if (next == cur) {
            const double4 c = com[cur];   // warp-uniform address: one broadcast transaction
            const uint2 mt = meta[cur];
            const double dx = c.x - px, dy = c.y - py, dz = c.z - pz;
            const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
            const bool massless = STATS && (__double2hiint(c.w) | __double2loint(c.w)) == 0;
            bool interact;
            if (mt.y & NB_LEAF_FLAG) {
                interact = (mt.y & NB_PAYLOAD_MASK) != me && !massless;  // own leaf skipped (:349)
                next = cur + 1;
                if (STATS) nvis += interact ? 1u : 0u;
            } else if (massless) {
                interact = false;
                next = max(mt.x, cur + 1);
            } else {
                const uint32_t depth = mt.y & NB_PAYLOAD_MASK;
                bool accept = d2 > scale_pow4(hi0, depth);
                if (!accept && !(d2 < scale_pow4(lo0, depth))) {
                    // borderline: the oracle's exact expression (BarnesHutAlgorithm.cpp:355-359), no contraction
                    const double d2o = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                    const double rs = __ddiv_rn(1.0, __dsqrt_rn(d2o));
                    accept = __dmul_rn(scale_pow2(edge0, depth), rs) < theta;
                }
                interact = accept;
                next = accept ? max(mt.x, cur + 1) : cur + 1;  // skip links always point forward
                if (STATS) nvis += 1u;
            }
            if (interact) {
                if (STATS) nacc += 1u;
                const double D = d2 + eps2;
                const double y0 = nb_rsqrt_seed(D);
                const double y2 = y0 * y0;
                const double e = fma(-D, y2, 1.0);
                const double y3 = y2 * y0;
                const double p = fma(1.875, e, 1.5);
                const double q = fma(p, e, 1.0);
                const double s = (y3 * c.w) * q;
                ax = fma(dx, s, ax);
                ay = fma(dy, s, ay);
                az = fma(dz, s, az);
            }
        }
        cur = __reduce_min_sync(0xffffffffu, next);
    }
    if (valid) {
        asx[b] = ax * G;
        asy[b] = ay * G;
        asz[b] = az * G;
        if (STATS) visits[b] = nvis;
    }
    if (STATS) {
        unsigned long long v = nvis, a = nacc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, o);
            a += __shfl_xor_sync(0xffffffffu, a, o);
        }
        if (lane == 0) { atomicAdd(&totals[0], v); atomicAdd(&totals[1], a); }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Production walk: acceptance test on the integer pipe, SM-local tile queues.
//
// Same traversal and the same interaction sets as bh_traverse_kernel above (kept as walk_variant 5 for A/B runs).
//   * D = dx^2 + dy^2 + dz^2 + eps2 is ONE fma chain (eps2 folded into the first term) and serves both the acceptance
//     test and the force;
//   * positive doubles order like their bit patterns, so "D > (edge_d/theta)^2" is decided by comparing HIGH WORDS:
//     W_d = hiword((edge_0/theta)^2) - (depth << 21)  (the edge halves exactly per level); accept when
//     hiword(D) >= W_d + 2, open when hiword(D) <= W_d - 2 (evaluated as V = hiword(D) + (depth << 21) against W_0 +- 1).  The band in between (relative width 2^-19) and every
//     depth for which eps2 is not negligible against (edge_d/theta)^2 take the oracle's exact expression
//     (BarnesHutAlgorithm.cpp:355-359) with correctly rounded operations, so the decision is the reference's.  No
//     per-depth fp64 thresholds, nothing spills under the 48-register budget (40 warps per SM);
//   * PERSIST: the grid only fills the machine and every WARP draws 32-body tiles from a queue that belongs to the SM
//     it runs on (queue c owns runs of RUN consecutive tiles of the sorted order, interleaved with the other queues'
//     runs; a warp whose queue is empty steals from the following queues).  The ~40 warps resident on an SM therefore
//     walk neighbouring bodies at the same time and share the node records they pull into that SM's L1 -- the hardware
//     block scheduler would hand an SM blocks that are 148 blocks apart -- while all SMs stay inside one moving window
//     of the tree (L2), and no CTA is ever re-launched.  ncu at N = 2^24: L1 hit rate 65 % -> 73 %, warps active
//     57 % -> 62 %, 89.6 ms -> 84.2 ms.  tile_counters: one uint32 per queue, zeroed before the launch.
//     The queue bookkeeping must not add live values to the cursor loop: with RUN as a run-time argument ptxas spilled
//     two loop invariants and the walk fell back to 90 ms, hence the template constant.
//   * the cursor step is issue bound, so every instruction taken out of it shows: 1.875 of the Taylor term arrives as a
//     kernel parameter (one LDC instead of two moves that ptxas re-materialised per node under the register budget), the
//     acceptance test is ONE multiply-add V = hiword(D) + (depth << 21) and two compares against W_0 + 1 and W_0 - 1
//     (kept opaque so they are not re-derived per node), the skip link is used as it is (it always points behind the
//     node's subtree, no max), and the depth guard t >= t_lim is compiled out of the loop when the deepest level of
//     THIS tree (flags[2]) cannot reach it.  56 -> 45 SASS instructions per accepted cell, 84.2 -> 77.2 ms at N = 2^24,
//     bit-identical results.
// Explored on top of this and rejected (N = 2^24, theta = 0.5, all parity-green): issuing the next node's loads before
// the force arithmetic (software pipelining: 95 ms, the 12 extra live registers cost 20 % of the resident warps);
// prefetch.global.L1 of the next node (98-102 ms); ticketed one-tile-per-warp assignment on a full grid (87 ms); one
// straight-line accept/skip decision for leaves and cells instead of the leaf branch (88.9 ms).
// ---------------------------------------------------------------------------------------------------------------------
// the oracle's acceptance expression (BarnesHutAlgorithm.cpp:355-359) with correctly rounded operations
__device__ __forceinline__ bool exact_accept(double dx, double dy, double dz, double edge0, uint32_t depth, double theta) {
    const double d2o = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const double rs = __ddiv_rn(1.0, __dsqrt_rn(d2o));
    return __dmul_rn(scale_pow2(edge0, depth), rs) < theta;
}

template <bool STATS, bool PERSIST, uint32_t RUN>
__global__ void __launch_bounds__(256, 5)
bh_traverse_iw_kernel(const double4 *__restrict__ com, const uint2 *__restrict__ meta, const uint32_t *__restrict__ flags,
                      uint64_t n_bodies, const double *__restrict__ aabb, const double *__restrict__ sx,
                      const double *__restrict__ sy, const double *__restrict__ sz, uint64_t s_begin, uint64_t s_end,
                      double theta, double eps2, double G, double *__restrict__ asx, double *__restrict__ asy,
                      double *__restrict__ asz, uint32_t *__restrict__ visits, unsigned long long *__restrict__ totals,
                      uint32_t *__restrict__ tile_counters, uint32_t n_chunks, double k1875) {
    const double edge0 = aabb[6];
    const double ratio0 = (edge0 / theta) * (edge0 / theta);
    const bool scalable = ratio0 > 1e-200 && ratio0 < 1e200;   // theta == 0 or absurd boxes: always the exact branch
    const int W0 = __double2hiint(ratio0);
    int Whi = W0 + 1, Wlo = W0 - 1;
    asm("" : "+r"(Whi), "+r"(Wlo));   // opaque: two plain compare operands, not W_0 plus an add per node
    // fast decisions need eps2 <= 2^-24 (edge_d/theta)^2 and a normal threshold: depth << 21 must stay below t_lim
    int w_min = __double2hiint(eps2) + (24 << 20);
    if (w_min < (64 << 20)) w_min = 64 << 20;
    const uint32_t t_lim = (scalable && W0 > w_min) ? (uint32_t) (W0 - w_min) : 0u;
    const uint32_t n_nodes = (uint32_t) n_bodies + flags[1];
    // a failed build (depth / pool flag) leaves no valid tree: produce zeros instead of walking garbage
    const bool tree_ok = flags[0] == 0;
    const int lane = threadIdx.x & 31;
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
        const uint32_t me = (uint32_t) b;
        double px = 0, py = 0, pz = 0;
        if (valid) { px = sx[b]; py = sy[b]; pz = sz[b]; }
        double ax = 0, ay = 0, az = 0;
        uint32_t next = (valid && tree_ok) ? 0u : 0xffffffffu;
        uint32_t nvis = 0, nacc = 0;

        uint32_t cur = __reduce_min_sync(0xffffffffu, next);
        while (cur < n_nodes) {
            if (next == cur) {
                const double4 c = com[cur];   // warp-uniform address: one broadcast transaction
                const uint2 mt = meta[cur];
                const double dx = c.x - px, dy = c.y - py, dz = c.z - pz;
                const double D = fma(dz, dz, fma(dy, dy, fma(dx, dx, eps2)));
                // SUM_MASSES == 0 nodes are invisible in the reference (BarnesHutAlgorithm.cpp:349): massless bodies,
                // and cells that hold only massless bodies, are neither counted nor opened.  Their contribution is
                // exactly 0.0 either way (the build stores a finite record for them), so only the instrumented build
                // pays for the check that keeps the visit counts identical to the reference's.
                const bool massless = STATS && (__double2hiint(c.w) | __double2loint(c.w)) == 0;
                bool interact;
                if (mt.y & NB_LEAF_FLAG) {
                    interact = (mt.y & NB_PAYLOAD_MASK) != me && !massless;  // own leaf skipped (:349)
                    next = cur + 1;
                    if (STATS) nvis += interact ? 1u : 0u;
                } else if (massless) {
                    interact = false;
                    next = max(mt.x, cur + 1);
                } else {
                    const uint32_t t = mt.y << 21;                       // depth << 21 (the rank bits shift out)
                    const int V = __double2hiint(D) + (int) t;           // one multiply-add; compare with W_0 -+ 1
                    bool accept = V > Whi;                               // hiword(D) - W_d >= 2
                    if ((GUARD && t >= t_lim) || (!accept && V >= Wlo)) {
                        // undecided (|hiword(D) - W_d| <= 1, or a depth where eps2 matters): the oracle's exact
                        // expression, no contraction
                        accept = exact_accept(dx, dy, dz, edge0, mt.y & NB_PAYLOAD_MASK, theta);
                    }
                    interact = accept;
                    next = accept ? mt.x : cur + 1;  // a skip link points behind the node's subtree: always > cur
                    if (STATS) nvis += 1u;
                }
                if (interact) {
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
                }
            }
            cur = __reduce_min_sync(0xffffffffu, next);
        }
        if (valid) {
            asx[b] = ax * G;
            asy[b] = ay * G;
            asz[b] = az * G;
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

// ---------------------------------------------------------------------------------------------------------------------
// Group traversal with exact per-body acceptance (optional variant, bh_variant = 3).
//
// A warp owns 32 consecutive bodies of the sorted order and their exact bounding box.  Work items are {node, lane
// mask} pairs ("this node must be looked at by these bodies") on a per-warp LIFO in shared memory.  Each round pops up
// to 32 items and classifies them NODE-PARALLEL (one item per lane) against the group box:
//     nearest point of the box farther than edge/theta  -> every masked body accepts   -> interaction list
//     farthest point of the box nearer than edge/theta  -> every masked body opens     -> children pushed (same mask)
//     body leaf                                         -> interaction list (mask minus the body itself)
//     otherwise ("mixed")                               -> the 32 bodies run the reference's per-body test on that node;
//                                                          accepting lanes -> interaction list, opening lanes -> children
// Because every body lies inside the box and the two group thresholds are widened by 1e-9 (>> rounding), a group
// decision is exactly the decision each masked body's own test (BarnesHutAlgorithm.cpp:355-359) would take, so every
// body interacts with exactly the reference's node set; only the summation order differs (~1e-16 relative).
// The interaction list {node, mask} is evaluated four entries at a time so independent rsqrt chains overlap.
// Stack bound: wide pops are only taken while 7*pops fit under CAP-301; otherwise one item is popped per round, which
// is a plain DFS needing <= 7 slots per level (<= 294 for 42 levels), so the LIFO can never overflow.
// ---------------------------------------------------------------------------------------------------------------------
#define NB_G_WARPS 4
#define NB_G_STACK 1024
#define NB_G_RESERVE 301
#define NB_G_ILIST 128

template <bool STATS>
__global__ void __launch_bounds__(NB_G_WARPS * 32)
bh_traverse3_kernel(const double4 *__restrict__ com, const uint2 *__restrict__ meta, const uint32_t *__restrict__ ctab,
                    const uint32_t *__restrict__ flags, uint64_t n_bodies, const double *__restrict__ aabb,
                    const double *__restrict__ sx, const double *__restrict__ sy, const double *__restrict__ sz,
                    uint64_t s_begin, uint64_t s_end, double theta, double eps2, double G, double *__restrict__ asx,
                    double *__restrict__ asy, double *__restrict__ asz, uint32_t *__restrict__ visits,
                    unsigned long long *__restrict__ totals) {
    __shared__ double t_hi[NB_BH_MAX_LEVELS], t_lo[NB_BH_MAX_LEVELS], t_edge[NB_BH_MAX_LEVELS];
    __shared__ double g_hi[NB_BH_MAX_LEVELS], g_lo[NB_BH_MAX_LEVELS];
    __shared__ uint2 s_stack[NB_G_WARPS][NB_G_STACK];
    __shared__ uint2 s_ilist[NB_G_WARPS][NB_G_ILIST];
    for (int t = threadIdx.x; t < NB_BH_MAX_LEVELS; t += blockDim.x) {
        const double e = ldexp(aabb[6], -t);
        const double ratio = (e / theta) * (e / theta);
        t_edge[t] = e;
        t_hi[t] = ratio * (1.0 + 1e-12);
        t_lo[t] = ratio * (1.0 - 1e-12);
        g_hi[t] = ratio * (1.0 + 1e-9);
        g_lo[t] = ratio * (1.0 - 1e-9);
    }
    __syncthreads();
    const uint32_t n_nodes = (uint32_t) n_bodies + flags[1];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const uint64_t warp_global = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t wbase = s_begin + warp_global * 32;
    const uint64_t b = wbase + lane;
    const bool valid = b < s_end;
    const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
    if (vmask == 0 || flags[0] != 0 || n_nodes == 0) {
        if (valid) { asx[b] = 0; asy[b] = 0; asz[b] = 0; if (STATS) visits[b] = 0; }
        return;
    }
    const uint64_t bsafe = valid ? b : wbase;  // lane 0 is always valid here
    const double px = sx[bsafe], py = sy[bsafe], pz = sz[bsafe];
    // exact bounding box of the group
    double lox = px, loy = py, loz = pz, hix = px, hiy = py, hiz = pz;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lox = fmin(lox, __shfl_xor_sync(0xffffffffu, lox, o)); hix = fmax(hix, __shfl_xor_sync(0xffffffffu, hix, o));
        loy = fmin(loy, __shfl_xor_sync(0xffffffffu, loy, o)); hiy = fmax(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
        loz = fmin(loz, __shfl_xor_sync(0xffffffffu, loz, o)); hiz = fmax(hiz, __shfl_xor_sync(0xffffffffu, hiz, o));
    }
    uint2 *stack = s_stack[wib];
    uint2 *ilist = s_ilist[wib];
    double ax = 0, ay = 0, az = 0;
    uint32_t nvis = 0, nacc = 0;
    uint32_t size = 0, cnt = 0;
    if (lane == 0) stack[0] = make_uint2(0u, vmask);
    size = 1;
    __syncwarp();

    auto evaluate = [&](uint32_t count) {
        for (uint32_t k = 0; k < count; k += 4) {
            double4 r[4];
            bool on[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint2 en = ilist[k + u < count ? k + u : k];
                r[u] = com[en.x];
                on[u] = k + u < count && ((en.y >> lane) & 1u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double dx = r[u].x - px, dy = r[u].y - py, dz = r[u].z - pz;
                const double D = fma(dz, dz, fma(dy, dy, fma(dx, dx, eps2)));
                const double y0 = nb_rsqrt_seed(D);
                const double y2 = y0 * y0;
                const double e = fma(-D, y2, 1.0);
                const double y3 = y2 * y0;
                const double q = fma(fma(1.875, e, 1.5), e, 1.0);
                const double sfac = (y3 * r[u].w) * q;
                if (on[u]) {
                    ax = fma(dx, sfac, ax);
                    ay = fma(dy, sfac, ay);
                    az = fma(dz, sfac, az);
                }
            }
        }
    };

    while (size > 0) {
        if (cnt > NB_G_ILIST - 64) {
            __syncwarp();
            evaluate(cnt);
            __syncwarp();
            cnt = 0;
        }
        // ---- pop: wide while the children are guaranteed to fit under the DFS reserve, else one item (plain DFS)
        uint32_t k = size < 32u ? size : 32u;
        const uint32_t wide_room = size + NB_G_RESERVE < NB_G_STACK ? (NB_G_STACK - NB_G_RESERVE - size) / 7u : 0u;
        if (k > wide_room) k = wide_room > 0 ? wide_room : 1u;
        const bool have = (uint32_t) lane < k;
        uint2 item = make_uint2(0u, 0u);
        if (have) item = stack[size - 1 - lane];
        size -= k;
        __syncwarp();
        // ---- node-parallel classification: 0 none, 1 accept (imask), 2 open, 3 mixed
        int cls = 0;
        uint32_t imask = 0, depth = 0;
        uint4 kid_lo = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu), kid_hi = kid_lo;
        if (have) {
            // independent loads issued together: one memory latency per round
            const uint2 mt = meta[item.x];
            const double4 c = com[item.x];
            kid_lo = reinterpret_cast<const uint4 *>(ctab)[2 * (size_t) item.x];
            kid_hi = reinterpret_cast<const uint4 *>(ctab)[2 * (size_t) item.x + 1];
            const bool massless = STATS && (__double2hiint(c.w) | __double2loint(c.w)) == 0;  // invisible in the reference (:349)
            if (massless) {
                cls = 0;
            } else if (mt.y & NB_LEAF_FLAG) {
                const uint64_t sidx = mt.y & NB_PAYLOAD_MASK;  // the leaf's own body never interacts with itself
                imask = item.y;
                if (sidx >= wbase && sidx < wbase + 32) imask &= ~(1u << (uint32_t) (sidx - wbase));
                cls = imask ? 1 : 0;
            } else {
                depth = mt.y & NB_PAYLOAD_MASK;
                const double ex = fmax(0.0, fmax(lox - c.x, c.x - hix));
                const double ey = fmax(0.0, fmax(loy - c.y, c.y - hiy));
                const double ez = fmax(0.0, fmax(loz - c.z, c.z - hiz));
                const double fx = fmax(c.x - lox, hix - c.x);
                const double fy = fmax(c.y - loy, hiy - c.y);
                const double fz = fmax(c.z - loz, hiz - c.z);
                const double dn2 = fma(ez, ez, fma(ey, ey, ex * ex));
                const double df2 = fma(fz, fz, fma(fy, fy, fx * fx));
                imask = item.y;
                cls = dn2 > g_hi[depth] ? 1 : (df2 < g_lo[depth] ? 2 : 3);
            }
        }
        if (STATS) {
            for (uint32_t j = 0; j < k; ++j) {
                const uint32_t vm = __shfl_sync(0xffffffffu, imask, j);
                const int cj = __shfl_sync(0xffffffffu, cls, j);
                nvis += (vm >> lane) & 1u;
                if (cj == 1) nacc += (vm >> lane) & 1u;
            }
        }
        if (STATS && lane == 0) {
            atomicAdd(&totals[2], 1ull);
            atomicAdd(&totals[3], (unsigned long long) k);
            atomicAdd(&totals[4], (unsigned long long) __popc(__ballot_sync(0xffffffffu, cls == 3) ));
        } else if (STATS) { __ballot_sync(0xffffffffu, cls == 3); }
        // ---- accepted by everybody: append to the interaction list
        {
            const uint32_t am = __ballot_sync(0xffffffffu, cls == 1);
            if (STATS && lane == 0) atomicAdd(&totals[5], (unsigned long long) __popc(am));
            if (cls == 1) ilist[cnt + __popc(am & lt)] = make_uint2(item.x, imask);
            cnt += __popc(am);
        }
        // ---- opened by everybody: push the children with the same mask
        {
            uint32_t kids[8];
            uint32_t nk = 0;
            if (cls == 2) {
                const uint4 c0 = kid_lo, c1 = kid_hi;
                kids[0] = c0.x; kids[1] = c0.y; kids[2] = c0.z; kids[3] = c0.w;
                kids[4] = c1.x; kids[5] = c1.y; kids[6] = c1.z; kids[7] = c1.w;
#pragma unroll
                for (int r = 0; r < 8; ++r) nk += kids[r] != 0xffffffffu ? 1u : 0u;
            }
            // exclusive prefix of nk over the warp
            uint32_t inc = nk;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
            if (cls == 2) {
                uint32_t pos = size + inc - nk;
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (kids[r] != 0xffffffffu) stack[pos++] = make_uint2(kids[r], item.y);
            }
            size += total;
        }
        // ---- mixed: the bodies apply the reference's own test to the node, one node at a time
        uint32_t mixed = __ballot_sync(0xffffffffu, cls == 3);
        while (mixed) {
            // two mixed nodes per trip so their loads and tests overlap
            const int s0 = __ffs(mixed) - 1;
            mixed &= mixed - 1;
            const bool two = mixed != 0;
            const int s1 = two ? __ffs(mixed) - 1 : s0;
            if (two) mixed &= mixed - 1;
            const uint32_t node0 = __shfl_sync(0xffffffffu, item.x, s0), node1 = __shfl_sync(0xffffffffu, item.x, s1);
            const uint32_t msk0 = __shfl_sync(0xffffffffu, item.y, s0);
            const uint32_t msk1 = two ? __shfl_sync(0xffffffffu, item.y, s1) : 0u;
            const uint32_t dep0 = __shfl_sync(0xffffffffu, depth, s0), dep1 = __shfl_sync(0xffffffffu, depth, s1);
            const double4 c0 = com[node0], c1 = com[node1];
            uint32_t kid = 0xffffffffu;
            if (lane < 16) kid = ctab[8 * (size_t) (lane < 8 ? node0 : node1) + (lane & 7)];
            const double dx0 = c0.x - px, dy0 = c0.y - py, dz0 = c0.z - pz;
            const double dx1 = c1.x - px, dy1 = c1.y - py, dz1 = c1.z - pz;
            const double d20 = fma(dz0, dz0, fma(dy0, dy0, dx0 * dx0));
            const double d21 = fma(dz1, dz1, fma(dy1, dy1, dx1 * dx1));
            bool acc0 = d20 > t_hi[dep0], acc1 = d21 > t_hi[dep1];
            if (!acc0 && !(d20 < t_lo[dep0])) {
                // borderline: the oracle's exact expression (BarnesHutAlgorithm.cpp:355-359), no contraction
                const double d2o = __dadd_rn(__dadd_rn(__dmul_rn(dx0, dx0), __dmul_rn(dy0, dy0)), __dmul_rn(dz0, dz0));
                acc0 = __dmul_rn(t_edge[dep0], __ddiv_rn(1.0, __dsqrt_rn(d2o))) < theta;
            }
            if (!acc1 && !(d21 < t_lo[dep1])) {
                const double d2o = __dadd_rn(__dadd_rn(__dmul_rn(dx1, dx1), __dmul_rn(dy1, dy1)), __dmul_rn(dz1, dz1));
                acc1 = __dmul_rn(t_edge[dep1], __ddiv_rn(1.0, __dsqrt_rn(d2o))) < theta;
            }
            const uint32_t am0 = __ballot_sync(0xffffffffu, ((msk0 >> lane) & 1u) && acc0);
            const uint32_t am1 = __ballot_sync(0xffffffffu, ((msk1 >> lane) & 1u) && acc1);
            const uint32_t om0 = msk0 & ~am0, om1 = msk1 & ~am1;
            if (STATS) {
                nacc += ((am0 >> lane) & 1u) + ((am1 >> lane) & 1u);
                if (lane == 0) atomicAdd(&totals[5], (unsigned long long) ((am0 != 0) + (am1 != 0)));
            }
            if (lane == 0) {
                if (am0) ilist[cnt] = make_uint2(node0, am0);
                if (am1) ilist[cnt + (am0 ? 1u : 0u)] = make_uint2(node1, am1);
            }
            cnt += (am0 ? 1u : 0u) + (am1 ? 1u : 0u);
            const uint32_t my_om = lane < 8 ? om0 : om1;
            const bool push = lane < 16 && kid != 0xffffffffu && my_om != 0;
            const uint32_t km = __ballot_sync(0xffffffffu, push);
            if (push) stack[size + __popc(km & lt)] = make_uint2(kid, my_om);
            size += __popc(km);
        }
        __syncwarp();
    }
    __syncwarp();
    evaluate(cnt);
    if (valid) {
        asx[b] = ax * G;
        asy[b] = ay * G;
        asz[b] = az * G;
        if (STATS) visits[b] = nvis;
    }
    if (STATS) {
        unsigned long long v = valid ? nvis : 0u, a = valid ? nacc : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, o);
            a += __shfl_xor_sync(0xffffffffu, a, o);
        }
        if (lane == 0) { atomicAdd(&totals[0], v); atomicAdd(&totals[1], a); }
    }
}

}  // namespace

// Accelerations of the bodies in storage slots [s_begin, s_end) (storage order == sorted order after nb_bh_build).
int nbk_bh_accel(nb_ctx *ctx, uint64_t s_begin, uint64_t s_end) {
    nb_bh_state &b = ctx->bh;
    if (!b.built) return nb_fail(ctx, NB_ERR_INVALID, "nb_bh_accel: call nb_bh_build first");
    if (s_end <= s_begin) return NB_OK;
    int threads = ctx->cfg.wg_size_barnes_hut;  // --wg_size_barnes_hut -> CTA size (multiple of 32, <= 256)
    if (threads < 32) threads = 32;
    if (threads > 256) threads = 256;
    threads = (threads + 31) & ~31;
    const uint64_t count = s_end - s_begin;
    const unsigned grid = (unsigned) ((count + threads - 1) / threads);
    const double4 *com = reinterpret_cast<const double4 *>(b.com);
    // reserved[1]: 0 = warp walk (default), 3 = group traversal (kept for A/B testing; measured slower on B200 at
    // N = 2^24, theta = 0.5: 91 ms walk vs 112 ms group)
    if (ctx->cfg.reserved[1] == 3) {
        if (!b.ctab_valid) return nb_fail(ctx, NB_ERR_INVALID, "group traversal needs a tree built with bh_variant = 3");
        const unsigned g3 = (unsigned) ((count + NB_G_WARPS * 32 - 1) / (NB_G_WARPS * 32));
        if (b.stats_enabled) {
            NB_CUDA(ctx, cudaMemsetAsync(b.stat_totals, 0, 8 * sizeof(unsigned long long), ctx->stream));
            NB_CUDA(ctx, cudaMemsetAsync(b.visits, 0, ctx->n * sizeof(uint32_t), ctx->stream));
            bh_traverse3_kernel<true><<<g3, NB_G_WARPS * 32, 0, ctx->stream>>>(com, b.meta, b.ctab, b.dev_flags, ctx->n, b.aabb_dev,
                                                                             ctx->x, ctx->y, ctx->z, s_begin, s_end, ctx->cfg.theta,
                                                                             ctx->cfg.epsilon2, ctx->cfg.G, ctx->ax, ctx->ay, ctx->az,
                                                                             b.visits, b.stat_totals);
        } else {
            bh_traverse3_kernel<false><<<g3, NB_G_WARPS * 32, 0, ctx->stream>>>(com, b.meta, b.ctab, b.dev_flags, ctx->n, b.aabb_dev,
                                                                              ctx->x, ctx->y, ctx->z, s_begin, s_end, ctx->cfg.theta,
                                                                              ctx->cfg.epsilon2, ctx->cfg.G, ctx->ax, ctx->ay, ctx->az,
                                                                              b.visits, b.stat_totals);
        }
        NB_LAUNCH_CHECK(ctx);
        return NB_OK;
    }
#define NB_LAUNCH_WALK(ST)                                                                                              \
    bh_traverse_kernel<ST><<<grid, threads, 0, ctx->stream>>>(com, b.meta, b.dev_flags, ctx->n, b.aabb_dev, ctx->x, ctx->y, \
                                                                  ctx->z, s_begin, s_end, ctx->cfg.theta, ctx->cfg.epsilon2,  \
                                                                  ctx->cfg.G, ctx->ax, ctx->ay, ctx->az, b.visits, b.stat_totals)
    // walk_variant (cfg.reserved[3]): 0 = production walk (integer-pipe acceptance test; SM-local tile queues from 2^19
    // bodies per call, below that the tail of the persistent form costs more than its locality gains); 20 / 50 force
    // the grid-mapped / persistent form; 5 = the earlier fp64-threshold walk, kept for A/B runs.
#define NB_LAUNCH_IW(ST, PERSIST, RUN, GRID)                                                                            \
    bh_traverse_iw_kernel<ST, PERSIST, RUN><<<GRID, threads, 0, ctx->stream>>>(                                         \
        com, b.meta, b.dev_flags, ctx->n, b.aabb_dev, ctx->x, ctx->y, ctx->z, s_begin, s_end, ctx->cfg.theta,           \
        ctx->cfg.epsilon2, ctx->cfg.G, ctx->ax, ctx->ay, ctx->az, b.visits, b.stat_totals, b.dev_flags + 8,             \
        (uint32_t) std::min<int>(ctx->sm_count, 1024), 1.875)
    const int wv = ctx->cfg.reserved[3];
    if (b.stats_enabled) {
        NB_CUDA(ctx, cudaMemsetAsync(b.stat_totals, 0, 8 * sizeof(unsigned long long), ctx->stream));
        NB_CUDA(ctx, cudaMemsetAsync(b.visits, 0, ctx->n * sizeof(uint32_t), ctx->stream));
        if (wv == 5) NB_LAUNCH_WALK(true);
        else NB_LAUNCH_IW(true, false, 0, grid);
    } else if (wv == 5) {
        NB_LAUNCH_WALK(false);
    } else if (wv == 50 || (wv != 20 && count >= (1ull << 19))) {
        if (b.walk_ctas_threads != threads) {   // resident CTAs per SM for this CTA size (queried once)
            int per_sm = 0;
            NB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bh_traverse_iw_kernel<false, true, 0>, threads, 0));
            b.walk_ctas_per_sm = per_sm < 1 ? 1 : per_sm;
            b.walk_ctas_threads = threads;
        }
        NB_CUDA(ctx, cudaMemsetAsync(b.dev_flags + 8, 0, 1024 * sizeof(uint32_t), ctx->stream));
        const unsigned pg = std::min<unsigned>(grid, (unsigned) (b.walk_ctas_per_sm * ctx->sm_count));
        // cfg.reserved[5] (walk_run_len): 0 = interleaved runs of 160 tiles per SM queue (84.2 ms at N = 2^24; 40 / 80 /
        // 320 tiles measured 84.4 / 84.2 / 84.5), 1 = one contiguous chunk per SM (85.2 ms: L1 hit rate up, but 148
        // distant windows at a time drop the L2 hit rate from 92 % to 71 %).  A rank's slice of an 8-GPU run (2^21 bodies,
        // nb_bh_accel_range on one GPU) takes 11.25 ms for runs of 20 / 40 / 80 / 160 tiles alike, against 84.2 / 8 = 10.5:
        // the 7 % are the tail -- a warp needs ~1 ms per tile, so the last round runs on partly empty SMs (smaller grids
        // that make the rounds come out even are slower: 11.5 ms at exactly 12 rounds)
        if (ctx->cfg.reserved[5] == 1) NB_LAUNCH_IW(false, true, 0, pg);
        else NB_LAUNCH_IW(false, true, 160, pg);
    } else {
        NB_LAUNCH_IW(false, false, 0, grid);
    }
#undef NB_LAUNCH_IW
#undef NB_LAUNCH_WALK
    NB_LAUNCH_CHECK(ctx);
    return NB_OK;
}
