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
//   * acceptance test edge*rsqrt(d2) < theta is decided by two compares of d2 against per-depth thresholds
//     (edge^2/theta^2 widened by 1e-12); the vanishing band in between is re-evaluated with correctly rounded
//     sqrt / reciprocal / multiply, i.e. exactly the oracle's expression, so the interaction set is identical;
//   * force: MUFU.RSQ64H seed + cubic Taylor refinement of (d2+eps2)^(-3/2) (see naive.cu).
// Payload per visited node: 32 B {com xyz, mass} + 8 B {skip, leaf|body / depth} = 40 B.
#include "common.cuh"

#define NB_BH_MAX_LEVELS 64

namespace {

template <bool STATS>
__global__ void __launch_bounds__(256)
bh_traverse_kernel(const double4 *__restrict__ com, const uint2 *__restrict__ meta, const uint32_t *__restrict__ flags,
                   uint64_t n_bodies, const double *__restrict__ aabb, const double *__restrict__ sx,
                   const double *__restrict__ sy, const double *__restrict__ sz, uint64_t s_begin, uint64_t s_end,
                   double theta, double eps2, double G, double *__restrict__ asx, double *__restrict__ asy,
                   double *__restrict__ asz, uint32_t *__restrict__ visits, unsigned long long *__restrict__ totals) {
    __shared__ double t_hi[NB_BH_MAX_LEVELS], t_lo[NB_BH_MAX_LEVELS], t_edge[NB_BH_MAX_LEVELS];
    for (int t = threadIdx.x; t < NB_BH_MAX_LEVELS; t += blockDim.x) {
        // edge of a depth-d cell: the root edge halved d times (exact), ParallelOctreeTopDownSubtrees.cpp:256
        const double e = ldexp(aabb[6], -t);
        const double ratio = (e / theta) * (e / theta);  // accept  <=>  d2 > (edge/theta)^2  (exact arithmetic)
        t_edge[t] = e;
        t_hi[t] = ratio * (1.0 + 1e-12);
        t_lo[t] = ratio * (1.0 - 1e-12);
    }
    __syncthreads();
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

    while (true) {
        const uint32_t cur = __reduce_min_sync(0xffffffffu, next);
        if (cur >= n_nodes) break;
        const double4 c = com[cur];   // warp-uniform address: one broadcast transaction
        const uint2 mt = meta[cur];
        if (next == cur) {
            const double dx = c.x - px, dy = c.y - py, dz = c.z - pz;
            const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
            bool interact;
            if (mt.y & NB_LEAF_FLAG) {
                interact = (mt.y & ~NB_LEAF_FLAG) != me;  // own leaf skipped (BarnesHutAlgorithm.cpp:349)
                next = cur + 1;
                if (STATS) nvis += interact ? 1u : 0u;
            } else {
                const uint32_t depth = mt.y;
                bool accept = d2 > t_hi[depth];
                if (!accept && !(d2 < t_lo[depth])) {
                    // borderline: the oracle's exact expression (BarnesHutAlgorithm.cpp:355-359), no contraction
                    const double d2o = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                    const double rs = __ddiv_rn(1.0, __dsqrt_rn(d2o));
                    accept = __dmul_rn(t_edge[depth], rs) < theta;
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

__global__ void __launch_bounds__(256)
scatter_accel_kernel(uint64_t n, const uint32_t *__restrict__ perm, const double *__restrict__ asx,
                     const double *__restrict__ asy, const double *__restrict__ asz, double *__restrict__ ax,
                     double *__restrict__ ay, double *__restrict__ az) {
    const uint64_t s = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t b = perm[s];
    ax[b] = asx[s]; ay[b] = asy[s]; az[b] = asz[s];
}

}  // namespace

// Accelerations of the sorted bodies [s_begin, s_end) into the sorted-order arrays asx/asy/asz.
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
    if (b.stats_enabled) {
        NB_CUDA(ctx, cudaMemsetAsync(b.stat_totals, 0, 2 * sizeof(unsigned long long), ctx->stream));
        NB_CUDA(ctx, cudaMemsetAsync(b.visits, 0, ctx->n * sizeof(uint32_t), ctx->stream));
        bh_traverse_kernel<true><<<grid, threads, 0, ctx->stream>>>(com, b.meta, b.dev_flags, ctx->n, b.aabb_dev, b.sx,
                                                                    b.sy, b.sz, s_begin, s_end, ctx->cfg.theta,
                                                                    ctx->cfg.epsilon2, ctx->cfg.G, b.asx, b.asy, b.asz,
                                                                    b.visits, b.stat_totals);
    } else {
        bh_traverse_kernel<false><<<grid, threads, 0, ctx->stream>>>(com, b.meta, b.dev_flags, ctx->n, b.aabb_dev, b.sx,
                                                                     b.sy, b.sz, s_begin, s_end, ctx->cfg.theta,
                                                                     ctx->cfg.epsilon2, ctx->cfg.G, b.asx, b.asy, b.asz,
                                                                     b.visits, b.stat_totals);
    }
    NB_LAUNCH_CHECK(ctx);
    return NB_OK;
}

// sorted order -> body-id order (ACC_X[i] = ..., BarnesHutAlgorithm.cpp:389-391)
int nbk_bh_scatter_accel(nb_ctx *ctx) {
    nb_bh_state &b = ctx->bh;
    const unsigned grid = (unsigned) ((ctx->n + 255) / 256);
    scatter_accel_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->n, b.perm, b.asx, b.asy, b.asz, ctx->ax, ctx->ay, ctx->az);
    NB_LAUNCH_CHECK(ctx);
    return NB_OK;
}
