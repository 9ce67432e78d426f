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
                interact = (mt.y & NB_PAYLOAD_MASK) != me;  // own leaf skipped (BarnesHutAlgorithm.cpp:349)
                next = cur + 1;
                if (STATS) nvis += interact ? 1u : 0u;
            } else {
                const uint32_t depth = mt.y & NB_PAYLOAD_MASK;
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


// ---------------------------------------------------------------------------------------------------------------------
// Two-phase variant (default).  The walk above keeps one long dependent chain per node: cursor -> load -> fp64 test ->
// rsqrt chain -> accumulate -> next cursor, which leaves the FP64 pipe ~45 % busy (profiles/).  Here the warp
//   phase 1 (walk):    decides accept/open with an FP32 test on pre-rounded coordinates relative to the root centre
//                      (FMA pipe, not the FP64 pipe).  The test is conservative: per depth the threshold is widened by a
//                      rigorous bound of the fp32 error (2.2e-7 * theta * 2^depth + 5e-7, doubled); anything inside the
//                      band falls back to the fp64 test of the kernel above, including its exact-oracle branch, so the
//                      interaction set stays identical to the reference's.  Accepted nodes are appended to a per-warp
//                      list in shared memory as {fp64 record, lane mask}; the next node (cursor+1) is prefetched.
//   phase 2 (evaluate): when the list is full the warp evaluates it branch-free, two entries at a time, so independent
//                      rsqrt chains overlap and the FP64 pipe stays fed.  Lanes outside an entry's mask use mass 0.
// Per-lane summation order is unchanged (list order == visit order).
// ---------------------------------------------------------------------------------------------------------------------
#define NB_BH_LIST 32

template <bool STATS>
__global__ void __launch_bounds__(256)
bh_traverse2_kernel(const double4 *__restrict__ com, const float4 *__restrict__ comf, const uint2 *__restrict__ meta,
                    const uint32_t *__restrict__ flags, uint64_t n_bodies, const double *__restrict__ aabb,
                    const double *__restrict__ sx, const double *__restrict__ sy, const double *__restrict__ sz,
                    uint64_t s_begin, uint64_t s_end, double theta, double eps2, double G, double *__restrict__ asx,
                    double *__restrict__ asy, double *__restrict__ asz, uint32_t *__restrict__ visits,
                    unsigned long long *__restrict__ totals) {
    __shared__ double t_hi[NB_BH_MAX_LEVELS], t_lo[NB_BH_MAX_LEVELS], t_edge[NB_BH_MAX_LEVELS];
    __shared__ float f_hi[NB_BH_MAX_LEVELS], f_lo[NB_BH_MAX_LEVELS];
    __shared__ uint32_t s_node[8][NB_BH_LIST];
    __shared__ uint32_t s_mask[8][NB_BH_LIST];
    for (int t = threadIdx.x; t < NB_BH_MAX_LEVELS; t += blockDim.x) {
        const double e = ldexp(aabb[6], -t);
        const double ratio = (e / theta) * (e / theta);
        t_edge[t] = e;
        t_hi[t] = ratio * (1.0 + 1e-12);
        t_lo[t] = ratio * (1.0 - 1e-12);
        // fp32 pre-test: relative error of the fp32 d2 at the acceptance boundary |d| = edge/theta is bounded by
        // 2*sqrt(3)*2^-24*(R/|d| + 1) + 4*2^-24 <= 2.2e-7*theta*2^depth + 5e-7 (R = root edge); doubled for safety
        const double marg = 2.0 * (2.2e-7 * fabs(theta) * ldexp(1.0, t) + 5e-7);
        float hi = __double2float_ru(ratio * (1.0 + marg));
        float lo = __double2float_rd(ratio * (1.0 - marg));
        if (!(marg < 0.25) || !(ratio < 1e37) || !(ratio > 1e-37)) { hi = __int_as_float(0x7f800000); lo = 0.0f; }  // fp64 only
        f_hi[t] = hi;
        f_lo[t] = lo;
    }
    __syncthreads();
    const uint32_t n_nodes = (uint32_t) n_bodies + flags[1];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t warp_global = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t b = s_begin + warp_global * 32 + lane;
    const bool valid = b < s_end;
    const uint32_t me = (uint32_t) b;
    double px = 0, py = 0, pz = 0;
    if (valid) { px = sx[b]; py = sy[b]; pz = sz[b]; }
    // fp32 coordinates relative to the centre of the root cube (same reference point as comf)
    const double hx = aabb[0] + 0.5 * aabb[6], hy = aabb[1] + 0.5 * aabb[6], hz = aabb[2] + 0.5 * aabb[6];
    const float pfx = (float) (px - hx), pfy = (float) (py - hy), pfz = (float) (pz - hz);
    double ax = 0, ay = 0, az = 0;
    uint32_t next = (valid && flags[0] == 0) ? 0u : 0xffffffffu;
    uint32_t nvis = 0, nacc = 0;
    uint32_t *my_node = s_node[wib];
    uint32_t *my_mask = s_mask[wib];
    uint32_t cnt = 0;

    // phase 2: the list holds {node, lane mask}; records are fetched here, four at a time, so the loads and the four
    // rsqrt chains are independent and overlap.  Entries past `count` are padded with mask 0 (mass 0 => no contribution).
    auto evaluate = [&](uint32_t count) {
        for (uint32_t k = 0; k < count; k += 4) {
            double4 r[4];
            double w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t kk = k + u < count ? k + u : k;
                r[u] = com[my_node[kk]];
                const uint32_t mk = k + u < count ? my_mask[kk] : 0u;
                w[u] = ((mk >> lane) & 1u) ? r[u].w : 0.0;
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
                const double sfac = (y3 * w[u]) * q;
                ax = fma(dx, sfac, ax);
                ay = fma(dy, sfac, ay);
                az = fma(dz, sfac, az);
            }
        }
    };

    uint32_t cur = __reduce_min_sync(0xffffffffu, next);
    float4 cf = make_float4(0, 0, 0, 0);
    uint2 mt = make_uint2(0, 0);
    if (cur < n_nodes) { cf = comf[cur]; mt = meta[cur]; }
    while (cur < n_nodes) {
        // speculative prefetch of the DFS successor (the most likely next cursor)
        const uint32_t nxt = cur + 1 < n_nodes ? cur + 1 : cur;
        const float4 cf_n = comf[nxt];
        const uint2 mt_n = meta[nxt];
        bool interact = false;
        if (next == cur) {
            if (mt.y & NB_LEAF_FLAG) {
                interact = (mt.y & NB_PAYLOAD_MASK) != me;
                next = cur + 1;
                if (STATS) nvis += interact ? 1u : 0u;
            } else {
                const uint32_t depth = mt.y & NB_PAYLOAD_MASK;
                const float dfx = cf.x - pfx, dfy = cf.y - pfy, dfz = cf.z - pfz;
                const float d2f = fmaf(dfz, dfz, fmaf(dfy, dfy, dfx * dfx));
                bool accept = d2f > f_hi[depth];
                if (!accept && !(d2f < f_lo[depth])) {
                    // inside the fp32 uncertainty band: fp64 test (and, inside ITS band, the oracle's exact expression)
                    const double4 c = com[cur];
                    const double dx = c.x - px, dy = c.y - py, dz = c.z - pz;
                    const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    accept = d2 > t_hi[depth];
                    if (!accept && !(d2 < t_lo[depth])) {
                        const double d2o = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                        const double rs = __ddiv_rn(1.0, __dsqrt_rn(d2o));
                        accept = __dmul_rn(t_edge[depth], rs) < theta;
                    }
                }
                interact = accept;
                next = accept ? max(mt.x, cur + 1) : cur + 1;
                if (STATS) nvis += 1u;
            }
        }
        const uint32_t msk = __ballot_sync(0xffffffffu, interact);
        if (msk) {
            if (STATS) nacc += interact ? 1u : 0u;
            if (lane == 0) {
                my_node[cnt] = cur;
                my_mask[cnt] = msk;
            }
            ++cnt;
            if (cnt == NB_BH_LIST) {
                __syncwarp();
                evaluate(cnt);
                __syncwarp();
                cnt = 0;
            }
        }
        const uint32_t ncur = __reduce_min_sync(0xffffffffu, next);
        if (ncur == cur + 1) {
            cf = cf_n; mt = mt_n;
        } else if (ncur < n_nodes) {
            cf = comf[ncur]; mt = meta[ncur];
        }
        cur = ncur;
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
    const bool two_phase = ctx->cfg.reserved[1] == 0;  // reserved[1] = 1 selects the single-phase walk (A/B testing)
    if (two_phase) {
        const float4 *comf = reinterpret_cast<const float4 *>(b.comf);
        if (b.stats_enabled) {
            NB_CUDA(ctx, cudaMemsetAsync(b.stat_totals, 0, 2 * sizeof(unsigned long long), ctx->stream));
            NB_CUDA(ctx, cudaMemsetAsync(b.visits, 0, ctx->n * sizeof(uint32_t), ctx->stream));
            bh_traverse2_kernel<true><<<grid, threads, 0, ctx->stream>>>(com, comf, b.meta, b.dev_flags, ctx->n, b.aabb_dev,
                                                                         b.sx, b.sy, b.sz, s_begin, s_end, ctx->cfg.theta,
                                                                         ctx->cfg.epsilon2, ctx->cfg.G, b.asx, b.asy, b.asz,
                                                                         b.visits, b.stat_totals);
        } else {
            bh_traverse2_kernel<false><<<grid, threads, 0, ctx->stream>>>(com, comf, b.meta, b.dev_flags, ctx->n, b.aabb_dev,
                                                                          b.sx, b.sy, b.sz, s_begin, s_end, ctx->cfg.theta,
                                                                          ctx->cfg.epsilon2, ctx->cfg.G, b.asx, b.asy, b.asz,
                                                                          b.visits, b.stat_totals);
        }
        NB_LAUNCH_CHECK(ctx);
        return NB_OK;
    }
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
