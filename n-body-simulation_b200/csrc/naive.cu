// All-pairs fp64 gravity for sm_100a.
//
// Replaces NaiveAlgorithm::computeAccelerations_opt_{0,1,2} (reference src/simulationBackend/NaiveAlgorithm.cpp:262-482):
//     a_i = G * sum_{j=0..N-1} m_j * r_ij * (|r_ij|^2 + eps2)^(-3/2),   r_ij = x_j - x_i,  j ascending, self term included.
//
// Design (B200-first, not a translation of the SYCL tiles):
//   * sources are pre-packed into 48-byte AoS records {x,y,z,m,1.5m,1.875m}; a tile of `tile_len` records is ONE
//     contiguous TMA bulk copy (cp.async.bulk, SASS UBLKCP) into shared memory, completion tracked by mbarriers;
//   * an NB_NAIVE_STAGES-deep full/empty mbarrier ring, no __syncthreads in the steady state.  Production form
//     (naive_accel_np_kernel): 4 consumer warps per CTA that take turns issuing the bulk copy two tiles ahead; the
//     round-1 form with a dedicated producer warp (naive_accel_kernel) is kept as naive_variant 1;
//   * register blocking: each consumer thread owns IPT target bodies, so one broadcast LDS.128 triple feeds
//     IPT*32 interactions;
//   * the FP64 pipe is the roofline.  Per interaction: 3 DADD (r), 3 DFMA (d2 = r.r + eps2), MUFU.RSQ64H seed
//     y0 ~ d2^(-1/2) (XU pipe, not DP), then   y2 = y0*y0;  e = 1 - d2*y2;  y3 = y2*y0;
//     s = m*d2^(-3/2) = y3 * (m + e*(1.5m + 1.875m*e))   [Taylor of (1-e)^(-3/2), |e| <~ 2^-20 -> error ~2e-18]
//     = 6 DP ops including the mass multiply, then 3 DFMA accumulate: 15 DP instructions for the 21 algorithmic flops
//     (SURVEY 8d).  precise_rsqrt=0 drops the quadratic term (14 DP ops, ~2e-12 relative).
//   * summation order per target: j ascending inside a source segment, segment partials added in ascending order.  Up to
//     2^24 bodies the source range is cut into up to 32 segments (a function of N only, see launch_naive) so that a
//     rank's slice of an 8-GPU run still fills the machine; the reference adds all j in one chain.  The differences to
//     the oracle are this blocking, the rsqrt refinement and FMA contraction (~1e-16 relative per term).
#include "common.cuh"

#define NB_NAIVE_CONSUMER_WARPS 4
#define NB_NAIVE_THREADS (32 * (NB_NAIVE_CONSUMER_WARPS + 1))
#define NB_NAIVE_STAGES 4

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void pack_sources_kernel(const double *__restrict__ m, const double *__restrict__ x,
                                    const double *__restrict__ y, const double *__restrict__ z,
                                    nb_src_rec *__restrict__ out, uint64_t n, uint64_t n_pad) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    nb_src_rec r;
    if (i < n) {
        double mm = m[i];
        r.x = x[i]; r.y = y[i]; r.z = z[i]; r.m = mm;
        r.c1 = 1.5 * mm; r.c2 = 1.875 * mm;
    } else {  // padding: zero mass => contributes exactly 0 (eps2 > 0 keeps d2 finite and positive)
        r.x = r.y = r.z = r.m = r.c1 = r.c2 = 0.0;
    }
    out[i] = r;
}

template <int IPT, bool PRECISE, int UNROLL, int MINB>
__global__ void __launch_bounds__(NB_NAIVE_THREADS, MINB)
naive_accel_kernel(const nb_src_rec *__restrict__ src, uint32_t n_tiles_total, uint32_t seg_tiles, uint32_t tile_len,
                   const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
                   uint64_t i_begin, uint64_t i_end, double eps2, double G, double *__restrict__ ax,
                   double *__restrict__ ay, double *__restrict__ az, double *__restrict__ partial) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + NB_NAIVE_STAGES;
    nb_src_rec *tiles = reinterpret_cast<nb_src_rec *>(smem_raw + 128);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t tile_bytes = tile_len * (uint32_t) sizeof(nb_src_rec);
    // blockIdx.y = source segment: this CTA sums the tiles [t0, t0 + n_tiles) only (finer work quanta for small target
    // counts; the segments are added in ascending order by naive_reduce_kernel)
    const uint32_t t0 = blockIdx.y * seg_tiles;
    const uint32_t n_tiles = n_tiles_total - t0 < seg_tiles ? n_tiles_total - t0 : seg_tiles;
    src += (size_t) t0 * tile_len;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NB_NAIVE_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NB_NAIVE_CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NB_NAIVE_CONSUMER_WARPS) {
        // ---- producer warp: one elected lane streams the source tiles through the ring ----
        if (lane == 0) {
            for (uint32_t t = 0; t < n_tiles; ++t) {
                const uint32_t s = t % NB_NAIVE_STAGES;
                const uint32_t use = t / NB_NAIVE_STAGES;
                if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
                mbar_expect_tx(&full[s], tile_bytes);
                bulk_g2s(tiles + (size_t) s * tile_len, src + (size_t) t * tile_len, tile_bytes, &full[s]);
            }
        }
        return;
    }

    // ---- consumer warps ----
    constexpr int TPB = NB_NAIVE_CONSUMER_WARPS * 32;  // targets per pass per IPT slot
    const uint64_t base = i_begin + (uint64_t) blockIdx.x * (TPB * IPT) + threadIdx.x;
    double px[IPT], py[IPT], pz[IPT], accx[IPT], accy[IPT], accz[IPT];
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        uint64_t i = base + (uint64_t) k * TPB;
        if (i >= i_end) i = i_end - 1;
        px[k] = x[i]; py[k] = y[i]; pz[k] = z[i];
        accx[k] = accy[k] = accz[k] = 0.0;
    }

    for (uint32_t t = 0; t < n_tiles; ++t) {
        const uint32_t s = t % NB_NAIVE_STAGES;
        mbar_wait(&full[s], (t / NB_NAIVE_STAGES) & 1);
        const double2 *tp = reinterpret_cast<const double2 *>(tiles + (size_t) s * tile_len);
#pragma unroll(UNROLL)
        for (uint32_t j = 0; j < tile_len; ++j) {
            const double2 xy = tp[3 * j + 0];
            const double2 zm = tp[3 * j + 1];
            const double2 cc = tp[3 * j + 2];
#pragma unroll
            for (int k = 0; k < IPT; ++k) {
                const double rx = xy.x - px[k];
                const double ry = xy.y - py[k];
                const double rz = zm.x - pz[k];
                double d2 = fma(rx, rx, eps2);
                d2 = fma(ry, ry, d2);
                d2 = fma(rz, rz, d2);
                const double y0 = nb_rsqrt_seed(d2);
                const double y2 = y0 * y0;
                const double e = fma(-d2, y2, 1.0);
                const double y3 = y2 * y0;
                double q;
                if (PRECISE) {
                    const double p = fma(cc.y, e, cc.x);
                    q = fma(p, e, zm.y);
                } else {
                    q = fma(cc.x, e, zm.y);
                }
                const double sfac = y3 * q;
                accx[k] = fma(rx, sfac, accx[k]);
                accy[k] = fma(ry, sfac, accy[k]);
                accz[k] = fma(rz, sfac, accz[k]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    const uint64_t count = i_end - i_begin;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        const uint64_t i = base + (uint64_t) k * TPB;
        if (i < i_end) {
            if (gridDim.y == 1) {
                ax[i] = accx[k] * G;
                ay[i] = accy[k] * G;
                az[i] = accz[k] * G;
            } else {
                double *p = partial + (size_t) blockIdx.y * 3 * count + (i - i_begin);
                p[0] = accx[k];
                p[count] = accy[k];
                p[2 * count] = accz[k];
            }
        }
    }
}

// The same kernel without a dedicated producer warp: 128 threads, all of them consumers.  Consumer warp (t mod 4) issues
// the bulk copy of tile t + 2 when it starts tile t (the stage it overwrites was last read for tile t - 2, which every
// warp has left behind once anybody starts tile t + ... see the empty barrier).  The producer warp of the kernel above
// holds 32 x 158 registers it never uses; without it a CTA needs 20 K registers instead of 25 K and THREE CTAs fit an
// SM: 12 consumer warps, 3 per scheduler instead of 2 (ncu of the 2-warp form: FP64 pipe 84.5 % active, 0.89 eligible
// warps per cycle).  Same arithmetic, same summation order, same bits.
template <int IPT, bool PRECISE, int UNROLL, int MINB>
__global__ void __launch_bounds__(NB_NAIVE_CONSUMER_WARPS * 32, MINB)
naive_accel_np_kernel(const nb_src_rec *__restrict__ src, uint32_t n_tiles_total, uint32_t seg_tiles, uint32_t tile_len,
                      const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
                      uint64_t i_begin, uint64_t i_end, double eps2, double G, double *__restrict__ ax,
                      double *__restrict__ ay, double *__restrict__ az, double *__restrict__ partial) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + NB_NAIVE_STAGES;
    nb_src_rec *tiles = reinterpret_cast<nb_src_rec *>(smem_raw + 128);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t tile_bytes = tile_len * (uint32_t) sizeof(nb_src_rec);
    const uint32_t t0 = blockIdx.y * seg_tiles;
    const uint32_t n_tiles = n_tiles_total - t0 < seg_tiles ? n_tiles_total - t0 : seg_tiles;
    src += (size_t) t0 * tile_len;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NB_NAIVE_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NB_NAIVE_CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    constexpr uint32_t AHEAD = 2;   // tiles in flight ahead of the one being consumed (NB_NAIVE_STAGES >= AHEAD + 2)
    if (threadIdx.x == 0) {         // prologue: tiles 0 .. AHEAD-1
        for (uint32_t t = 0; t < AHEAD && t < n_tiles; ++t) {
            mbar_expect_tx(&full[t], tile_bytes);
            bulk_g2s(tiles + (size_t) t * tile_len, src + (size_t) t * tile_len, tile_bytes, &full[t]);
        }
    }
    constexpr int TPB = NB_NAIVE_CONSUMER_WARPS * 32;
    const uint64_t base = i_begin + (uint64_t) blockIdx.x * (TPB * IPT) + threadIdx.x;
    double px[IPT], py[IPT], pz[IPT], accx[IPT], accy[IPT], accz[IPT];
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        uint64_t i = base + (uint64_t) k * TPB;
        if (i >= i_end) i = i_end - 1;
        px[k] = x[i]; py[k] = y[i]; pz[k] = z[i];
        accx[k] = accy[k] = accz[k] = 0.0;
    }
    for (uint32_t t = 0; t < n_tiles; ++t) {
        const uint32_t s = t % NB_NAIVE_STAGES;
        // this warp's turn to fetch: tile t + AHEAD into the stage tile t + AHEAD - STAGES was read from
        if ((t % NB_NAIVE_CONSUMER_WARPS) == (uint32_t) warp && lane == 0 && t + AHEAD < n_tiles) {
            const uint32_t tn = t + AHEAD, sn = tn % NB_NAIVE_STAGES, use = tn / NB_NAIVE_STAGES;
            if (use > 0) mbar_wait(&empty[sn], (use - 1) & 1);
            mbar_expect_tx(&full[sn], tile_bytes);
            bulk_g2s(tiles + (size_t) sn * tile_len, src + (size_t) tn * tile_len, tile_bytes, &full[sn]);
        }
        __syncwarp();
        mbar_wait(&full[s], (t / NB_NAIVE_STAGES) & 1);
        const double2 *tp = reinterpret_cast<const double2 *>(tiles + (size_t) s * tile_len);
#pragma unroll(UNROLL)
        for (uint32_t j = 0; j < tile_len; ++j) {
            const double2 xy = tp[3 * j + 0];
            const double2 zm = tp[3 * j + 1];
            const double2 cc = tp[3 * j + 2];
#pragma unroll
            for (int k = 0; k < IPT; ++k) {
                const double rx = xy.x - px[k];
                const double ry = xy.y - py[k];
                const double rz = zm.x - pz[k];
                double d2 = fma(rx, rx, eps2);
                d2 = fma(ry, ry, d2);
                d2 = fma(rz, rz, d2);
                const double y0 = nb_rsqrt_seed(d2);
                const double y2 = y0 * y0;
                const double e = fma(-d2, y2, 1.0);
                const double y3 = y2 * y0;
                double q;
                if (PRECISE) {
                    const double p = fma(cc.y, e, cc.x);
                    q = fma(p, e, zm.y);
                } else {
                    q = fma(cc.x, e, zm.y);
                }
                const double sfac = y3 * q;
                accx[k] = fma(rx, sfac, accx[k]);
                accy[k] = fma(ry, sfac, accy[k]);
                accz[k] = fma(rz, sfac, accz[k]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    const uint64_t count = i_end - i_begin;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        const uint64_t i = base + (uint64_t) k * TPB;
        if (i < i_end) {
            if (gridDim.y == 1) {
                ax[i] = accx[k] * G;
                ay[i] = accy[k] * G;
                az[i] = accz[k] * G;
            } else {
                double *p = partial + (size_t) blockIdx.y * 3 * count + (i - i_begin);
                p[0] = accx[k];
                p[count] = accy[k];
                p[2 * count] = accz[k];
            }
        }
    }
}

// deterministic sum of the source segments (ascending), then the final scale by G (NaiveAlgorithm.cpp:349-351)
__global__ void __launch_bounds__(256)
naive_reduce_kernel(const double *__restrict__ partial, uint32_t segments, uint64_t i_begin, uint64_t count, double G,
                    double *__restrict__ ax, double *__restrict__ ay, double *__restrict__ az) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double sx = 0, sy = 0, sz = 0;
    for (uint32_t s = 0; s < segments; ++s) {
        const double *p = partial + (size_t) s * 3 * count + i;
        sx += p[0];
        sy += p[count];
        sz += p[2 * count];
    }
    ax[i_begin + i] = sx * G;
    ay[i_begin + i] = sy * G;
    az[i_begin + i] = sz * G;
}

// DFMA-chain microbenchmark: CHAINS independent dependent-FMA chains per thread, 2 flops per DFMA.
template <int CHAINS>
__global__ void __launch_bounds__(1024) fp64_peak_kernel(double *out, int iters, double a, double b) {
    double v[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) v[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) v[c] = fma(v[c], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += v[c];
    out[(size_t) blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int IPT, bool PRECISE, int UNROLL = 4, int MINB = 2, bool NOPROD = false>
int launch_naive(nb_ctx *ctx, uint32_t n_tiles, uint32_t tile_len, uint64_t i_begin, uint64_t i_end) {
    const uint64_t per_cta = (uint64_t) NB_NAIVE_CONSUMER_WARPS * 32 * IPT;
    const uint64_t count = i_end - i_begin;
    const uint64_t grid = (count + per_cta - 1) / per_cta;
    const size_t smem = 128 + (size_t) NB_NAIVE_STAGES * tile_len * sizeof(nb_src_rec);
    // work quanta: aim for >= 32 CTAs per SM so the tail of the last wave is a few percent; split the source range
    // into segments when there are too few target tiles (small N, or a rank's slice on many GPUs).  The number of
    // segments fixes the summation order of a target (segment partials are added in ascending order), so it is a
    // function of N alone -- sized for a rank's slice of an 8-GPU run of a 148-SM part, whatever this launch is: one
    // GPU and eight produce the same bits, on any device.
    uint32_t segments = 1;
    const uint64_t grid_full = (ctx->n + per_cta - 1) / per_cta;
    const uint64_t want = (uint64_t) NB_SM_COUNT_FALLBACK * 32 * NB_MAX_PEERS;
    if (ctx->cfg.reserved[4] > 0) segments = (uint32_t) ctx->cfg.reserved[4];
    else if (grid_full < want) segments = (uint32_t) ((want + grid_full - 1) / grid_full);
    if (segments > 32) segments = 32;
    if (segments > n_tiles) segments = n_tiles;
    uint32_t seg_tiles = (n_tiles + segments - 1) / segments;
    segments = (n_tiles + seg_tiles - 1) / seg_tiles;
    if (segments > 1) {
        const size_t need = (size_t) segments * 3 * count;
        if (need > ctx->naive_partial_cap) {
            NB_CHECK(nb_alloc(ctx, &ctx->naive_partial, need));
            ctx->naive_partial_cap = need;
        }
    }
    void (*kern)(const nb_src_rec *, uint32_t, uint32_t, uint32_t, const double *, const double *, const double *, uint64_t,
                 uint64_t, double, double, double *, double *, double *, double *);
    if constexpr (NOPROD) kern = naive_accel_np_kernel<IPT, PRECISE, UNROLL, MINB>;
    else kern = naive_accel_kernel<IPT, PRECISE, UNROLL, MINB>;
    NB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    kern<<<dim3((unsigned) grid, segments), NOPROD ? NB_NAIVE_CONSUMER_WARPS * 32 : NB_NAIVE_THREADS, smem, ctx->stream>>>(
        ctx->src, n_tiles, seg_tiles, tile_len, ctx->x, ctx->y, ctx->z, i_begin, i_end, ctx->cfg.epsilon2, ctx->cfg.G, ctx->ax,
        ctx->ay, ctx->az, ctx->naive_partial);
    NB_LAUNCH_CHECK(ctx);
    if (segments > 1) {
        naive_reduce_kernel<<<(unsigned) ((count + 255) / 256), 256, 0, ctx->stream>>>(ctx->naive_partial, segments, i_begin,
                                                                                    count, ctx->cfg.G, ctx->ax, ctx->ay, ctx->az);
        NB_LAUNCH_CHECK(ctx);
    }
    return NB_OK;
}

}  // namespace

int nbk_naive_accel(nb_ctx *ctx, uint64_t i_begin, uint64_t i_end) {
    if (ctx->n == 0 || i_begin >= i_end) return NB_OK;
    // --block_size -> shared-memory tile length (reference: tile = work-group = blockSize, NaiveAlgorithm.cpp:275-296)
    uint32_t tile_len = (uint32_t) ctx->cfg.block_size;
    if (tile_len < 16) tile_len = 16;
    if (tile_len > 1024) tile_len = 1024;
    tile_len = (tile_len + 3u) & ~3u;
    const uint64_t n_pad = (ctx->n + tile_len - 1) / tile_len * tile_len;
    if (n_pad > ctx->src_cap) {
        NB_CHECK(nb_alloc(ctx, &ctx->src, n_pad + 1024));
        ctx->src_cap = n_pad + 1024;
    }
    {
        const unsigned threads = 256;
        const unsigned blocks = (unsigned) ((n_pad + threads - 1) / threads);
        pack_sources_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->m, ctx->x, ctx->y, ctx->z, ctx->src, ctx->n, n_pad);
        NB_LAUNCH_CHECK(ctx);
    }
    const uint32_t n_tiles = (uint32_t) (n_pad / tile_len);
    const int ipt = ctx->cfg.reserved[0] > 0 ? ctx->cfg.reserved[0] : 4;  // register blocking (tuning knob)
    const bool precise = ctx->cfg.precise_rsqrt != 0;
    // Production form: no dedicated producer warp (naive_accel_np_kernel), IPT 4, unroll 2, two 128-thread CTAs per SM.
    // Measured at N = 2^19 on B200 (tools/dev_naive_sweep.py, all bit-identical): 258.6 ms, against 262.5 ms for the
    // producer-warp form of round 1 (naive_variant 1), 260.9 ms with three CTAs per SM (2: a third warp per scheduler
    // does not help -- the DP pipe, not latency, is what the kernel waits for), 265.5 ms with four (3, 128 registers);
    // IPT 3 / IPT 2 forms 263 - 267 ms.  naive_variant (cfg.reserved[2]) keeps these three for A/B runs.
    switch (precise ? ctx->cfg.reserved[2] : 0) {
        case 1: return launch_naive<4, true, 2, 2, false>(ctx, n_tiles, tile_len, i_begin, i_end);
        case 2: return launch_naive<4, true, 2, 3, true>(ctx, n_tiles, tile_len, i_begin, i_end);
        case 3: return launch_naive<4, true, 2, 4, true>(ctx, n_tiles, tile_len, i_begin, i_end);
        default: break;
    }
    if (ipt == 1) return precise ? launch_naive<1, true, 4, 2, true>(ctx, n_tiles, tile_len, i_begin, i_end)
                                 : launch_naive<1, false, 4, 2, true>(ctx, n_tiles, tile_len, i_begin, i_end);
    if (ipt == 4) return precise ? launch_naive<4, true, 2, 2, true>(ctx, n_tiles, tile_len, i_begin, i_end)
                                 : launch_naive<4, false, 2, 2, true>(ctx, n_tiles, tile_len, i_begin, i_end);
    return precise ? launch_naive<2, true, 4, 2, true>(ctx, n_tiles, tile_len, i_begin, i_end)
                   : launch_naive<2, false, 4, 2, true>(ctx, n_tiles, tile_len, i_begin, i_end);
}

int nbk_fp64_peak(nb_ctx *ctx, double *tflops) {
    // best of several launch shapes (chains per thread x CTA size), each timed over >= 10 ms so clocks settle
    double *buf = nullptr;
    const size_t max_threads = (size_t) ctx->sm_count * 2048;
    NB_CUDA(ctx, cudaMalloc(&buf, max_threads * sizeof(double)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0;
    const int shapes[4][2] = {{8, 256}, {16, 256}, {8, 512}, {16, 1024}};  // {chains, threads}
    for (int sidx = 0; sidx < 4; ++sidx) {
        const int chains = shapes[sidx][0], threads = shapes[sidx][1];
        const int blocks = ctx->sm_count * (2048 / threads);
        const int iters = 16384 / chains * 8;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0, ctx->stream);
            if (chains == 8) fp64_peak_kernel<8><<<blocks, threads, 0, ctx->stream>>>(buf, iters, 1.0000001, 1e-9);
            else fp64_peak_kernel<16><<<blocks, threads, 0, ctx->stream>>>(buf, iters, 1.0000001, 1e-9);
            ctx->launches++;
            cudaEventRecord(e1, ctx->stream);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double fl = 2.0 * 8.0 * chains * iters * (double) blocks * threads;
            const double tf = fl / (ms * 1e-3) / 1e12;
            if (rep > 0 && tf > best) best = tf;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return nb_fail(ctx, NB_ERR_CUDA, "fp64 peak kernel: %s", cudaGetErrorString(err));
    *tflops = best;
    return NB_OK;
}
