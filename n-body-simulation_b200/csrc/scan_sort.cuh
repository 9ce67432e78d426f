// Hand-written device primitives used by the Barnes-Hut build: exclusive prefix sum over uint32 and a stable
// LSD radix sort of (uint64 key, uint32 value) pairs or of packed 64-bit words (one-sweep form, decoupled look-back).  Replaces the reference's serial single_task scan
// (ParallelOctreeTopDownSubtrees.cpp:461-474), its O(S^2) prefix (:491-500) and the linear-search scatter (:512-532).
#pragma once
#include "common.cuh"

namespace nbprim {
namespace {  // internal linkage: this header is included by several translation units

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048 elements per block

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across a block of NT threads; returns exclusive prefix, total in *total
template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total, uint32_t *smem /* NT/32 + 1 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t inc = warp_incl_scan(v);
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < NT / 32 ? smem[lane] : 0;
        uint32_t wi = warp_incl_scan(w);
        if (lane < NT / 32) smem[lane] = wi - w;
        if (lane == 31) smem[NT / 32] = wi;
    }
    __syncthreads();
    const uint32_t res = smem[warp] + inc - v;
    *total = smem[NT / 32];
    __syncthreads();
    return res;
}

// phase 1: per-tile sums
__global__ void __launch_bounds__(SCAN_THREADS)
scan_reduce_kernel(const uint32_t *__restrict__ in, uint64_t n, uint32_t *__restrict__ tile_sums) {
    __shared__ uint32_t sm[SCAN_THREADS / 32 + 1];
    const uint64_t base = (uint64_t) blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const uint64_t i = base + (uint64_t) k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    uint32_t tot;
    block_excl_scan<SCAN_THREADS>(s, &tot, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// phase 2: one block scans the tile sums in place (exclusive), writes the grand total to *total_out
__global__ void __launch_bounds__(1024)
scan_tiles_kernel(uint32_t *__restrict__ tile_sums, uint32_t n_tiles, uint32_t *__restrict__ total_out) {
    __shared__ uint32_t sm[1024 / 32 + 1];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n_tiles; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_tiles ? tile_sums[i] : 0;
        uint32_t tot;
        const uint32_t ex = block_excl_scan<1024>(v, &tot, sm);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// phase 3: per-tile exclusive scan + tile offset.  Thread-contiguous item layout for the local scan.
__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const uint32_t *in, uint64_t n, const uint32_t *__restrict__ tile_offs, uint32_t *out) {  // in may alias out
    __shared__ uint32_t sm[SCAN_THREADS / 32 + 1];
    const uint64_t base = (uint64_t) blockIdx.x * SCAN_TILE + (uint64_t) threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    uint32_t tot;
    uint32_t ex = block_excl_scan<SCAN_THREADS>(s, &tot, sm) + tile_offs[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
}

inline uint64_t scan_tiles_for(uint64_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

// out[i] = sum_{k<i} in[i] for i in [0,n); grand total written to *total_dev (device pointer, may be null).
// tile_tmp: device scratch of scan_tiles_for(n) uint32.  in may alias out.
inline int exclusive_scan_u32(nb_ctx *ctx, const uint32_t *in, uint32_t *out, uint64_t n, uint32_t *tile_tmp,
                              uint32_t *total_dev) {
    if (n == 0) {
        if (total_dev) NB_CUDA(ctx, cudaMemsetAsync(total_dev, 0, sizeof(uint32_t), ctx->stream));
        return NB_OK;
    }
    const uint32_t tiles = (uint32_t) scan_tiles_for(n);
    scan_reduce_kernel<<<tiles, SCAN_THREADS, 0, ctx->stream>>>(in, n, tile_tmp);
    NB_LAUNCH_CHECK(ctx);
    scan_tiles_kernel<<<1, 1024, 0, ctx->stream>>>(tile_tmp, tiles, total_dev);
    NB_LAUNCH_CHECK(ctx);
    scan_apply_kernel<<<tiles, SCAN_THREADS, 0, ctx->stream>>>(in, n, tile_tmp, out);
    NB_LAUNCH_CHECK(ctx);
    return NB_OK;
}

constexpr int RS_BINS = 256;                      // 8-bit digits

// ---------------------------------------------------------------------------------------------------------------
// Stable LSD radix sort, 8-bit digits, one-sweep form.  The digit histograms of ALL passes are taken in one read of
// the keys before the first pass (a stable sort does not change them), so a pass is a single kernel: every tile
// ranks its keys, publishes its per-digit counts and obtains the counts of the tiles before it by decoupled
// look-back over a status table ({flag, count} packed in one 32-bit word per (tile, digit): 1 = the tile's own
// count, 2 = inclusive prefix over tiles 0..t).  Tiles are numbered by an atomic ticket, so a tile only ever waits
// for tiles that are already running.  (Round 1 also carried a histogram / scan / scatter form with three kernels and
// a 256 x tiles table scan per pass; it produced the same permutation bit for bit and was removed.)
// ---------------------------------------------------------------------------------------------------------------
constexpr int OS_MAX_PASSES = 8;
constexpr uint32_t OS_FLAG_COUNT = 1u << 30, OS_FLAG_PREFIX = 2u << 30, OS_VALUE_MASK = (1u << 30) - 1u;

// Packed sort word of the Barnes-Hut build: the 63-bit octant-path key in the upper bits, the body's storage slot in the
// lower idx_bits (the lowest key bits are dropped; bodies that agree on all kept key bits are ordered afterwards by
// their full keys).  idx_bits == 0: plain keys.
__device__ __forceinline__ uint64_t os_pack(uint64_t key, uint64_t i, int idx_bits) {
    if (idx_bits == 0) return key;
    const uint64_t mask = (1ull << idx_bits) - 1ull;
    return ((key << 1) & ~mask) | i;
}

// all-pass digit histograms; adjacent equal digits inside a warp are added as one run (the keys of the high passes are
// nearly sorted from the previous step, so a warp usually holds one or two distinct high digits).  Pass p looks at the
// 8 bits from bit first_shift + 8 p.
__global__ void __launch_bounds__(256)
os_hist_kernel(const uint64_t *__restrict__ keys, uint64_t n, int passes, int first_shift, int idx_bits,
               uint32_t *__restrict__ ghist /* [passes][256] */) {
    __shared__ uint32_t h[OS_MAX_PASSES][RS_BINS];
    for (int b = threadIdx.x; b < OS_MAX_PASSES * RS_BINS; b += blockDim.x) (&h[0][0])[b] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    constexpr int U = 4;                                                  // keys in flight per thread
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    const uint64_t n_round = (n + 31) & ~31ull;                           // whole warps stay in the loop together
    for (uint64_t i0 = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i0 < n_round; i0 += U * stride) {
        uint64_t key[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = i0 + (uint64_t) u * stride;
            key[u] = i < n ? os_pack(keys[i], i, idx_bits) : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = i0 + (uint64_t) u * stride;
            if (i >= n_round) break;                                      // warp-uniform
            const bool valid = i < n;
            for (int p = 0; p < passes; ++p) {
                const uint32_t d = valid ? (uint32_t) ((key[u] >> (first_shift + 8 * p)) & 0xff) : 0x100u;
                const uint32_t prev = __shfl_up_sync(0xffffffffu, d, 1);
                const bool head = lane == 0 || d != prev;
                const uint32_t heads = __ballot_sync(0xffffffffu, head);
                if (head && valid) {
                    const uint32_t later = lane == 31 ? 0u : (heads >> (lane + 1));
                    const uint32_t run = later ? (uint32_t) __ffs(later) : (uint32_t) (32 - lane);
                    atomicAdd(&h[p][d], run);
                }
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < passes * RS_BINS; b += blockDim.x) {
        const uint32_t c = (&h[0][0])[b];
        if (c) atomicAdd(&ghist[b], c);
    }
}

// ghist[p][d] -> exclusive prefix over d (the global start of digit d in pass p); one block, thread d owns digit d
__global__ void __launch_bounds__(RS_BINS)
os_scan_kernel(uint32_t *__restrict__ ghist, int passes) {
    __shared__ uint32_t sm[RS_BINS / 32 + 1];
    for (int p = 0; p < passes; ++p) {
        const uint32_t v = ghist[p * RS_BINS + threadIdx.x];
        uint32_t tot;
        const uint32_t ex = block_excl_scan<RS_BINS>(v, &tot, sm);
        ghist[p * RS_BINS + threadIdx.x] = ex;
    }
}

// MODE: OS_PAIRS = (key, value) pairs; OS_PAIRS_IOTA = the same, values of the first pass are 0..n-1; OS_WORDS = 64-bit
// words alone (no value arrays: the packed (key | slot) words of the Barnes-Hut build); OS_WORDS_PACK = the first pass
// of a packed sort: reads the plain keys and packs the slot index into the low idx_bits on the fly.
enum { OS_PAIRS = 0, OS_PAIRS_IOTA = 1, OS_WORDS = 2, OS_WORDS_PACK = 3 };

template <int THREADS, int ITEMS, int MODE>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 5 : 4)
os_scatter_kernel(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in, uint64_t n, int shift,
                  const uint32_t *__restrict__ gstart /* [256] of this pass */, uint32_t *status /* [tiles][256] */,
                  uint32_t *ticket, uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int idx_bits) {
    constexpr int TILE = THREADS * ITEMS;
    constexpr int WARPS = THREADS / 32;
    constexpr bool WORDS = MODE == OS_WORDS || MODE == OS_WORDS_PACK;
    static_assert(THREADS >= RS_BINS, "thread b owns digit b");
    extern __shared__ __align__(16) unsigned char os_smem[];
    uint64_t *skey = reinterpret_cast<uint64_t *>(os_smem);                       // TILE keys staged in digit order
    uint32_t *sval = reinterpret_cast<uint32_t *>(skey + TILE);                   // TILE values (pairs only)
    uint32_t(*cnt)[RS_BINS] = reinterpret_cast<uint32_t(*)[RS_BINS]>(sval + (WORDS ? 0 : TILE));  // [WARPS][RS_BINS]
    uint32_t *gbase = &cnt[0][0] + WARPS * RS_BINS;                            // [RS_BINS]
    uint32_t *wsum = gbase + RS_BINS;                                             // [WARPS + 1]
    uint32_t *s_tile = wsum + WARPS + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) *s_tile = atomicAdd(ticket, 1u);
    for (int b = threadIdx.x; b < WARPS * RS_BINS; b += THREADS) (&cnt[0][0])[b] = 0;
    __syncthreads();
    const uint32_t tile = *s_tile;

    // warp w owns the contiguous sub-tile [w*32*ITEMS, (w+1)*32*ITEMS), processed in ITEMS rounds of 32 (stable order)
    const uint64_t tile_base = (uint64_t) tile * TILE;
    const uint64_t wbase = tile_base + (uint64_t) warp * (32 * ITEMS);
    uint64_t key[ITEMS];
    uint32_t rank[ITEMS], prior_of[ITEMS];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint64_t i = wbase + (uint64_t) k * 32 + lane;
        key[k] = i < n ? (MODE == OS_WORDS_PACK ? os_pack(keys_in[i], i, idx_bits) : keys_in[i]) : ~0ull;
    }
    // rank of a key among the keys of its warp with the same digit: the leader of each group of equal digits adds the
    // group size to the warp's counter with ONE shared-memory atomic that returns the count so far.  The ITEMS atomics
    // of a thread do not depend on each other through registers, so they pipeline; instructions of one warp reach
    // shared memory in program order, which keeps the ranking stable.  (Measured: splitting this into separate
    // match / atomic / shuffle loops costs registers under the 48-register budget and is 2 % slower.)
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint64_t i = wbase + (uint64_t) k * 32 + lane;
        const bool valid = i < n;
        const uint32_t d = valid ? (uint32_t) ((key[k] >> shift) & 0xff) : 0x100u;  // invalid lanes never match a real digit
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t prior = 0;
        if (valid && lane == leader) prior = atomicAdd(&cnt[warp][d], (uint32_t) __popc(peers));
        rank[k] = (uint32_t) __popc(peers & lt) | ((uint32_t) leader << 16);
        prior_of[k] = prior;   // shuffled to the group in the second loop so the atomics issue back to back
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t prior = __shfl_sync(0xffffffffu, prior_of[k], (int) (rank[k] >> 16));
        rank[k] = prior + (rank[k] & 0xffffu);
    }
    __syncthreads();
    // thread b owns digit b: exclusive prefix of the digit over the warps of the tile
    const int b = threadIdx.x;
    const bool owner = THREADS == RS_BINS || b < RS_BINS;
    uint32_t digit_total = 0;
    if (owner) {
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t c = cnt[w][b];
            cnt[w][b] = digit_total;
            digit_total += c;
        }
    }
    // publish this tile's count first, so the tiles behind it can look through it while it is still looking back
    uint32_t *my_status = status + (size_t) tile * RS_BINS + b;
    if (owner) __stcg(my_status, (tile == 0 ? OS_FLAG_PREFIX : OS_FLAG_COUNT) | digit_total);
    uint32_t tile_total;
    const uint32_t digit_start = block_excl_scan<THREADS>(digit_total, &tile_total, wsum);
    if (owner) {
#pragma unroll
        for (int w = 0; w < WARPS; ++w) cnt[w][b] += digit_start;
    }
    // decoupled look-back: bodies of digit b in the tiles before this one, four status words in flight at a time
    uint32_t before = 0;
    if (owner && tile > 0) {
        uint32_t t = tile;
        bool done = false;
        while (!done) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                w[j] = t > (uint32_t) j ? *(const volatile uint32_t *) (status + (size_t) (t - 1 - j) * RS_BINS + b) : OS_FLAG_PREFIX;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!done) {
                    while ((w[j] & ~OS_VALUE_MASK) == 0) w[j] = *(const volatile uint32_t *) (status + (size_t) (t - 1 - j) * RS_BINS + b);
                    before += w[j] & OS_VALUE_MASK;
                    if (w[j] & OS_FLAG_PREFIX) done = true;
                }
            }
            t = t > 4 ? t - 4 : 0;
        }
        __stcg(my_status, OS_FLAG_PREFIX | (before + digit_total));
    }
    if (owner) gbase[b] = gstart[b] + before - digit_start;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint64_t i = wbase + (uint64_t) k * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t) ((key[k] >> shift) & 0xff);
            const uint32_t pos = cnt[warp][d] + rank[k];
            skey[pos] = key[k];
            if (!WORDS) sval[pos] = MODE == OS_PAIRS_IOTA ? (uint32_t) i : vals_in[i];
        }
    }
    __syncthreads();
    const uint32_t tile_count = (uint32_t) (n - tile_base < (uint64_t) TILE ? n - tile_base : (uint64_t) TILE);
    for (uint32_t pos = threadIdx.x; pos < tile_count; pos += THREADS) {
        const uint64_t kk = skey[pos];
        const uint32_t d = (uint32_t) ((kk >> shift) & 0xff);
        const uint64_t g = (uint64_t) gbase[d] + pos;
        keys_out[g] = kk;
        if (!WORDS) vals_out[g] = sval[pos];
    }
}

// tile geometry: OS_THREADS x OS_ITEMS keys per tile
#ifndef NB_OS_THREADS
#define NB_OS_THREADS 256
#endif
#ifndef NB_OS_ITEMS
#define NB_OS_ITEMS 8
#endif
constexpr int OS_THREADS = NB_OS_THREADS, OS_ITEMS = NB_OS_ITEMS;
inline uint32_t os_tiles_for(uint64_t n) { return (uint32_t) ((n + OS_THREADS * OS_ITEMS - 1) / (OS_THREADS * OS_ITEMS)); }
// scratch in uint32 elements: [passes][256] histograms, 64 words of tickets, one status table per pass
inline size_t os_scratch_elems(uint64_t n) {
    return (size_t) OS_MAX_PASSES * RS_BINS + 64 + (size_t) OS_MAX_PASSES * os_tiles_for(n) * RS_BINS;
}
constexpr size_t os_smem_bytes(int threads, int items, bool words = false) {
    return (size_t) threads * items * (words ? 8 : 12) + (size_t) ((threads / 32) * RS_BINS + RS_BINS + threads / 32 + 1 + 3) * 4;
}

// same contract as radix_sort_pairs
inline int onesweep_sort_pairs(nb_ctx *ctx, uint64_t *keys_a, uint32_t *vals_a, uint64_t *keys_b, uint32_t *vals_b,
                               uint64_t n, int key_bits, uint32_t *scratch, uint64_t **keys_sorted,
                               uint32_t **vals_sorted, bool iota_first) {
    uint64_t *kin = keys_a, *kout = keys_b;
    uint32_t *vin = vals_a, *vout = vals_b;
    if (n == 0) { *keys_sorted = kin; *vals_sorted = vin; return NB_OK; }
    const int passes = (key_bits + 7) / 8;
    if (passes > OS_MAX_PASSES) return nb_fail(ctx, NB_ERR_INVALID, "onesweep_sort_pairs: more than 64 key bits");
    const uint32_t tiles = os_tiles_for(n);
    uint32_t *ghist = scratch;
    uint32_t *tickets = scratch + (size_t) OS_MAX_PASSES * RS_BINS;
    uint32_t *status = tickets + 64;
    NB_CUDA(ctx, cudaMemsetAsync(scratch, 0, ((size_t) OS_MAX_PASSES * RS_BINS + 64 + (size_t) passes * tiles * RS_BINS) * sizeof(uint32_t),
                                 ctx->stream));
    const unsigned hgrid = (unsigned) std::min<uint64_t>((n + 2047) / 2048, (uint64_t) ctx->sm_count * 8);
    os_hist_kernel<<<hgrid, 256, 0, ctx->stream>>>(kin, n, passes, 0, 0, ghist);
    NB_LAUNCH_CHECK(ctx);
    os_scan_kernel<<<1, RS_BINS, 0, ctx->stream>>>(ghist, passes);
    NB_LAUNCH_CHECK(ctx);
    constexpr size_t smem = os_smem_bytes(OS_THREADS, OS_ITEMS);
    static_assert(smem <= 48 * 1024, "one-sweep tile must fit the default shared-memory window");
    for (int p = 0; p < passes; ++p) {
        uint32_t *st = status + (size_t) p * tiles * RS_BINS;
        if (p == 0 && iota_first)
            os_scatter_kernel<OS_THREADS, OS_ITEMS, OS_PAIRS_IOTA><<<tiles, OS_THREADS, smem, ctx->stream>>>(
                kin, vin, n, 8 * p, ghist + p * RS_BINS, st, tickets + p, kout, vout, 0);
        else
            os_scatter_kernel<OS_THREADS, OS_ITEMS, OS_PAIRS><<<tiles, OS_THREADS, smem, ctx->stream>>>(
                kin, vin, n, 8 * p, ghist + p * RS_BINS, st, tickets + p, kout, vout, 0);
        NB_LAUNCH_CHECK(ctx);
        uint64_t *tk = kin; kin = kout; kout = tk;
        uint32_t *tv = vin; vin = vout; vout = tv;
    }
    *keys_sorted = kin;
    *vals_sorted = vin;
    return NB_OK;
}

// Packed form for the Barnes-Hut build: ONE 64-bit word per body, {upper bits of the 63-bit key | storage slot in the
// low idx_bits}, sorted on its upper `passes` x 8 bits only.  A pass moves 16 B per body instead of 24 B, and 5 passes
// (40 key bits = 13 octree levels) replace 8; bodies that agree on all sorted bits are put in order afterwards from
// their full keys (bh_build.cu).  Sorting the words is sorting (key prefix, slot): stable by construction.
// keys: plain keys (read by the first pass only); words_a / words_b: ping-pong buffers.  Returns the sorted words.
// (Tiles of 12 and 16 keys per thread were measured against the 8 used here: 2.095 / 2.107 / 2.102 ms for the whole sort
// phase at N = 2^24 -- the pass is bound by the ranking chain per key, not by the per-tile look-back.)
inline int onesweep_sort_packed(nb_ctx *ctx, const uint64_t *keys, uint64_t *words_a, uint64_t *words_b, uint64_t n,
                                int idx_bits, int passes, uint32_t *scratch, uint64_t **words_sorted) {
    constexpr int ITEMS = OS_ITEMS;
    if (passes < 1 || passes > OS_MAX_PASSES) return nb_fail(ctx, NB_ERR_INVALID, "onesweep_sort_packed: bad pass count");
    const int first_shift = 64 - 8 * passes;
    const uint32_t tiles = (uint32_t) ((n + OS_THREADS * ITEMS - 1) / (OS_THREADS * ITEMS));
    uint32_t *ghist = scratch;
    uint32_t *tickets = scratch + (size_t) OS_MAX_PASSES * RS_BINS;
    uint32_t *status = tickets + 64;
    NB_CUDA(ctx, cudaMemsetAsync(scratch, 0, ((size_t) OS_MAX_PASSES * RS_BINS + 64 + (size_t) passes * tiles * RS_BINS) * sizeof(uint32_t),
                                 ctx->stream));
    const unsigned hgrid = (unsigned) std::min<uint64_t>((n + 2047) / 2048, (uint64_t) ctx->sm_count * 8);
    os_hist_kernel<<<hgrid, 256, 0, ctx->stream>>>(keys, n, passes, first_shift, idx_bits, ghist);
    NB_LAUNCH_CHECK(ctx);
    os_scan_kernel<<<1, RS_BINS, 0, ctx->stream>>>(ghist, passes);
    NB_LAUNCH_CHECK(ctx);
    constexpr size_t smem = os_smem_bytes(OS_THREADS, ITEMS, true);
    static_assert(smem <= 48 * 1024, "one-sweep tile must fit the default shared-memory window");
    const uint64_t *kin = keys;
    uint64_t *kout = words_a;
    for (int p = 0; p < passes; ++p) {
        uint32_t *st = status + (size_t) p * tiles * RS_BINS;
        if (p == 0)
            os_scatter_kernel<OS_THREADS, ITEMS, OS_WORDS_PACK><<<tiles, OS_THREADS, smem, ctx->stream>>>(
                kin, nullptr, n, first_shift + 8 * p, ghist + p * RS_BINS, st, tickets + p, kout, nullptr, idx_bits);
        else
            os_scatter_kernel<OS_THREADS, ITEMS, OS_WORDS><<<tiles, OS_THREADS, smem, ctx->stream>>>(
                kin, nullptr, n, first_shift + 8 * p, ghist + p * RS_BINS, st, tickets + p, kout, nullptr, idx_bits);
        NB_LAUNCH_CHECK(ctx);
        kin = kout;
        kout = kout == words_a ? words_b : words_a;
    }
    *words_sorted = const_cast<uint64_t *>(kin);
    return NB_OK;
}

}  // namespace
}  // namespace nbprim
