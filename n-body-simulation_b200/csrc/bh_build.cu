// Barnes-Hut tree pipeline for sm_100a: AABB -> octant-path keys -> radix sort -> node construction -> centre of mass.
//
// Replaces BarnesHutOctree::buildOctree and everything below it (reference
// src/simulationBackend/BarnesHutTreeAlgorithms/ParallelOctreeTopDownSubtrees.cpp:15-813, BarnesHutOctree.cpp:45-613).
// The reference inserts bodies one by one under per-node spin locks (one work-group for the top tree, one per
// subtree) and computes the centre of mass in a single spinning work-group.  None of that is kept.  What IS kept,
// bit for bit, is the canonical tree those kernels produce:
//   * AABB seeded with 0.0 (origin always inside), grown to a cube on the two shorter axes;
//   * a cell with >= 2 bodies is internal and has all 8 children; leaves hold 0 or 1 body;
//   * octant = 4*(y > mid) + 2*(x > mid) + 1*(z < mid), cell bounds by repeated fp64 `min + edge/2` down the path;
//   * centre of mass = sum over children in octant order 0..7 of (m*x, m*y, m*z, m).
//
// B200-first construction (lock-free, O(N) work, no host round trips):
//   1. every body descends the cube arithmetically (exactly the reference's fp64 compares) and records its path as a
//      63-bit key, 3 bits per level, 21 levels; a second key word (levels 21..41) is computed only for the rare bodies
//      that share all 21 upper levels with a neighbour (closer than edge*2^-21).  The digit is the octant's VISIT RANK 4u+2b+(1-r) in the
//      reference's traversal order [2,0,3,1,6,4,7,5] (BarnesHutAlgorithm.cpp:370-385), so sorted order == DFS order.
//   2. stable LSD radix sort of (key_hi, slot) (scan_sort.cuh); rare equal-key_hi runs are ordered by key_lo; then the
//      whole body state (m, x, v, a, id) is physically permuted into the sorted order, so everything downstream streams
//      and the next step's permutation is a near-identity map (bodies move little per step).
//   3. delta[i] = common-prefix digits of sorted neighbours.  Body i is the FIRST body of the internal cells of depth
//      delta[i-1]+1 .. delta[i]; an exclusive scan of those counts gives every node its index in a DFS pre-order
//      array (internal chain of body i, then leaf i).  Empty leaves are implied, never materialised.
//   4. one thread per body emits its chain + leaf: skip links (first node after the subtree) by galloping searches
//      over the sorted keys; each node records its visit rank inside its parent.
//   5. centre of mass level by level, deepest first (one launch per depth, no atomics): a node sums its children in
//      octant order (deterministic, bit-identical to the reference's order).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cooperative_groups.h>

#include "bh_accept.cuh"
#include "scan_sort.cuh"

#define NB_NONE 0xffffffffu
#define NB_FLAG_DEPTH 1u
#define NB_FLAG_POOL 2u
#define NB_LEVELS 64

namespace {

// ---- 1. AABB ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
aabb_partial_kernel(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
                    uint64_t n, double *__restrict__ partial /* [6][gridDim.x] */, uint32_t *__restrict__ flags) {
    if (blockIdx.x == 0 && threadIdx.x == 0) flags[6] = 0;   // run statistic of this build's sort (tie_fix_kernel)
    // scratch starts at 0.0 like the reference's value-initialised per-work-item arrays (BarnesHutOctree.cpp:58-72)
    double mnx = 0, mny = 0, mnz = 0, mxx = 0, mxy = 0, mxz = 0;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const double a = x[i], b = y[i], c = z[i];
        mnx = fmin(mnx, a); mxx = fmax(mxx, a);
        mny = fmin(mny, b); mxy = fmax(mxy, b);
        mnz = fmin(mnz, c); mxz = fmax(mxz, c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        mnz = fmin(mnz, __shfl_xor_sync(0xffffffffu, mnz, o)); mxz = fmax(mxz, __shfl_xor_sync(0xffffffffu, mxz, o));
    }
    __shared__ double s[6][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s[0][warp] = mnx; s[1][warp] = mny; s[2][warp] = mnz; s[3][warp] = mxx; s[4][warp] = mxy; s[5][warp] = mxz; }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = s[threadIdx.x][0];
        for (int w = 1; w < 8; ++w) v = threadIdx.x < 3 ? fmin(v, s[threadIdx.x][w]) : fmax(v, s[threadIdx.x][w]);
        partial[(size_t) threadIdx.x * gridDim.x + blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(256)
aabb_final_kernel(const double *__restrict__ partial, int n_partials, double theta, double *__restrict__ out /* 8 + 64 */) {
    __shared__ double s[6][256];
    __shared__ double s_edge;
    for (int c = 0; c < 6; ++c) {
        double v = 0.0;
        for (int i = threadIdx.x; i < n_partials; i += 256) {
            const double p = partial[(size_t) c * n_partials + i];
            v = c < 3 ? fmin(v, p) : fmax(v, p);
        }
        s[c][threadIdx.x] = v;
    }
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if ((int) threadIdx.x < st)
            for (int c = 0; c < 6; ++c)
                s[c][threadIdx.x] = c < 3 ? fmin(s[c][threadIdx.x], s[c][threadIdx.x + st])
                                          : fmax(s[c][threadIdx.x], s[c][threadIdx.x + st]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double min_x = s[0][0], min_y = s[1][0], min_z = s[2][0], max_x = s[3][0], max_y = s[4][0], max_z = s[5][0];
        // cube growth, BarnesHutOctree.cpp:162-190 (same operations, same tie order x, y, z)
        const double x_length = fabs(__dsub_rn(max_x, min_x));
        const double y_length = fabs(__dsub_rn(max_y, min_y));
        const double z_length = fabs(__dsub_rn(max_z, min_z));
        const double maxEdgeLength = fmax(x_length, fmax(y_length, z_length));
        const double gx = __ddiv_rn(__dsub_rn(maxEdgeLength, x_length), 2.0);
        const double gy = __ddiv_rn(__dsub_rn(maxEdgeLength, y_length), 2.0);
        const double gz = __ddiv_rn(__dsub_rn(maxEdgeLength, z_length), 2.0);
        if (maxEdgeLength == x_length) {
            min_z = __dsub_rn(min_z, gz); min_y = __dsub_rn(min_y, gy);
            max_z = __dadd_rn(max_z, gz); max_y = __dadd_rn(max_y, gy);
        } else if (maxEdgeLength == y_length) {
            min_x = __dsub_rn(min_x, gx); min_z = __dsub_rn(min_z, gz);
            max_x = __dadd_rn(max_x, gx); max_z = __dadd_rn(max_z, gz);
        } else {
            min_x = __dsub_rn(min_x, gx); min_y = __dsub_rn(min_y, gy);
            max_x = __dadd_rn(max_x, gx); max_y = __dadd_rn(max_y, gy);
        }
        out[0] = min_x; out[1] = min_y; out[2] = min_z;
        out[3] = max_x; out[4] = max_y; out[5] = max_z;
        out[6] = maxEdgeLength;
        s_edge = maxEdgeLength;
    }
    // the walk's exact acceptance thresholds of this box, one per depth (bh_accept.cuh)
    __syncthreads();
    if (threadIdx.x < NB_ACCEPT_DEPTHS)
        out[NB_ACCEPT_TABLE_OFFSET + threadIdx.x] = nb_accept_threshold(s_edge, theta, threadIdx.x);
}

// the same table for a theta that changed after the build (nb_set_theta between nb_bh_build and nb_bh_accel)
__global__ void __launch_bounds__(NB_ACCEPT_DEPTHS)
accept_table_kernel(double theta, double *__restrict__ aabb) {
    aabb[NB_ACCEPT_TABLE_OFFSET + threadIdx.x] = nb_accept_threshold(aabb[6], theta, threadIdx.x);
}

// ---- 2. octant-path keys --------------------------------------------------------------------------------------------
// Descent test of ParallelOctreeTopDownSubtrees.cpp:400-406 with the child bounds of :256-315.  One call advances the
// cell (mn*, edge) by 21 levels and returns the 63-bit key word of those levels.
__device__ __forceinline__ uint64_t descend21(double px, double py, double pz, double &mnx, double &mny, double &mnz,
                                             double &edge) {
    uint64_t k = 0;
#pragma unroll
    for (int l = 0; l < 21; ++l) {
        const double h = __dmul_rn(edge, 0.5);           // parentEdgeLength / 2 (exact halving)
        const double midx = __dadd_rn(mnx, h);
        const double midy = __dadd_rn(mny, h);
        const double midz = __dadd_rn(mnz, h);
        const bool upper = py > midy;
        const bool right = px > midx;
        const bool back = pz < midz;
        // visit-rank digit: 4u + 2b + (1-r)
        const uint64_t digit = (upper ? 4u : 0u) | (back ? 2u : 0u) | (right ? 0u : 1u);
        k = (k << 3) | digit;
        mny = upper ? midy : mny;
        mnx = right ? midx : mnx;
        mnz = back ? mnz : midz;                         // z gets +h when back == 0 (:256-315)
        edge = h;
    }
    return k;
}

// levels 0..20 of every body (the sort key)
__global__ void __launch_bounds__(256)
keys_kernel(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z, uint64_t n,
            const double *__restrict__ aabb, uint64_t *__restrict__ key_hi) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double mnx = aabb[0], mny = aabb[1], mnz = aabb[2], edge = aabb[6];
    key_hi[i] = descend21(x[i], y[i], z[i], mnx, mny, mnz, edge);
}

// levels 21..41, only ever needed for bodies that share all 21 upper levels with a neighbour (closer than edge*2^-21)
__device__ __forceinline__ uint64_t key_lo_of(double px, double py, double pz, const double *__restrict__ aabb) {
    double mnx = aabb[0], mny = aabb[1], mnz = aabb[2], edge = aabb[6];
    descend21(px, py, pz, mnx, mny, mnz, edge);
    return descend21(px, py, pz, mnx, mny, mnz, edge);
}

// ---- 2b. after the sort: slots and full keys of the packed words; order runs the sort left undecided ----------------------
// packed words (scan_sort.cuh os_pack) in sorted order -> perm (storage slot of sorted body i) and the body's full key
__global__ void __launch_bounds__(256)
unpack_kernel(const uint64_t *__restrict__ words, const uint64_t *__restrict__ key_by_slot, uint64_t n, int idx_bits,
              uint32_t *__restrict__ perm, uint64_t *__restrict__ hi_sorted) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = (uint32_t) (words[i] & ((1ull << idx_bits) - 1ull));
    perm[i] = p;
    hi_sorted[i] = key_by_slot[p];
}

#define NB_PACKED_KEY_SHIFT 23   /* a 5-pass packed sort orders the upper 40 of the 63 key bits: bits 23..62 */
#define NB_PACKED4_KEY_SHIFT 31  /* a 4-pass packed sort orders the upper 32: bits 31..62 (10 octree levels and a bit) */
#define NB_PACKED_RUN_LIMIT 64   /* longer undecided runs make the host take more passes (4 -> 5 -> the full 8-pass sort) */

// Bodies whose keys agree above `run_shift` form a run the sort left in slot order (run_shift = 23 after the packed
// sort, 0 after the full sort: only bodies closer than edge * 2^-21 remain).  The head of each run orders it by the
// full (key_hi, key_lo) with an insertion sort -- runs are rare and short (two or three bodies) unless thousands of
// bodies share a 13-level cell, which flags[3] reports so that the host uses the full sort from the next build on.
// key_lo (levels 21..41) is computed on demand, only for bodies that agree on all of key_hi.
__global__ void __launch_bounds__(256)
tie_fix_kernel(uint64_t *__restrict__ hi_sorted, const double *__restrict__ x, const double *__restrict__ y,
               const double *__restrict__ z, const double *__restrict__ aabb, uint32_t *__restrict__ perm, uint64_t n,
               int run_shift, uint32_t *__restrict__ flags) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    const uint64_t hi_i = hi_sorted[i];
    // statistics for the host's choice of sort: longest run of equal upper-40-bit prefixes (flags[3]) and of equal
    // upper-32-bit prefixes (flags[6]), counted up to the limit + 1
    {
        const uint64_t h_next = hi_sorted[i + 1], h_prev = i > 0 ? hi_sorted[i - 1] : ~hi_i;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const int sh = w ? NB_PACKED4_KEY_SHIFT : NB_PACKED_KEY_SHIFT;
            uint32_t *stat = w ? &flags[6] : &flags[3];
            const uint64_t k = hi_i >> sh;
            if ((h_next >> sh) == k && (h_prev >> sh) != k) {
                uint32_t len = 2;
                while (len <= NB_PACKED_RUN_LIMIT && i + len < n && (hi_sorted[i + len] >> sh) == k) ++len;
                if (len > *(volatile uint32_t *) stat) atomicMax(stat, len);
            }
        }
    }
    const uint64_t k = hi_i >> run_shift;
    if ((hi_sorted[i + 1] >> run_shift) != k) return;
    if (i > 0 && (hi_sorted[i - 1] >> run_shift) == k) return;
    uint64_t e = i + 1;
    while (e + 1 < n && (hi_sorted[e + 1] >> run_shift) == k) ++e;
    for (uint64_t a = i + 1; a <= e; ++a) {  // insertion sort of the run by (key_hi, key_lo)
        const uint32_t pa = perm[a];
        const uint64_t ha = hi_sorted[a];
        uint64_t la = 0;
        bool have_la = false;
        uint64_t b = a;
        while (b > i) {
            const uint64_t hb = hi_sorted[b - 1];
            if (hb < ha) break;
            const uint32_t pb = perm[b - 1];
            if (hb == ha) {
                if (!have_la) { la = key_lo_of(x[pa], y[pa], z[pa], aabb); have_la = true; }
                const uint64_t lb = key_lo_of(x[pb], y[pb], z[pb], aabb);
                if (lb == la) atomicOr(&flags[0], NB_FLAG_DEPTH);  // identical 42-level paths: coincident bodies
                if (lb <= la) break;
            }
            perm[b] = pb;
            hi_sorted[b] = hb;
            --b;
        }
        perm[b] = pa;
        hi_sorted[b] = ha;
    }
}

// lower key word in sorted order: computed where a neighbour shares the upper word, 0 elsewhere (never consulted there)
__global__ void __launch_bounds__(256)
keys_lo_kernel(const uint64_t *__restrict__ hi_sorted, const double *__restrict__ x, const double *__restrict__ y,
               const double *__restrict__ z, const double *__restrict__ aabb, uint64_t n, uint64_t *__restrict__ lo_sorted) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = hi_sorted[i];
    const bool tied = (i > 0 && hi_sorted[i - 1] == k) || (i + 1 < n && hi_sorted[i + 1] == k);
    lo_sorted[i] = tied ? key_lo_of(x[i], y[i], z[i], aabb) : 0ull;
}

// ---- 3. physical reorder of the whole state into sorted order ----------------------------------------------------------------
struct reorder_args {
    const double *src[10];
    double *dst[10];
};
__global__ void __launch_bounds__(256)
reorder_kernel(const uint32_t *__restrict__ perm, uint64_t n, int count, reorder_args a, const uint32_t *__restrict__ id_in,
               uint32_t *__restrict__ id_out, int identity) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = perm[i];
#pragma unroll
    for (int k = 0; k < 10; ++k)
        if (k < count) a.dst[k][i] = a.src[k][p];
    id_out[i] = identity ? p : id_in[p];
}

// read-back / upload helpers between storage order and body-id order
__global__ void __launch_bounds__(256)
unpermute_kernel(uint64_t n, const uint32_t *__restrict__ id, int count, reorder_args a) {
    const uint64_t s = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t b = id[s];
    for (int k = 0; k < count; ++k) a.dst[k][b] = a.src[k][s];
}
__global__ void __launch_bounds__(256)
permute_in_kernel(uint64_t n, const uint32_t *__restrict__ id, int count, reorder_args a) {
    const uint64_t s = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t b = id[s];
    for (int k = 0; k < count; ++k) a.dst[k][s] = a.src[k][b];
}

__device__ __forceinline__ int common_digits(uint64_t hi_a, uint64_t lo_a, uint64_t hi_b, uint64_t lo_b) {
    const uint64_t xh = hi_a ^ hi_b;
    if (xh) return (__clzll((long long) xh) - 1) / 3;
    const uint64_t xl = lo_a ^ lo_b;
    if (xl) return 21 + (__clzll((long long) xl) - 1) / 3;
    return NB_MAX_TREE_DEPTH;
}

// does sorted body j share its first d digits with the key (hi_i, lo_i)?  the lower key word of j is only loaded when the question
// reaches below level 21 (the searches of emit_kernel are dependent loads; nearly all of them stop at the upper word)
__device__ __forceinline__ bool shares_prefix_with(const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo, uint64_t j,
                                                   uint64_t hi_i, uint64_t lo_i, int d) {
    const uint64_t hj = hi[j];
    if (d <= 21) return ((hj ^ hi_i) >> (63 - 3 * d)) == 0;
    return hj == hi_i && ((lo[j] ^ lo_i) >> (63 - 3 * (d - 21))) == 0;
}

__device__ __forceinline__ uint32_t digit_at(uint64_t hi, uint64_t lo, int level) {
    return level < 21 ? (uint32_t) ((hi >> (60 - 3 * level)) & 7) : (uint32_t) ((lo >> (60 - 3 * (level - 21))) & 7);
}

// ---- 3b. neighbour prefix lengths, per-body internal-node counts, per-depth node counts ----------------------------------
// level[0..63] += internal nodes per depth (the sizes of the dense per-depth work lists of the centre-of-mass pass)
__global__ void __launch_bounds__(256)
delta_kernel(const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo, uint64_t n, int32_t *__restrict__ delta,
             uint32_t *__restrict__ cnt, uint32_t *__restrict__ flags, uint32_t *__restrict__ level) {
    __shared__ uint32_t lcnt[NB_LEVELS];
    __shared__ int smax;
    if (threadIdx.x < NB_LEVELS) lcnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) smax = -1;
    __syncthreads();
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    int d_cur = -1;
    if (i < n) {
        // the lower key words matter only between bodies that agree on the whole upper word
        const uint64_t h = hi[i];
        const uint64_t hp = i > 0 ? hi[i - 1] : 0, hn = i + 1 < n ? hi[i + 1] : 0;
        const bool tie_p = i > 0 && hp == h, tie_n = i + 1 < n && hn == h;
        const uint64_t l = (tie_p || tie_n) ? lo[i] : 0;
        const int d_prev = i > 0 ? common_digits(hp, tie_p ? lo[i - 1] : 0, h, l) : -1;
        d_cur = i + 1 < n ? common_digits(h, l, hn, tie_n ? lo[i + 1] : 0) : -1;
        delta[i] = d_cur;
        cnt[i] = d_cur > d_prev ? (uint32_t) (d_cur - d_prev) : 0u;
        if (d_cur >= NB_MAX_TREE_DEPTH) atomicOr(&flags[0], NB_FLAG_DEPTH);
        for (int d = d_prev + 1; d <= d_cur && d < NB_LEVELS; ++d) atomicAdd(&lcnt[d], 1u);
    }
    // max depth of the tree = deepest leaf = max(delta) + 1: one atomic per block, and only when it raises the value
    int mx = d_cur;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx >= 0) atomicMax(&smax, mx);
    __syncthreads();
    if (threadIdx.x == 0 && smax >= 0 && (uint32_t) (smax + 1) > *(volatile uint32_t *) &flags[2])
        atomicMax(&flags[2], (uint32_t) (smax + 1));
    if (threadIdx.x < NB_LEVELS && lcnt[threadIdx.x]) atomicAdd(&level[threadIdx.x], lcnt[threadIdx.x]);
}

// ---- 4. node emission ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
emit_kernel(const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo, const int32_t *__restrict__ delta,
            const uint32_t *__restrict__ base, uint64_t n, uint64_t cap_nodes, uint32_t *__restrict__ flags,
            uint2 *__restrict__ meta, uint32_t *__restrict__ body_count, uint32_t *__restrict__ leaf_node) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t total_internal = flags[1];
    const uint64_t M = n + total_internal;
    if (M > cap_nodes) {  // reference: silent overflow of the 16N pool (README.md:117-118); here: reported
        if (i == 0) atomicOr(&flags[0], NB_FLAG_POOL);
        return;
    }
    const uint64_t hi_i = hi[i], lo_i = lo[i];
    const int d_prev = i > 0 ? delta[i - 1] : -1;
    const int d_cur = delta[i];
    const uint32_t c = d_cur > d_prev ? (uint32_t) (d_cur - d_prev) : 0u;
    const uint64_t b = base[i];
    const uint64_t head = i + b;          // first node that starts at body i
    const uint64_t leaf = head + c;

    // chain of internal nodes first-bodied by i, deepest first so the right boundary only moves outwards
    uint64_t r = i + 1;  // i+1 shares d_cur digits when c > 0
    for (int k = (int) c - 1; k >= 0; --k) {
        const int d = d_prev + 1 + k;
        uint64_t step = 1;
        while (r + step < n && shares_prefix_with(hi, lo, r + step, hi_i, lo_i, d)) { r += step; step <<= 1; }
        while (step > 1) {
            step >>= 1;
            if (r + step < n && shares_prefix_with(hi, lo, r + step, hi_i, lo_i, d)) r += step;
        }
        const uint64_t node = head + k;
        const uint64_t skip = r + 1 < n ? (r + 1) + base[r + 1] : M;
        const uint32_t dig = d > 0 ? digit_at(hi_i, lo_i, d - 1) : 0u;  // visit rank of this cell inside its parent
        meta[node] = make_uint2((uint32_t) skip, ((uint32_t) d << NB_DEPTH_SHIFT) | dig);
        body_count[node] = (uint32_t) (r - i + 1);
    }
    const int leaf_parent_depth = d_cur > d_prev ? d_cur : d_prev;  // -1 only when N == 1
    const uint32_t leaf_dig = leaf_parent_depth >= 0 ? digit_at(hi_i, lo_i, leaf_parent_depth) : 0u;
    meta[leaf] = make_uint2((uint32_t) (leaf + 1), NB_LEAF_FLAG | (leaf_dig << NB_DIGIT_SHIFT) | (uint32_t) i);
    body_count[leaf] = 1;
    leaf_node[i] = (uint32_t) leaf;
}

// The same emission with the chain nodes of a warp's 32 bodies spread evenly over its lanes.  In emit_kernel a thread
// loops over the internal nodes first-bodied by its body: 62 % of the bodies have none, a few have three or more, and each
// node costs a galloping search of its own length -- ncu: 6 of 32 lanes active, issue bound.  Here every lane first writes
// its body's leaf, then the warp lists its (body, chain index) items in shared memory and each lane takes one item per
// round: one search per lane, lanes busy as long as there are items.  Same nodes, same values; 0.83 -> 0.73 ms for the
// "Build subtrees" phase at N = 2^24 (environment NB_EMIT_PER_BODY=1 keeps the per-body loop for A/B).
__global__ void __launch_bounds__(128)
emit_balanced_kernel(const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo, const int32_t *__restrict__ delta,
                     const uint32_t *__restrict__ base, uint64_t n, uint64_t cap_nodes, uint32_t *__restrict__ flags,
                     uint2 *__restrict__ meta, uint32_t *__restrict__ body_count, uint32_t *__restrict__ leaf_node) {
    constexpr int SLOTS = 128;                       // items listed per pass and warp
    __shared__ uint16_t s_owner[4][SLOTS];           // item -> source lane | chain index << 5
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total_internal = flags[1];
    const uint64_t M = n + total_internal;
    if (M > cap_nodes) {  // reference: silent overflow of the 16N pool (README.md:117-118); here: reported
        if (i == 0) atomicOr(&flags[0], NB_FLAG_POOL);
        return;
    }
    const bool valid = i < n;
    uint64_t hi_i = 0;
    int d_prev = -1, d_cur = -1;
    uint32_t c = 0;
    uint64_t head = 0;
    if (valid) {
        hi_i = hi[i];
        d_prev = i > 0 ? delta[i - 1] : -1;
        d_cur = delta[i];
        c = d_cur > d_prev ? (uint32_t) (d_cur - d_prev) : 0u;
        head = i + base[i];
        const uint64_t leaf = head + c;
        const int leaf_parent_depth = d_cur > d_prev ? d_cur : d_prev;  // -1 only when N == 1
        const uint64_t lo_i = leaf_parent_depth >= 21 ? lo[i] : 0ull;   // the lower key word only below level 21
        const uint32_t leaf_dig = leaf_parent_depth >= 0 ? digit_at(hi_i, lo_i, leaf_parent_depth) : 0u;
        meta[leaf] = make_uint2((uint32_t) (leaf + 1), NB_LEAF_FLAG | (leaf_dig << NB_DIGIT_SHIFT) | (uint32_t) i);
        body_count[leaf] = 1;
        leaf_node[i] = (uint32_t) leaf;
    }
    // items of the warp: exclusive prefix of the chain lengths
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const uint32_t off = inc - c, W = __shfl_sync(0xffffffffu, inc, 31);
    const uint32_t ilo = (uint32_t) i, ihi = (uint32_t) (i >> 32);
    for (uint32_t pass = 0; pass < W; pass += SLOTS) {
        __syncwarp();
        for (uint32_t k = 0; k < c; ++k) {
            const uint32_t t = off + k;
            if (t >= pass && t < pass + SLOTS) s_owner[w][t - pass] = (uint16_t) (lane | (k << 5));
        }
        __syncwarp();
        const uint32_t count = W - pass < (uint32_t) SLOTS ? W - pass : (uint32_t) SLOTS;
        for (uint32_t t0 = 0; t0 < count; t0 += 32) {     // warp-uniform trip count: the shuffles need every lane
            const uint32_t t = t0 + lane;
            const bool have = t < count;
            const uint32_t o = have ? s_owner[w][t] : 0u;
            const int src = (int) (o & 31u), k = (int) (o >> 5);
            const uint64_t body = ((uint64_t) __shfl_sync(0xffffffffu, ihi, src) << 32) | __shfl_sync(0xffffffffu, ilo, src);
            const uint64_t key = ((uint64_t) __shfl_sync(0xffffffffu, (uint32_t) (hi_i >> 32), src) << 32) |
                                 __shfl_sync(0xffffffffu, (uint32_t) hi_i, src);
            const int dp = __shfl_sync(0xffffffffu, d_prev, src);
            const uint32_t head_lo = __shfl_sync(0xffffffffu, (uint32_t) head, src), head_hi = __shfl_sync(0xffffffffu, (uint32_t) (head >> 32), src);
            if (!have) continue;
            const int d = dp + 1 + k;
            const uint64_t key_lo = d > 21 ? lo[body] : 0ull;
            uint64_t r = body + 1, step = 1;   // body + 1 shares d_cur >= d digits
            while (r + step < n && shares_prefix_with(hi, lo, r + step, key, key_lo, d)) { r += step; step <<= 1; }
            while (step > 1) {
                step >>= 1;
                if (r + step < n && shares_prefix_with(hi, lo, r + step, key, key_lo, d)) r += step;
            }
            const uint64_t node = (((uint64_t) head_hi << 32) | head_lo) + (uint64_t) k;
            const uint64_t skip = r + 1 < n ? (r + 1) + base[r + 1] : M;
            const uint32_t dig = d > 0 ? digit_at(key, key_lo, d - 1) : 0u;  // visit rank of this cell inside its parent
            meta[node] = make_uint2((uint32_t) skip, ((uint32_t) d << NB_DEPTH_SHIFT) | dig);
            body_count[node] = (uint32_t) (r - body + 1);
        }
    }
}

// ---- 4b. per-depth lists of internal nodes (dense work lists for the centre-of-mass levels) ---------------------------------
// level[0..63] = node count per depth (delta_kernel), level[64..127] = start of the depth's segment in `list`,
// level[128..191] = fill cursor
__global__ void level_scan_kernel(uint32_t *__restrict__ level) {
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int d = 0; d < NB_LEVELS; ++d) {
            level[NB_LEVELS + d] = run;
            level[2 * NB_LEVELS + d] = 0;
            run += level[d];
        }
    }
}

__global__ void __launch_bounds__(256)
level_fill_kernel(const int32_t *__restrict__ delta, const uint32_t *__restrict__ base, uint64_t n,
                  const uint32_t *__restrict__ flags, uint32_t *__restrict__ level, uint32_t *__restrict__ list) {
    __shared__ uint32_t cnt[NB_LEVELS], start[NB_LEVELS];
    if (flags[0] & NB_FLAG_POOL) return;
    if (threadIdx.x < NB_LEVELS) cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    int d_prev = 0, d_cur = -1;
    if (i < n) {
        d_prev = i > 0 ? delta[i - 1] : -1;
        d_cur = delta[i];
        for (int d = d_prev + 1; d <= d_cur; ++d) atomicAdd(&cnt[d], 1u);
    }
    __syncthreads();
    if (threadIdx.x < NB_LEVELS) {
        const uint32_t c = cnt[threadIdx.x];
        start[threadIdx.x] = level[NB_LEVELS + threadIdx.x] + (c ? atomicAdd(&level[2 * NB_LEVELS + threadIdx.x], c) : 0u);
        cnt[threadIdx.x] = 0;
    }
    __syncthreads();
    if (i < n) {
        const uint32_t head = (uint32_t) (i + base[i]);
        for (int d = d_prev + 1; d <= d_cur; ++d)
            list[start[d] + atomicAdd(&cnt[d], 1u)] = head + (uint32_t) (d - d_prev - 1);
    }
}

// ---- 5. centre of mass ------------------------------------------------------------------------------------------------------
// leaf: prepareCenterOfMass (BarnesHutOctree.cpp:216-226); internal: the octant-ordered sum of :299-317.
// Level-synchronous, deepest level first: one launch per depth, one thread per body; the thread owns the internal node
// of that depth whose first body it is (if any).  No locks, flags, atomics or fences (the reference spins on
// SUM_MASSES != 0 inside one work-group, :263-384): kernel boundaries order the levels.  Every node writes
//   msum4[n] = {sum m*x, sum m*y, sum m*z, sum m}   (the reference's massCenters_* / sumOfMasses; what parents read)
//   com[n]   = {sums / mass, mass}                   (the quotient the reference forms per visit, BarnesHutAlgorithm.cpp:351-353)
__device__ __forceinline__ void store_node(double *com, double *msum4, uint32_t node, double sx, double sy, double sz,
                                           double m) {
    double2 *s2 = reinterpret_cast<double2 *>(msum4 + 4 * (size_t) node);
    s2[0] = make_double2(sx, sy);
    s2[1] = make_double2(sz, m);
    // massless cells are skipped by the traversal; keep their record finite
    const bool has_mass = m != 0.0;
    const double cx = has_mass ? __ddiv_rn(sx, m) : 0.0, cy = has_mass ? __ddiv_rn(sy, m) : 0.0, cz = has_mass ? __ddiv_rn(sz, m) : 0.0;
    double2 *c2 = reinterpret_cast<double2 *>(com + 4 * (size_t) node);
    c2[0] = make_double2(cx, cy);
    c2[1] = make_double2(cz, m);
}

__global__ void __launch_bounds__(256)
com_leaf_kernel(uint64_t n, uint32_t *flags_in, const double *__restrict__ px,
                const double *__restrict__ py, const double *__restrict__ pz, const double *__restrict__ pm,
                const uint32_t *__restrict__ leaf_node, double *__restrict__ com, double *__restrict__ msum4) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    // every kernel that can raise an error bit of this build has run: keep the bits in the sticky word, which the next
    // build does not clear (a failed build inside a batch of steps must still be reported at the end of the batch)
    if (i == 0 && flags_in[0]) atomicOr(&flags_in[4], flags_in[0]);
    if (i >= n || (flags_in[0] & NB_FLAG_POOL)) return;
    const double m = pm[i];
    store_node(com, msum4, leaf_node[i], __dmul_rn(px[i], m), __dmul_rn(py[i], m), __dmul_rn(pz[i], m), m);
}

// One level of the bottom-up pass: internal node = sum over its children in octant order 0..7 (BarnesHutOctree.cpp:299-317).
// `tid` / `nthreads`: this thread's index in, and the size of, the set of threads that share the level.
// sum of the children of one node in octant order (octant code o = 4u + 2r + b  <->  visit rank 4u + 2b + (1-r):
// octants 0..7 are ranks 1,3,0,2,5,7,4,6); an empty octant adds 0.0 in the reference: a no-op
template <bool COHERENT>
__device__ __forceinline__ void com_sum_and_store(const uint32_t (&child)[8], uint32_t p, double *com, double *msum4) {
    const int rank_of_octant[8] = {1, 3, 0, 2, 5, 7, 4, 6};
    double sumMasses = 0, cx = 0, cy = 0, cz = 0;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        const uint32_t c = child[rank_of_octant[o]];
        if (c != NB_NONE) {
            const double2 *s2 = reinterpret_cast<const double2 *>(msum4 + 4 * (size_t) c);
            // COHERENT: the sums of the level below were written by other CTAs of the SAME launch; read them from L2
            const double2 a = COHERENT ? __ldcg(s2) : s2[0], b = COHERENT ? __ldcg(s2 + 1) : s2[1];
            cx = __dadd_rn(cx, a.x);
            cy = __dadd_rn(cy, a.y);
            cz = __dadd_rn(cz, b.x);
            sumMasses = __dadd_rn(sumMasses, b.y);
        }
    }
    store_node(com, msum4, p, cx, cy, cz, sumMasses);
}

// A thread works on TWO nodes of the level at a time: finding a node's children is a chain of dependent loads along the
// skip links (child k+1 = skip[child k], up to 8 hops), and the pass is bound by that latency (round 1, ncu: 77 cycles of
// long-scoreboard stall per issue, DRAM 41 % active); two interleaved chains per thread double the loads in flight.
template <bool COHERENT>
__device__ __forceinline__ void com_level(int depth, const uint32_t *__restrict__ level, const uint32_t *__restrict__ list,
                                          const uint2 *__restrict__ meta, double *com, double *msum4,
                                          uint32_t tid, uint32_t nthreads) {
    const uint32_t count = level[depth];
    const uint32_t *nodes = list + level[NB_LEVELS + depth];
    for (uint32_t k = tid; k < count; k += 2 * nthreads) {
        const uint32_t k2 = k + nthreads;
        const bool two = k2 < count;
        const uint32_t pa = nodes[k], pb = two ? nodes[k2] : pa;
        const uint32_t enda = meta[pa].x, endb = two ? meta[pb].x : 0u;
        // children (leaves, or depth+1 nodes finished before this level), indexed by their visit rank
        uint32_t ca[8], cb[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) ca[r] = cb[r] = NB_NONE;
        uint32_t cha = pa + 1, chb = two ? pb + 1 : 0u;
        while (cha < enda || chb < endb) {
            const bool ga = cha < enda, gb = chb < endb;
            uint2 ma = make_uint2(0u, 0u), mb = make_uint2(0u, 0u);
            if (ga) ma = meta[cha];
            if (gb) mb = meta[chb];
            if (ga) {
                const uint32_t rank = nb_meta_rank(ma.y);
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (rank == (uint32_t) r) ca[r] = cha;
                cha = ma.x > cha ? ma.x : enda;  // skip links always point forward
            }
            if (gb) {
                const uint32_t rank = nb_meta_rank(mb.y);
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (rank == (uint32_t) r) cb[r] = chb;
                chb = mb.x > chb ? mb.x : endb;
            }
        }
        com_sum_and_store<COHERENT>(ca, pa, com, msum4);
        if (two) com_sum_and_store<COHERENT>(cb, pb, com, msum4);
    }
}

// one launch per level (fallback when a cooperative launch is not possible)
__global__ void __launch_bounds__(256)
com_level_kernel(int depth, const uint32_t *__restrict__ flags_in, const uint32_t *__restrict__ level,
                 const uint32_t *__restrict__ list, const uint2 *__restrict__ meta, double *com, double *msum4) {
    // flags[2] = deepest leaf = 1 + deepest internal node: nothing to do for the levels below the tree
    if ((uint32_t) depth >= flags_in[2] || (flags_in[0] & NB_FLAG_POOL)) return;
    com_level<false>(depth, level, list, meta, com, msum4, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// All levels in ONE cooperative launch: the grid is resident as a whole and walks the levels from the deepest internal
// level up to the root with a grid-wide barrier between two levels (the reference spins on per-node flags inside one
// work-group, BarnesHutOctree.cpp:327-383).  Replaces 42 dependent launches, most of them for levels below the tree.
// Measured at N = 2^24 (gpurun_out/r2_build_ab.log): one chain per thread with 5 / 6 CTAs per SM 1.27 / 1.30 ms, with 8
// CTAs per SM (32 registers) 1.43 ms, two chains per thread 1.27 ms -- the pass does not respond to more loads in flight:
// it moves 3.2 GB of scattered 32-byte sectors (2.68 GB read: 8-byte node words and 32-byte sums of the children) at 41 %
// DRAM-active, which is what random sector traffic gets.
__global__ void __launch_bounds__(256, 4)
com_levels_kernel(const uint32_t *__restrict__ flags_in, const uint32_t *__restrict__ level, const uint32_t *__restrict__ list,
                  const uint2 *__restrict__ meta, double *com, double *msum4) {
    if (flags_in[0] & NB_FLAG_POOL) return;   // the same decision in every CTA
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    int depth = (int) flags_in[2] - 1;
    if (depth > NB_MAX_TREE_DEPTH - 1) depth = NB_MAX_TREE_DEPTH - 1;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    for (; depth >= 0; --depth) {
        com_level<true>(depth, level, list, meta, com, msum4, tid, nthreads);
        if (depth > 0) grid.sync();
    }
}

}  // namespace

// -----------------------------------------------------------------------------------------------------------------------------
int nbk_bh_reserve(nb_ctx *ctx) {
    nb_bh_state &b = ctx->bh;
    const uint64_t n = ctx->n;
    // node pool: the reference holds storage_size_param*N canonical nodes (1 + 8 per internal node); the same budget
    // in internal nodes is param*N/8 (2N for the default 16).
    uint64_t param = ctx->cfg.storage_size_param > 0 ? (uint64_t) ctx->cfg.storage_size_param : 16;
    const uint64_t cap_nodes = n + (param * n + 7) / 8 + 8;
    if (n <= b.cap_bodies && cap_nodes <= b.cap_nodes) return NB_OK;
    nbk_bh_release(ctx);
    const uint64_t nb = n + 32;
    NB_CHECK(nb_alloc(ctx, &b.key_hi, nb));
    NB_CHECK(nb_alloc(ctx, &b.key_hi_alt, nb));
    NB_CHECK(nb_alloc(ctx, &b.word_a, nb));
    NB_CHECK(nb_alloc(ctx, &b.word_b, nb));
    NB_CHECK(nb_alloc(ctx, &b.perm, nb));
    NB_CHECK(nb_alloc(ctx, &b.perm_alt, nb));
    NB_CHECK(nb_alloc(ctx, &b.delta, nb));
    NB_CHECK(nb_alloc(ctx, &b.chain_cnt, nb));
    NB_CHECK(nb_alloc(ctx, &b.chain_base, nb));
    NB_CHECK(nb_alloc(ctx, &b.leaf_node, nb));
    NB_CHECK(nb_alloc(ctx, &b.visits, nb));
    NB_CHECK(nb_alloc(ctx, &b.com, 4 * cap_nodes));
    NB_CHECK(nb_alloc(ctx, &b.msum, 4 * cap_nodes));
    NB_CHECK(nb_alloc(ctx, &b.meta, cap_nodes));
    NB_CHECK(nb_alloc(ctx, &b.level, 3 * NB_LEVELS));
    NB_CHECK(nb_alloc(ctx, &b.level_list, cap_nodes - n + 8));
    NB_CHECK(nb_alloc(ctx, &b.body_count, cap_nodes));
    const size_t scratch = nbprim::os_scratch_elems(nb) + nbprim::scan_tiles_for(nb) + 64;
    NB_CHECK(nb_alloc(ctx, &b.hist, scratch));
    NB_CHECK(nb_alloc(ctx, &b.aabb_dev, NB_ACCEPT_TABLE_OFFSET + NB_ACCEPT_DEPTHS));
    NB_CHECK(nb_alloc(ctx, &b.aabb_partial, 6 * 1024));
    NB_CHECK(nb_alloc(ctx, &b.dev_flags, 8 + 1024));   // 8 flag words + per-SM tile counters of the traversal
    NB_CHECK(nb_alloc(ctx, &b.stat_totals, 8));
    if (!b.stat_host) {
        NB_CUDA(ctx, cudaMallocHost((void **) &b.stat_host, 4 * sizeof(uint32_t)));
        NB_CUDA(ctx, cudaEventCreateWithFlags(&b.stat_event, cudaEventDisableTiming));
    }
    b.stat_pending = b.stat_known = b.long_runs = b.long_runs32 = false;   // a new problem size: nothing is known about its distribution
    NB_CUDA(ctx, cudaMemsetAsync(b.dev_flags, 0, 8 * sizeof(uint32_t), ctx->stream));
    b.cap_bodies = n;
    b.cap_nodes = cap_nodes;
    b.built = false;
    return NB_OK;
}

void nbk_bh_release(nb_ctx *ctx) {
    nb_bh_state &b = ctx->bh;
    nb_free(&b.key_hi); nb_free(&b.key_hi_alt); nb_free(&b.word_a); nb_free(&b.word_b); nb_free(&b.perm); nb_free(&b.perm_alt);
    nb_free(&b.delta); nb_free(&b.chain_cnt);
    nb_free(&b.chain_base); nb_free(&b.leaf_node); 
    nb_free(&b.visits); nb_free(&b.com); nb_free(&b.msum); nb_free(&b.meta); nb_free(&b.level); nb_free(&b.level_list);
    nb_free(&b.body_count); nb_free(&b.hist); nb_free(&b.aabb_dev);
    nb_free(&b.aabb_partial); nb_free(&b.dev_flags); nb_free(&b.stat_totals);
    if (b.stat_host) {
        cudaStreamSynchronize(ctx->stream);   // a pending read-back of the statistic targets stat_host
        cudaFreeHost(b.stat_host); b.stat_host = nullptr;
        cudaEventDestroy(b.stat_event); b.stat_event = nullptr;
    }
    b.stat_pending = b.stat_known = b.long_runs = b.long_runs32 = false;
    b.cap_bodies = b.cap_nodes = 0;
    b.built = false;
}

int nbk_bh_aabb(nb_ctx *ctx) {
    nb_bh_state &b = ctx->bh;
    const uint64_t n = ctx->n;
    int blocks = (int) ((n + 255) / 256);
    if (blocks > 1024) blocks = 1024;
    if (blocks < 1) blocks = 1;
    aabb_partial_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->x, ctx->y, ctx->z, n, b.aabb_partial, b.dev_flags);
    NB_LAUNCH_CHECK(ctx);
    aabb_final_kernel<<<1, 256, 0, ctx->stream>>>(b.aabb_partial, blocks, ctx->cfg.theta, b.aabb_dev);
    NB_LAUNCH_CHECK(ctx);
    b.accept_theta = ctx->cfg.theta;
    return NB_OK;
}

// the walk calls this: the table must belong to the theta it is about to use
int nbk_bh_accept_table(nb_ctx *ctx) {
    nb_bh_state &b = ctx->bh;
    if (memcmp(&b.accept_theta, &ctx->cfg.theta, sizeof(double)) == 0) return NB_OK;
    accept_table_kernel<<<1, NB_ACCEPT_DEPTHS, 0, ctx->stream>>>(ctx->cfg.theta, b.aabb_dev);
    NB_LAUNCH_CHECK(ctx);
    b.accept_theta = ctx->cfg.theta;
    return NB_OK;
}

// Passes of this build's sort: 8 = the full (key, slot) sort, 5 / 4 = packed words ordered on their upper 40 / 32 bits
// (see nbk_bh_build).  The statistics of the latest finished build arrive through pinned words and an event; nothing
// here waits for the device.
static int bh_choose_sort_passes(nb_ctx *ctx) {
    nb_bh_state &b = ctx->bh;
    const int forced = ctx->cfg.reserved[6];
    if (forced == 1) return 8;
    if (forced == 2) return 5;
    if (forced == 3) return 4;
    if (b.stat_pending && !ctx->capturing && cudaEventQuery(b.stat_event) == cudaSuccess) {
        b.stat_pending = false;
        b.stat_known = true;
        b.long_runs = b.stat_host[0] > NB_PACKED_RUN_LIMIT;
        b.long_runs32 = b.stat_host[1] > NB_PACKED_RUN_LIMIT;
    }
    if (!b.stat_known || b.long_runs) return 8;
    return b.long_runs32 ? 5 : 4;
}

int nbk_bh_build(nb_ctx *ctx) {
    nb_bh_state &b = ctx->bh;
    const uint64_t n = ctx->n;
    if (n == 0) return nb_fail(ctx, NB_ERR_INVALID, "nb_bh_build: no bodies");
    if (n >= (1ull << NB_DIGIT_SHIFT)) return nb_fail(ctx, NB_ERR_UNSUPPORTED, "nb_bh_build: N must be < 2^28");
    NB_CHECK(nbk_bh_reserve(ctx));
    const unsigned g256 = (unsigned) ((n + 255) / 256), g128 = (unsigned) ((n + 127) / 128);
    nb_timer_scope total(ctx, NB_T_TREE_TOTAL);
    // flags: [0] error bits, [1] internal node count, [2] max depth
    NB_CUDA(ctx, cudaMemsetAsync(b.dev_flags, 0, 4 * sizeof(uint32_t), ctx->stream));
    {
        nb_timer_scope t(ctx, NB_T_AABB);
        NB_CHECK(nbk_bh_aabb(ctx));
    }
    uint64_t *hi_sorted = nullptr;
    uint32_t *perm_sorted = nullptr;
    {
        nb_timer_scope t(ctx, NB_T_KEYS_SORT);
        keys_kernel<<<g256, 256, 0, ctx->stream>>>(ctx->x, ctx->y, ctx->z, n, b.aabb_dev, b.key_hi);
        NB_LAUNCH_CHECK(ctx);
        // Forms of the same sort (cfg.reserved[6], sort_variant: 0 = chosen per build, 1 = full, 2 = packed with 5 passes,
        // 3 = packed with 4 passes):
        //   packed: one 64-bit word {upper key bits | storage slot} per body, one-sweep passes over its upper 40 bits
        //           (13 octree levels) or 32 bits (10 levels), 16 B per body and pass; the bodies that share such a cell
        //           (a few hundred pairs / a few hundred thousand pairs at N = 2^24) are then ordered from their full
        //           keys by tie_fix_kernel;
        //   full:   (key, slot) pairs, 8 passes over all 63 key bits, 24 B per body and pass.
        // All end in the same order.  A step uses the packed form with as few passes as the previous build's run
        // statistics allow (flags[3] / flags[6]: the longest run of bodies that agree on 40 / 32 key bits, read back
        // asynchronously): a cluster far denser than the box makes those runs long, and the single thread that orders a run
        // slow.  The first build of a context, whose distribution nobody has seen yet, is always the full sort.
        const int passes = bh_choose_sort_passes(ctx);
        const bool packed = passes < 8;
        b.sort_passes = passes;
        if (packed) {
            int idx_bits = 1;
            while ((1ull << idx_bits) < n) ++idx_bits;
            uint64_t *ws = nullptr;
            NB_CHECK(nbprim::onesweep_sort_packed(ctx, b.key_hi, b.word_a, b.word_b, n, idx_bits, passes, b.hist, &ws));
            unpack_kernel<<<g256, 256, 0, ctx->stream>>>(ws, b.key_hi, n, idx_bits, b.perm, b.key_hi_alt);
            NB_LAUNCH_CHECK(ctx);
            hi_sorted = b.key_hi_alt;
            perm_sorted = b.perm;
        } else {
            NB_CHECK(nbprim::onesweep_sort_pairs(ctx, b.key_hi, b.perm, b.key_hi_alt, b.perm_alt, n, 63, b.hist, &hi_sorted,
                                                 &perm_sorted, true));
        }
        tie_fix_kernel<<<g256, 256, 0, ctx->stream>>>(hi_sorted, ctx->x, ctx->y, ctx->z, b.aabb_dev, perm_sorted, n,
                                                      packed ? 63 - 8 * passes : 0, b.dev_flags);
        NB_LAUNCH_CHECK(ctx);
        // key_hi_alt / perm_alt are reused below: make the sorted data live in (key_hi, perm)
        if (hi_sorted != b.key_hi) { uint64_t *tk = b.key_hi; b.key_hi = b.key_hi_alt; b.key_hi_alt = tk; }
        if (perm_sorted != b.perm) { uint32_t *tp = b.perm; b.perm = b.perm_alt; b.perm_alt = tp; }
        // move the whole state into sorted order; the arrays swap roles with their partners.  Accelerations that belong
        // to earlier positions (the integrator has consumed them; the walk that follows overwrites them) stay behind:
        // 7 arrays instead of 10
        {
            double **cur[10] = {&ctx->m, &ctx->x, &ctx->y, &ctx->z, &ctx->vx, &ctx->vy, &ctx->vz, &ctx->ax, &ctx->ay, &ctx->az};
            const int count = ctx->a_fresh ? 10 : 7;
            if (!ctx->a_fresh) ctx->a_order_ok = false;
            reorder_args ra;
            for (int k = 0; k < 10; ++k) { ra.src[k] = *cur[k]; ra.dst[k] = ctx->alt[k]; }
            reorder_kernel<<<g256, 256, 0, ctx->stream>>>(b.perm, n, count, ra, ctx->id, ctx->id_alt, ctx->identity_order ? 1 : 0);
            NB_LAUNCH_CHECK(ctx);
            for (int k = 0; k < count; ++k) { double *t = *cur[k]; *cur[k] = ctx->alt[k]; ctx->alt[k] = t; }
            uint32_t *ti = ctx->id; ctx->id = ctx->id_alt; ctx->id_alt = ti;
            ctx->identity_order = false;
        }
        // lower key word (levels 21..41) of the sorted bodies, where needed (key_hi_alt is free after the sort)
        keys_lo_kernel<<<g256, 256, 0, ctx->stream>>>(b.key_hi, ctx->x, ctx->y, ctx->z, b.aabb_dev, n, b.key_hi_alt);
        NB_LAUNCH_CHECK(ctx);
    }
    const uint64_t *hi = b.key_hi, *lo = b.key_hi_alt;
    {
        nb_timer_scope t(ctx, NB_T_BUILD);
        NB_CUDA(ctx, cudaMemsetAsync(b.level, 0, 3 * NB_LEVELS * sizeof(uint32_t), ctx->stream));
        delta_kernel<<<g256, 256, 0, ctx->stream>>>(hi, lo, n, b.delta, b.chain_cnt, b.dev_flags, b.level);
        NB_LAUNCH_CHECK(ctx);
        uint32_t *tile_tmp = b.hist;  // scan scratch (sort is finished)
        NB_CHECK(nbprim::exclusive_scan_u32(ctx, b.chain_cnt, b.chain_base, n, tile_tmp, b.dev_flags + 1));
        if (getenv("NB_EMIT_PER_BODY"))
            emit_kernel<<<g128, 128, 0, ctx->stream>>>(hi, lo, b.delta, b.chain_base, n, b.cap_nodes, b.dev_flags, b.meta,
                                                       b.body_count, b.leaf_node);
        else
            emit_balanced_kernel<<<g128, 128, 0, ctx->stream>>>(hi, lo, b.delta, b.chain_base, n, b.cap_nodes, b.dev_flags, b.meta,
                                                   b.body_count, b.leaf_node);
        NB_LAUNCH_CHECK(ctx);
        // dense per-depth node lists for the centre-of-mass levels
        level_scan_kernel<<<1, 32, 0, ctx->stream>>>(b.level);
        NB_LAUNCH_CHECK(ctx);
        level_fill_kernel<<<g256, 256, 0, ctx->stream>>>(b.delta, b.chain_base, n, b.dev_flags, b.level, b.level_list);
        NB_LAUNCH_CHECK(ctx);
    }
    {
        nb_timer_scope t(ctx, NB_T_COM);
        com_leaf_kernel<<<g256, 256, 0, ctx->stream>>>(n, b.dev_flags, ctx->x, ctx->y, ctx->z, ctx->m, b.leaf_node, b.com, b.msum);
        NB_LAUNCH_CHECK(ctx);
        const unsigned level_grid = (unsigned) std::min<uint64_t>(g256, (uint64_t) ctx->sm_count * 32);
        // cfg.reserved[7] (com_variant): 0 = all levels in one cooperative launch, 1 = one launch per level (A/B, and the
        // fallback when the device or the driver refuses the cooperative launch)
        bool done = false;
        if (ctx->cfg.reserved[7] != 1 && ctx->coop_launch) {
            const void *kern = (const void *) com_levels_kernel;
            if (b.com_ctas_per_sm == 0) {
                int per_sm = 0;
                NB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
                b.com_ctas_per_sm = per_sm < 1 ? 1 : per_sm;
            }
            // no more CTAs than there are internal nodes to share (< n): small systems get a cheap barrier
            const unsigned cgrid = (unsigned) std::min<uint64_t>(g256, (uint64_t) b.com_ctas_per_sm * ctx->sm_count);
            const uint32_t *a_flags = b.dev_flags, *a_level = b.level, *a_list = b.level_list;
            const uint2 *a_meta = b.meta;
            double *a_com = b.com, *a_msum = b.msum;
            void *args[] = {&a_flags, &a_level, &a_list, &a_meta, &a_com, &a_msum};
            const cudaError_t e = cudaLaunchCooperativeKernel(kern, dim3(cgrid), dim3(256), args, 0, ctx->stream);
            if (e == cudaSuccess) {
                ctx->launches++;
                done = true;
            } else {
                (void) cudaGetLastError();   // fall back to the per-level launches below
                ctx->coop_launch = false;
            }
        }
        if (!done) {
            for (int depth = NB_MAX_TREE_DEPTH - 1; depth >= 0; --depth) {
                com_level_kernel<<<level_grid, 256, 0, ctx->stream>>>(depth, b.dev_flags, b.level, b.level_list, b.meta, b.com,
                                                                      b.msum);
                NB_LAUNCH_CHECK(ctx);
            }
        }
    }
    if (!ctx->capturing && !b.stat_pending) {   // longest undecided run of this build -> the next builds' choice of sort
        NB_CUDA(ctx, cudaMemcpyAsync(b.stat_host, b.dev_flags + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        NB_CUDA(ctx, cudaMemcpyAsync(b.stat_host + 1, b.dev_flags + 6, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        NB_CUDA(ctx, cudaEventRecord(b.stat_event, ctx->stream));
        b.stat_pending = true;
    }
    b.built = true;
    return NB_OK;
}

int nbk_unpermute(nb_ctx *ctx, int count, const double *const *src, double *const *dst) {
    reorder_args ra = {};
    for (int k = 0; k < count; ++k) { ra.src[k] = src[k]; ra.dst[k] = dst[k]; }
    unpermute_kernel<<<(unsigned) ((ctx->n + 255) / 256), 256, 0, ctx->stream>>>(ctx->n, ctx->id, count, ra);
    NB_LAUNCH_CHECK(ctx);
    return NB_OK;
}

int nbk_permute_in(nb_ctx *ctx, int count, const double *const *src, double *const *dst) {
    reorder_args ra = {};
    for (int k = 0; k < count; ++k) { ra.src[k] = src[k]; ra.dst[k] = dst[k]; }
    permute_in_kernel<<<(unsigned) ((ctx->n + 255) / 256), 256, 0, ctx->stream>>>(ctx->n, ctx->id, count, ra);
    NB_LAUNCH_CHECK(ctx);
    return NB_OK;
}
