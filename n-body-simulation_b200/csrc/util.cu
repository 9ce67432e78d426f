// Device restatement of the two subtree helpers of the reference's default builder, kept because the reference's own
// tests pin them (tests/BarnesHutTest.cpp:129-220):
//   prepareSubtrees        (ParallelOctreeTopDownSubtrees.cpp:436-476): per-node body histogram + ordered compaction of
//                          the non-empty subtree roots (node 0 = "already placed", skipped);
//   sortBodiesForSubtrees  (:478-534): start offsets (exclusive prefix of the counts) + grouping of bodies by subtree.
// The reference uses a serial single_task scan, an O(S^2) prefix and a linear search per body; here the same outputs
// come from the scan / stable radix sort primitives of the build (scan_sort.cuh).  Bodies inside a subtree are emitted
// in ascending body id (the reference's order is whatever the atomics produce; its test expects ascending).
#include "scan_sort.cuh"

namespace {

__global__ void hist_kernel(const uint32_t *__restrict__ subtree_of_body, uint32_t n, uint32_t *__restrict__ counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&counts[subtree_of_body[i]], 1u);
}
__global__ void flag_kernel(const uint32_t *__restrict__ counts, uint32_t node_count, uint32_t *__restrict__ flag,
                            uint32_t *__restrict__ masked) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < node_count) {
        const bool is_root = i >= 1 && counts[i] > 0;
        flag[i] = is_root ? 1u : 0u;
        masked[i] = is_root ? counts[i] : 0u;
    }
}
__global__ void compact_kernel(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ rank,
                               const uint32_t *__restrict__ start, uint32_t node_count, uint32_t *__restrict__ subtrees,
                               uint32_t *__restrict__ start_index) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 1 && i < node_count && counts[i] > 0) {
        subtrees[rank[i]] = i;
        start_index[rank[i]] = start[i];
    }
}
__global__ void body_key_kernel(const uint32_t *__restrict__ subtree_of_body, const uint32_t *__restrict__ rank,
                                uint32_t n, uint64_t *__restrict__ keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint32_t s = subtree_of_body[i];
        keys[i] = s == 0 ? 0xffffffffull : (uint64_t) rank[s];
    }
}

}  // namespace

extern "C" int nb_util_group_by_subtree(nb_ctx *ctx, uint32_t n, const uint32_t *subtree_of_body, uint32_t node_count,
                                        uint32_t *body_count_subtree, uint32_t *subtrees, uint32_t *subtree_count,
                                        uint32_t *start_index, uint32_t *sorted_bodies) {
    if (!ctx || !n || !node_count || !subtree_of_body) return nb_fail(ctx, NB_ERR_INVALID, "nb_util_group_by_subtree: bad arguments");
    for (uint32_t i = 0; i < n; ++i)
        if (subtree_of_body[i] >= node_count) return nb_fail(ctx, NB_ERR_INVALID, "subtree id out of range");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t *d_sub = nullptr, *d_cnt = nullptr, *d_flag = nullptr, *d_masked = nullptr, *d_rank = nullptr,
             *d_start = nullptr, *d_subtrees = nullptr, *d_startidx = nullptr, *d_tmp = nullptr, *d_total = nullptr,
             *d_va = nullptr, *d_vb = nullptr, *d_scratch = nullptr;
    uint64_t *d_ka = nullptr, *d_kb = nullptr;
    int rc = NB_OK;
    auto cleanup = [&]() {
        nb_free(&d_sub); nb_free(&d_cnt); nb_free(&d_flag); nb_free(&d_masked); nb_free(&d_rank); nb_free(&d_start);
        nb_free(&d_subtrees); nb_free(&d_startidx); nb_free(&d_tmp); nb_free(&d_total); nb_free(&d_va); nb_free(&d_vb);
        nb_free(&d_scratch); nb_free(&d_ka); nb_free(&d_kb);
    };
#define NB_TRY(expr) do { rc = (expr); if (rc != NB_OK) { cleanup(); return rc; } } while (0)
#define NB_TRY_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return nb_fail(ctx, NB_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); } } while (0)
    NB_TRY(nb_alloc(ctx, &d_sub, (size_t) n));
    NB_TRY(nb_alloc(ctx, &d_cnt, (size_t) node_count));
    NB_TRY(nb_alloc(ctx, &d_flag, (size_t) node_count));
    NB_TRY(nb_alloc(ctx, &d_masked, (size_t) node_count));
    NB_TRY(nb_alloc(ctx, &d_rank, (size_t) node_count));
    NB_TRY(nb_alloc(ctx, &d_start, (size_t) node_count));
    NB_TRY(nb_alloc(ctx, &d_subtrees, (size_t) node_count));
    NB_TRY(nb_alloc(ctx, &d_startidx, (size_t) node_count));
    NB_TRY(nb_alloc(ctx, &d_tmp, (size_t) nbprim::scan_tiles_for(node_count) + 8));
    NB_TRY(nb_alloc(ctx, &d_total, (size_t) 2));
    NB_TRY(nb_alloc(ctx, &d_va, (size_t) n));
    NB_TRY(nb_alloc(ctx, &d_vb, (size_t) n));
    NB_TRY(nb_alloc(ctx, &d_ka, (size_t) n));
    NB_TRY(nb_alloc(ctx, &d_kb, (size_t) n));
    NB_TRY(nb_alloc(ctx, &d_scratch, nbprim::os_scratch_elems(n)));
    NB_TRY_CUDA(cudaMemcpyAsync(d_sub, subtree_of_body, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    NB_TRY_CUDA(cudaMemsetAsync(d_cnt, 0, node_count * sizeof(uint32_t), ctx->stream));
    const unsigned gb = (n + 255) / 256, gn = (node_count + 255) / 256;
    hist_kernel<<<gb, 256, 0, ctx->stream>>>(d_sub, n, d_cnt);
    ctx->launches++;
    flag_kernel<<<gn, 256, 0, ctx->stream>>>(d_cnt, node_count, d_flag, d_masked);
    ctx->launches++;
    NB_TRY(nbprim::exclusive_scan_u32(ctx, d_flag, d_rank, node_count, d_tmp, d_total));
    NB_TRY(nbprim::exclusive_scan_u32(ctx, d_masked, d_start, node_count, d_tmp, d_total + 1));
    compact_kernel<<<gn, 256, 0, ctx->stream>>>(d_cnt, d_rank, d_start, node_count, d_subtrees, d_startidx);
    ctx->launches++;
    body_key_kernel<<<gb, 256, 0, ctx->stream>>>(d_sub, d_rank, n, d_ka);
    ctx->launches++;
    uint64_t *ks = nullptr;
    uint32_t *vs = nullptr;
    NB_TRY(nbprim::onesweep_sort_pairs(ctx, d_ka, d_va, d_kb, d_vb, n, 32, d_scratch, &ks, &vs, true));
    uint32_t totals[2] = {0, 0};
    NB_TRY_CUDA(cudaMemcpyAsync(totals, d_total, sizeof totals, cudaMemcpyDeviceToHost, ctx->stream));
    NB_TRY_CUDA(cudaStreamSynchronize(ctx->stream));
    if (body_count_subtree) NB_TRY_CUDA(cudaMemcpy(body_count_subtree, d_cnt, node_count * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (subtrees && totals[0]) NB_TRY_CUDA(cudaMemcpy(subtrees, d_subtrees, totals[0] * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (start_index && totals[0]) NB_TRY_CUDA(cudaMemcpy(start_index, d_startidx, totals[0] * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (sorted_bodies && totals[1]) NB_TRY_CUDA(cudaMemcpy(sorted_bodies, vs, totals[1] * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (subtree_count) *subtree_count = totals[0];
#undef NB_TRY
#undef NB_TRY_CUDA
    cleanup();
    return NB_OK;
}
