// Kinetic / potential energy of the system.
//
// Replaces nBodyAlgorithm::computeEnergy (reference nBodyAlgorithm.cpp:11-86):
//   E_kin[j] = 0.5 * m_j * |v_j|^2
//   E_pot[j] = sum_{i<j} G * m_i * m_j / sqrt(|x_j - x_i|^2)        (no softening, i < j only)
//   E_kin = sum_j E_kin[j];  E_pot = -sum_j E_pot[j];  E_tot = E_kin + E_pot;  virial = 2 E_kin / |E_pot|.
// The reference runs one work-item per j with a j-long serial inner loop accumulating into global memory and
// sums the N partials serially on the host.  Here: a shared-memory tiled triangular pair kernel (targets j in
// registers, 2 per thread; sources i streamed through smem tiles; only the diagonal tiles evaluate the i<j mask;
// heavy CTAs are scheduled first), then a deterministic two-level device reduction.  fp64 throughout;
// 1/sqrt is MUFU.RSQ64H + one cubic refinement (<= ~1 ulp).  FP64-pipe bound: 12 DP instructions per pair.
#include "common.cuh"

#define NB_EN_THREADS 128
#define NB_EN_IPT 2
#define NB_EN_TILE 128

namespace {

__device__ __forceinline__ double inv_sqrt_refined(double r2) {
    const double y0 = nb_rsqrt_seed(r2);
    const double y2 = y0 * y0;
    const double e = fma(-r2, y2, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double t = y0 * e;
    return fma(p, t, y0);
}

__global__ void __launch_bounds__(NB_EN_THREADS)
energy_pair_kernel(uint64_t n, uint64_t j_begin, uint64_t j_end, double G, const double *__restrict__ m,
                   const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
                   const double *__restrict__ vx, const double *__restrict__ vy, const double *__restrict__ vz,
                   double *__restrict__ e_kin, double *__restrict__ e_pot) {
    __shared__ double4 tile[2][NB_EN_TILE];
    constexpr uint64_t PER_CTA = (uint64_t) NB_EN_THREADS * NB_EN_IPT;
    // reverse order: the CTA with the largest j (longest inner loop) is scheduled first
    const uint64_t n_cta = gridDim.x;
    const uint64_t cta = n_cta - 1 - blockIdx.x;
    const uint64_t jb = j_begin + cta * PER_CTA;
    uint64_t je = jb + PER_CTA;
    if (je > j_end) je = j_end;

    double px[NB_EN_IPT], py[NB_EN_IPT], pz[NB_EN_IPT], pot[NB_EN_IPT];
    uint64_t jj[NB_EN_IPT];
#pragma unroll
    for (int k = 0; k < NB_EN_IPT; ++k) {
        jj[k] = jb + threadIdx.x + (uint64_t) k * NB_EN_THREADS;
        const uint64_t j = jj[k] < n ? jj[k] : n - 1;
        px[k] = x[j]; py[k] = y[j]; pz[k] = z[j];
        pot[k] = 0.0;
    }
    const uint64_t n_tiles = (je + NB_EN_TILE - 1) / NB_EN_TILE;  // sources 0 .. je-1
    // prefetch tile 0
    {
        const uint64_t i = threadIdx.x;
        double4 r = make_double4(0, 0, 0, 0);
        if (i < n) r = make_double4(x[i], y[i], z[i], m[i]);
        tile[0][threadIdx.x] = r;
    }
    __syncthreads();
    for (uint64_t t = 0; t < n_tiles; ++t) {
        const int cur = (int) (t & 1);
        double4 nxt = make_double4(0, 0, 0, 0);
        if (t + 1 < n_tiles) {
            const uint64_t i = (t + 1) * NB_EN_TILE + threadIdx.x;
            if (i < n) nxt = make_double4(x[i], y[i], z[i], m[i]);
        }
        const uint64_t i0 = t * NB_EN_TILE;
        if (i0 + NB_EN_TILE <= jb) {
            // all sources of the tile are below every target of this CTA: no mask
#pragma unroll 4
            for (int s = 0; s < NB_EN_TILE; ++s) {
                const double4 sr = tile[cur][s];
#pragma unroll
                for (int k = 0; k < NB_EN_IPT; ++k) {
                    const double rx = px[k] - sr.x, ry = py[k] - sr.y, rz = pz[k] - sr.z;
                    const double r2 = fma(rz, rz, fma(ry, ry, rx * rx));
                    pot[k] = fma(sr.w, inv_sqrt_refined(r2), pot[k]);
                }
            }
        } else {
#pragma unroll 2
            for (int s = 0; s < NB_EN_TILE; ++s) {
                const double4 sr = tile[cur][s];
                const uint64_t i = i0 + s;
#pragma unroll
                for (int k = 0; k < NB_EN_IPT; ++k) {
                    const double rx = px[k] - sr.x, ry = py[k] - sr.y, rz = pz[k] - sr.z;
                    const double r2 = fma(rz, rz, fma(ry, ry, rx * rx));
                    const double c = sr.w * inv_sqrt_refined(r2);
                    if (i < jj[k]) pot[k] += c;  // i < j only (nBodyAlgorithm.cpp:55)
                }
            }
        }
        tile[cur ^ 1][threadIdx.x] = nxt;
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < NB_EN_IPT; ++k) {
        const uint64_t j = jj[k];
        if (j < je) {
            const double mj = m[j];
            const double v2 = vx[j] * vx[j] + vy[j] * vy[j] + vz[j] * vz[j];
            e_kin[j] = 0.5 * mj * v2;
            e_pot[j] = G * mj * pot[k];
        }
    }
}

// deterministic reduction: fixed chunking, fixed tree inside the block
__global__ void __launch_bounds__(256)
reduce2_kernel(const double *__restrict__ a, const double *__restrict__ b, uint64_t begin, uint64_t end,
               double *__restrict__ out_a, double *__restrict__ out_b) {
    __shared__ double sa[256], sb[256];
    const uint64_t len = end - begin;
    const uint64_t chunk = (len + gridDim.x - 1) / gridDim.x;
    uint64_t lo = begin + (uint64_t) blockIdx.x * chunk;
    uint64_t hi = lo + chunk < end ? lo + chunk : end;
    double va = 0, vb = 0;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += 256) { va += a[i]; vb += b[i]; }
    sa[threadIdx.x] = va; sb[threadIdx.x] = vb;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int) threadIdx.x < s) { sa[threadIdx.x] += sa[threadIdx.x + s]; sb[threadIdx.x] += sb[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out_a[blockIdx.x] = sa[0]; out_b[blockIdx.x] = sb[0]; }
}

}  // namespace

// Computes per-body partials for targets [j_begin, j_end) and leaves {sum E_kin, sum E_pot(positive)} of that
// slice in e_partial[2n .. 2n+1].
int nbk_energy(nb_ctx *ctx, uint64_t j_begin, uint64_t j_end) {
    const uint64_t n = ctx->n;
    if (!n) return NB_OK;
    double *e_kin = ctx->e_partial, *e_pot = ctx->e_partial + n;
    double *blk = ctx->e_partial + 2 * n + 8;  // 2 x 1024 block partials
    double *fin = ctx->e_partial + 2 * n;
    if (j_end > j_begin) {
        constexpr uint64_t PER_CTA = (uint64_t) NB_EN_THREADS * NB_EN_IPT;
        const unsigned grid = (unsigned) ((j_end - j_begin + PER_CTA - 1) / PER_CTA);
        energy_pair_kernel<<<grid, NB_EN_THREADS, 0, ctx->stream>>>(n, j_begin, j_end, ctx->cfg.G, ctx->m, ctx->x, ctx->y,
                                                                    ctx->z, ctx->vx, ctx->vy, ctx->vz, e_kin, e_pot);
        NB_LAUNCH_CHECK(ctx);
        reduce2_kernel<<<1024, 256, 0, ctx->stream>>>(e_kin, e_pot, j_begin, j_end, blk, blk + 1024);
        NB_LAUNCH_CHECK(ctx);
        reduce2_kernel<<<1, 256, 0, ctx->stream>>>(blk, blk + 1024, 0, 1024, fin, fin + 1);
        NB_LAUNCH_CHECK(ctx);
    } else {
        NB_CUDA(ctx, cudaMemsetAsync(fin, 0, 2 * sizeof(double), ctx->stream));
    }
    return NB_OK;
}
