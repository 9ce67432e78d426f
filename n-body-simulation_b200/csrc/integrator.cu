// Leapfrog (kick-drift-kick) integrator and |a| capture.
//
// Replaces the two parallel_for kernels of the time loop (reference NaiveAlgorithm.cpp:140-164,206-222 =
// BarnesHutAlgorithm.cpp:157-182,223-239) and nBodyAlgorithm::storeAccelerations (nBodyAlgorithm.cpp:88-102).
// The reference keeps a separate half-step velocity array v_k1_2; here the half-step velocity lives in place in v
// (same values: v_half = v + a*(dt/2), later v = v_half + a_new*(dt/2)).  Arithmetic uses explicit
// round-to-nearest multiplies and adds (no FMA contraction) so a step is bit-identical to the oracle given
// bit-identical accelerations.  Pure streaming kernels: HBM-bound, 9 reads + 6 writes (part 1), 6 + 3 (part 2),
// 9 + 6 for the fused part2+part1 (saves one pass over v and a per non-visualised step).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
leapfrog1_kernel(uint64_t n, double dt, double *__restrict__ x, double *__restrict__ y, double *__restrict__ z,
                 double *__restrict__ vx, double *__restrict__ vy, double *__restrict__ vz,
                 const double *__restrict__ ax, const double *__restrict__ ay, const double *__restrict__ az) {
    const double h = dt / 2.0;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const double wx = __dadd_rn(vx[i], __dmul_rn(ax[i], h));
        const double wy = __dadd_rn(vy[i], __dmul_rn(ay[i], h));
        const double wz = __dadd_rn(vz[i], __dmul_rn(az[i], h));
        vx[i] = wx; vy[i] = wy; vz[i] = wz;
        x[i] = __dadd_rn(x[i], __dmul_rn(wx, dt));
        y[i] = __dadd_rn(y[i], __dmul_rn(wy, dt));
        z[i] = __dadd_rn(z[i], __dmul_rn(wz, dt));
    }
}

__global__ void __launch_bounds__(256)
leapfrog2_kernel(uint64_t n, double dt, double *__restrict__ vx, double *__restrict__ vy, double *__restrict__ vz,
                 const double *__restrict__ ax, const double *__restrict__ ay, const double *__restrict__ az) {
    const double h = dt / 2.0;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        vx[i] = __dadd_rn(vx[i], __dmul_rn(ax[i], h));
        vy[i] = __dadd_rn(vy[i], __dmul_rn(ay[i], h));
        vz[i] = __dadd_rn(vz[i], __dmul_rn(az[i], h));
    }
}

__global__ void __launch_bounds__(256)
leapfrog21_kernel(uint64_t n, double dt, double *__restrict__ x, double *__restrict__ y, double *__restrict__ z,
                  double *__restrict__ vx, double *__restrict__ vy, double *__restrict__ vz,
                  const double *__restrict__ ax, const double *__restrict__ ay, const double *__restrict__ az) {
    const double h = dt / 2.0;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const double kx = __dmul_rn(ax[i], h), ky = __dmul_rn(ay[i], h), kz = __dmul_rn(az[i], h);
        const double wx = __dadd_rn(__dadd_rn(vx[i], kx), kx);  // part 2 of step k, then part 1 of step k+1
        const double wy = __dadd_rn(__dadd_rn(vy[i], ky), ky);
        const double wz = __dadd_rn(__dadd_rn(vz[i], kz), kz);
        vx[i] = wx; vy[i] = wy; vz[i] = wz;
        x[i] = __dadd_rn(x[i], __dmul_rn(wx, dt));
        y[i] = __dadd_rn(y[i], __dmul_rn(wy, dt));
        z[i] = __dadd_rn(z[i], __dmul_rn(wz, dt));
    }
}

__global__ void __launch_bounds__(256)
accel_norm_kernel(uint64_t n, const double *__restrict__ ax, const double *__restrict__ ay,
                  const double *__restrict__ az, double *__restrict__ out) {
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const double s = __dadd_rn(__dadd_rn(__dmul_rn(ax[i], ax[i]), __dmul_rn(ay[i], ay[i])), __dmul_rn(az[i], az[i]));
        out[i] = __dsqrt_rn(s);
    }
}

inline unsigned stream_grid(const nb_ctx *ctx, uint64_t n) {
    uint64_t blocks = (n + 255) / 256;
    uint64_t cap = (uint64_t) ctx->sm_count * 16;
    return (unsigned) (blocks < cap ? (blocks ? blocks : 1) : cap);
}

}  // namespace

int nbk_leapfrog_part1(nb_ctx *ctx, double dt) {
    if (!ctx->n) return NB_OK;
    leapfrog1_kernel<<<stream_grid(ctx, ctx->n), 256, 0, ctx->stream>>>(ctx->n, dt, ctx->x, ctx->y, ctx->z, ctx->vx,
                                                                        ctx->vy, ctx->vz, ctx->ax, ctx->ay, ctx->az);
    NB_LAUNCH_CHECK(ctx);
    return NB_OK;
}

int nbk_leapfrog_part2(nb_ctx *ctx, double dt) {
    if (!ctx->n) return NB_OK;
    leapfrog2_kernel<<<stream_grid(ctx, ctx->n), 256, 0, ctx->stream>>>(ctx->n, dt, ctx->vx, ctx->vy, ctx->vz, ctx->ax,
                                                                        ctx->ay, ctx->az);
    NB_LAUNCH_CHECK(ctx);
    return NB_OK;
}

int nbk_leapfrog_part2_part1(nb_ctx *ctx, double dt) {
    if (!ctx->n) return NB_OK;
    leapfrog21_kernel<<<stream_grid(ctx, ctx->n), 256, 0, ctx->stream>>>(ctx->n, dt, ctx->x, ctx->y, ctx->z, ctx->vx,
                                                                         ctx->vy, ctx->vz, ctx->ax, ctx->ay, ctx->az);
    NB_LAUNCH_CHECK(ctx);
    return NB_OK;
}

int nbk_accel_norm(nb_ctx *ctx) {
    if (!ctx->n) return NB_OK;
    accel_norm_kernel<<<stream_grid(ctx, ctx->n), 256, 0, ctx->stream>>>(ctx->n, ctx->ax, ctx->ay, ctx->az, ctx->anorm);
    NB_LAUNCH_CHECK(ctx);
    return NB_OK;
}
