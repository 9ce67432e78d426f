// The reference's acceptance test as one fp64 comparison per tree depth.
//
// BarnesHutAlgorithm.cpp:355-359 accepts a cell of edge e_d = edge_0 * 2^-d for a body at squared distance u when
//     f(u) = RN(e_d * RN(1 / RN(sqrt(u)))) < theta.
// Correctly rounded sqrt, reciprocal and multiplication are monotone, so f is non-increasing in u, and the set of
// accepted distances of a depth is exactly { u > T_d } for ONE double T_d: the largest u with f(u) >= theta.  T_d is
// found by bisection over the bit patterns of the non-negative doubles (they order like the values) with the very
// operations of the reference's expression, so "u > T_d" decides every case -- including the last-ulp ones -- as the
// reference does, without a square root or a division in the walk.  theta <= 0 or NaN: nothing is ever accepted
// (T_d = +inf); theta = +inf: everything but u = 0 is (T_d = 0).
// The table (one entry per depth, NB_ACCEPT_DEPTHS of them) lives behind the AABB scalars: aabb_dev[8 + d].
#pragma once
#include <stdint.h>

#define NB_ACCEPT_DEPTHS 64
#define NB_ACCEPT_TABLE_OFFSET 8   /* doubles: aabb_dev = {min xyz, max xyz, edge, unused, T_0 .. T_63} */

__device__ __forceinline__ double nb_scale_pow2(double v, uint32_t depth) {  // v * 2^-depth, exact (exponent arithmetic)
    return __hiloint2double(__double2hiint(v) - (int) (depth << 20), __double2loint(v));
}

// the reference's expression for the squared distance with bit pattern ub
__device__ __forceinline__ bool nb_accepts(unsigned long long ub, double edge_d, double theta) {
    const double rs = __ddiv_rn(1.0, __dsqrt_rn(__longlong_as_double((long long) ub)));
    return __dmul_rn(edge_d, rs) < theta;
}

__device__ inline double nb_accept_threshold(double edge0, double theta, uint32_t depth) {
    const double e = nb_scale_pow2(edge0, depth);
    // invariant: lo is not accepted (u = 0 never is: f = inf or NaN), hi is accepted or the virtual pattern behind +inf
    unsigned long long lo = 0ull, hi = 0x7ff0000000000001ull;
    // the real-valued threshold (e / theta)^2 is within a few ulp of T_d: try a narrow bracket around it first
    const double g = (e / theta) * (e / theta);
    if (g > 1e-290 && g < 1e290) {
        const unsigned long long gb = (unsigned long long) __double_as_longlong(g);
        if (!nb_accepts(gb - 8192ull, e, theta) && nb_accepts(gb + 8192ull, e, theta)) { lo = gb - 8192ull; hi = gb + 8192ull; }
    }
    while (hi - lo > 1ull) {
        const unsigned long long mid = lo + ((hi - lo) >> 1);
        if (nb_accepts(mid, e, theta)) hi = mid; else lo = mid;
    }
    return __longlong_as_double((long long) lo);
}
