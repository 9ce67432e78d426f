"""The walk decides the undecided band of the acceptance test with ONE comparison per depth, d2 > T_d
(n-body-simulation_b200/csrc/bh_accept.cuh).  That is exact only if the reference's expression
(BarnesHutAlgorithm.cpp:355-359)

    f(d2) = RN(edge_d * RN(1 / RN(sqrt(d2)))) < theta

is monotone in d2 and T_d is the largest d2 that is NOT accepted.  This test restates the device's bisection in numpy
(IEEE double sqrt / divide / multiply round exactly like the __d*_rn intrinsics the kernel uses) and checks the
equivalence  f(d2) < theta  <=>  d2 > T_d  around the threshold ulp by ulp, and far from it, for many boxes, depths
and opening angles -- including theta <= 0, NaN and +inf, where nothing / everything but d2 = 0 is accepted.  The CUDA
path itself is checked against the oracle's per-body visit and accept counts in tests/test_gpu_parity.py."""
import random
import struct

import numpy as np


def _bits(x):
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def _from_bits(b):
    return struct.unpack("<d", struct.pack("<Q", b))[0]


def _accepts(ub, edge_d, theta):
    """the reference's expression for the squared distance with bit pattern ub"""
    with np.errstate(all="ignore"):
        rs = np.float64(1.0) / np.sqrt(np.float64(_from_bits(ub)))
        return bool(np.float64(edge_d) * rs < np.float64(theta))


def _threshold(edge0, theta, depth):
    """nb_accept_threshold of bh_accept.cuh, statement by statement"""
    e = edge0 * 2.0 ** -depth
    lo, hi = 0, 0x7FF0000000000001
    with np.errstate(all="ignore"):
        q = np.float64(e) / np.float64(theta)
        g = q * q
    if g > 1e-290 and g < 1e290:
        gb = _bits(g)
        if not _accepts(gb - 8192, e, theta) and _accepts(gb + 8192, e, theta):
            lo, hi = gb - 8192, gb + 8192
    steps = 0
    while hi - lo > 1:
        mid = lo + ((hi - lo) >> 1)
        if _accepts(mid, e, theta):
            hi = mid
        else:
            lo = mid
        steps += 1
    return lo, e, steps


def test_one_comparison_per_depth_decides_like_the_reference():
    rng = random.Random(20261018)
    for _ in range(120):
        edge0 = rng.choice([1.0, 3.7, 1e-3, 123456.789, rng.uniform(0.1, 100.0), 2.0 ** rng.randrange(-30, 30)])
        theta = rng.choice([0.5, 0.2, 1.05, 1e-8, 3.0, rng.uniform(0.05, 2.0)])
        depth = rng.randrange(0, 43)
        t, e, steps = _threshold(edge0, theta, depth)
        assert steps <= 15, "the narrow bracket around (edge_d / theta)^2 must hold for ordinary boxes"
        for off in range(-200, 201):          # ulp by ulp around the threshold
            ub = t + off
            if ub >= 0:
                assert _accepts(ub, e, theta) == (ub > t)
        for _ in range(40):                   # and anywhere else
            ub = rng.randrange(0, 0x7FF0000000000000)
            assert _accepts(ub, e, theta) == (ub > t)


def test_degenerate_opening_angles():
    inf_bits = 0x7FF0000000000000
    for theta in (0.0, -1.0, float("nan")):   # nothing is accepted: the threshold is +inf
        t, e, _ = _threshold(2.0, theta, 3)
        assert t == inf_bits
    t, e, _ = _threshold(2.0, float("inf"), 3)  # everything but d2 = 0 is accepted
    assert t == 0
    # absurd boxes fall back to the full bisection and still satisfy the invariant
    for edge0 in (1e-300, 1e300):
        t, e, steps = _threshold(edge0, 0.5, 0)
        assert not _accepts(t, e, 0.5)
        assert t + 1 > inf_bits or _accepts(t + 1, e, 0.5)
