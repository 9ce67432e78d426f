"""Pins the CPU oracle to the reference's own golden vectors (reference tests/BarnesHutTest.cpp:11-220) and to
analytic known answers.  CPU only."""
import numpy as np
import pytest

X3 = np.array([0.0, 0.0, 2.0])
Y3 = np.array([1.0, 0.0, 0.0])
Z3 = np.array([0.0, 2.0, 0.0])
M3 = np.array([10.0, 10.0, 10.0])


def test_constants(oracle):
    # nBodyAlgorithm.hpp:55-61 and Configuration.cpp:6 (values recorded in SURVEY.md fact 3)
    assert oracle.gravitational_constant().hex() == "0x1.8ba054b950b18p-113"
    assert oracle.epsilon2() == 10.0 ** -22


def test_init_config_values(oracle):
    # Configuration.cpp:24-33: storage = param*N; stack = param*ceil(log2 N) (+500 below 15000 bodies)
    assert oracle.init_config(1 << 20, 16, 16) == (16 << 20, 320)
    assert oracle.init_config(1 << 24, 16, 16) == (16 << 24, 384)
    assert oracle.init_config(3, 3, 3) == (9, 3 * 2 + 500)
    assert oracle.init_config(14999, 16, 16)[1] == 16 * 14 + 500


def test_aabb_creation(oracle):
    # TEST(TestTreeCreation, AABB_creation), BarnesHutTest.cpp:11-33
    a = oracle.aabb(X3, Y3, Z3)
    assert a[6] == 2
    assert tuple(a[:3]) == (0, -0.5, 0)
    assert tuple(a[3:6]) == (2, 1.5, 2)


def test_aabb_contains_origin(oracle):
    # scratch arrays start at 0.0 (BarnesHutOctree.cpp:58-72): a cloud far from the origin still spans it
    a = oracle.aabb([5.0, 6.0], [5.0, 7.0], [5.0, 5.5])
    assert a[0] <= 0 <= a[3] and a[1] <= 0 <= a[4] and a[2] <= 0 <= a[5]
    assert a[6] == 7.0


@pytest.mark.parametrize("storage_param", [16, 3])
def test_build_octree(oracle, storage_param):
    # buildOctreeTest (:35-73) and buildOctreeSubtreesTest (:75-127): same canonical tree from both builders
    t = oracle.Tree(M3, X3, Y3, Z3, storage_param=storage_param)
    assert t.num_nodes == 9
    assert list(t.body_of_node) == [3, 0, 3, 3, 3, 3, 2, 1, 3]
    assert t.sum_masses[0] == pytest.approx(30.0)      # :118-120
    assert list(t.sorted_bodies) == [1, 2, 0]           # :122-126


def test_prepare_subtrees(oracle):
    # prepareSubtreesTest, BarnesHutTest.cpp:129-168
    counts, subtrees, n = oracle.prepare_subtrees([1, 1, 1, 1, 0, 4, 4, 5, 5, 7], 9)
    assert n == 4
    assert list(subtrees[:4]) == [1, 4, 5, 7]
    assert list(counts) == [1, 4, 0, 0, 2, 2, 0, 1, 0]


def test_sort_bodies_for_subtrees(oracle):
    # TestSortBodiesForSubtrees, BarnesHutTest.cpp:170-220
    sob = [1, 1, 1, 1, 0, 4, 4, 5, 5, 7]
    counts, subtrees, n = oracle.prepare_subtrees(sob, 9)
    start, sorted_bodies = oracle.sort_bodies_for_subtrees(sob, counts, subtrees, n)
    assert list(start) == [0, 4, 6, 8]
    assert list(sorted_bodies[:9]) == [0, 1, 2, 3, 5, 6, 7, 8, 9]


def test_octant_code_strict_compares(oracle):
    """o = 4*(y > mid) + 2*(x > mid) + 1*(z < mid) with strict compares (BarnesHutOctree.cpp:586-591): a body exactly on
    every mid-plane of the root cube [0,2]^3 falls in octant 0 (not upper, not right, not back)."""
    x = [0.0, 2.0, 1.0]; y = [0.0, 2.0, 1.0]; z = [0.0, 2.0, 1.0]
    t = oracle.Tree([1.0, 1.0, 1.0], x, y, z)
    c = t.canonical()
    kids = [(int(k), int(b)) for d, k, b in zip(c["depth"], c["kind"], c["body"]) if d == 1]
    # octant order 0..7: body 2 (centre) shares octant 0 with nobody: body 0 at the origin has z < mid -> octant 1;
    # body 1 at (2,2,2): upper, right, not back -> octant 6
    assert kids[0] == (1, 2) and kids[1] == (1, 0) and kids[6] == (1, 1)


def test_two_body_force_known_answer(oracle):
    # a = G m / (r^2 + eps2)^(3/2) * r along the separation (NaiveAlgorithm.cpp:332-351)
    G, eps2 = oracle.gravitational_constant(), oracle.epsilon2()
    m = np.array([2.0e30, 6.0e24])
    x = np.array([0.0, 1.5]); y = np.zeros(2); z = np.zeros(2)
    ax, ay, az = oracle.naive_accel(m, x, y, z)
    expect0 = G * m[1] * 1.5 / (1.5 * 1.5 + eps2) ** 1.5
    expect1 = -G * m[0] * 1.5 / (1.5 * 1.5 + eps2) ** 1.5
    assert ax[0] == pytest.approx(expect0, rel=1e-14) and ax[1] == pytest.approx(expect1, rel=1e-14)
    assert ay[0] == 0 and az[1] == 0


def test_naive_momentum_conservation(oracle, nb):
    m, x, y, z, *_ = nb.generators.plummer(512, seed=4)
    ax, ay, az = oracle.naive_accel(m, x, y, z)
    for a in (ax, ay, az):
        assert abs((m * a).sum()) <= 1e-12 * np.abs(m * a).sum()


def test_naive_rows_match_full(oracle, nb):
    m, x, y, z, *_ = nb.generators.uniform_sphere(300, seed=2)
    full = oracle.naive_accel(m, x, y, z)
    part = oracle.naive_accel(m, x, y, z, rows=(100, 117))
    for f, p in zip(full, part):
        assert np.array_equal(f[100:117], p[100:117])


def test_bh_theta_zero_equals_naive_up_to_rounding(oracle, nb):
    # theta = 0: nothing is ever accepted except body leaves -> the all-pairs sum (without the self term, which is 0)
    m, x, y, z, *_ = nb.generators.plummer(400, seed=9)
    t = oracle.Tree(m, x, y, z)
    bh = t.accel(0.0)
    nv = oracle.naive_accel(m, x, y, z)
    num = np.sqrt(sum((a - b) ** 2 for a, b in zip(bh, nv)))
    den = np.sqrt(sum(b ** 2 for b in nv))
    assert (num / den).max() < 1e-12


def test_canonical_tree_invariants(oracle, nb):
    m, x, y, z, *_ = nb.generators.plummer(2000, seed=11)
    t = oracle.Tree(m, x, y, z)
    c = t.canonical()
    kind, count = c["kind"], c["count"]
    n_internal = int((kind == 2).sum())
    assert t.num_nodes == 1 + 8 * n_internal            # every internal node has exactly 8 children (fact 8)
    assert int((kind == 1).sum()) == 2000               # one body leaf per body
    assert np.all(count[kind == 2] >= 2) and np.all(count[kind == 1] == 1) and np.all(count[kind == 0] == 0)
    assert c["mass"][0] == pytest.approx(m.sum(), rel=1e-12)
    # the in-order sort is a permutation
    assert sorted(t.sorted_bodies.tolist()) == list(range(2000))


def test_leapfrog_energy_drift_small(oracle, nb):
    m, x, y, z, vx, vy, vz = nb.generators.solar_like(30, seed=5)
    out = oracle.simulate("naive", m, x, y, z, vx, vy, vz, dt=0.25, t_end=20.0, vs=5.0, energy=True)
    e = out["energy"][:, 2]
    assert out["n_snap"] == 5 and out["n_steps"] == 80
    assert abs(e[-1] - e[0]) < 1e-5 * abs(e[0])


def test_simulate_step0_velocity_quirk(oracle, nb):
    # step-0 OUTPUT velocities are adjusted, the integrator uses the unadjusted ones
    m, x, y, z, vx, vy, vz = nb.generators.solar_like(20, seed=6)
    vx = vx + 0.01
    out = oracle.simulate("naive", m, x, y, z, vx, vy, vz, dt=1.0, t_end=2.0, vs=1.0)
    adj = oracle.adjust_velocities(m, vx, vy, vz)
    assert np.array_equal(out["vx"][0], adj[0])
    assert abs((m * out["vx"][0]).sum()) < 1e-9 * (m * np.abs(vx)).sum()
    # drift of the barycentre shows the integrator kept the +0.01 AU/day bulk motion
    com0 = (m * out["px"][0]).sum() / m.sum()
    com2 = (m * out["px"][2]).sum() / m.sum()
    assert com2 - com0 == pytest.approx(2.0 * (m * vx).sum() / m.sum(), rel=1e-6)
