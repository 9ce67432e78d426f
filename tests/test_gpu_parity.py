"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Run with `pytest -m gpu` on a B200.  Nothing here reads /root/reference.

Tolerances (BASELINE.json north_star): fp64 accelerations within 1e-10 relative of the reference's path; Barnes-Hut at
equal theta: the same octree node set with per-node mass / COM, accelerations within 1e-10 relative; bit-exact for the
integer / index work (node kinds, body assignment, counts, body order) and for the leapfrog arithmetic."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-10


def relerr(a, b):
    a = np.stack(a, 1); b = np.stack(b, 1)
    return float((np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)).max())


@pytest.fixture()
def ctx(nb):
    c = nb.Context(device=0)
    yield c
    c.close()


def load_csv(path):
    rows = [l.rstrip("\n").split(",") for l in open(path)][1:]
    cols = list(zip(*rows))
    f = lambda k: np.array(cols[k], dtype=np.float64)
    return f(3), f(4), f(5), f(6), f(7), f(8), f(9)


# ---- naive all-pairs --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,gen,seed", [(1, "plummer", 1), (2, "plummer", 1), (33, "uniform_sphere", 2),
                                        (257, "plummer", 3), (4096, "plummer", 4), (5000, "uniform_sphere", 5),
                                        (16384, "plummer", 6)])
def test_naive_matches_oracle(nb, oracle, ctx, n, gen, seed):
    m, x, y, z, vx, vy, vz = getattr(nb.generators, gen)(n, seed=seed)
    ctx.set_bodies(m, x, y, z, vx, vy, vz)
    ctx.naive_accel()
    got = ctx.accelerations()
    ref = oracle.naive_accel(m, x, y, z)
    if n == 1:
        assert all(np.array_equal(g, r) for g, r in zip(got, ref))  # only the self term: exactly 0
    else:
        assert relerr(got, ref) <= TOL


@pytest.mark.parametrize("block_size", [16, 64, 100, 128, 256, 512, 1024])
@pytest.mark.parametrize("ipt", [1, 2, 4])
def test_naive_block_size_and_register_blocking(nb, oracle, block_size, ipt):
    """--block_size (tile length) and the register blocking factor are launch geometry only."""
    m, x, y, z, vx, vy, vz = nb.generators.plummer(3001, seed=7)
    ref = oracle.naive_accel(m, x, y, z)
    c = nb.Context(device=0, block_size=block_size, ipt=ipt)
    c.set_bodies(m, x, y, z, vx, vy, vz)
    c.naive_accel()
    assert relerr(c.accelerations(), ref) <= TOL
    c.close()


@pytest.mark.parametrize("segments", [1, 2, 5, 32])
def test_naive_source_segments(nb, oracle, segments):
    """Splitting the source range across CTAs (work quanta for small target counts) only reorders the partial sums."""
    m, x, y, z, *_ = nb.generators.plummer(7000, seed=17)
    c = nb.Context(device=0, naive_segments=segments, block_size=64)
    c.set_bodies(m, x, y, z)
    c.naive_accel()
    assert relerr(c.accelerations(), oracle.naive_accel(m, x, y, z)) <= TOL
    c.close()


@pytest.mark.parametrize("opt_stage", [0, 1, 2])
def test_naive_opt_stages_agree(nb, oracle, opt_stage):
    m, x, y, z, *_ = nb.generators.uniform_sphere(777, seed=8)
    c = nb.Context(device=0, opt_stage=opt_stage)
    c.set_bodies(m, x, y, z)
    c.naive_accel()
    assert relerr(c.accelerations(), oracle.naive_accel(m, x, y, z)) <= TOL
    c.close()


def test_naive_fast_rsqrt_within_tolerance(nb, oracle):
    m, x, y, z, *_ = nb.generators.plummer(4096, seed=9)
    c = nb.Context(device=0, precise_rsqrt=0)
    c.set_bodies(m, x, y, z)
    c.naive_accel()
    assert relerr(c.accelerations(), oracle.naive_accel(m, x, y, z)) <= TOL
    c.close()


def test_naive_solar_system_fixture(nb, oracle, ctx, golden_dir):
    """BASELINE config 1 input: unequal masses over 19 orders of magnitude."""
    m, x, y, z, vx, vy, vz = load_csv(os.path.join(golden_dir, "solar_178.csv"))
    ctx.set_bodies(m, x, y, z, vx, vy, vz)
    ctx.naive_accel()
    assert relerr(ctx.accelerations(), oracle.naive_accel(m, x, y, z)) <= TOL


def test_naive_operator_form_with_host_buffers(nb, oracle, ctx):
    m, x, y, z, *_ = nb.generators.plummer(2000, seed=10)
    out = ctx.op_naive_accelerations(m, x, y, z)
    assert relerr(out, oracle.naive_accel(m, x, y, z)) <= TOL


def test_naive_full_size_sample(nb, oracle, ctx):
    """N = 2^20 (BASELINE config 2): the oracle evaluates a 512-row sample (5.4e8 interactions)."""
    n = 1 << 20
    m, x, y, z, vx, vy, vz = nb.generators.plummer(n, seed=1)
    ctx.set_block_size(256)
    ctx.set_bodies(m, x, y, z, vx, vy, vz)
    ctx.naive_accel()
    got = ctx.accelerations()
    lo, hi = 700000, 700512
    ref = oracle.naive_accel(m, x, y, z, rows=(lo, hi))
    assert relerr([g[lo:hi] for g in got], [r[lo:hi] for r in ref]) <= TOL
    # size-independent property: total momentum change is zero (Newton's third law), to rounding
    for a in got:
        assert abs((m * a).sum()) <= 1e-11 * np.abs(m * a).sum()


# ---- leapfrog / energy ----------------------------------------------------------------------------------------------------
def test_leapfrog_is_bit_exact(nb, oracle, ctx):
    m, x, y, z, vx, vy, vz = nb.generators.plummer(3000, seed=11)
    ctx.set_bodies(m, x, y, z, vx, vy, vz)
    ctx.naive_accel()
    ax, ay, az = ctx.accelerations()
    dt = 1.0 / 24
    X, Y, Z = x.copy(), y.copy(), z.copy()
    vh = oracle.leapfrog_part1(dt, X, Y, Z, vx, vy, vz, ax, ay, az)
    ctx.leapfrog_part1(dt)
    assert all(np.array_equal(a, b) for a, b in zip(ctx.positions(), (X, Y, Z)))
    assert all(np.array_equal(a, b) for a, b in zip(ctx.velocities(), vh))
    V = [vx.copy(), vy.copy(), vz.copy()]
    oracle.leapfrog_part2(dt, *V, *vh, ax, ay, az)
    ctx.leapfrog_part2(dt)
    assert all(np.array_equal(a, b) for a, b in zip(ctx.velocities(), V))
    assert np.array_equal(ctx.acceleration_norms(), oracle.accel_norm(ax, ay, az))


def test_fused_kick_matches_two_passes(nb, ctx):
    m, x, y, z, vx, vy, vz = nb.generators.uniform_sphere(1500, seed=12, velocity_scale=0.5)
    other = nb.Context(device=0)
    for c in (ctx, other):
        c.set_bodies(m, x, y, z, vx, vy, vz)
        c.naive_accel()
        c.leapfrog_part1(0.5)
    ctx.leapfrog_part2(0.5); ctx.leapfrog_part1(0.5)
    other.leapfrog_part2_part1(0.5)
    assert all(np.array_equal(a, b) for a, b in zip(ctx.positions() + ctx.velocities(),
                                                    other.positions() + other.velocities()))
    other.close()


@pytest.mark.parametrize("n", [2, 100, 2049, 6000])
def test_energy_matches_oracle(nb, oracle, ctx, n):
    m, x, y, z, vx, vy, vz = nb.generators.plummer(n, seed=13)
    ctx.set_bodies(m, x, y, z, vx, vy, vz)
    got = ctx.energy()
    ref = oracle.energy(m, x, y, z, vx, vy, vz)
    assert np.all(np.abs(got - ref) <= 1e-11 * np.abs(ref))


# ---- Barnes-Hut tree ---------------------------------------------------------------------------------------------------------
X3 = np.array([0.0, 0.0, 2.0]); Y3 = np.array([1.0, 0.0, 0.0]); Z3 = np.array([0.0, 2.0, 0.0]); M3 = np.full(3, 10.0)


def test_golden_aabb(ctx):
    # reference tests/BarnesHutTest.cpp:11-33
    ctx.set_bodies(M3, X3, Y3, Z3)
    a = ctx.bh_aabb()
    assert a[6] == 2 and tuple(a[:3]) == (0, -0.5, 0) and tuple(a[3:6]) == (2, 1.5, 2)


def test_golden_three_body_tree(ctx):
    # reference tests/BarnesHutTest.cpp:35-127: leaf assignment, root mass 30, in-order sort {1,2,0}.
    # Node IDs are an artefact of the reference's atomic allocation; the (depth, path) set is what is compared.
    ctx.set_bodies(M3, X3, Y3, Z3)
    ctx.bh_build()
    info = ctx.bh_tree_info()
    assert info.num_nodes_canonical == 9 and info.num_internal == 1
    c = ctx.bh_export_canonical()
    # children of the root in octant order 0..7: body 1 in octant 0, body 2 in octant 3, body 0 in octant 5
    assert list(c["kind"]) == [2, 1, 0, 0, 1, 0, 1, 0, 0]
    assert list(c["body"]) == [3, 1, 3, 3, 2, 3, 0, 3, 3]
    assert c["mass"][0] == 30.0
    assert list(ctx.bh_sorted_bodies()) == [1, 2, 0]


def test_golden_subtree_helpers(ctx):
    # reference tests/BarnesHutTest.cpp:129-220
    counts, subtrees, n, start, sorted_bodies = ctx.util_group_by_subtree([1, 1, 1, 1, 0, 4, 4, 5, 5, 7], 9)
    assert n == 4 and list(subtrees) == [1, 4, 5, 7]
    assert list(counts) == [1, 4, 0, 0, 2, 2, 0, 1, 0]
    assert list(start) == [0, 4, 6, 8]
    assert list(sorted_bodies[:9]) == [0, 1, 2, 3, 5, 6, 7, 8, 9]


def assert_same_tree(cg, co):
    assert len(cg["depth"]) == len(co["depth"])
    for k in ("depth", "path_hi", "path_lo", "kind", "body", "count"):
        assert np.array_equal(cg[k], co[k]), k
    for k in ("edge", "minx", "miny", "minz", "mass", "comx", "comy", "comz"):
        assert np.array_equal(cg[k], co[k]), k   # bitwise: same fp64 operations in the same order


@pytest.mark.parametrize("n,gen,seed", [(1, "plummer", 1), (2, "plummer", 2), (3, "uniform_sphere", 3),
                                        (64, "plummer", 4), (1000, "uniform_sphere", 5), (4096, "plummer", 6),
                                        (30000, "plummer", 7), (100000, "uniform_sphere", 8)])
def test_tree_node_set_and_com_match_oracle(nb, oracle, ctx, n, gen, seed):
    m, x, y, z, *_ = getattr(nb.generators, gen)(n, seed=seed)
    m = m * (1.0 + 0.5 * np.sin(np.arange(n)))  # unequal masses
    ctx.set_bodies(m, x, y, z)
    ctx.bh_build()
    t = oracle.Tree(m, x, y, z)
    info = ctx.bh_tree_info()
    assert info.num_nodes_canonical == t.num_nodes and info.max_depth == t.max_depth
    assert np.array_equal(ctx.bh_aabb(), t.aabb())
    assert_same_tree(ctx.bh_export_canonical(), t.canonical())
    assert np.array_equal(ctx.bh_sorted_bodies(), t.sorted_bodies)


def test_tree_with_bodies_on_cell_boundaries(nb, oracle, ctx):
    """Lattice positions (multiples of 2^-4 in a cube that the AABB reproduces exactly) put many bodies exactly on cell
    mid-planes: the strict compares `y > mid`, `x > mid`, `z < mid` (ParallelOctreeTopDownSubtrees.cpp:400-406) decide."""
    g = np.arange(17) / 16.0
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    x, y, z = X.ravel().copy(), Y.ravel().copy(), Z.ravel().copy()
    rng = np.random.default_rng(5)
    keep = rng.random(x.size) < 0.6
    keep[0] = keep[-1] = True                      # keep the corners so the cube is [0, 1]^3
    x, y, z = x[keep], y[keep], z[keep]
    m = 1.0e24 * (1.0 + rng.random(x.size))
    ctx.set_theta(0.5)
    ctx.set_bodies(m, x, y, z)
    ctx.bh_enable_stats(True)
    ctx.bh_build(); ctx.bh_accel()
    t = oracle.Tree(m, x, y, z)
    assert np.array_equal(ctx.bh_aabb(), t.aabb()) and t.aabb()[6] == 1.0
    assert_same_tree(ctx.bh_export_canonical(), t.canonical())
    assert np.array_equal(ctx.bh_sorted_bodies(), t.sorted_bodies)
    ax, ay, az, st = t.accel(0.5, stats=True)
    assert np.array_equal(ctx.bh_stats(per_body=True)[2], st[:, 1].astype(np.uint32))
    assert relerr(ctx.accelerations(), (ax, ay, az)) <= TOL


def test_tree_two_clusters_negative_coordinates(nb, oracle, ctx):
    """Two well separated clumps (deep, unbalanced tree; long single-child chains) in the negative octants."""
    m1, x1, y1, z1, *_ = nb.generators.plummer(3000, seed=31, a=0.01, r_max=20.0)
    m2, x2, y2, z2, *_ = nb.generators.uniform_sphere(2000, seed=32, radius=0.05)
    x = np.concatenate([x1 - 40.0, x2 - 3.0]); y = np.concatenate([y1 - 25.0, y2 - 30.0]); z = np.concatenate([z1 - 7.0, z2 - 9.0])
    m = np.concatenate([m1, m2])
    c = nb.Context(device=0, theta=0.7, storage_size_param=64)
    c.set_bodies(m, x, y, z)
    c.bh_build(); c.bh_accel()
    t = oracle.Tree(m, x, y, z, storage_param=64)
    assert c.bh_tree_info().max_depth == t.max_depth
    assert_same_tree(c.bh_export_canonical(), t.canonical())
    assert relerr(c.accelerations(), t.accel(0.7)) <= TOL
    c.close()


def test_tree_with_cloud_far_from_origin(nb, oracle, ctx):
    """The AABB always contains the origin."""
    m, x, y, z, *_ = nb.generators.uniform_sphere(500, seed=9)
    x = x + 7.0; y = y - 3.0
    ctx.set_bodies(m, x, y, z)
    ctx.bh_build()
    t = oracle.Tree(m, x, y, z)
    a = ctx.bh_aabb()
    assert np.array_equal(a, t.aabb())
    assert a[0] == 0.0 and a[6] == x.max()       # x in [0, max]: the origin is the low corner of the longest axis
    assert a[1] < y.min() and a[4] > 0.0         # y spans [min, 0], then grown symmetrically to the cube
    assert_same_tree(ctx.bh_export_canonical(), t.canonical())


def test_deep_tree_uses_second_key_word(nb, oracle):
    """Pairs closer than edge * 2^-21 need the 42-level path (key_lo).  Long single-child chains need more node
    storage than 16 N (the reference would overflow silently): --storage_size_param=64 on both sides."""
    ctx = nb.Context(device=0, storage_size_param=64)
    m, x, y, z, *_ = nb.generators.uniform_sphere(300, seed=10)
    x = np.concatenate([x, x[:40] + 3e-9]); y = np.concatenate([y, y[:40] - 2e-9]); z = np.concatenate([z, z[:40] + 1e-9])
    m = np.concatenate([m, m[:40]])
    ctx.set_bodies(m, x, y, z)
    ctx.bh_build()
    t = oracle.Tree(m, x, y, z, storage_param=64)
    assert t.max_depth > 21 and ctx.bh_tree_info().max_depth == t.max_depth
    assert_same_tree(ctx.bh_export_canonical(), t.canonical())
    ctx.set_theta(0.5)
    ctx.bh_accel()
    assert relerr(ctx.accelerations(), t.accel(0.5)) <= TOL
    ctx.close()


@pytest.mark.parametrize("sort_variant", [0, 1, 2, 3])
@pytest.mark.parametrize("n_cluster", [40, 700])
def test_dense_cluster_in_a_wide_box(nb, oracle, sort_variant, n_cluster):
    """A cluster 2^-18 of the box wide plus far outliers: every cluster body shares its first ~17 octree levels, i.e. all 40
    key bits the packed sort orders (or all 32 of its 4-pass form, sort_variant 3), so the whole cluster is one run the
    sort leaves undecided (tie_fix_kernel orders it from the full keys; beyond 64 bodies per run the per-build choice goes
    back to more passes and finally to the full sort).  Same canonical tree, same order, over several builds, whichever
    form sorts."""
    rng = np.random.default_rng(77)
    m0, x0, y0, z0, *_ = nb.generators.uniform_sphere(200, seed=13)
    xc = 0.3 + 4e-6 * rng.random(n_cluster); yc = -0.2 + 4e-6 * rng.random(n_cluster); zc = 0.1 + 4e-6 * rng.random(n_cluster)
    x = np.concatenate([x0, xc]); y = np.concatenate([y0, yc]); z = np.concatenate([z0, zc])
    m = np.concatenate([m0, np.full(n_cluster, m0[0])])
    c = nb.Context(device=0, theta=0.5, storage_size_param=64, sort_variant=sort_variant)
    c.set_bodies(m, x, y, z)
    t = oracle.Tree(m, x, y, z, storage_param=64)
    for _ in range(3):       # build 1: full sort (nothing known yet), then packed unless the run statistic forbids it
        c.bh_build()
        c.synchronize()
        assert c.bh_tree_info().max_depth == t.max_depth
        assert np.array_equal(c.bh_sorted_bodies(), t.sorted_bodies)
    assert_same_tree(c.bh_export_canonical(), t.canonical())
    c.bh_accel()
    assert relerr(c.accelerations(), t.accel(0.5)) <= TOL
    c.close()


def test_failed_build_inside_a_batch_is_reported(nb):
    """A build that fails in the middle of nb_advance (here: the node pool is too small from the start) zeroes that
    step's accelerations; later builds clear the per-build error word, so the sticky word must carry the status to the
    next synchronising call -- once."""
    m, x, y, z, vx, vy, vz = nb.generators.plummer(2000, seed=12)
    c = nb.Context(device=0, storage_size_param=1)
    c.set_bodies(m, x, y, z, vx, vy, vz)
    c.advance("BarnesHut", 1e-3, 8)
    with pytest.raises(nb.NBodyError) as e:
        c.synchronize()
    assert e.value.status == -5
    c.close()


def test_coincident_bodies_are_reported(nb, ctx):
    """Reference: unbounded splitting (UB).  Here: an explicit status."""
    m, x, y, z, *_ = nb.generators.uniform_sphere(100, seed=11)
    x[5], y[5], z[5] = x[6], y[6], z[6]
    ctx.set_bodies(m, x, y, z)
    ctx.bh_build()
    with pytest.raises(nb.NBodyError) as e:
        ctx.synchronize()
    assert e.value.status == -4


def test_node_pool_overflow_is_reported(nb):
    """Reference: silent overflow of the storage_size_param*N pool (README.md:117-118).  Here: NB_ERR_NODE_POOL."""
    m, x, y, z, *_ = nb.generators.plummer(2000, seed=12)
    c = nb.Context(device=0, storage_size_param=1)   # 1*N canonical nodes = N/8 internal nodes: too few
    c.set_bodies(m, x, y, z)
    c.bh_build()
    with pytest.raises(nb.NBodyError) as e:
        c.synchronize()
    assert e.value.status == -5
    c.close()


# ---- Barnes-Hut accelerations ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("theta", [0.0, 0.2, 0.5, 1.05])
@pytest.mark.parametrize("n,gen,seed", [(2, "plummer", 1), (100, "uniform_sphere", 2), (5000, "plummer", 3),
                                        (20000, "uniform_sphere", 4)])
def test_bh_accelerations_and_visit_counts(nb, oracle, n, gen, seed, theta):
    m, x, y, z, *_ = getattr(nb.generators, gen)(n, seed=seed)
    c = nb.Context(device=0, theta=theta)
    c.set_bodies(m, x, y, z)
    c.bh_enable_stats(True)
    c.bh_build()
    c.bh_accel()
    got = c.accelerations()
    tv, ta, per_body = c.bh_stats(per_body=True)
    ax, ay, az, st = oracle.Tree(m, x, y, z).accel(theta, stats=True)
    # identical interaction sets: the per-body count of non-empty visits (BarnesHutAlgorithm.cpp:349) and accepts
    assert np.array_equal(per_body, st[:, 1].astype(np.uint32))
    assert (tv, ta) == (int(st[:, 1].sum()), int(st[:, 2].sum()))
    assert relerr(got, (ax, ay, az)) <= TOL
    c.close()


@pytest.mark.parametrize("n,gen", [(3000, "plummer"), (700000, "uniform_sphere")])
def test_bh_walk_forms_agree(nb, oracle, n, gen):
    """The production walk in its grid-mapped (20) and SM-queue (50) forms is the same arithmetic per body: bitwise equal
    accelerations.  walk_variant 0 picks the form by size (SM queues from 2^19 bodies).  All forms must match the oracle."""
    m, x, y, z, *_ = getattr(nb.generators, gen)(n, seed=21)
    got = {}
    for wv in (0, 20, 50):
        c = nb.Context(device=0, theta=0.5, walk_variant=wv)
        c.set_bodies(m, x, y, z)
        c.bh_build()
        c.bh_accel()
        got[wv] = np.stack(c.accelerations())
        c.close()
    assert np.array_equal(got[20], got[50]) and np.array_equal(got[0], got[20])
    want = oracle.Tree(m, x, y, z).accel(0.5)
    assert relerr(tuple(got[0]), want) <= TOL


@pytest.mark.parametrize("n,gen", [(1, "plummer"), (2047, "plummer"), (2049, "uniform_sphere"), (300001, "plummer"),
                                   (1 << 21, "uniform_sphere")])
def test_sort_forms_give_the_same_order(nb, oracle, n, gen):
    """The forms of the build's sort -- packed {upper key bits | slot} words, 5 or 4 one-sweep passes over 40 or 32 key
    bits, ties ordered from the full keys (sort_variant 2, 3) and (key, slot) pairs, 8 passes over all 63 key bits (1) --
    and the per-build choice between them (0: full for the first build, then as few passes as the run statistics of the
    previous build allow) end in the same order: identical in-order permutation
    (BarnesHutOctree.cpp:550-613), identical storage order, identical accelerations -- also over repeated builds of
    moving bodies (the look-back status table is reused) and for tile counts of 1, 2 and many.  The permutation must be
    the oracle's."""
    m, x, y, z, vx, vy, vz = getattr(nb.generators, gen)(n, seed=33)
    got = {}
    for sv in (0, 1, 2, 3):
        c = nb.Context(device=0, theta=0.5, sort_variant=sv)
        c.set_bodies(m, x, y, z, vx, vy, vz)
        for _ in range(3):
            c.leapfrog_part1(0.01); c.bh_build(); c.bh_accel(); c.leapfrog_part2(0.01)
        got[sv] = (np.asarray(c.bh_sorted_bodies()), np.stack(c.accelerations()), np.stack(c.positions()))
        c.close()
    for sv in (1, 2, 3):
        for a, b in zip(got[0], got[sv]):
            assert np.array_equal(a, b)
    if n <= 300001:
        px, py, pz = got[0][2]
        assert np.array_equal(got[0][0], oracle.Tree(m, px, py, pz).sorted_bodies)


@pytest.mark.parametrize("n,world", [(5000, 3), (600001, 8), (1 << 21, 2)])
def test_bh_slices_assemble_to_the_full_traversal(nb, n, world):
    """nb_bh_accel_range evaluates the storage slots one rank of a world_size-P run owns (nb_slice_bounds); the slices of
    all ranks, evaluated one after the other on one GPU, give exactly the accelerations of the full traversal."""
    m, x, y, z, *_ = nb.generators.uniform_sphere(n, seed=29)
    c = nb.Context(device=0, theta=0.5)
    c.set_bodies(m, x, y, z)
    c.bh_build(); c.bh_accel()
    want = np.stack(c.accelerations())
    c.bh_build()
    covered = 0
    for r in reversed(range(world)):
        b, e = nb.slice_bounds(n, world, r)
        c.bh_accel_range(b, e)
        covered += e - b
    assert covered == n
    assert np.array_equal(np.stack(c.accelerations()), want)
    with pytest.raises(nb.NBodyError):
        c.bh_accel_range(5, n + 1)
    c.close()


def test_bh_massless_bodies_are_invisible(nb, oracle):
    """The reference skips nodes with SUM_MASSES == 0 (BarnesHutAlgorithm.cpp:349): massless bodies exert no force and
    are not counted as visits, but they are still accelerated (tracer particles)."""
    m, x, y, z, *_ = nb.generators.plummer(3000, seed=41)
    m[::7] = 0.0
    m[100:140] = 0.0   # a few cells that hold only massless bodies
    c = nb.Context(device=0, theta=0.5)
    c.set_bodies(m, x, y, z)
    c.bh_enable_stats(True)
    c.bh_build(); c.bh_accel()
    got = c.accelerations()
    ax, ay, az, st = oracle.Tree(m, x, y, z).accel(0.5, stats=True)
    assert np.array_equal(c.bh_stats(per_body=True)[2], st[:, 1].astype(np.uint32))
    assert all(np.isfinite(g).all() for g in got)
    assert relerr(got, (ax, ay, az)) <= TOL
    c.bh_enable_stats(False)          # the production (uninstrumented) kernel must give the same accelerations
    c.bh_build(); c.bh_accel()
    again = c.accelerations()
    assert all(np.array_equal(u, v) for u, v in zip(got, again))
    c.naive_accel()
    assert relerr(c.accelerations(), oracle.naive_accel(m, x, y, z)) <= TOL
    c.close()


@pytest.mark.parametrize("wg", [32, 64, 128, 256])
def test_bh_work_group_size_is_geometry_only(nb, oracle, wg):
    m, x, y, z, *_ = nb.generators.plummer(3000, seed=5)
    c = nb.Context(device=0, theta=0.6, wg_size_barnes_hut=wg)
    c.set_bodies(m, x, y, z)
    c.bh_build(); c.bh_accel()
    assert relerr(c.accelerations(), oracle.Tree(m, x, y, z).accel(0.6)) <= TOL
    c.close()


@pytest.mark.parametrize("theta", [0.3, 0.7])
def test_bh_traversal_gives_identical_interaction_sets(nb, oracle, theta):
    """The warp walk must reproduce the reference's per-body node sets (visit / accept counts) and accelerations."""
    m, x, y, z, *_ = nb.generators.plummer(12345, seed=21)
    c = nb.Context(device=0, theta=theta)
    c.set_bodies(m, x, y, z)
    c.bh_enable_stats(True)
    c.bh_build(); c.bh_accel()
    tv, ta, per_body = c.bh_stats(per_body=True)
    ax, ay, az, st = oracle.Tree(m, x, y, z).accel(theta, stats=True)
    assert np.array_equal(per_body, st[:, 1].astype(np.uint32))
    assert (tv, ta) == (int(st[:, 1].sum()), int(st[:, 2].sum()))
    assert relerr(c.accelerations(), (ax, ay, az)) <= TOL
    c.close()


def test_bh_default_theta_solar_fixture(nb, oracle, ctx, golden_dir):
    m, x, y, z, vx, vy, vz = load_csv(os.path.join(golden_dir, "solar_178.csv"))
    out = ctx.op_barnes_hut_accelerations(m, x, y, z)
    assert relerr(out, oracle.Tree(m, x, y, z).accel(1.05)) <= TOL


def test_bh_full_size_properties(nb, oracle, ctx):
    """N = 2^20 Plummer, theta = 0.5 (BASELINE config 3): structural invariants + a sampled oracle comparison."""
    n = 1 << 20
    m, x, y, z, *_ = nb.generators.plummer(n, seed=1)
    ctx.set_theta(0.5)
    ctx.set_bodies(m, x, y, z)
    ctx.bh_enable_stats(True)
    ctx.bh_build(); ctx.bh_accel()
    got = ctx.accelerations()
    info = ctx.bh_tree_info()
    t = oracle.Tree(m, x, y, z)
    assert info.num_nodes_canonical == t.num_nodes and info.max_depth == t.max_depth
    ax, ay, az, st = t.accel(0.5, stats=True)
    tv, ta = ctx.bh_stats()
    assert (tv, ta) == (int(st[:, 1].sum()), int(st[:, 2].sum()))
    assert relerr(got, (ax, ay, az)) <= TOL
    c = ctx.bh_export_canonical()
    assert c["mass"][0] == t.sum_masses[0] and c["count"][0] == n
    assert sorted(np.unique(ctx.bh_sorted_bodies()).tolist()) == list(range(n))


# ---- trajectories ----------------------------------------------------------------------------------------------------------------
def gpu_simulate(nb, algorithm, m, x, y, z, vx, vy, vz, dt, t_end, vs, theta=1.05, energy=False):
    """The reference's time loop (NaiveAlgorithm.cpp:82-259) driven through the C ABI."""
    c = nb.Context(device=0, theta=theta)
    c.set_bodies(m, x, y, z, vx, vy, vz)

    def forces():
        if algorithm == "naive":
            c.naive_accel()
        else:
            c.bh_build(); c.bh_accel()

    snaps = {"px": [x.copy()], "anorm": [], "vx": [], "energy": []}
    forces()
    if energy:
        snaps["energy"].append(c.energy())
    snaps["anorm"].append(c.acceleration_norms())
    time, since, step = dt, dt, 1
    while time <= t_end + 0.000001:
        vis = abs(since - vs) < 0.000001
        c.leapfrog_part1(dt)
        if vis:
            snaps["px"].append(c.positions()[0])
        forces()
        c.leapfrog_part2(dt)
        if vis:
            snaps["anorm"].append(c.acceleration_norms())
            snaps["vx"].append(c.velocities()[0])
            if energy:
                snaps["energy"].append(c.energy())
            step += 1
            since = 0.0
        time += dt
        since += dt
    final = c.positions() + c.velocities()
    c.close()
    return snaps, final


def test_trajectory_naive_100_steps(nb, oracle):
    """K = 100 leapfrog steps at N = 4096: positions within 1e-9 of the system radius."""
    m, x, y, z, vx, vy, vz = nb.generators.plummer(4096, seed=14)
    dt, K = 1.0 / 24, 100
    snaps, final = gpu_simulate(nb, "naive", m, x, y, z, vx, vy, vz, dt, K * dt, 25 * dt, energy=True)
    ref = oracle.simulate("naive", m, x, y, z, vx, vy, vz, dt, K * dt, 25 * dt, energy=True)
    assert ref["n_steps"] == K and ref["n_snap"] == len(snaps["px"]) == 5
    radius = 50.0
    for s in range(5):
        assert np.abs(snaps["px"][s] - ref["px"][s]).max() <= 1e-9 * radius
    assert np.abs(np.array(snaps["energy"]) - ref["energy"]).max() <= 1e-9 * np.abs(ref["energy"]).max()
    assert np.allclose(snaps["anorm"][-1], ref["anorm"][-1], rtol=1e-8, atol=0)


@pytest.mark.parametrize("algorithm,n,unfused", [("naive", 300, 0), ("BarnesHut", 300, 0), ("BarnesHut", 20000, 0),
                                                 ("BarnesHut", 20000, 1), ("BarnesHut", 700000, 0)])
@pytest.mark.parametrize("timers", [False, True])
def test_advance_equals_single_steps(nb, algorithm, n, unfused, timers):
    """nb_advance (batch of steps, inner steps replayed from a CUDA graph of two steps) is bit-identical to issuing
    part 1 / forces / part 2 one by one, for any batch length, also across repeated batches (graph reuse) and after the
    configuration changed in between (graph re-capture).  Barnes-Hut batches run fused (the leapfrog half-steps in the
    epilogue of the walk, state updated in place); unfused_advance=1 keeps the separate integrator kernels."""
    m, x, y, z, vx, vy, vz = nb.generators.plummer(n, seed=31)
    dt = 1.0 / 24

    def forces(c):
        if algorithm == "naive":
            c.naive_accel()
        else:
            c.bh_build(); c.bh_accel()

    a = nb.Context(device=0, theta=0.6, unfused_advance=unfused)
    b = nb.Context(device=0, theta=0.6)
    for c in (a, b):
        c.set_bodies(m, x, y, z, vx, vy, vz)
        c.enable_timers(timers)
        forces(c)
    total = 0
    for k in ((1, 2, 5, 6, 13, 40, 7) if n < 100000 else (1, 3, 8)):
        if k == 13:   # a configuration change invalidates the captured graph
            a.set_theta(0.45); b.set_theta(0.45)
        ms = a.advance(algorithm, dt, k, timers=timers)
        for _ in range(k):
            b.leapfrog_part1(dt); forces(b); b.leapfrog_part2(dt)
        total += k
        for ga, gb in zip(a.positions() + a.velocities() + a.accelerations(), b.positions() + b.velocities() + b.accelerations()):
            assert np.array_equal(ga, gb), (k, total)
        if timers:
            assert ms[0] > 0 and ms[1] > 0    # acceleration, leapfrog 1 of the sampled step
            if algorithm == "naive" or unfused:
                assert ms[2] > 0              # a fused Barnes-Hut batch has no separate leapfrog part 2
    # the graph stands for real launches: same count as the eager sequence, within the fused half-kicks
    assert abs(a.launch_count() - b.launch_count()) <= 2 * total
    a.close(); b.close()


def test_trajectory_barnes_hut(nb, oracle):
    m, x, y, z, vx, vy, vz = nb.generators.plummer(2048, seed=15)
    dt, K = 1.0 / 24, 40
    snaps, final = gpu_simulate(nb, "BarnesHut", m, x, y, z, vx, vy, vz, dt, K * dt, 10 * dt, theta=0.5)
    ref = oracle.simulate("BarnesHut", m, x, y, z, vx, vy, vz, dt, K * dt, 10 * dt, theta=0.5)
    assert ref["n_snap"] == len(snaps["px"])
    for s in range(ref["n_snap"]):
        assert np.abs(snaps["px"][s] - ref["px"][s]).max() <= 1e-9 * 50.0
