#!/usr/bin/env python
"""Generates tests/golden/reference_vectors.npz from the reference ITSELF (oracle/_ref: the unmodified sources of
/root/reference compiled with g++ through oracle/sycl_shim, see oracle/Makefile).

    python tools/make_golden.py            # rewrites the fixture (needs /root/reference or a prebuilt oracle/_ref)
    python tools/make_golden.py --check    # regenerates in memory and compares with the committed file, bit for bit

Every array in the file is an output of a reference function (named in the key) on the inputs stored next to it.  The
fixture travels to the GPU box where /root/reference does not exist; tests/test_reference_oracle.py pins the CPU oracle
to it and tests/test_gpu_reference.py checks the CUDA path against it.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
OUT = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
SOLAR = os.path.join(ROOT, "tests", "golden", "solar_178.csv")

N_SMALL = 256
SEED = 21
THETAS = (0.2, 0.5, 1.05)


def generate():
    import refimpl as R
    nb = importlib.import_module("n-body-simulation_b200")
    g = {}
    g["G"] = np.array([R.gravitational_constant()])
    g["epsilon2"] = np.array([R.epsilon2()])
    g["config_2p20"] = np.array(R.configure(1 << 20, 16, 16), dtype=np.uint32)      # storage size, stack size
    g["config_3_3_3"] = np.array(R.configure(3, 3, 3), dtype=np.uint32)

    for tag, gen in (("plummer", nb.generators.plummer), ("uniform", nb.generators.uniform_sphere)):
        m, x, y, z, vx, vy, vz = gen(N_SMALL, seed=SEED)
        for k, a in zip(("m", "x", "y", "z", "vx", "vy", "vz"), (m, x, y, z, vx, vy, vz)):
            g["%s_in_%s" % (tag, k)] = a
        for stage in (0, 1, 2):
            a = R.naive_accel(m, x, y, z, opt_stage=stage, block_size=64)
            g["%s_naive_opt%d" % (tag, stage)] = np.stack(a)
        g["%s_energy" % tag] = R.energy(m, x, y, z, vx, vy, vz)
        for builder in ("subtrees", "synchronized"):
            t = R.Tree(m, x, y, z, builder=builder)
            c = t.canonical()
            for k, a in c.items():
                g["%s_tree_%s_%s" % (tag, builder, k)] = a
            g["%s_tree_%s_aabb" % (tag, builder)] = t.aabb()
            g["%s_tree_%s_sorted" % (tag, builder)] = t.sorted_bodies
        for theta in THETAS:
            ax, ay, az, nodes = R.bh_accel(m, x, y, z, theta)
            g["%s_bh_theta%g" % (tag, theta)] = np.stack([ax, ay, az])
        ax, ay, az, _ = R.bh_accel(m, x, y, z, 0.5, sort_bodies=False)
        g["%s_bh_theta0.5_unsorted" % tag] = np.stack([ax, ay, az])
        # 24 leapfrog steps of one hour, a snapshot every 6 hours, energies on
        for alg, theta in (("naive", 1.05), ("BarnesHut", 0.5)):
            s = R.simulate(alg, m, x, y, z, vx, vy, vz, dt=1.0 / 24, t_end=1.0, vs=0.25, theta=theta, energy=True)
            for k in ("px", "py", "pz", "vx", "vy", "vz", "anorm", "energy"):
                g["%s_sim_%s_%s" % (tag, alg, k)] = s[k]

    # BASELINE config 1 input through the reference's own InputParser: 30 days, dt = 1h, a snapshot every 10 days
    for alg, theta in (("naive", 1.05), ("BarnesHut", 1.05)):
        s = R.simulate_csv(alg, SOLAR, dt=1.0 / 24, t_end=30.0, vs=10.0, theta=theta, energy=True)
        for k in ("px", "py", "pz", "vx", "vy", "vz", "anorm", "energy"):
            g["solar_sim_%s_%s" % (alg, k)] = s[k]

    # the three-body case of the reference's tests/BarnesHutTest.cpp
    x3, y3, z3, m3 = [0.0, 0.0, 2.0], [1.0, 0.0, 0.0], [0.0, 2.0, 0.0], [10.0, 10.0, 10.0]
    t = R.Tree(m3, x3, y3, z3, builder="subtrees", storage_param=3, stack_param=3, num_wi_octree=3,
               num_wi_top_octree=3, max_level_top_octree=1)
    g["three_body_of_node"] = t.body_of_node
    g["three_sum_masses"] = t.sum_masses
    g["three_sorted"] = t.sorted_bodies
    g["three_aabb"] = t.aabb()
    return g


def main():
    g = generate()
    if "--check" in sys.argv:
        old = np.load(OUT)
        assert sorted(old.files) == sorted(g), "key sets differ"
        bad = [k for k in g if not np.array_equal(np.asarray(g[k]), old[k])]
        assert not bad, "fixture differs from the reference for: %s" % bad
        print("reference_vectors.npz matches the reference bit for bit (%d arrays)" % len(g))
        return
    np.savez_compressed(OUT, **g)
    print("wrote %s: %d arrays, %.1f KB" % (OUT, len(g), os.path.getsize(OUT) / 1e3))


if __name__ == "__main__":
    main()
