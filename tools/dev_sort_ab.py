"""A/B of the Barnes-Hut build's radix sort on one GPU: sort_variant 0 (one-sweep) against 1 (three kernels per pass).
Checks that both give the same sorted order and bit-identical accelerations, prints the build phase timers.
usage: python tools/dev_sort_ab.py [N] [generator] [steps] [variants comma separated]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nb = importlib.import_module("n-body-simulation_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
gen = sys.argv[2] if len(sys.argv) > 2 else "uniform_sphere"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
m, x, y, z, vx, vy, vz = getattr(nb.generators, gen)(n, seed=1)
ref = None
variants = [int(v) for v in (sys.argv[4] if len(sys.argv) > 4 else "1,0,1,0").split(",")]
for sv in variants:
    c = nb.Context(theta=0.5, sort_variant=sv)
    c.set_bodies(m, x, y, z, vx, vy, vz)
    c.enable_timers(True)
    rows = []
    for _ in range(steps):
        c.leapfrog_part1(1e-3); c.bh_build(); c.bh_accel(); c.leapfrog_part2(1e-3); c.synchronize()
        t = c.timers()
        rows.append((t["Sort bodies for subtrees"], t["Octree creation"]))
    out = (np.asarray(c.bh_sorted_bodies()), np.stack(c.accelerations()), np.stack(c.positions()))
    if ref is None:
        ref = out
    same = all(np.array_equal(a, b) for a, b in zip(out, ref))
    print("N=%d %s sort_variant=%d  sort ms: %s | build ms: %s  identical=%s" %
          (n, gen, sv, " ".join("%.3f" % r[0] for r in rows), " ".join("%.3f" % r[1] for r in rows), same), flush=True)
    c.close()
