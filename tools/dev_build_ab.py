"""A/B of Barnes-Hut build variants on one GPU (sort_variant: 0 one-sweep / 1 three kernels per pass; com_variant: 0 one
cooperative launch for all centre-of-mass levels / 1 one launch per level).  Checks that every variant gives the same
sorted order and bit-identical accelerations over a few steps of moving bodies; prints the build phase timers.
usage: python tools/dev_build_ab.py [N] [generator] [steps] [variants, e.g. "sort_variant=1;;com_variant=1"]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nb = importlib.import_module("n-body-simulation_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
gen = sys.argv[2] if len(sys.argv) > 2 else "uniform_sphere"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
specs = (sys.argv[4] if len(sys.argv) > 4 else "sort_variant=1,com_variant=1;;sort_variant=1,com_variant=1;").split(";")
m, x, y, z, vx, vy, vz = getattr(nb.generators, gen)(n, seed=1)
ref = None
for spec in specs:
    kw = {k: int(v) for k, v in (kv.split("=") for kv in spec.split(",") if kv)}
    c = nb.Context(theta=0.5, **kw)
    c.set_bodies(m, x, y, z, vx, vy, vz)
    c.enable_timers(True)
    rows = []
    for _ in range(steps):
        c.leapfrog_part1(1e-3); c.bh_build(); c.bh_accel(); c.leapfrog_part2(1e-3); c.synchronize()
        t = c.timers()
        rows.append((t["Sort bodies for subtrees"], t["Build subtrees"], t["Compute center of mass"], t["Octree creation"]))
    out = (np.asarray(c.bh_sorted_bodies()), np.stack(c.accelerations()), np.stack(c.positions()))
    if ref is None:
        ref = out
    same = all(np.array_equal(a, b) for a, b in zip(out, ref))
    r = rows[-1]
    print("N=%d %s %-28s sort %.3f emit %.3f com %.3f | build ms: %s  identical=%s" %
          (n, gen, spec or "(default)", r[0], r[1], r[2], " ".join("%.3f" % q[3] for q in rows), same), flush=True)
    c.close()
