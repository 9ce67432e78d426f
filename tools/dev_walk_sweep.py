"""Times the Barnes-Hut traversal variants on one GPU and checks that they produce identical accelerations.
usage: python tools/dev_walk_sweep.py [N] [variants comma separated] [wg sizes comma separated]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nb = importlib.import_module("n-body-simulation_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
variants = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "20,50,0").split(",")]
wgs = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "128").split(",")]
gen = sys.argv[4] if len(sys.argv) > 4 else "uniform_sphere"
m, x, y, z, vx, vy, vz = getattr(nb.generators, gen)(n, seed=1)
base = None
for wv in variants:
    for wg in wgs:
        c = nb.Context(theta=0.5, wg_size_barnes_hut=wg, walk_variant=wv)
        c.set_bodies(m, x, y, z, vx, vy, vz)
        c.enable_timers(True)
        ts = []
        for _ in range(4):
            c.bh_build(); c.bh_accel(); c.synchronize()
            ts.append(c.timers()["Acceleration Kernel Time"])
        a = np.stack(c.accelerations())
        if base is None:
            base = a
        same = bool(np.array_equal(a, base))
        err = float(np.abs(a - base).max() / np.abs(base).max())
        print("%s N=%d %s walk_variant=%d wg=%d traversal ms: %s  identical=%s maxdiff=%.2e sum|a|=%.17g" %
              (os.environ.get("NB_LIB", "current"), n, gen, wv, wg, " ".join("%.2f" % t for t in ts), same, err,
               float(np.abs(a).sum())), flush=True)
        c.close()
