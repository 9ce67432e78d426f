"""Times the Barnes-Hut traversal for several (walk_variant, walk_run_len) pairs: python tools/dev_cfg_sweep.py N v:r,v:r,..."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nb = importlib.import_module("n-body-simulation_b200")
n = int(sys.argv[1]); cfgs = [tuple(int(t) for t in v.split(":")) for v in sys.argv[2].split(",")]
gen = sys.argv[3] if len(sys.argv) > 3 else "uniform_sphere"
m, x, y, z, vx, vy, vz = getattr(nb.generators, gen)(n, seed=1)
base = None
for wv, r in cfgs:
    c = nb.Context(theta=0.5, wg_size_barnes_hut=128, walk_variant=wv, walk_run_len=r)
    c.set_bodies(m, x, y, z, vx, vy, vz); c.enable_timers(True)
    ts = []
    for _ in range(3):
        c.bh_build(); c.bh_accel(); c.synchronize(); ts.append(c.timers()["Acceleration Kernel Time"])
    a = np.stack(c.accelerations())
    base = a if base is None else base
    print("N=%d %s walk_variant=%d run_len=%d: %s identical=%s" % (n, gen, wv, r, " ".join("%.2f" % t for t in ts), bool(np.array_equal(a, base))), flush=True)
    c.close()
