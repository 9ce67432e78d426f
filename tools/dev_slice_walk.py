"""Times the Barnes-Hut walk of ONE rank's slice of a world_size-P run on a single GPU (nb_bh_accel_range), and checks the slice against the full traversal bit for bit.
usage: python tools/dev_slice_walk.py [N] [P] [run lengths comma separated] [ranks comma separated]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nb = importlib.import_module("n-body-simulation_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
P = int(sys.argv[2]) if len(sys.argv) > 2 else 8
runs = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "0").split(",")]   # walk_run_len: 0 = runs of 160 tiles, 1 = one chunk per SM
ranks = [int(v) for v in (sys.argv[4] if len(sys.argv) > 4 else "0,3,7").split(",")]
m, x, y, z, vx, vy, vz = nb.generators.uniform_sphere(n, seed=1)
c = nb.Context(theta=0.5, wg_size_barnes_hut=128)
c.set_bodies(m, x, y, z, vx, vy, vz); c.enable_timers(True)
c.bh_build(); c.bh_accel(); c.synchronize()
full_ms = c.timers()["Acceleration Kernel Time"]
order = np.asarray(c.bh_sorted_bodies())
full = np.stack(c.accelerations())
c.close()
print("N=%d full traversal %.2f ms -> ideal share at P=%d: %.2f ms" % (n, full_ms, P, full_ms / P), flush=True)
for run in runs:
    c = nb.Context(theta=0.5, wg_size_barnes_hut=128, walk_run_len=run)
    c.set_bodies(m, x, y, z, vx, vy, vz); c.enable_timers(True)
    c.bh_build()
    for r in ranks:
        b, e = nb.slice_bounds(n, P, r)
        ts = []
        for _ in range(3):
            c.bh_accel_range(b, e); c.synchronize()
            ts.append(c.timers()["Acceleration Kernel Time"])
        print("run=%d rank %d slots [%d,%d): %s ms" % (run, r, b, e, " ".join("%.2f" % t for t in ts)), flush=True)
    c.close()
