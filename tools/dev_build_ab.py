"""A/B of the tree-build phases on one GPU: com_variant and the tile size of the packed sort (NB_OS_ITEMS).
usage: NB_OS_ITEMS=8|12|16 python tools/dev_build_ab.py [N] [com variants comma separated]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nb = importlib.import_module("n-body-simulation_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
variants = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0").split(",")]
m, x, y, z, vx, vy, vz = nb.generators.uniform_sphere(n, seed=1, velocity_scale=0.3)
base = None
for cv in variants:
    c = nb.Context(theta=0.5, wg_size_barnes_hut=128, com_variant=cv)
    c.set_bodies(m, x, y, z, vx, vy, vz)
    c.bh_build(); c.bh_accel_range(0, 4096); c.synchronize()
    c.enable_timers(True)
    acc = {}
    for s in range(6):
        c.leapfrog_part1(1e-3); c.bh_build(); c.bh_accel_range(0, 64); c.synchronize()
        t = c.timers()
        if s >= 2:
            for k, v in t.items():
                acc.setdefault(k, []).append(v)
    c.bh_accel_range(0, 65536)
    a = np.stack(c.accelerations())[:, :1]   # placeholder read: forces a sync
    info = c.bh_tree_info()
    c2 = c.bh_export_canonical() if n <= (1 << 20) else None
    chk = (int(info.num_internal), int(info.max_depth))
    print("N=%d NB_OS_ITEMS=%s emit=%s com_variant=%d: %s tree %s" % (n, os.environ.get("NB_OS_ITEMS", "8"), "per-body" if os.environ.get("NB_EMIT_PER_BODY") else "balanced", cv,
          {k: round(float(np.median(v)), 3) for k, v in acc.items() if np.median(v) > 0 and k not in ("Leapfrog Part 1", "Acceleration Kernel Time")}, chk), flush=True)
    c.close()
