"""Times the all-pairs kernel instantiations (naive_variant) on one GPU and checks them against the default.
usage: python tools/dev_naive_sweep.py [N] [variants comma separated] [tile lengths comma separated]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nb = importlib.import_module("n-body-simulation_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
variants = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0,1,2,3").split(",")]
tiles = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "256").split(",")]
m, x, y, z, vx, vy, vz = nb.generators.plummer(n, seed=1)
base = None
for tl in tiles:
    for nv in variants:
        c = nb.Context(block_size=tl, naive_variant=nv)
        c.set_bodies(m, x, y, z, vx, vy, vz)
        c.enable_timers(True)
        ts = []
        for _ in range(4):
            c.naive_accel(); c.synchronize()
            ts.append(c.timers()["Acceleration Kernel Time"])
        a = np.stack(c.accelerations())
        if base is None:
            base = a
        t = min(ts[1:])
        print("N=%d tile=%d naive_variant=%2d  %.3f ms  %.4g interactions/s  %.2f TFLOP/s (21-flop)  identical=%s maxrel=%.1e" %
              (n, tl, nv, t, n * float(n) / (t * 1e-3), 21.0 * n * n / (t * 1e-3) / 1e12, bool(np.array_equal(a, base)),
               float(np.abs(a - base).max() / np.abs(base).max())), flush=True)
        c.close()
