"""How does the time of the persistent walk depend on the number of tile rounds?  One GPU, N = 2^24: slices of the sorted
order whose tile count is a chosen multiple of the resident warps (148 SMs x 40 warps) are walked with nb_bh_accel_range.
A staircase in rounds = quantisation of whole rounds; a line with an offset = a tail of constant length.
usage: python tools/dev_tail.py [N]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nb = importlib.import_module("n-body-simulation_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
m, x, y, z, vx, vy, vz = nb.generators.uniform_sphere(n, seed=1)
c = nb.Context(theta=0.5, wg_size_barnes_hut=128, walk_variant=50)
c.set_bodies(m, x, y, z, vx, vy, vz); c.enable_timers(True)
c.bh_build(); c.bh_accel(); c.synchronize()
full = c.timers()["Acceleration Kernel Time"]
warps = 148 * 40
print("N=%d full walk %.2f ms = %.4f ms per round of %d tiles" % (n, full, full / (n / 32 / warps), warps))
start = (n // 3) // 32 * 32
for rounds in (1, 2, 4, 8, 10, 10.5, 10.9, 11.0, 11.07, 11.25, 11.5, 11.9, 12.0, 12.1, 16, 32):
    tiles = int(round(rounds * warps))
    b, e = start, min(n, start + tiles * 32)
    ts = []
    for _ in range(3):
        c.bh_accel_range(b, e); c.synchronize()
        ts.append(c.timers()["Acceleration Kernel Time"])
    t = min(ts)
    print("rounds %6.2f tiles %7d: %7.3f ms  (%.4f ms per round, %.3f ms above rounds x full-walk rate)" %
          (rounds, tiles, t, t / rounds, t - rounds * full / (n / 32 / warps)), flush=True)
