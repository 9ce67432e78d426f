"""Per-step time of the time loop at the size of BASELINE configs[0] (178 bodies): nb_advance (CUDA graph) against
call-by-call issue, naive and Barnes-Hut.  usage: python tools/dev_small_n.py [steps]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nb = importlib.import_module("n-body-simulation_b200")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
m, x, y, z, vx, vy, vz = nb.generators.solar_like(178)
dt = 1.0 / 24
for alg in ("naive", "BarnesHut"):
    c = nb.Context(theta=1.05)
    c.set_bodies(m, x, y, z, vx, vy, vz)
    if alg == "naive":
        c.naive_accel()
    else:
        c.bh_build(); c.bh_accel()
    c.advance(alg, dt, 64); c.synchronize()
    l0 = c.launch_count()
    t0 = time.perf_counter(); c.advance(alg, dt, steps); c.synchronize(); t1 = time.perf_counter()
    launches = c.launch_count() - l0
    t2 = time.perf_counter()
    for _ in range(200):
        c.leapfrog_part1(dt)
        if alg == "naive":
            c.naive_accel()
        else:
            c.bh_build(); c.bh_accel()
        c.leapfrog_part2(dt)
    c.synchronize(); t3 = time.perf_counter()
    print("N=178 %-9s nb_advance %.1f us/step (%.1f kernel launches per step)   call by call %.1f us/step" %
          (alg, (t1 - t0) / steps * 1e6, launches / steps, (t3 - t2) / 200 * 1e6), flush=True)
    c.close()
