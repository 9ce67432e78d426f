"""A/B of a full Barnes-Hut step on one GPU: phase timers of separate calls and the per-step time of nb_advance.
usage: python tools/dev_ab_step.py [N] [generator] [theta] [steps]   (NB_LIB selects the library build)"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nb = importlib.import_module("n-body-simulation_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
gen = sys.argv[2] if len(sys.argv) > 2 else "uniform_sphere"
theta = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 6
kw = dict(velocity_scale=0.3) if gen == "uniform_sphere" else {}
m, x, y, z, vx, vy, vz = getattr(nb.generators, gen)(n, seed=1, **kw)
c = nb.Context(theta=theta, wg_size_barnes_hut=128)
c.set_bodies(m, x, y, z, vx, vy, vz)
dt = 1e-3
c.bh_build(); c.bh_accel(); c.synchronize()
c.enable_timers(True)
acc = {}
for s in range(steps):
    c.leapfrog_part1(dt); c.bh_build(); c.bh_accel(); c.leapfrog_part2(dt)
    t = c.timers()
    if s >= 2:
        for k, v in t.items():
            acc.setdefault(k, []).append(v)
print(os.environ.get("NB_LIB", "current"), "N", n, gen, "theta", theta)
print("  separate calls, phase ms (median):", {k: round(float(np.median(v)), 3) for k, v in acc.items() if np.median(v) > 0})
c.enable_timers(False)
for label, k in (("advance", steps),):
    c.event_record(0); c.advance("BarnesHut", dt, k); c.event_record(1)
    ms = c.event_elapsed_ms(0, 1)
    c.event_record(0); c.advance("BarnesHut", dt, k); c.event_record(1)
    ms = c.event_elapsed_ms(0, 1)
    print("  nb_advance(%d steps): %.3f ms per step" % (k, ms / k))
a = np.stack(c.accelerations()); p = np.stack(c.positions())
print("  checksum sum|a| %.17g sum|x| %.17g" % (np.abs(a).sum(), np.abs(p).sum()))
c.close()
