"""Scale stress (BASELINE config 5 shape): Barnes-Hut theta=0.2 at N = 2^26 on one GPU; size-independent checks."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
nb = importlib.import_module("n-body-simulation_b200")
import oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
theta = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
t0 = time.time()
m, x, y, z, vx, vy, vz = nb.generators.uniform_sphere(n, seed=2, velocity_scale=0.3)
print("generated", n, "bodies in %.1f s" % (time.time() - t0), flush=True)
c = nb.Context(theta=theta, wg_size_barnes_hut=128)
c.set_bodies(m, x, y, z, vx, vy, vz); c.enable_timers(True); c.bh_enable_stats(True)
c.bh_build(); c.bh_accel(); c.synchronize()
info = c.bh_tree_info(); tv, ta = c.bh_stats()
print("timers", {k: round(v, 2) for k, v in c.timers().items() if v}, flush=True)
print("nodes", info.num_nodes_materialised, "internal/body %.3f" % (info.num_internal / n), "depth", info.max_depth,
      "visits/body %.1f accepts/body %.1f" % (tv / n, ta / n), flush=True)
a = c.accelerations()
# sampled comparison with the all-pairs oracle (Barnes-Hut error at this theta, not a parity bound)
rows = (n // 2, n // 2 + 64)
ref = O.naive_accel(m, x, y, z, rows=rows)
num = np.sqrt(sum((u[rows[0]:rows[1]] - v[rows[0]:rows[1]]) ** 2 for u, v in zip(a, ref)))
den = np.sqrt(sum(v[rows[0]:rows[1]] ** 2 for v in ref))
print("BH vs all-pairs on 64 sampled bodies: max rel %.3e median %.3e" % ((num / den).max(), np.median(num / den)), flush=True)
for comp in a:
    print("sum m*a / sum |m*a| = %.3e" % (abs((m * comp).sum()) / np.abs(m * comp).sum()))
c.leapfrog_part1(1e-3); c.bh_build(); c.bh_accel(); c.leapfrog_part2(1e-3); c.synchronize()
print("second step timers", {k: round(v, 2) for k, v in c.timers().items() if v}, flush=True)
p = c.positions()
print("positions finite", all(np.isfinite(q).all() for q in p))
